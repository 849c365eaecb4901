"""Shared helpers for the parity tests."""
import ast
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")
REFERENCE = "/root/reference"


def load_golden(tag):
    d = np.load(os.path.join(GOLDEN, f"micro_{tag}.npz"), allow_pickle=False)
    cfg = ast.literal_eval(str(d["cfg_json"]))
    weights = {k[3:]: d[k] for k in d.files if k.startswith("w::")}
    return d, cfg, weights


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def have_reference():
    return os.path.isdir(os.path.join(REFERENCE, "src", "models"))


def import_reference():
    """The unmodified reference modules (CPU).  Only available in the build container."""
    import sys
    import types
    for p in (REFERENCE, os.path.join(REFERENCE, "src")):
        if p not in sys.path:
            sys.path.insert(0, p)
    sys.modules.setdefault("dac", types.ModuleType("dac"))
    from src.models.jat_audiosr_v2 import JaT_AudioSR_V2
    from src.models.jat_audiosr_v3 import JaT_AudioSR_V3
    import infer_test_v3m2
    return JaT_AudioSR_V2, JaT_AudioSR_V3, infer_test_v3m2.flow_matching_sample


def rerandomise_zero_init(model, seed=1, bf16_exact=True):
    """Re-randomise the zero-initialised layers (adaLN-Zero + final Linear), otherwise the model outputs 0
    and any parity test is vacuous."""
    import torch
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, prm in model.named_parameters():
            if "adaLN_modulation.1" in name or name.startswith("final_layer.1"):
                prm.copy_((torch.randn(prm.shape, generator=g) * 0.02).to(prm.device))
            if name.endswith(("norm1.weight", "norm2.weight")) or name == "final_layer.0.weight":
                prm.copy_((1.0 + 0.1 * torch.randn(prm.shape, generator=g)).to(prm.device))
        if bf16_exact:
            for prm in model.parameters():
                prm.copy_(prm.to(torch.bfloat16).to(torch.float32))
    return model


def state_dict_numpy(model):
    return {k: v.detach().float().cpu().numpy() for k, v in model.state_dict().items()}
