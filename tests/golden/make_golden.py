"""Generates tests/golden/*.npz from the UNMODIFIED reference, imported from /root/reference on CPU.

Run in the build container only (the GPU box has no /root/reference):
    python tests/golden/make_golden.py

The reference has no golden vectors of its own, so these fixtures are outputs of the
reference's own modules (fp32, eval mode) on seeded inputs:
  * JaT_AudioSR_V2 (LayerNorm) and JaT_AudioSR_V3 (RMSNorm) at a micro configuration,
    forward output + every block's output (forward hooks on model.blocks[i]);
  * flow_matching_sample (infer_test_v3m2.py) with CFG 3.0 / 5 steps and CFG 1.0 / 3 steps.
Zero-initialised layers (adaLN_modulation.1, final_layer.1 -- jat_audiosr_v2.py:372-381) are
re-randomised, otherwise the model outputs exactly 0.  All weights are rounded to bf16-representable
values so the same fixture can drive the bf16 CUDA path without a weight-quantisation term.
"""
import contextlib
import io
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))

MICRO = dict(input_channels=32, cond_channels=32, patch_len=4, hidden_size=128, depth=2, num_q_heads=2,
             num_kv_heads=1, bottleneck_dim=128, mlp_ratio=2.0, dropout=0.1, drop_path_rate=0.05)


def import_reference():
    sys.path.insert(0, REF)
    sys.path.insert(0, os.path.join(REF, "src"))
    sys.modules.setdefault("dac", types.ModuleType("dac"))  # infer_test_v3m2.py:30 imports it at module level
    from src.models.jat_audiosr_v2 import JaT_AudioSR_V2
    from src.models.jat_audiosr_v3 import JaT_AudioSR_V3
    import infer_test_v3m2
    return JaT_AudioSR_V2, JaT_AudioSR_V3, infer_test_v3m2.flow_matching_sample


def rerandomise(model, seed=1):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, prm in model.named_parameters():
            if "adaLN_modulation.1" in name or name.startswith("final_layer.1"):
                prm.copy_(torch.randn(prm.shape, generator=g) * 0.02)
            if name.endswith("norm1.weight") or name.endswith("norm2.weight") or name == "final_layer.0.weight":
                prm.copy_(1.0 + 0.1 * torch.randn(prm.shape, generator=g))
        for prm in model.parameters():  # bf16-representable weights
            prm.copy_(prm.to(torch.bfloat16).to(torch.float32))


def build(cls, cfg, seed):
    torch.manual_seed(seed)
    with contextlib.redirect_stdout(io.StringIO()):
        m = cls(**cfg)
    rerandomise(m, seed + 1)
    return m.eval()


def main():
    V2, V3, ref_sample = import_reference()
    for tag, cls in (("v2_layernorm", V2), ("v3_rmsnorm", V3)):
        model = build(cls, MICRO, seed=10)
        g = torch.Generator().manual_seed(5)
        B, C, T = 2, MICRO["input_channels"], 86  # 86 -> padded to 88 -> 22 tokens (BASELINE config C1's T)
        x_t = torch.randn(B, C, T, generator=g)
        cond = torch.randn(B, C, T, generator=g)
        t = torch.tensor([0.13, 0.98])
        blocks = []
        hooks = [blk.register_forward_hook(lambda _m, _i, o: blocks.append(o.detach().clone())) for blk in model.blocks]
        with torch.no_grad():
            out = model(x_t, t, cond)
        for h in hooks:
            h.remove()
        out_sd = {"w::" + k: v.numpy() for k, v in model.state_dict().items()
                  if not k.endswith(("cos_cached", "sin_cached"))}
        data = dict(out_sd)
        data.update(x_t=x_t.numpy(), cond=cond.numpy(), t=t.numpy(), out=out.numpy(),
                    blocks=torch.stack(blocks).numpy())
        # sampler goldens: the reference draws z0 = torch.randn(B, C, T) from the global generator
        lr = torch.randn(1, C, T, generator=g)
        for name, steps, cfg_scale in (("cfg3", 5, 3.0), ("cfg1", 3, 1.0)):
            torch.manual_seed(123)
            z0 = torch.randn(1, C, T)
            torch.manual_seed(123)
            z = ref_sample(model, lr, num_steps=steps, cfg_scale=cfg_scale, device="cpu", verbose=False)
            data.update({f"s_{name}_z0": z0.numpy(), f"s_{name}_out": z.numpy(),
                         f"s_{name}_steps": np.int64(steps), f"s_{name}_cfg": np.float32(cfg_scale)})
        data["s_lr"] = lr.numpy()
        data["cfg_json"] = np.array(repr(MICRO))
        path = os.path.join(HERE, f"micro_{tag}.npz")
        np.savez_compressed(path, **data)
        print(tag, "out absmax", float(out.abs().max()), "->", path, os.path.getsize(path) // 1024, "KiB")

    # torch bf16-autocast error yardstick on the reference itself (CPU), v1-tiny config (train_ddp_v1.py:64-76)
    V1CFG = dict(input_channels=1024, cond_channels=1024, patch_len=4, hidden_size=512, depth=12, num_q_heads=8,
                 num_kv_heads=4, bottleneck_dim=512, mlp_ratio=4.0, dropout=0.1, drop_path_rate=0.0)
    yard = {}
    for tag, cls in (("v2_layernorm", V2), ("v3_rmsnorm", V3)):
        model = build(cls, V1CFG, seed=20)
        g = torch.Generator().manual_seed(6)
        x_t, cond = torch.randn(2, 1024, 345 * 4 - 2, generator=g), torch.randn(2, 1024, 345 * 4 - 2, generator=g)
        t = torch.tensor([0.3, 0.9])
        res = {}
        for mode in ("fp32", "bf16"):
            blocks = []
            hooks = [blk.register_forward_hook(lambda _m, _i, o: blocks.append(o.detach().float().clone()))
                     for blk in model.blocks]
            with torch.no_grad(), torch.autocast("cpu", dtype=torch.bfloat16, enabled=(mode == "bf16")):
                out = model(x_t, t, cond).float()
            for h in hooks:
                h.remove()
            res[mode] = (out, blocks)
        rel = lambda a, b: float((a - b).norm() / b.norm())
        yard[tag] = dict(out_rel_l2=rel(res["bf16"][0], res["fp32"][0]),
                         block_rel_l2=[rel(a, b) for a, b in zip(res["bf16"][1], res["fp32"][1])])
        print(tag, "torch bf16-autocast vs fp32:", yard[tag]["out_rel_l2"], yard[tag]["block_rel_l2"][::4])
    import json
    json.dump(yard, open(os.path.join(HERE, "bf16_autocast_yardstick.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
