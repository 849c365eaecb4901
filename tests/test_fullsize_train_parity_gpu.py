"""Gradient parity of the training step at the FULL v3mod2 size (766 M parameters, hidden 1280, depth 28, 20Q/4KV, G = 5,
N = 345 tokens) -- SURVEY.md 8a row a14 / BASELINE config C4.  This is the shape at which the shipped training path makes
its size-dependent choices: the weight-gradient GEMMs pick their split-K factor from the makespan model, dQ leaves the
attention backward by TMA reduce-add over three key tiles, and the key-stationary loop walks 15 (head, q-tile) pairs.

Truth: fp32 autograd of tests/_torch_dit.py on the same GPU (TF32 off) -- the restatement tests/test_oracle.py pins
against the unmodified reference modules (eval mode and train mode with injected masks).  Yardstick, measured in the same
test: the gradients torch's own bf16 autocast produces for the same step.

Stated tolerance (bf16 operands, fp32 accumulate / statistics / residual stream): prediction rel-L2, global gradient
rel-L2 and the worst per-parameter rel-L2 <= 2 x the corresponding bf16-autocast error.  Two cases: the deterministic step
(dropout = drop_path = 0) and the reference's training regularisers (Dropout 0.1 / DropPath 0.05) with the CUDA path's
counter-based masks materialised and injected into the restatement."""
import json
import os

import pytest
import torch

from tests._util import rerandomise_zero_init

pytestmark = pytest.mark.gpu
CFG = dict(input_channels=1024, cond_channels=1024, patch_len=4, hidden_size=1280, depth=28, num_q_heads=20, num_kv_heads=4,
           bottleneck_dim=512, mlp_ratio=4.0)


def rel_l2(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


def _masks(model, seed, B, N, p, dev):
    from jat_b200 import _lib as L, ops
    D, F_ = model.hidden_size, model.blocks[0].mlp[0].out_features
    Hq = model.blocks[0].attn.num_q_heads
    m = {"attn": None, "hid": None, "out": None}
    if p > 0:
        m = {"attn": [], "hid": [], "out": []}
        for i in range(len(model.blocks)):
            m["attn"].append(ops.dropout_scale_mask(B * Hq * N, N, p, ops.dropout_site_seed(seed, i, L.DROP_SITE_ATTN),
                                                    dev).view(B, Hq, N, N))
            m["hid"].append(ops.dropout_scale_mask(B * N, F_, p, ops.dropout_site_seed(seed, i, L.DROP_SITE_MLP_HIDDEN), dev))
            m["out"].append(ops.dropout_scale_mask(B * N, D, p, ops.dropout_site_seed(seed, i, L.DROP_SITE_MLP_OUT), dev))
    rates = torch.tensor([float(getattr(b.drop_path, "drop_prob", 0.0)) for b in model.blocks], device=dev)
    m["path"] = ops.drop_path_scales(rates, B, seed) if (rates > 0).any() else None
    return m


def _grad_errors(got, want):
    num = den = 0.0
    worst = ("", 0.0)
    for k in want:
        e = rel_l2(got[k], want[k])
        if e > worst[1]:
            worst = (k, e)
        num += (got[k].double() - want[k].double()).pow(2).sum().item()
        den += want[k].double().pow(2).sum().item()
    return (num / den) ** 0.5, worst


@pytest.mark.parametrize("cls,rms,p,dpr", [("JaT_AudioSR_V2", False, 0.0, 0.0), ("JaT_AudioSR_V2", False, 0.1, 0.05),
                                           ("JaT_AudioSR_V3", True, 0.1, 0.05)])
def test_fullsize_training_step_gradients(cls, rms, p, dpr):
    import jat_b200
    from tests._torch_dit import dit_forward
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    dev = torch.device("cuda", 0)
    cfg = dict(CFG, dropout=p, drop_path_rate=dpr)
    torch.manual_seed(0)
    model = rerandomise_zero_init(getattr(jat_b200, cls)(**cfg), bf16_exact=False).to(dev).train()
    B, T = 2, 1378
    N = (T + 3) // 4
    g = torch.Generator(device=dev).manual_seed(17)
    hr, lr, eps = (torch.randn(B, 1024, T, generator=g, device=dev) for _ in range(3))
    t = torch.tensor([0.23, 0.71], device=dev)
    z_t = t.view(B, 1, 1) * hr + (1 - t.view(B, 1, 1)) * eps

    torch.manual_seed(5)
    pred = model(z_t, t, lr)
    torch.nn.functional.mse_loss(pred, hr).backward()
    got = {k: q.grad.detach().clone() for k, q in model.named_parameters()}
    assert all(torch.isfinite(v).all() for v in got.values())
    masks = _masks(model, model._train_seed, B, N, p, dev) if (p > 0 or dpr > 0) else None
    pred = pred.detach().clone()
    model.zero_grad(set_to_none=True)

    def truth(autocast):
        prm = {k: v.detach().float().clone().requires_grad_(v.dtype.is_floating_point and "rope" not in k)
               for k, v in model.state_dict().items()}
        if autocast:
            with torch.autocast("cuda", dtype=torch.bfloat16):
                out = dit_forward(prm, cfg, z_t, t, lr, rms=rms, masks=masks)
                loss = torch.nn.functional.mse_loss(out.float(), hr)
        else:
            out = dit_forward(prm, cfg, z_t, t, lr, rms=rms, masks=masks)
            loss = torch.nn.functional.mse_loss(out, hr)
        loss.backward()
        return out.detach().float(), {k: prm[k].grad for k in got}

    want_pred, want = truth(False)
    ac_pred, ac = truth(True)
    assert want_pred.abs().max() > 0.05
    ours_glob, ours_worst = _grad_errors(got, want)
    yard_glob, yard_worst = _grad_errors(ac, want)
    report = {"class": cls, "B": B, "T": T, "dropout": p, "drop_path": dpr,
              "pred_rel_l2": rel_l2(pred, want_pred), "pred_autocast_rel_l2": rel_l2(ac_pred, want_pred),
              "grad_global_rel_l2": ours_glob, "grad_global_autocast_rel_l2": yard_glob,
              "grad_worst_param": ours_worst, "grad_worst_param_autocast": yard_worst}
    print("FULLSIZE_TRAIN_PARITY " + json.dumps(report))
    out_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(out_dir):
        json.dump(report, open(os.path.join(out_dir, f"fullsize_train_parity_{cls}_p{p}.json"), "w"), indent=1)
    assert report["pred_rel_l2"] <= 2 * report["pred_autocast_rel_l2"], report
    assert ours_glob <= 2 * yard_glob, report
    assert ours_worst[1] <= 2 * yard_worst[1], report
