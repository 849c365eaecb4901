"""CPU tests pinning the numpy oracle (oracle/dit_oracle.py) to the reference:
  * against the committed golden fixtures (outputs of the reference modules, tests/golden/make_golden.py);
  * directly against the reference modules when /root/reference is present (build container only).
Tolerance: the oracle and the reference are both fp32 restatements of the same formulas; they differ
only by summation order inside BLAS, so max-abs 2e-5 / rel-L2 1e-5 on O(1) activations."""
import contextlib
import io
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import dit_oracle as O  # noqa: E402
from tests._util import (have_reference, import_reference, load_golden, rel_l2, rerandomise_zero_init,  # noqa: E402
                         state_dict_numpy)


@pytest.mark.parametrize("tag", ["v2_layernorm", "v3_rmsnorm"])
def test_oracle_forward_matches_golden(tag):
    d, cfg, w = load_golden(tag)
    out, blocks = O.dit_forward(w, d["x_t"], d["t"], d["cond"], num_q_heads=cfg["num_q_heads"],
                                num_kv_heads=cfg["num_kv_heads"], patch_len=cfg["patch_len"], return_blocks=True)
    assert out.shape == d["out"].shape
    assert np.abs(d["out"]).max() > 0.1  # not the vacuous all-zero output of an un-randomised adaLN-Zero model
    assert np.abs(out - d["out"]).max() < 2e-5
    assert rel_l2(out, d["out"]) < 1e-5
    for i, b in enumerate(blocks):
        assert rel_l2(b, d["blocks"][i]) < 1e-5, i


@pytest.mark.parametrize("tag", ["v2_layernorm", "v3_rmsnorm"])
@pytest.mark.parametrize("name", ["cfg3", "cfg1"])
def test_oracle_sampler_matches_golden(tag, name):
    d, cfg, w = load_golden(tag)
    z = O.flow_matching_sample(w, d["s_lr"], d[f"s_{name}_z0"], num_steps=int(d[f"s_{name}_steps"]),
                               cfg_scale=float(d[f"s_{name}_cfg"]), num_q_heads=cfg["num_q_heads"],
                               num_kv_heads=cfg["num_kv_heads"], patch_len=cfg["patch_len"])
    assert np.abs(z - d[f"s_{name}_out"]).max() < 5e-5
    assert rel_l2(z, d[f"s_{name}_out"]) < 1e-5


def test_oracle_schedule_is_torch_linspace():
    import torch
    for n in (1, 2, 5, 25, 50, 100, 1000):
        assert np.array_equal(O.sampler_schedule(n), torch.linspace(0.0, 1.0, n + 1).numpy())


def test_oracle_euler_update_matches_torch_expression():
    import torch
    g = torch.Generator().manual_seed(0)
    z, xc, xu = (torch.randn(3, 5, 7, generator=g) for _ in range(3))
    ts = torch.linspace(0.0, 1.0, 51)
    for i in (0, 13, 49):
        t, dt = ts[i], ts[i + 1] - ts[i]
        x = xu + 3.0 * (xc - xu)
        want = z + (x - z) / (1 - t + 1e-5) * dt  # infer_test_v3m2.py:164,175-176
        got = O.euler_cfg_update(z.numpy(), xc.numpy(), xu.numpy(), 3.0, t.numpy(), dt.numpy())
        assert np.array_equal(got, want.numpy())


def test_oracle_rejects_long_sequences():
    _, cfg, w = load_golden("v2_layernorm")
    x = np.zeros((1, cfg["input_channels"], 4 * 2049), np.float32)
    with pytest.raises(ValueError):
        O.dit_forward(w, x, np.zeros(1, np.float32), x, num_q_heads=cfg["num_q_heads"],
                      num_kv_heads=cfg["num_kv_heads"])


# ------------------------------------------------------------------ direct comparison with the reference
REF_CASES = [
    # (class idx, cfg, B, T)  -- BASELINE config C1 (v1 tiny, [1,1024,86]) and a GQA case with 2 KV groups
    (0, dict(input_channels=1024, cond_channels=1024, patch_len=4, hidden_size=512, depth=12, num_q_heads=8,
             num_kv_heads=4, bottleneck_dim=512), 1, 86),
    (1, dict(input_channels=64, cond_channels=64, patch_len=4, hidden_size=256, depth=3, num_q_heads=4,
             num_kv_heads=2, bottleneck_dim=64, mlp_ratio=2.0), 2, 517),
    (0, dict(input_channels=64, cond_channels=64, patch_len=4, hidden_size=320, depth=2, num_q_heads=5,
             num_kv_heads=1, bottleneck_dim=64), 2, 88),
    # condition latent with its own channel count (constructor contract, jat_audiosr_v2.py:297-308)
    (1, dict(input_channels=32, cond_channels=96, patch_len=4, hidden_size=256, depth=2, num_q_heads=4,
             num_kv_heads=2, bottleneck_dim=64), 2, 171),
]


@pytest.mark.skipif(not have_reference(), reason="/root/reference not present (GPU box)")
@pytest.mark.parametrize("case", range(len(REF_CASES)))
def test_oracle_matches_reference_module(case):
    import torch
    cls_idx, cfg, B, T = REF_CASES[case]
    classes = import_reference()
    torch.manual_seed(case)
    with contextlib.redirect_stdout(io.StringIO()):
        model = classes[cls_idx](**cfg).eval()
    rerandomise_zero_init(model, seed=case + 1, bf16_exact=False)
    g = torch.Generator().manual_seed(100 + case)
    x_t = torch.randn(B, cfg["input_channels"], T, generator=g)
    cond = torch.randn(B, cfg["cond_channels"], T, generator=g)
    t = torch.rand(B, generator=g)
    with torch.no_grad():
        want = model(x_t, t, cond).numpy()
    got = O.dit_forward(state_dict_numpy(model), x_t.numpy(), t.numpy(), cond.numpy(),
                        num_q_heads=cfg["num_q_heads"], num_kv_heads=cfg["num_kv_heads"])
    assert np.abs(want).max() > 0.05
    assert np.abs(got - want).max() < 3e-5
    assert rel_l2(got, want) < 1e-5
    # the torch restatement (fp32 truth of the GPU parity tests) on the same case
    from tests._torch_dit import dit_forward
    with torch.no_grad():
        got_t = dit_forward({k: v.detach() for k, v in model.state_dict().items()}, dict(cfg, patch_len=cfg.get("patch_len", 4)),
                            x_t, t, cond, rms=(cls_idx == 1)).numpy()
    assert rel_l2(got_t, want) < 1e-5


@pytest.mark.skipif(not have_reference(), reason="/root/reference not present (GPU box)")
def test_oracle_sampler_matches_reference_sampler():
    import torch
    cfg = dict(input_channels=64, cond_channels=64, patch_len=4, hidden_size=256, depth=2, num_q_heads=4,
               num_kv_heads=2, bottleneck_dim=64, mlp_ratio=2.0)
    V2, V3, ref_sample = import_reference()
    torch.manual_seed(3)
    with contextlib.redirect_stdout(io.StringIO()):
        model = V3(**cfg).eval()
    rerandomise_zero_init(model, seed=4, bf16_exact=False)
    lr = torch.randn(2, 64, 86, generator=torch.Generator().manual_seed(5))
    torch.manual_seed(77)
    z0 = torch.randn(2, 64, 86)
    torch.manual_seed(77)
    want = ref_sample(model, lr, num_steps=6, cfg_scale=3.0, device="cpu", verbose=False).numpy()
    got = O.flow_matching_sample(state_dict_numpy(model), lr.numpy(), z0.numpy(), num_steps=6, cfg_scale=3.0,
                                 num_q_heads=4, num_kv_heads=2)
    assert np.abs(got - want).max() < 1e-4
    assert rel_l2(got, want) < 1e-5


# ------------------------------------------------------------------------------------------------
# tests/_torch_dit.py (the differentiable restatement used by the train-mode dropout parity tests on the GPU) pinned
# against the unmodified reference: eval mode, and TRAIN mode with the reference's nn.Dropout / drop_path draws replaced
# by given masks -- which pins where each mask enters the computation (jat_audiosr_v2.py:158,250,252,281,287).
# ------------------------------------------------------------------------------------------------
@pytest.mark.skipif(not have_reference(), reason="reference not mounted")
@pytest.mark.parametrize("cls_idx,rms", [(0, False), (1, True)])
def test_torch_restatement_with_masks_matches_reference(cls_idx, rms):
    import torch
    from tests._torch_dit import dit_forward
    ref_cls = import_reference()[cls_idx]
    cfg = dict(input_channels=16, cond_channels=16, patch_len=4, hidden_size=128, depth=2, num_q_heads=2, num_kv_heads=1,
               bottleneck_dim=64, mlp_ratio=2.0, dropout=0.25, drop_path_rate=0.5)
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        model = rerandomise_zero_init(ref_cls(**cfg), bf16_exact=False)
    B, T = 3, 30
    N, D, F_ = 8, 128, 256
    g = torch.Generator().manual_seed(5)
    x_t, cond = torch.randn(B, 16, T, generator=g), torch.randn(B, 16, T, generator=g)
    t = torch.rand(B, generator=g)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    model.eval()
    with torch.no_grad():
        want = model(x_t, t, cond)
        got = dit_forward(sd, cfg, x_t, t, cond, rms=rms)
    assert rel_l2(got.numpy(), want.numpy()) < 1e-5

    def bern(shape, p):
        return (torch.rand(shape, generator=g) >= p).float() / (1 - p)
    masks = {"attn": [bern((B, 2, N, N), 0.25) for _ in range(2)], "hid": [bern((B * N, F_), 0.25) for _ in range(2)],
             "out": [bern((B * N, D), 0.25) for _ in range(2)], "path": bern((2, 2, B), 0.5)}

    class Mask(torch.nn.Module):
        def __init__(self, m):
            super().__init__()
            self.m = m

        def forward(self, x):
            return x * self.m.view(x.shape)

    class PathMask(torch.nn.Module):  # called twice per block forward: attention branch, then MLP branch
        def __init__(self, ms):
            super().__init__()
            self.ms, self.n = ms, 0

        def forward(self, x):
            m = self.ms[self.n % 2]
            self.n += 1
            return x * m[:, None, None]
    mod = sys.modules[ref_cls.__module__]
    for i, blk in enumerate(model.blocks):
        blk.attn.dropout = Mask(masks["attn"][i])
        blk.mlp[2] = Mask(masks["hid"][i])
        blk.mlp[4] = Mask(masks["out"][i])
        blk.drop_path = PathMask([masks["path"][i, 0], masks["path"][i, 1]])
    assert hasattr(mod, "drop_path")
    model.train()
    params = {k: v.clone().requires_grad_(v.dtype.is_floating_point and "rope" not in k) for k, v in sd.items()}
    got = dit_forward(params, cfg, x_t, t, cond, rms=rms, masks=masks)
    want = model(x_t, t, cond)
    assert rel_l2(got.detach().numpy(), want.detach().numpy()) < 1e-5
    hr = torch.randn(B, 16, T, generator=g)
    torch.nn.functional.mse_loss(got, hr).backward()
    torch.nn.functional.mse_loss(want, hr).backward()
    for k, prm in model.named_parameters():
        assert rel_l2(params[k].grad.numpy(), prm.grad.numpy()) < 2e-4, k


@pytest.mark.skipif(not have_reference(), reason="reference not mounted")
def test_torch_restatement_sampler_matches_reference_sampler():
    """tests/_torch_dit.flow_matching_sample (the fp32 truth of the full-size GPU parity test) == the reference's
    flow_matching_sample (infer_test_v3m2.py:108-185) under the same seed, CFG and no-CFG."""
    import torch
    from tests._torch_dit import flow_matching_sample as restated
    ref_cls, _, ref_sampler = import_reference()
    cfg = dict(input_channels=16, cond_channels=16, patch_len=4, hidden_size=128, depth=2, num_q_heads=2, num_kv_heads=1,
               bottleneck_dim=64, mlp_ratio=2.0, dropout=0.0, drop_path_rate=0.0)
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        model = rerandomise_zero_init(ref_cls(**cfg), bf16_exact=False).eval()
    sd = {k: v.detach() for k, v in model.state_dict().items()}
    lr = torch.randn(2, 16, 30, generator=torch.Generator().manual_seed(1))
    for scale in (3.0, 1.0):
        torch.manual_seed(42)
        z0 = torch.randn(2, 16, 30)
        torch.manual_seed(42)
        with contextlib.redirect_stdout(io.StringIO()):
            want = ref_sampler(model, lr, num_steps=7, cfg_scale=scale, device="cpu", verbose=False)
        got = restated(sd, cfg, lr, z0, num_steps=7, cfg_scale=scale)
        assert rel_l2(got.numpy(), want.numpy()) < 1e-5, scale


def test_optimizer_oracle_matches_torch_clip_plus_adamw():
    """oracle/optimizer_oracle.py against what the reference's loop calls (train_ddp_v3mod2.py:709, 926-928) on CPU:
    clip_grad_norm_(params, 1.0) + torch.optim.AdamW.step(), several steps, one parameter frozen for the first two."""
    import torch
    from oracle import optimizer_oracle as OO
    rng = np.random.default_rng(0)
    shapes = [(64, 48), (48,), (7,), (3, 5, 2)]
    params = [torch.nn.Parameter(torch.from_numpy(rng.standard_normal(s).astype(np.float32))) for s in shapes]
    opt = torch.optim.AdamW(params, lr=3e-3, betas=(0.9, 0.95), eps=1e-8, weight_decay=0.1, foreach=False)
    mine = [p.detach().numpy().copy() for p in params]
    m = [np.zeros_like(x) for x in mine]
    v = [np.zeros_like(x) for x in mine]
    steps = [0] * len(mine)
    for step in range(5):
        grads = [rng.standard_normal(s).astype(np.float32) * (5.0 if step % 2 else 0.05) for s in shapes]
        live = [i for i in range(len(shapes)) if not (i == 1 and step < 2)]
        for i, p in enumerate(params):
            p.grad = torch.from_numpy(grads[i].copy()) if i in live else None
        want_norm = float(torch.nn.utils.clip_grad_norm_(params, 1.0))
        opt.step()
        sub = lambda xs: [xs[i] for i in live]
        st = [steps[i] for i in live]
        got_norm = OO.adamw_step(sub(mine), sub(grads), sub(m), sub(v), st, lr=3e-3, betas=(0.9, 0.95), eps=1e-8,
                                weight_decay=0.1, max_norm=1.0)
        for k, i in enumerate(live):
            steps[i] = st[k]
        assert abs(got_norm - want_norm) <= 1e-5 * want_norm
        for i, p in enumerate(params):
            np.testing.assert_allclose(mine[i], p.detach().numpy(), rtol=2e-6, atol=1e-6)
            if i in live or steps[i] > 0:
                np.testing.assert_allclose(m[i], opt.state[p]["exp_avg"].numpy(), rtol=2e-6, atol=1e-7)
                np.testing.assert_allclose(v[i], opt.state[p]["exp_avg_sq"].numpy(), rtol=2e-6, atol=1e-7)
    assert steps == [5, 3, 5, 5]
