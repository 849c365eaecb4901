"""GPU parity tests of the product path (nn.Module facade -> C-ABI -> sm_100a kernels) against
(a) the committed golden fixtures produced by the reference itself and (b) the numpy oracle on the
same seeded inputs.

Stated bf16 tolerance: torch's own bf16-autocast of the reference differs from
the fp32 reference by rel-L2 0.42 % after block 0, 0.77 % after block 8 and 0.90 % at the output of the
12-block v1 model (tests/golden/bf16_autocast_yardstick.json, measured with make_golden.py).  The
CUDA path (bf16 operands, fp32 accumulate, fp32 residual stream, fp32 norm / softmax statistics) is
held to <= 1.5x that yardstick: per-block rel-L2 <= 1.2 %, output rel-L2 <= 1.35 %, and max-abs
<= 0.05 on O(1) activations.
"""
import contextlib
import io
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import dit_oracle as O  # noqa: E402  (checker only)
from tests._util import load_golden, rel_l2, rerandomise_zero_init, state_dict_numpy  # noqa: E402

pytestmark = pytest.mark.gpu

BLOCK_TOL = 0.012
OUT_TOL = 0.0135
MAXABS_TOL = 0.05


def dev():
    return torch.device("cuda", 0)


def build(cls_name, cfg, seed, bf16_exact=True):
    import jat_b200
    torch.manual_seed(seed)
    m = getattr(jat_b200, cls_name)(**cfg)
    rerandomise_zero_init(m, seed + 1, bf16_exact=bf16_exact)
    return m.eval()


@pytest.mark.parametrize("tag,cls", [("v2_layernorm", "JaT_AudioSR_V2"), ("v3_rmsnorm", "JaT_AudioSR_V3")])
def test_forward_matches_reference_golden(tag, cls):
    import jat_b200
    d, cfg, w = load_golden(tag)
    model = getattr(jat_b200, cls)(**cfg)
    sd = {k: torch.from_numpy(v) for k, v in w.items()}
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected and all(k.endswith(("cos_cached", "sin_cached")) for k in missing)
    model = model.to(dev()).eval()
    x_t, cond, t = (torch.from_numpy(d[k]).to(dev()) for k in ("x_t", "cond", "t"))
    out, blocks = model.forward_with_blocks(x_t, t, cond)
    B, N = x_t.shape[0], d["blocks"].shape[2]
    for i in range(cfg["depth"]):
        got = blocks[i].view(B, N, -1).cpu().numpy()
        assert rel_l2(got, d["blocks"][i]) <= BLOCK_TOL, (i, rel_l2(got, d["blocks"][i]))
    got = out.cpu().numpy()
    assert got.shape == d["out"].shape
    assert rel_l2(got, d["out"]) <= OUT_TOL, rel_l2(got, d["out"])
    assert np.abs(got - d["out"]).max() <= MAXABS_TOL
    # the public forward() returns the same tensor
    assert torch.equal(model(x_t, t, cond), out)


@pytest.mark.parametrize("tag,cls", [("v2_layernorm", "JaT_AudioSR_V2"), ("v3_rmsnorm", "JaT_AudioSR_V3")])
@pytest.mark.parametrize("name", ["cfg3", "cfg1"])
@pytest.mark.parametrize("use_graph", [False, True])
def test_sampler_matches_reference_golden(tag, cls, name, use_graph):
    import jat_b200
    d, cfg, w = load_golden(tag)
    model = getattr(jat_b200, cls)(**cfg)
    model.load_state_dict({k: torch.from_numpy(v) for k, v in w.items()}, strict=False)
    model = model.to(dev()).eval()
    lr = torch.from_numpy(d["s_lr"]).to(dev())
    steps, scale = int(d[f"s_{name}_steps"]), float(d[f"s_{name}_cfg"])
    # drop-in contract: same global-generator draw as the reference (infer_test_v3m2.py:133) ...
    torch.manual_seed(123)
    z_seeded = torch.randn(lr.shape, device=dev())
    # ... but the golden z0 came from the CPU generator, so inject it for the value comparison
    z0 = torch.from_numpy(d[f"s_{name}_z0"])
    got = jat_b200.flow_matching_sample(model, lr, num_steps=steps, cfg_scale=scale, device="cuda", verbose=False,
                                        use_graph=use_graph, z0=z0)
    want = d[f"s_{name}_out"]
    assert got.shape == tuple(want.shape) and got.dtype == torch.float32
    assert rel_l2(got.cpu().numpy(), want) <= 0.02, rel_l2(got.cpu().numpy(), want)
    torch.manual_seed(123)
    got_seeded = jat_b200.flow_matching_sample(model, lr, num_steps=steps, cfg_scale=scale, device="cuda",
                                               verbose=False, use_graph=use_graph)
    again = jat_b200.flow_matching_sample(model, lr, num_steps=steps, cfg_scale=scale, device="cuda",
                                          verbose=False, use_graph=use_graph, z0=z_seeded)
    assert torch.equal(got_seeded, again)  # consumes exactly one torch.randn(B, C, T) from the global generator


CASES = [
    # BASELINE config C1: v1 tiny (train_ddp_v1.py:64-76), [1, 1024, 86]
    ("c1_v1tiny", "JaT_AudioSR_V2", dict(hidden_size=512, depth=12, num_q_heads=8, num_kv_heads=4, bottleneck_dim=512), 1, 86),
    # v3mod2 width / heads (1280, 20Q/4KV), 2 blocks, the headline chunk length T=1378 -> N=345
    ("v3mod2_2blk_ln", "JaT_AudioSR_V2", dict(hidden_size=1280, depth=2, num_q_heads=20, num_kv_heads=4, bottleneck_dim=512), 2, 1378),
    ("v3mod2_2blk_rms", "JaT_AudioSR_V3", dict(hidden_size=1280, depth=2, num_q_heads=20, num_kv_heads=4, bottleneck_dim=512), 2, 1378),
    # v2 width (1024, 16Q/4KV), ragged T (517 -> pad 3), 3 batch items
    ("v2_ragged", "JaT_AudioSR_V3", dict(hidden_size=1024, depth=2, num_q_heads=16, num_kv_heads=4, bottleneck_dim=512), 3, 517),
    # single-frame input: T=1 -> one token
    ("one_token", "JaT_AudioSR_V2", dict(hidden_size=512, depth=1, num_q_heads=8, num_kv_heads=4, bottleneck_dim=512), 2, 1),
]


@pytest.mark.parametrize("name,cls,cfg,B,T", CASES, ids=[c[0] for c in CASES])
def test_forward_matches_oracle(name, cls, cfg, B, T):
    model = build(cls, cfg, seed=1000 + [c[0] for c in CASES].index(name))
    w = state_dict_numpy(model)
    g = torch.Generator().manual_seed(9)
    x_t, cond = torch.randn(B, 1024, T, generator=g), torch.randn(B, 1024, T, generator=g)
    t = torch.rand(B, generator=g)
    want, want_blocks = O.dit_forward(w, x_t.numpy(), t.numpy(), cond.numpy(), num_q_heads=cfg["num_q_heads"],
                                      num_kv_heads=cfg["num_kv_heads"], return_blocks=True)
    model = model.to(dev())
    out, blocks = model.forward_with_blocks(x_t.to(dev()), t.to(dev()), cond.to(dev()))
    N = want_blocks[0].shape[1]
    errs = [rel_l2(blocks[i].view(B, N, -1).cpu().numpy(), want_blocks[i]) for i in range(cfg["depth"])]
    assert max(errs) <= BLOCK_TOL, errs
    got = out.cpu().numpy()
    assert np.abs(want).max() > 0.05
    assert rel_l2(got, want) <= OUT_TOL, rel_l2(got, want)
    assert np.abs(got - want).max() <= MAXABS_TOL * max(1.0, np.abs(want).max())


def test_sampler_matches_oracle_v1tiny_cfg():
    """BASELINE config C1 widened to a 4-step CFG=3.0 run: final-latent rel-L2 vs the fp32 oracle."""
    import jat_b200
    cfg = dict(hidden_size=512, depth=12, num_q_heads=8, num_kv_heads=4, bottleneck_dim=512)
    model = build("JaT_AudioSR_V2", cfg, seed=31)
    w = state_dict_numpy(model)
    g = torch.Generator().manual_seed(10)
    lr, z0 = torch.randn(1, 1024, 86, generator=g), torch.randn(1, 1024, 86, generator=g)
    want = O.flow_matching_sample(w, lr.numpy(), z0.numpy(), num_steps=4, cfg_scale=3.0, num_q_heads=8, num_kv_heads=4)
    got = jat_b200.flow_matching_sample(model.to(dev()), lr.to(dev()), num_steps=4, cfg_scale=3.0, verbose=False, z0=z0)
    assert rel_l2(got.cpu().numpy(), want) <= 0.02, rel_l2(got.cpu().numpy(), want)


def test_full_size_properties_v3mod2():
    """BASELINE config C3 at FULL size (766 M params, B_eff = 56, M = 19320 rows): properties that do
    not need the oracle -- batch-position invariance (a row's result does not depend on which M tile
    it lands in), CFG halves == separate cond / zero-cond forwards, determinism, finiteness."""
    cfg = dict(hidden_size=1280, depth=28, num_q_heads=20, num_kv_heads=4, bottleneck_dim=512)
    model = build("JaT_AudioSR_V2", cfg, seed=77, bf16_exact=False).to(dev())
    assert sum(p.numel() for p in model.parameters()) == 766_125_312
    g = torch.Generator().manual_seed(11)
    B = 28
    z = torch.randn(B, 1024, 1378, generator=g).to(dev())
    lr = torch.randn(B, 1024, 1378, generator=g).to(dev())
    t = torch.full((2 * B,), 0.37, device=dev())
    big = model(torch.cat([z, z]), t, torch.cat([lr, torch.zeros_like(lr)]))
    assert torch.isfinite(big).all() and big.abs().max() > 0.01
    assert torch.equal(big, model(torch.cat([z, z]), t, torch.cat([lr, torch.zeros_like(lr)])))  # deterministic
    small_c = model(z[5:7], t[:2], lr[5:7])
    small_u = model(z[26:28], t[:2], torch.zeros_like(lr[26:28]))
    assert torch.equal(big[5:7], small_c)
    assert torch.equal(big[B + 26:B + 28], small_u)


def test_cpu_tensors_fail_loudly():
    model = build("JaT_AudioSR_V2", dict(hidden_size=512, depth=1, num_q_heads=8, num_kv_heads=4), seed=1)
    with pytest.raises(RuntimeError):
        model(torch.zeros(1, 1024, 8), torch.zeros(1), torch.zeros(1, 1024, 8))


def test_max_len_raises_value_error():
    model = build("JaT_AudioSR_V2", dict(input_channels=32, cond_channels=32, hidden_size=128, depth=1, num_q_heads=2,
                                         num_kv_heads=1, bottleneck_dim=128), seed=1).to(dev())
    x = torch.zeros(1, 32, 4 * 2049, device=dev())
    with pytest.raises(ValueError):
        model(x, torch.zeros(1, device=dev()), x)


@pytest.mark.parametrize("cls", ["JaT_AudioSR_V2", "JaT_AudioSR_V3"])
def test_forward_long_sequence_matches_oracle(cls):
    """T = 2000 frames -> N = 500 tokens (> 352: the chunked attention path inside the whole-model plan), eval forward
    against the numpy oracle; and the training forward + backward run at that length."""
    from oracle import dit_oracle as O
    cfg = dict(hidden_size=256, depth=2, num_q_heads=4, num_kv_heads=2, bottleneck_dim=128, input_channels=64, cond_channels=64)
    model = build(cls, cfg, seed=21)
    g = torch.Generator().manual_seed(4)
    B, T = 2, 2000
    x, c, t = torch.randn(B, 64, T, generator=g), torch.randn(B, 64, T, generator=g), torch.rand(B, generator=g)
    w = {k: v.detach().float().cpu().numpy() for k, v in model.state_dict().items()}
    want = O.dit_forward(w, x.numpy(), t.numpy(), c.numpy(), num_q_heads=4, num_kv_heads=2, patch_len=4)
    model = model.to(dev())
    got = model(x.to(dev()), t.to(dev()), c.to(dev()))
    assert rel_l2(got.cpu().numpy(), want) < 1.5e-2
    model.train()
    for p_ in model.parameters():
        p_.requires_grad_(True)
    model.dropout_p = 0.0
    out = model(x.to(dev()), t.to(dev()), c.to(dev()))
    out.float().pow(2).mean().backward()
    assert all(p_.grad is not None and torch.isfinite(p_.grad).all() for p_ in model.parameters())


@pytest.mark.parametrize("cls", ["JaT_AudioSR_V2", "JaT_AudioSR_V3"])
def test_cond_channels_differ_from_input_channels(cls):
    """Constructor contract `cond_channels != input_channels` (jat_audiosr_v2.py:297-308; the patch embed sees
    input + cond channels, :198-210): eval forward and the CFG sampler against the numpy oracle, training step gradients
    against fp32 autograd of the torch restatement."""
    import jat_b200
    from tests._torch_dit import dit_forward
    cfg = dict(input_channels=32, cond_channels=96, hidden_size=256, depth=2, num_q_heads=4, num_kv_heads=2, bottleneck_dim=128)
    model = build(cls, cfg, seed=5)
    assert model.patch_embed.proj[0].in_features == (32 + 96) * 4
    w = state_dict_numpy(model)
    g = torch.Generator().manual_seed(6)
    B, T = 3, 171
    x, c, t = torch.randn(B, 32, T, generator=g), torch.randn(B, 96, T, generator=g), torch.rand(B, generator=g)
    want = O.dit_forward(w, x.numpy(), t.numpy(), c.numpy(), num_q_heads=4, num_kv_heads=2)
    model = model.to(dev())
    got = model(x.to(dev()), t.to(dev()), c.to(dev()))
    assert got.shape == (B, 32, T) and rel_l2(got.detach().cpu().numpy(), want) <= OUT_TOL
    with pytest.raises(ValueError):
        model(x.to(dev()), t.to(dev()), x.to(dev()))        # condition with the wrong channel count
    z0 = torch.randn(B, 32, T, generator=g)
    want_z = O.flow_matching_sample(w, c.numpy(), z0.numpy(), num_steps=3, cfg_scale=2.0, num_q_heads=4, num_kv_heads=2)
    got_z = jat_b200.flow_matching_sample(model, c.to(dev()), num_steps=3, cfg_scale=2.0, verbose=False, z0=z0)
    assert got_z.shape == (B, 32, T) and rel_l2(got_z.cpu().numpy(), want_z) <= 0.02
    # training step
    model.train()
    model.dropout_p = 0.0
    hr = torch.randn(B, 32, T, generator=g).to(dev())
    torch.nn.functional.mse_loss(model(x.to(dev()), t.to(dev()), c.to(dev())), hr).backward()
    prm = {k: v.detach().float().clone().requires_grad_(v.dtype.is_floating_point and "rope" not in k)
           for k, v in model.state_dict().items()}
    full = dict(cfg, patch_len=4)
    torch.nn.functional.mse_loss(dit_forward(prm, full, x.to(dev()), t.to(dev()), c.to(dev()), rms=(cls == "JaT_AudioSR_V3")), hr).backward()
    for k, q in model.named_parameters():
        e = (q.grad - prm[k].grad).norm() / prm[k].grad.norm().clamp_min(1e-20)
        assert e < 0.03, (k, float(e))


def test_default_tail_split_mode_matches_the_bit_reproducible_schedule():
    """Library default outside this suite: GEMM tail split mode 2 (the K-split parts of the tiles of a partial last wave add
    their partial sums straight into the residual stream).  At the headline token count (M = 19320: 380 out_proj / fc2 tiles
    on 74 CTA pairs = 5 waves + 10 tiles) the f32 residual stream differs from the bit-reproducible schedule in the last bit
    of some elements; downstream that flips an occasional bf16 rounding of the next GEMM operand, so the OUTPUT differs by a
    few 1e-5 relative -- two orders of magnitude below the bf16 error budget of the path (3e-3, see the module docstring).
    Stated tolerance: rel-L2 <= 3e-4, max-abs <= 5e-3 on O(1) outputs.  torch.use_deterministic_algorithms(True) selects the
    bit-reproducible schedule."""
    from jat_b200 import ops
    cfg = dict(hidden_size=1280, depth=2, num_q_heads=20, num_kv_heads=4, bottleneck_dim=512)
    model = build("JaT_AudioSR_V2", cfg, seed=3, bf16_exact=False).to(dev())
    g = torch.Generator().manual_seed(2)
    B = 56
    x, c = torch.randn(B, 1024, 1378, generator=g).to(dev()), torch.randn(B, 1024, 1378, generator=g).to(dev())
    t = torch.rand(B, generator=g).to(dev())
    try:
        ops.set_gemm_tail_split(dev(), 0)
        want = model(x, t, c)
        assert torch.equal(want, model(x, t, c))
        ops.set_gemm_tail_split(dev(), 2)
        got = model(x, t, c)
        assert rel_l2(got.cpu().numpy(), want.cpu().numpy()) <= 3e-4
        assert (got - want).abs().max().item() <= 5e-3 * max(1.0, want.abs().max().item())
        torch.use_deterministic_algorithms(True)
        assert torch.equal(model(x, t, c), want)          # the engine pushes mode 0 into the context
    finally:
        torch.use_deterministic_algorithms(False)
        ops.set_gemm_tail_split(dev(), 0)
