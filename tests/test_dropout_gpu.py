"""Train-mode Dropout / DropPath (jat_audiosr_v2.py:21-47 drop_path, :158 attention-probability
dropout, :250/:252 MLP dropouts, :281/:287 DropPath on the gated branches).

The reference draws its masks from torch's Philox stream, so the masks themselves cannot be reproduced; what is checked:
  * the counter-based masks are Bernoulli(1-p) with 1/(1-p) scaling, independent across sites / blocks / seeds / shifts;
  * every fused kernel applies exactly the mask `jat_dropout_scale_mask` materialises for its site (forward and backward);
  * the whole training step (forward + parameter gradients) equals tests/_torch_dit.py -- which tests/test_oracle.py pins
    against the unmodified reference with injected masks -- run in fp32 with the same masks, to the training-step
    tolerance of tests/test_train_gpu.py (2x the bf16-autocast yardstick);
  * eval mode ignores dropout; p = 0 train mode equals the no-dropout path bit for bit."""
import json
import math
import os

import pytest
import torch

from tests._util import GOLDEN

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from jat_b200 import ops as _ops
    return _ops


@pytest.fixture(scope="module")
def L():
    from jat_b200 import _lib
    return _lib


def dev():
    return torch.device("cuda", 0)


def rel_l2(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


# ------------------------------------------------------------------------------------------------ the masks
@pytest.mark.parametrize("p", [0.1, 0.5, 0.03])
def test_mask_is_bernoulli_and_sites_are_independent(ops, L, p):
    rows, cols = 4096, 1280
    seeds = [ops.dropout_site_seed(1234, blk, site) for blk in (0, 1) for site in (L.DROP_SITE_ATTN, L.DROP_SITE_MLP_HIDDEN,
                                                                                    L.DROP_SITE_MLP_OUT)]
    seeds.append(ops.dropout_site_seed(1235, 0, L.DROP_SITE_ATTN))
    assert len(set(seeds)) == len(seeds)
    ms = [ops.dropout_scale_mask(rows, cols, p, s, dev()) for s in seeds]
    n = rows * cols
    sigma = math.sqrt(p * (1 - p) / n)
    for m in ms:
        vals = torch.unique(m)
        p_eff = round(p * 65536) / 65536  # 16-bit lanes: the effective rate is within 2^-17 of the requested one
        assert vals.numel() == 2 and vals[0] == 0 and abs(vals[1].item() - 1 / (1 - p_eff)) < 1e-6
        keep = (m != 0).float().mean().item()
        assert abs(keep - (1 - p)) < 6 * sigma, (keep, 1 - p)
        # per-row and per-column keep rates: no structure along either axis
        assert ((m != 0).float().mean(1) - (1 - p)).abs().max() < 7 * math.sqrt(p * (1 - p) / cols)
        assert ((m != 0).float().mean(0) - (1 - p)).abs().max() < 7 * math.sqrt(p * (1 - p) / rows)
    k = [(m != 0).float() - (1 - p) for m in ms]
    var = p * (1 - p)
    lim = 6 / math.sqrt(n)   # correlation coefficient of independent masks ~ N(0, 1/n)
    for i in range(len(k)):
        for j in range(i + 1, len(k)):
            assert abs((k[i] * k[j]).mean().item() / var) < lim, (i, j)
        # shifted copies of the same site: neighbouring elements are independent
        assert abs((k[i][:, 1:] * k[i][:, :-1]).mean().item() / var) < lim
        assert abs((k[i][1:] * k[i][:-1]).mean().item() / var) < lim
        assert abs((k[i][1:, 1:] * k[i][:-1, :-1]).mean().item() / var) < lim
    assert torch.equal(ops.dropout_scale_mask(8, 16, 0.0, seeds[0], dev()), torch.ones(8, 16, device=dev()))


def test_drop_path_scales(ops):
    depth, B = 28, 4096
    rates = torch.linspace(0, 0.2, depth, device=dev())
    s = ops.drop_path_scales(rates, B, 77)
    assert s.shape == (depth, 2, B)
    assert torch.equal(s[0], torch.ones(2, B, device=dev()))  # rate 0 -> identity
    for i in (1, 13, 27):
        r = rates[i].item()
        for br in (0, 1):
            vals = torch.unique(s[i, br])
            assert vals.numel() == 2 and vals[0] == 0 and abs(vals[1].item() - 1 / (1 - r)) < 1e-6
            assert abs((s[i, br] != 0).float().mean().item() - (1 - r)) < 6 * math.sqrt(r * (1 - r) / B)
        a, b = (s[i, 0] != 0).float() - (1 - r), (s[i, 1] != 0).float() - (1 - r)
        assert abs((a * b).mean().item() / (r * (1 - r))) < 6 / math.sqrt(B)  # the two branches draw independently
    assert not torch.equal(s, ops.drop_path_scales(rates, B, 78))
    assert torch.equal(s, ops.drop_path_scales(rates, B, 77))


# ------------------------------------------------------------------------------------------------ fused kernels
def test_gemm_bias_act_dropout_and_dgrad(ops, L):
    torch.manual_seed(3)
    M, K, N, p = 700, 256, 512, 0.2
    seed = ops.dropout_site_seed(9, 3, L.DROP_SITE_MLP_HIDDEN)
    A = torch.randn(M, K, device=dev()).to(torch.bfloat16)
    W = (torch.randn(N, K, device=dev()) / math.sqrt(K)).to(torch.bfloat16)
    bias = torch.randn(N, device=dev())
    mask = ops.dropout_scale_mask(M, N, p, seed, dev())
    u = torch.empty(M, N, dtype=torch.bfloat16, device=dev())
    got = ops.gemm(A, W, bias=bias, act=L.ACT_GELU_ERF, aux=u, drop_p=p, drop_seed=seed)
    pre = A.float() @ W.float().T + bias
    want = torch.nn.functional.gelu(pre) * mask
    assert torch.equal(got == 0, (mask == 0) | (got == 0))
    assert ((got != 0) & (mask == 0)).sum() == 0
    assert rel_l2(got.float(), want) < 6e-3
    assert rel_l2(u.float(), pre) < 4e-3  # the saved pre-activation is NOT masked
    # backward through dropout + GELU: dU = (dY W2) * mask * gelu'(u)
    W2 = (torch.randn(256, N, device=dev()) / math.sqrt(N)).to(torch.bfloat16)   # [out, hidden]
    dY = torch.randn(M, 256, device=dev()).to(torch.bfloat16)
    du = ops.gemm(dY, W2, kind=L.EPI_DACT, act=L.ACT_GELU_ERF, aux=u, w_transposed=True, drop_p=p, drop_seed=seed)
    uf = u.float().requires_grad_(True)
    (torch.nn.functional.gelu(uf) * mask).backward(dY.float() @ W2.float())
    assert rel_l2(du.float(), uf.grad) < 6e-3


@pytest.mark.parametrize("cta_pair,block_n", [(1, 256), (0, 128)])
def test_training_shape_fc1_and_gelu_dgrad_epilogues(ops, L, cta_pair, block_n):
    """The training step's fc1 forward (pre-activation copy + dropout) and GELU' dgrad at their real shape (M = 9660 token
    rows, 1280 -> 5120): 760 tiles on the persistent grid.  With 256-wide CTA-pair tiles these two epilogues run on 16
    epilogue warps with one single-buffered staging slab per warp (gemm_tcgen05_kernel, EW = 16); (0, 128) is the 8-warp
    schedule on the same data.  Ragged last row block (9660 = 37 x 256 + 188)."""
    torch.manual_seed(13)
    M, K, N, p = 9660, 1280, 5120, 0.1
    seed = ops.dropout_site_seed(5, 7, L.DROP_SITE_MLP_HIDDEN)
    A = torch.randn(M, K, device=dev()).to(torch.bfloat16)
    W = (torch.randn(N, K, device=dev()) / math.sqrt(K)).to(torch.bfloat16)
    bias = torch.randn(N, device=dev())
    mask = ops.dropout_scale_mask(M, N, p, seed, dev())
    u = torch.full((M, N), float("nan"), dtype=torch.bfloat16, device=dev())
    got = ops.gemm(A, W, bias=bias, act=L.ACT_GELU_ERF, aux=u, drop_p=p, drop_seed=seed, cta_pair=cta_pair, block_n=block_n)
    pre = A.float() @ W.float().T + bias
    assert torch.isfinite(u.float()).all() and rel_l2(u.float(), pre) < 4e-3
    assert rel_l2(got.float(), torch.nn.functional.gelu(pre) * mask) < 6e-3
    plain = ops.gemm(A, W, bias=bias, act=L.ACT_GELU_ERF, cta_pair=cta_pair, block_n=block_n)      # 8-warp epilogue, no extras
    assert torch.equal(got[mask != 0], (plain.float() * mask).to(torch.bfloat16)[mask != 0]) or \
        rel_l2(got.float(), plain.float() * mask) < 4e-3
    W2 = (torch.randn(K, N, device=dev()) / math.sqrt(N)).to(torch.bfloat16)   # [out, hidden]
    dY = torch.randn(M, K, device=dev()).to(torch.bfloat16)
    du = ops.gemm(dY, W2, kind=L.EPI_DACT, act=L.ACT_GELU_ERF, aux=u, w_transposed=True, drop_p=p, drop_seed=seed,
                  cta_pair=cta_pair, block_n=block_n)
    uf = u.float().requires_grad_(True)
    (torch.nn.functional.gelu(uf) * mask).backward(dY.float() @ W2.float())
    assert rel_l2(du.float(), uf.grad) < 6e-3


@pytest.mark.parametrize("with_path", [False, True])
def test_gemm_gate_residual_dropout_and_gate_bwd(ops, L, with_path):
    torch.manual_seed(4)
    B, Ntok, K, D, p = 3, 97, 512, 1280, 0.15
    M = B * Ntok
    seed = ops.dropout_site_seed(9, 1, L.DROP_SITE_MLP_OUT)
    A = torch.randn(M, K, device=dev()).to(torch.bfloat16)
    W = (torch.randn(D, K, device=dev()) / math.sqrt(K)).to(torch.bfloat16)
    bias = torch.randn(D, device=dev())
    gate = torch.randn(B, D, device=dev())
    rs = torch.tensor([1.25, 0.0, 1.25], device=dev()) if with_path else None
    x0 = torch.randn(M, D, device=dev())
    x = x0.clone()
    y = torch.empty(M, D, dtype=torch.bfloat16, device=dev())
    mask = ops.dropout_scale_mask(M, D, p, seed, dev())
    ops.gemm(A, W, kind=L.EPI_GATE_RESIDUAL, bias=bias, out=x, gate=gate, gate_batch_stride=D, tokens_per_batch=Ntok, aux=y,
             drop_p=p, drop_seed=seed, gate_rowscale=rs)
    yd = (A.float() @ W.float().T + bias) * mask
    geff = gate * (rs[:, None] if with_path else 1.0)
    want = x0 + geff.repeat_interleave(Ntok, 0) * yd
    assert rel_l2(x - x0, want - x0) < 2e-5
    assert rel_l2(y.float(), yd) < 4e-3
    if with_path:
        assert torch.equal(x[Ntok:2 * Ntok], x0[Ntok:2 * Ntok])  # the dropped sample's branch contributes exactly 0
    # gate backward with the same mask / row scale
    dx = torch.randn(M, D, device=dev())
    dgate = torch.zeros(B, D, device=dev())
    dbias = torch.zeros(D, device=dev())
    dy = ops.gate_bwd(dx, y, gate, B, Ntok, dgate, mod_batch_stride=D, dmod_batch_stride=D, dbias=dbias, drop_p=p,
                      drop_seed=seed, gate_rowscale=rs)
    pre = (A.float() @ W.float().T).requires_grad_(True)
    b_ = bias.clone().requires_grad_(True)
    g_ = gate.clone().requires_grad_(True)
    ge = g_ * (rs[:, None] if with_path else 1.0)
    out = ge.repeat_interleave(Ntok, 0) * ((pre + b_) * mask)
    out.backward(dx)
    assert rel_l2(dy.float(), pre.grad) < 4e-3
    assert rel_l2(dbias, b_.grad) < 1e-5
    # dgate uses the saved (bf16) y
    want_dg = ((dx * y.float()).view(B, Ntok, D).sum(1)) * (rs[:, None] if with_path else 1.0)
    assert rel_l2(dgate, want_dg) < 1e-5
    assert rel_l2(dgate, g_.grad) < 4e-3


def _rope_tables(npos):
    inv = 1.0 / (10000 ** (torch.arange(0, 64, 2, device=dev()).float() / 64))
    f = torch.outer(torch.arange(npos, device=dev()).float(), inv)
    e = torch.cat([f, f], -1)
    return e.cos().contiguous(), e.sin().contiguous()


def _rope(x, cos, sin):
    x1, x2 = x[..., :32], x[..., 32:]
    return x * cos[None, :, None, :] + torch.cat([-x2, x1], -1) * sin[None, :, None, :]


@pytest.mark.parametrize("B,N,Hq,Hkv", [(2, 345, 20, 4), (1, 22, 8, 4), (3, 129, 4, 2), (1, 300, 5, 1)])
def test_gqa_attention_dropout_fwd_bwd(ops, L, B, N, Hq, Hkv):
    torch.manual_seed(41)
    p = 0.1
    seed = ops.dropout_site_seed(5, 2, L.DROP_SITE_ATTN)
    G = Hq // Hkv
    cos, sin = _rope_tables(N)
    raw = (torch.randn(B, N, Hq + 2 * Hkv, 64, device=dev())).to(torch.bfloat16).float().requires_grad_(True)
    q = _rope(raw[:, :, :Hq], cos, sin)
    k = _rope(raw[:, :, Hq:Hq + Hkv], cos, sin)
    v = raw[:, :, Hq + Hkv:]
    qkv = torch.cat([q, k, v], 2).detach().reshape(B * N, -1).to(torch.bfloat16)
    lse = torch.empty(B, Hq, N, device=dev())
    out = ops.gqa_attention_fwd(qkv, B, N, Hq, Hkv, lse=lse, drop_p=p, drop_seed=seed)
    mask = ops.dropout_scale_mask(B * Hq * N, N, p, seed, dev()).view(B, Hq, N, N)
    s = torch.einsum("bnhd,bmhd->bhnm", q, k.repeat_interleave(G, dim=2)) / 8.0
    o = torch.einsum("bhnm,bmhd->bnhd", torch.softmax(s, -1) * mask, v.repeat_interleave(G, dim=2)).reshape(B * N, Hq * 64)
    assert rel_l2(out.float(), o.detach()) < 8e-3
    # the normaliser is the undropped softmax sum (jat_audiosr_v2.py:157-158: softmax, THEN dropout)
    assert (lse - torch.logsumexp(s, -1).detach() * math.log2(math.e)).abs().max() < 2e-2
    no_drop = ops.gqa_attention_fwd(qkv, B, N, Hq, Hkv)
    assert rel_l2(out.float(), no_drop.float()) > 0.05
    d_out = torch.randn(B * N, Hq * 64, device=dev()).to(torch.bfloat16)
    got = ops.gqa_attention_bwd(qkv, d_out, out, lse, cos, sin, B, N, Hq, Hkv, drop_p=p, drop_seed=seed)
    got = got.float().view(B, N, Hq + 2 * Hkv, 64)
    o.backward(d_out.float())
    for name, sl in (("dq", slice(0, Hq)), ("dk", slice(Hq, Hq + Hkv)), ("dv", slice(Hq + Hkv, Hq + 2 * Hkv))):
        err = rel_l2(got[:, :, sl], raw.grad[:, :, sl])
        assert err < 1.5e-2, (name, err)


# ------------------------------------------------------------------------------------------------ the training step
def _masks(ops, L, model, seed, B, N, p):
    depth = len(model.blocks)
    D, F_ = model.hidden_size, model.blocks[0].mlp[0].out_features
    Hq = model.blocks[0].attn.num_q_heads
    m = {"attn": [], "hid": [], "out": []}
    for i in range(depth):
        m["attn"].append(ops.dropout_scale_mask(B * Hq * N, N, p, ops.dropout_site_seed(seed, i, L.DROP_SITE_ATTN),
                                                dev()).view(B, Hq, N, N))
        m["hid"].append(ops.dropout_scale_mask(B * N, F_, p, ops.dropout_site_seed(seed, i, L.DROP_SITE_MLP_HIDDEN), dev()))
        m["out"].append(ops.dropout_scale_mask(B * N, D, p, ops.dropout_site_seed(seed, i, L.DROP_SITE_MLP_OUT), dev()))
    rates = torch.tensor([float(getattr(b.drop_path, "drop_prob", 0.0)) for b in model.blocks], device=dev())
    m["path"] = ops.drop_path_scales(rates, B, seed) if (rates > 0).any() else None
    return m


@pytest.mark.parametrize("cls,rms", [("JaT_AudioSR_V2", False), ("JaT_AudioSR_V3", True)])
@pytest.mark.parametrize("p,dpr", [(0.1, 0.0), (0.0, 0.3), (0.1, 0.3)])
def test_training_step_with_dropout_matches_masked_restatement(ops, L, cls, rms, p, dpr):
    import jat_b200
    from tests._torch_dit import dit_forward
    from tests._util import rerandomise_zero_init
    yard = json.load(open(os.path.join(GOLDEN, "bf16_autocast_yardstick.json")))["grads"]
    yard = yard["v3_rmsnorm" if rms else "v2_layernorm"]
    cfg = dict(input_channels=32, cond_channels=32, patch_len=4, hidden_size=256, depth=4, num_q_heads=4, num_kv_heads=2,
               bottleneck_dim=128, mlp_ratio=4.0, dropout=p, drop_path_rate=dpr)
    torch.manual_seed(0)
    model = rerandomise_zero_init(getattr(jat_b200, cls)(**cfg)).to(dev()).train()
    B, T = 6, 170
    N = (T + 3) // 4
    g = torch.Generator(device=dev()).manual_seed(3)
    hr, lr, eps = (torch.randn(B, 32, T, generator=g, device=dev()) for _ in range(3))
    t = torch.rand(B, generator=g, device=dev())
    z_t = t.view(B, 1, 1) * hr + (1 - t.view(B, 1, 1)) * eps
    torch.manual_seed(11)
    pred = model(z_t, t, lr)
    torch.nn.functional.mse_loss(pred, hr).backward()
    seed = model._train_seed
    masks = _masks(ops, L, model, seed, B, N, p)
    if p == 0.0:
        masks["attn"] = masks["hid"] = masks["out"] = None
    prm = {k: v.detach().float().clone().requires_grad_(v.dtype.is_floating_point and "rope" not in k)
           for k, v in model.state_dict().items()}
    want = dit_forward(prm, cfg, z_t, t, lr, rms=rms, masks=masks)
    torch.nn.functional.mse_loss(want, hr).backward()
    assert rel_l2(pred.detach(), want.detach()) <= 2 * yard["pred_rel_l2"]
    num = den = 0.0
    worst = ("", 0.0)
    for k, q in model.named_parameters():
        e = rel_l2(q.grad, prm[k].grad)
        worst = max(worst, (k, e), key=lambda kv: kv[1])
        num += (q.grad.double() - prm[k].grad.double()).pow(2).sum().item()
        den += prm[k].grad.double().pow(2).sum().item()
    assert (num / den) ** 0.5 <= 2 * yard["global_rel_l2"], ((num / den) ** 0.5, worst)
    assert worst[1] <= 2 * yard["max_param_rel_l2"], worst
    # masks really were applied: the unmasked restatement is far away
    plain = dit_forward({k: v.detach() for k, v in prm.items()}, cfg, z_t, t, lr, rms=rms)
    assert rel_l2(pred.detach(), plain) > 10 * rel_l2(pred.detach(), want.detach())


def test_dropout_is_off_in_eval_and_reseeds_every_step():
    import jat_b200
    from tests._util import rerandomise_zero_init
    cfg = dict(input_channels=32, cond_channels=32, patch_len=4, hidden_size=128, depth=2, num_q_heads=2, num_kv_heads=1,
               bottleneck_dim=128, mlp_ratio=2.0, dropout=0.1, drop_path_rate=0.1)
    torch.manual_seed(0)
    model = rerandomise_zero_init(jat_b200.JaT_AudioSR_V2(**cfg)).to(dev())
    ref = jat_b200.JaT_AudioSR_V2(**dict(cfg, dropout=0.0, drop_path_rate=0.0)).to(dev())
    ref.load_state_dict(model.state_dict())
    x, c = torch.randn(2, 32, 86, device=dev()), torch.randn(2, 32, 86, device=dev())
    t = torch.rand(2, device=dev())
    model.eval(), ref.eval()
    with torch.no_grad():
        assert torch.equal(model(x, t, c), ref(x, t, c))
    model.train(), ref.train()
    torch.manual_seed(1)
    a = model(x, t, c).detach().clone()
    b = model(x, t, c).detach().clone()
    assert not torch.equal(a, b)                       # new masks every step
    torch.manual_seed(1)
    assert torch.equal(model(x, t, c).detach(), a)     # ... controlled by torch.manual_seed like the reference's
    r = ref(x, t, c).detach()
    assert rel_l2(a, r) > 1e-2
