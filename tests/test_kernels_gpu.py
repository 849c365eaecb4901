"""Kernel-level parity tests (GPU): every C-ABI kernel against a plain torch fp32 restatement of the
reference expression it replaces.  Tolerances are stated per test; bf16 outputs are compared after
the same bf16 rounding of the operands, so what is measured is the kernel, not the quantisation."""
import math

import pytest
import torch

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


# ------------------------------------------------------------------------------------------- elementwise
@pytest.mark.parametrize("cfg", [3.0, 1.0])
@pytest.mark.parametrize("shape", [(2, 1024, 1378), (1, 1024, 86), (3, 7, 5)])
def test_cfg_euler_update_bit_exact(ops, cfg, shape):
    g = torch.Generator(device="cpu").manual_seed(0)
    z = torch.randn(shape, generator=g).to(dev())
    xc = torch.randn(shape, generator=g).to(dev())
    xu = torch.randn(shape, generator=g).to(dev())
    steps = 50
    ts = torch.linspace(0.0, 1.0, steps + 1, device=dev())
    t_dt = torch.stack([ts[:-1], ts[1:] - ts[:-1]], dim=1).contiguous()
    for step in (0, 17, 49):
        t, dt = ts[step], ts[step + 1] - ts[step]
        # reference expression, infer_test_v3m2.py:164,175-176
        x = xu + cfg * (xc - xu) if cfg != 1.0 else xc
        want = z + (x - z) / (1 - t + 1e-5) * dt
        got = ops.cfg_euler_update(z.clone(), xc, xu if cfg != 1.0 else None, cfg, t_dt, step)
        assert torch.equal(got, want), (got - want).abs().max().item()


def test_cfg_euler_update_direct_branch(ops):
    z = torch.randn(4, 33, device=dev())
    xc = torch.randn(4, 33, device=dev())
    t_dt = torch.tensor([[0.9995, 0.0005]], device=dev())
    got = ops.cfg_euler_update(z.clone(), xc, None, 1.0, t_dt, 0)
    assert torch.equal(got, xc)  # t >= 0.999 -> z = x_pred (infer_test_v3m2.py:177-179)


@pytest.mark.parametrize("norm_kind", [0, 1])
@pytest.mark.parametrize("D", [512, 1024, 1280])
@pytest.mark.parametrize("mode", ["per_batch", "shared", "none"])
def test_adaln_norm_modulate(ops, norm_kind, D, mode):
    torch.manual_seed(1)
    B, N = 3, 45
    M = B * N
    x = (torch.randn(M, D, device=dev()) * 1.7 + 0.3)
    mod = torch.randn(B, 6 * D, device=dev()) * 0.5
    w = torch.rand(D, device=dev()) + 0.5 if norm_kind == 1 else None
    if norm_kind == 0:
        xh = torch.nn.functional.layer_norm(x, (D,), eps=1e-6)
    else:
        xh = x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + 1e-6) * w
    if mode == "none":
        want = xh
        got = ops.adaln_norm_modulate(x, None, None, 0, w, norm_kind, 1e-6, N)
    else:
        stride = 6 * D if mode == "per_batch" else 0
        shift, scale = mod[:, 3 * D:4 * D], mod[:, 4 * D:5 * D]
        if mode == "shared":
            sh, sc = shift[:1].expand(B, D), scale[:1].expand(B, D)
        else:
            sh, sc = shift, scale
        want = (xh.view(B, N, D) * (1 + sc.unsqueeze(1)) + sh.unsqueeze(1)).view(M, D)
        got = ops.adaln_norm_modulate(x, shift, scale, stride, w, norm_kind, 1e-6, N)
    err = (got.float() - want).abs().max().item()
    # bf16 output rounding: half ulp = 2^-9 relative
    assert err <= 2 ** -8 * want.abs().max().item() + 1e-5, err
    assert rel_l2(got.float(), want) < 3e-3


@pytest.mark.parametrize("T", [1378, 86, 516, 7])
def test_patchify_cast(ops, T):
    torch.manual_seed(2)
    B, Cc, P = 4, 64, 4
    xt = torch.randn(2, Cc, T, device=dev())      # xt_batch = 2 -> rows b read x_t[b % 2]
    cond = torch.randn(2, Cc, T, device=dev())    # cond_batch = 2 -> b >= 2 read zeros
    got = ops.patchify_cast(xt, cond, B)
    pad = (P - T % P) % P
    xt4 = torch.cat([xt, xt], 0)
    c4 = torch.cat([cond, torch.zeros_like(cond)], 0)
    xin = torch.nn.functional.pad(torch.cat([xt4, c4], dim=1), (0, pad))  # jat_audiosr_v2.py:411-421
    N = xin.shape[-1] // P
    want = xin.reshape(B, 2 * Cc, N, P).permute(0, 2, 1, 3).reshape(B * N, 2 * Cc * P)  # :225-227
    assert torch.equal(got, want.to(torch.bfloat16))
    got0 = ops.patchify_cast(xt, None, 2)
    assert torch.equal(got0[:, Cc * P:], torch.zeros_like(got0[:, Cc * P:]))


def test_timestep_features(ops):
    t = torch.tensor([0.0, 0.13, 0.5, 0.98, 1.0], device=dev())
    D = 1280
    half = D // 2
    k = math.log(10000) / (half - 1)
    f = torch.exp(torch.arange(half, device=dev()) * -k)
    e = t[:, None] * f[None, :]
    want = torch.cat([e.sin(), e.cos()], -1)  # jat_audiosr_v2.py:185-189
    got = ops.timestep_features(t, D)
    assert (got.float() - want).abs().max().item() <= 2 ** -8


# ------------------------------------------------------------------------------------------- GEMM
def _ab(M, N, K, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    A = (torch.randn(M, K, generator=g) * 0.5).to(dev()).to(torch.bfloat16)
    W = (torch.randn(N, K, generator=g) * (1.0 / math.sqrt(K))).to(dev()).to(torch.bfloat16)
    bias = torch.randn(N, generator=g).to(dev())
    return A, W, bias


GEMM_CFGS = [pytest.param(0, 256, id="cg1n256"), pytest.param(0, 128, id="cg1n128"),
             pytest.param(1, 256, id="cg2n256"), pytest.param(1, 128, id="cg2n128")]


@pytest.mark.parametrize("cta_pair,block_n", GEMM_CFGS)
@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (300, 512, 1280), (1000, 1280, 512), (56, 7680, 1280)])
def test_gemm_bias_f32(ops, L, cta_pair, block_n, M, N, K):
    A, W, bias = _ab(M, N, K)
    want = A.float() @ W.float().t() + bias
    got = ops.gemm(A, W, bias=bias, out_dtype=L.DTYPE_F32, cta_pair=cta_pair, block_n=block_n)
    err = (got - want).abs().max().item()
    assert err < 2e-3 * max(1.0, want.abs().max().item()), err
    assert rel_l2(got, want) < 1e-4


@pytest.mark.parametrize("cta_pair,block_n", GEMM_CFGS)
@pytest.mark.parametrize("act", [0, 1, 2])
def test_gemm_bias_act_bf16(ops, L, cta_pair, block_n, act):
    M, N, K = 345 * 2, 1024, 256
    A, W, bias = _ab(M, N, K, seed=3)
    y = A.float() @ W.float().t() + bias
    want = [y, torch.nn.functional.gelu(y), torch.nn.functional.silu(y)][act]
    got = ops.gemm(A, W, bias=bias, act=act, out_dtype=L.DTYPE_BF16, cta_pair=cta_pair, block_n=block_n)
    assert (got.float() - want).abs().max().item() <= 2 ** -8 * want.abs().max().item() + 1e-3
    assert rel_l2(got.float(), want) < 4e-3


def _rope_tables(max_pos=4096, dim=64):
    inv_freq = 1.0 / (10000 ** (torch.arange(0, dim, 2).float() / dim))  # jat_audiosr_v2.py:60
    freqs = torch.outer(torch.arange(max_pos).float(), inv_freq)
    emb = torch.cat([freqs, freqs], -1)
    return emb.cos().to(dev()).contiguous(), emb.sin().to(dev()).contiguous()


def _rope_ref(x, cos, sin):
    # x [B, N, H, 64]; jat_audiosr_v2.py:70-91
    N = x.shape[1]
    x1, x2 = x[..., :32], x[..., 32:]
    rot = torch.cat([-x2, x1], -1)
    return x * cos[:N].unsqueeze(0).unsqueeze(2) + rot * sin[:N].unsqueeze(0).unsqueeze(2)


@pytest.mark.parametrize("cta_pair,block_n", GEMM_CFGS)
def test_gemm_qkv_rope(ops, L, cta_pair, block_n):
    B, N, D, Hq, Hkv = 3, 129, 1280, 20, 4
    M, NQ = B * N, (Hq + 2 * Hkv) * 64
    A, W, _ = _ab(M, NQ, D, seed=4)
    cos, sin = _rope_tables()
    y = (A.float() @ W.float().t()).view(B, N, Hq + 2 * Hkv, 64)
    want = torch.cat([_rope_ref(y[:, :, :Hq + Hkv], cos, sin), y[:, :, Hq + Hkv:]], 2).reshape(M, NQ)
    got = ops.gemm(A, W, kind=L.EPI_QKV_ROPE, tokens_per_batch=N, rope_cos=cos, rope_sin=sin,
                   rope_cols=(Hq + Hkv) * 64, cta_pair=cta_pair, block_n=block_n)
    assert (got.float() - want).abs().max().item() <= 2 ** -8 * want.abs().max().item() + 1e-3
    assert rel_l2(got.float(), want) < 4e-3


@pytest.mark.parametrize("cta_pair,block_n", GEMM_CFGS)
@pytest.mark.parametrize("with_bias", [False, True])
@pytest.mark.parametrize("stride_mode", ["per_batch", "shared"])
def test_gemm_gate_residual(ops, L, cta_pair, block_n, with_bias, stride_mode):
    B, N, D, K = 3, 120, 1280, 512
    M = B * N
    A, W, bias = _ab(M, D, K, seed=5)
    x0 = torch.randn(M, D, device=dev())
    mod = torch.randn(B, 6 * D, device=dev())
    gate = mod[:, 2 * D:3 * D]
    gb = gate if stride_mode == "per_batch" else gate[:1].expand(B, D)
    y = A.float() @ W.float().t() + (bias if with_bias else 0.0)
    want = x0 + (gb.unsqueeze(1) * y.view(B, N, D)).view(M, D)  # jat_audiosr_v2.py:281,287
    x = x0.clone()
    ops.gemm(A, W, kind=L.EPI_GATE_RESIDUAL, out=x, bias=bias if with_bias else None, gate=gate,
             gate_batch_stride=6 * D if stride_mode == "per_batch" else 0, tokens_per_batch=N,
             cta_pair=cta_pair, block_n=block_n)
    assert (x - want).abs().max().item() < 5e-3
    assert rel_l2(x, want) < 1e-4


@pytest.mark.parametrize("cta_pair,block_n", GEMM_CFGS)
@pytest.mark.parametrize("T", [1378, 516, 86, 85])
def test_gemm_unpatchify(ops, L, cta_pair, block_n, T):
    B, Cc, P, K = 2, 256, 4, 256
    N = (T + P - 1) // P
    M = B * N
    A, W, bias = _ab(M, Cc * P, K, seed=6)
    y = A.float() @ W.float().t() + bias
    want = y.view(B, N, Cc, P).permute(0, 2, 1, 3).reshape(B, Cc, N * P)[:, :, :T]  # jat_audiosr_v2.py:383-397
    out = torch.full((B, Cc, T), float("nan"), device=dev())
    ops.gemm(A, W, kind=L.EPI_UNPATCHIFY, out=out, bias=bias, tokens_per_batch=N, patch_len=P, t_out=T,
             cta_pair=cta_pair, block_n=block_n)
    assert torch.isfinite(out).all()
    assert (out - want).abs().max().item() < 5e-3
    assert rel_l2(out, want) < 1e-4


def test_gemm_headline_shape_tail(ops, L):
    """M = 19320 = 150*128 + 120: the ragged last M tile of the headline config."""
    M, N, K = 19320, 1280, 1280
    A, W, bias = _ab(M, N, K, seed=7)
    got = ops.gemm(A, W, bias=bias, out_dtype=L.DTYPE_F32)
    for lo, hi in [(0, 256), (9000, 9300), (19200, 19320)]:
        want = A[lo:hi].float() @ W.float().t() + bias
        assert rel_l2(got[lo:hi], want) < 1e-4


def test_gemm_rejects_bad_shapes(ops, L):
    A = torch.zeros(8, 100, dtype=torch.bfloat16, device=dev())
    W = torch.zeros(128, 100, dtype=torch.bfloat16, device=dev())
    with pytest.raises(L.JatError):
        ops.gemm(A, W)


# ------------------------------------------------------------------------------------------- GEMM tail split
@pytest.mark.parametrize("cta_pair,block_n", GEMM_CFGS)
@pytest.mark.parametrize("kind", ["bias_f32", "bias_gelu_bf16", "qkv_rope", "gate_residual", "unpatchify", "dgrad_dact"])
def test_gemm_tail_split_matches_unsplit_and_is_deterministic(ops, L, cta_pair, block_n, kind):
    """Headline-sized problems whose persistent schedule ends in a partial wave: the tiles of that wave are cut along K
    and fixed up in part order (jat_set_gemm_tail_split).  Same result as the unsplit schedule up to f32 summation
    order, bit-identical run to run, and correct against torch on the rows of the split tiles (the last M blocks)."""
    B, Ntok = 56, 345
    M = B * Ntok
    shapes = {"bias_f32": (1280, 512), "bias_gelu_bf16": (512, 1280), "qkv_rope": (1792, 1280), "gate_residual": (1280, 1280),
              "unpatchify": (4096, 256), "dgrad_dact": (1280, 512)}
    N, K = shapes[kind]
    A, W, bias = _ab(M, N, K, seed=11)
    cos, sin = (t.to(dev()) for t in _rope_tables())
    gate = torch.randn(B, N, device=dev())
    x0 = torch.randn(M, N, device=dev())
    u = torch.randn(M, N, device=dev()).to(torch.bfloat16)
    Wt = W.t().contiguous()  # [K, N] for the dgrad case: dX[M, N] = dY[M, K] Wt[K, N]

    def run():
        if kind == "bias_f32":
            return ops.gemm(A, W, bias=bias, out_dtype=L.DTYPE_F32, cta_pair=cta_pair, block_n=block_n)
        if kind == "bias_gelu_bf16":
            return ops.gemm(A, W, bias=bias, act=L.ACT_GELU_ERF, cta_pair=cta_pair, block_n=block_n)
        if kind == "qkv_rope":
            return ops.gemm(A, W, kind=L.EPI_QKV_ROPE, tokens_per_batch=Ntok, rope_cos=cos, rope_sin=sin, rope_cols=1536,
                            cta_pair=cta_pair, block_n=block_n)
        if kind == "gate_residual":
            x = x0.clone()
            ops.gemm(A, W, kind=L.EPI_GATE_RESIDUAL, out=x, bias=bias, gate=gate, gate_batch_stride=N,
                     tokens_per_batch=Ntok, cta_pair=cta_pair, block_n=block_n)
            return x
        if kind == "unpatchify":
            out = torch.empty(B, N // 4, 1378, device=dev())
            ops.gemm(A, W, kind=L.EPI_UNPATCHIFY, out=out, bias=bias, tokens_per_batch=Ntok, patch_len=4, t_out=1378,
                     cta_pair=cta_pair, block_n=block_n)
            return out
        return ops.gemm(A, Wt, kind=L.EPI_DACT, act=L.ACT_GELU_ERF, aux=u, w_transposed=True, cta_pair=cta_pair,
                        block_n=block_n)
    try:
        ops.set_gemm_tail_split(dev(), False)
        plain = run()
        ops.set_gemm_tail_split(dev(), True)
        split = run()
        again = run()
    finally:
        ops.set_gemm_tail_split(dev(), False)  # the library default
    assert torch.equal(split, again)
    if split.dtype == torch.bfloat16:
        # one bf16 ulp where the f32 sums straddle a rounding boundary
        assert (split.float() - plain.float()).abs().max() <= 2 ** -7 * plain.float().abs().max()
        assert rel_l2(split.float(), plain.float()) < 1e-3
    else:
        assert rel_l2(split, plain) < 2e-6
    if kind == "gate_residual":  # against torch on the last rows (split tiles live in the last M blocks)
        lo = M - 700
        y = A[lo:].float() @ W.float().t() + bias
        want = x0[lo:] + gate.repeat_interleave(Ntok, 0)[lo:] * y
        assert rel_l2(split[lo:], want) < 1e-4


@pytest.mark.parametrize("cta_pair,block_n", GEMM_CFGS)
@pytest.mark.parametrize("N,K", [(1280, 1280), (1280, 5120)])
def test_gemm_tail_split_direct_gate_residual(ops, L, cta_pair, block_n, N, K):
    """Tail split mode 2 (out_proj / fc2 shapes of the headline step): the parts of a tail tile reduce-add their partial
    sums straight into the residual stream (bias with part 0).  Same result as the unsplit schedule up to f32 summation
    order; other epilogues are left unsplit in this mode (bit-identical to mode 0)."""
    B, Ntok = 56, 345
    M = B * Ntok
    A, W, bias = _ab(M, N, K, seed=12)
    gate = torch.randn(B, N, device=dev())
    x0 = torch.randn(M, N, device=dev())

    def run():
        x = x0.clone()
        ops.gemm(A, W, kind=L.EPI_GATE_RESIDUAL, out=x, bias=bias, gate=gate, gate_batch_stride=N, tokens_per_batch=Ntok,
                 cta_pair=cta_pair, block_n=block_n)
        return x
    try:
        ops.set_gemm_tail_split(dev(), 0)
        plain = run()
        plain_bf16 = ops.gemm(A, W, bias=bias, act=L.ACT_GELU_ERF, cta_pair=cta_pair, block_n=block_n)
        ops.set_gemm_tail_split(dev(), 2)
        split = run()
        split_bf16 = ops.gemm(A, W, bias=bias, act=L.ACT_GELU_ERF, cta_pair=cta_pair, block_n=block_n)
    finally:
        ops.set_gemm_tail_split(dev(), 0)
    assert torch.equal(split_bf16, plain_bf16)
    assert rel_l2(split, plain) < 2e-6
    lo = M - 700
    y = A[lo:].float() @ W.float().t() + bias
    want = x0[lo:] + gate.repeat_interleave(Ntok, 0)[lo:] * y
    assert rel_l2(split[lo:], want) < 1e-4
    head = slice(0, 256 * 60)  # whole tiles of the full waves: untouched by the split
    assert torch.equal(split[head], plain[head])


@pytest.mark.parametrize("kind", ["bias_f32", "bias_gelu_bf16", "qkv_rope", "gate_residual", "unpatchify"])
@pytest.mark.parametrize("M", [19320, 9660])
def test_gemm_multicast_cluster_is_bit_identical(ops, L, kind, M):
    """cta_pair = 2: clusters of two CTA pairs, the W tile multicast between the pairs (one L2 read instead of two).  The
    per-element accumulation order is unchanged, so the result equals the independent-pair schedule bit for bit."""
    Ntok = 345
    B = M // Ntok
    shapes = {"bias_f32": (1280, 512), "bias_gelu_bf16": (5120, 1280), "qkv_rope": (1792, 1280), "gate_residual": (1280, 5120),
              "unpatchify": (4096, 1280)}
    N, K = shapes[kind]
    A, W, bias = _ab(M, N, K, seed=13)
    cos, sin = _rope_tables()
    gate = torch.randn(B, N, device=dev())
    x0 = torch.randn(M, N, device=dev())

    def run(cta_pair):
        if kind == "bias_f32":
            return ops.gemm(A, W, bias=bias, out_dtype=L.DTYPE_F32, cta_pair=cta_pair, block_n=256)
        if kind == "bias_gelu_bf16":
            return ops.gemm(A, W, bias=bias, act=L.ACT_GELU_ERF, cta_pair=cta_pair, block_n=256)
        if kind == "qkv_rope":
            return ops.gemm(A, W, kind=L.EPI_QKV_ROPE, tokens_per_batch=Ntok, rope_cos=cos, rope_sin=sin, rope_cols=1536,
                            cta_pair=cta_pair, block_n=256)
        if kind == "gate_residual":
            x = x0.clone()
            ops.gemm(A, W, kind=L.EPI_GATE_RESIDUAL, out=x, bias=bias, gate=gate, gate_batch_stride=N,
                     tokens_per_batch=Ntok, cta_pair=cta_pair, block_n=256)
            return x
        out = torch.full((B, N // 4, 1378), float("nan"), device=dev())
        ops.gemm(A, W, kind=L.EPI_UNPATCHIFY, out=out, bias=bias, tokens_per_batch=Ntok, patch_len=4, t_out=1378,
                 cta_pair=cta_pair, block_n=256)
        return out
    want = run(1)
    got = run(2)
    assert torch.isfinite(got.float()).all()
    assert torch.equal(got, want), (got.float() - want.float()).abs().max().item()


# ------------------------------------------------------------------------------------------- attention
# heads per CTA picked by the launch's makespan model: B = 28 -> parts of 3 + 2 heads, B = 8 -> 4 + 1, B = 4 -> 2 + 2 + 1,
# B = 56 -> whole groups (5), B <= 2 -> single heads; (6, 100, 7, 1): 7 heads in ragged parts
@pytest.mark.parametrize("B,N,Hq,Hkv", [(2, 345, 20, 4), (1, 22, 8, 4), (3, 129, 16, 4), (2, 256, 4, 4), (1, 352, 5, 1),
                                        (28, 345, 20, 4), (8, 345, 20, 4), (4, 345, 20, 4), (56, 345, 20, 4), (6, 100, 7, 1),
                                        (40, 130, 7, 1)])
def test_gqa_attention(ops, B, N, Hq, Hkv):
    torch.manual_seed(8)
    hd = 64
    qkv = (torch.randn(B * N, (Hq + 2 * Hkv) * hd, device=dev()) * 1.5).to(torch.bfloat16)
    got = ops.gqa_attention_fwd(qkv, B, N, Hq, Hkv)
    f = qkv.float().view(B, N, Hq + 2 * Hkv, hd)
    Q, K, V = f[:, :, :Hq], f[:, :, Hq:Hq + Hkv], f[:, :, Hq + Hkv:]
    G = Hq // Hkv
    K = K.repeat_interleave(G, dim=2)  # jat_audiosr_v2.py:147-148
    V = V.repeat_interleave(G, dim=2)
    s = torch.matmul(Q.transpose(1, 2), K.transpose(1, 2).transpose(-2, -1)) / math.sqrt(hd)
    want = torch.matmul(torch.softmax(s, -1), V.transpose(1, 2)).transpose(1, 2).reshape(B * N, Hq * hd)
    err = (got.float() - want).abs().max().item()
    assert err < 2e-2 * max(1.0, want.abs().max().item()), err  # P is rounded to bf16 before P.V
    assert rel_l2(got.float(), want) < 8e-3


def test_gqa_attention_single_pass_entry_rejects_long(ops, L):
    """The scratch-free entry point covers one 352-key chunk; longer sequences must use jat_gqa_attention_fwd_long."""
    import ctypes as C
    qkv = torch.zeros(400, 3 * 64, dtype=torch.bfloat16, device=dev())
    out = torch.zeros(400, 64, dtype=torch.bfloat16, device=dev())
    rc = L.load().jat_gqa_attention_fwd(L.context(0), qkv.data_ptr(), out.data_ptr(), None, 1, 400, 1, 1, 64,
                                        torch.cuda.current_stream().cuda_stream)
    assert rc == -2 and b"jat_gqa_attention_fwd_long" in L.load().jat_last_error()


@pytest.mark.parametrize("B,N,Hq,Hkv", [(1, 353, 4, 2), (2, 400, 8, 4), (1, 704, 5, 1), (1, 1000, 4, 4), (1, 2048, 2, 1)])
def test_gqa_attention_long_sequences(ops, B, N, Hq, Hkv):
    """More than 352 tokens (the reference allows 2048, jat_audiosr_v2.py:428): one pass per 352-key chunk merged by the
    chunks' log-sum-exps == the full softmax; the merged LSE feeds the (length-agnostic) backward kernel."""
    torch.manual_seed(9)
    hd = 64
    G = Hq // Hkv
    cos, sin = _bwd_rope_tables(N)
    raw = (torch.randn(B, N, Hq + 2 * Hkv, hd, device=dev()) * 1.2).to(torch.bfloat16).float().requires_grad_(True)
    q = _bwd_rope(raw[:, :, :Hq], cos, sin)
    k = _bwd_rope(raw[:, :, Hq:Hq + Hkv], cos, sin)
    v = raw[:, :, Hq + Hkv:]
    qkv = torch.cat([q, k, v], 2).detach().reshape(B * N, -1).to(torch.bfloat16)
    lse = torch.empty(B, Hq, N, device=dev())
    got = ops.gqa_attention_fwd(qkv, B, N, Hq, Hkv, lse=lse)
    s = torch.einsum("bnhd,bmhd->bhnm", q, k.repeat_interleave(G, dim=2)) / 8.0
    o = torch.einsum("bhnm,bmhd->bnhd", torch.softmax(s, -1), v.repeat_interleave(G, dim=2)).reshape(B * N, Hq * hd)
    assert rel_l2(got.float(), o.detach()) < 1e-2
    assert (lse - torch.logsumexp(s, -1).detach() * math.log2(math.e)).abs().max() < 2e-2
    d_out = torch.randn(B * N, Hq * hd, device=dev()).to(torch.bfloat16)
    dq = ops.gqa_attention_bwd(qkv, d_out, got, lse, cos, sin, B, N, Hq, Hkv).float().view(B, N, Hq + 2 * Hkv, hd)
    o.backward(d_out.float())
    for name, sl in (("dq", slice(0, Hq)), ("dk", slice(Hq, Hq + Hkv)), ("dv", slice(Hq + Hkv, Hq + 2 * Hkv))):
        assert rel_l2(dq[:, :, sl], raw.grad[:, :, sl]) < 2e-2, name


def test_gqa_attention_extreme_scores(ops):
    """Keys with huge norm in one key half only: the two halves' row maxima differ by hundreds of nats, so
    the split-KV combine weights 2^(m_half - m) underflow to exactly 0 on one side; results must stay
    finite and equal the fp32 softmax (which is ~one-hot on that key)."""
    torch.manual_seed(9)
    B, N, Hq, Hkv, hd = 2, 345, 20, 4, 64
    f = torch.randn(B, N, Hq + 2 * Hkv, hd, device=dev()) * 1.5
    f[:, 300, Hq:Hq + Hkv] *= 25.0   # one key row with huge norm in every KV head
    f[:, 171, Hq:Hq + Hkv] *= 60.0   # and a larger one in the other half
    qkv = f.reshape(B * N, -1).to(torch.bfloat16)
    got = ops.gqa_attention_fwd(qkv, B, N, Hq, Hkv)
    assert torch.isfinite(got.float()).all()
    g = qkv.float().view(B, N, Hq + 2 * Hkv, hd)
    Q, K, V = g[:, :, :Hq], g[:, :, Hq:Hq + Hkv], g[:, :, Hq + Hkv:]
    K = K.repeat_interleave(Hq // Hkv, dim=2)
    V = V.repeat_interleave(Hq // Hkv, dim=2)
    s = torch.matmul(Q.transpose(1, 2), K.transpose(1, 2).transpose(-2, -1)) / math.sqrt(hd)
    want = torch.matmul(torch.softmax(s, -1), V.transpose(1, 2)).transpose(1, 2).reshape(B * N, Hq * hd)
    assert rel_l2(got.float(), want) < 8e-3


# ------------------------------------------------------------------------------------------------ backward GEMMs
def _gelu_grad(u):
    return 0.5 * (1 + torch.erf(u / math.sqrt(2))) + u * torch.exp(-0.5 * u * u) / math.sqrt(2 * math.pi)


@pytest.mark.parametrize("cta_pair,block_n", GEMM_CFGS)
@pytest.mark.parametrize("M,Nfwd,Kfwd", [(300, 512, 1280), (1000, 1280, 256), (56, 128, 128)])
def test_gemm_dgrad_transposed_weight(ops, L, cta_pair, block_n, M, Nfwd, Kfwd):
    """dX[M, K'] = dY[M, N'] W[N', K'] with W passed as stored (nn.Linear [out, in]) via w_transposed."""
    torch.manual_seed(3)
    dY = torch.randn(M, Nfwd, device=dev()).to(torch.bfloat16)
    W = (torch.randn(Nfwd, Kfwd, device=dev()) / math.sqrt(Nfwd)).to(torch.bfloat16)
    got = ops.gemm(dY, W, w_transposed=True, cta_pair=cta_pair, block_n=block_n)
    want = dY.float() @ W.float()
    assert got.shape == (M, Kfwd) and rel_l2(got.float(), want) < 4e-3
    got32 = ops.gemm(dY, W, w_transposed=True, out_dtype=L.DTYPE_F32, cta_pair=cta_pair, block_n=block_n)
    assert rel_l2(got32, want) < 1e-5


@pytest.mark.parametrize("cta_pair,block_n", GEMM_CFGS)
@pytest.mark.parametrize("act", [1, 2])
def test_gemm_dgrad_through_activation(ops, L, cta_pair, block_n, act):
    """du = (dY W) * act'(u): dgrad of mlp.3 fused with the GELU backward (and SiLU for t_embedder)."""
    torch.manual_seed(4)
    M, Nfwd, Kfwd = 333, 256, 1280
    dY = torch.randn(M, Nfwd, device=dev()).to(torch.bfloat16)
    W = (torch.randn(Nfwd, Kfwd, device=dev()) / math.sqrt(Nfwd)).to(torch.bfloat16)
    u = (torch.randn(M, Kfwd, device=dev()) * 1.5).to(torch.bfloat16)
    got = ops.gemm(dY, W, kind=L.EPI_DACT, act=act, aux=u, w_transposed=True, cta_pair=cta_pair, block_n=block_n)
    uf = u.float()
    d = _gelu_grad(uf) if act == 1 else torch.sigmoid(uf) * (1 + uf * (1 - torch.sigmoid(uf)))
    want = (dY.float() @ W.float()) * d
    assert rel_l2(got.float(), want) < 4e-3


@pytest.mark.parametrize("cta_pair,block_n", GEMM_CFGS)
@pytest.mark.parametrize("Mtok,Nfwd,Kfwd,splits", [(690, 512, 1280, 1), (1000, 256, 128, 3), (9660, 128, 256, 7), (28, 640, 128, 4)])
def test_gemm_wgrad_accumulate(ops, L, cta_pair, block_n, Mtok, Nfwd, Kfwd, splits):
    """dW[N', K'] += dY^T[N', M] X[M, K'] (both operands passed as stored), reduction over tokens split `splits`-way;
    the token count need not be a multiple of 64 (TMA zero-fills)."""
    torch.manual_seed(5)
    dY = torch.randn(Mtok, Nfwd, device=dev()).to(torch.bfloat16)
    X = torch.randn(Mtok, Kfwd, device=dev()).to(torch.bfloat16)
    base = torch.randn(Nfwd, Kfwd, device=dev())
    out = base.clone()
    ops.gemm(dY, X, kind=L.EPI_ACCUM, out=out, a_transposed=True, w_transposed=True, k_splits=splits,
             cta_pair=cta_pair, block_n=block_n)
    want = base + dY.float().t() @ X.float()
    assert rel_l2(out, want) < 1e-5


def test_gemm_forward_keeps_preactivation(ops, L):
    torch.manual_seed(6)
    M, N, K = 500, 512, 256
    A = torch.randn(M, K, device=dev()).to(torch.bfloat16)
    W = (torch.randn(N, K, device=dev()) / 16).to(torch.bfloat16)
    b = torch.randn(N, device=dev())
    u = torch.zeros(M, N, dtype=torch.bfloat16, device=dev())
    got = ops.gemm(A, W, act=L.ACT_GELU_ERF, bias=b, aux=u)
    pre = A.float() @ W.float().t() + b
    assert rel_l2(u.float(), pre) < 4e-3
    assert rel_l2(got.float(), torch.nn.functional.gelu(pre)) < 5e-3


@pytest.mark.parametrize("cta_pair,block_n", GEMM_CFGS)
def test_gemm_gate_residual_split_k(ops, L, cta_pair, block_n):
    torch.manual_seed(7)
    B, Ntok, N, K = 3, 100, 256, 1280
    M = B * Ntok
    A = torch.randn(M, K, device=dev()).to(torch.bfloat16)
    W = (torch.randn(N, K, device=dev()) / math.sqrt(K)).to(torch.bfloat16)
    bias, gate = torch.randn(N, device=dev()), torch.randn(B, N, device=dev())
    x0 = torch.randn(M, N, device=dev())
    x = x0.clone()
    ops.gemm(A, W, kind=L.EPI_GATE_RESIDUAL, out=x, bias=bias, gate=gate, gate_batch_stride=N, tokens_per_batch=Ntok,
             k_splits=5, cta_pair=cta_pair, block_n=block_n)
    want = x0 + gate.repeat_interleave(Ntok, 0) * (A.float() @ W.float().t() + bias)
    assert rel_l2(x, want) < 1e-5


# ------------------------------------------------------------------------------------------------ backward elementwise
@pytest.mark.parametrize("norm_kind", [0, 1])
@pytest.mark.parametrize("D", [512, 1280])
@pytest.mark.parametrize("mode", ["per_batch", "none"])
def test_adaln_bwd_matches_autograd(ops, norm_kind, D, mode):
    torch.manual_seed(21)
    B, Ntok = 3, 77
    M = B * Ntok
    x = (torch.randn(M, D, device=dev()) * 1.3 + 0.2).requires_grad_(True)
    w = (1 + 0.1 * torch.randn(D, device=dev())).requires_grad_(True)
    shift = (0.3 * torch.randn(B, D, device=dev())).requires_grad_(True)
    scale = (0.3 * torch.randn(B, D, device=dev())).requires_grad_(True)
    dh = torch.randn(M, D, device=dev()).to(torch.bfloat16)
    if norm_kind == 0:
        y = torch.nn.functional.layer_norm(x, (D,), eps=1e-6)
    else:
        y = x * torch.rsqrt((x * x).mean(-1, keepdim=True) + 1e-6) * w
    h = y
    if mode == "per_batch":
        h = y * (1 + scale.repeat_interleave(Ntok, 0)) + shift.repeat_interleave(Ntok, 0)
    h.backward(dh.float())
    base = torch.randn(M, D, device=dev())
    dx = base.clone()
    dmod = torch.zeros(B, 2 * D, device=dev())
    dw = torch.zeros(D, device=dev())
    kw = dict(norm_kind=norm_kind, weight=w.detach() if norm_kind else None, dweight=dw if norm_kind else None)
    if mode == "per_batch":
        kw.update(scale=scale.detach(), mod_batch_stride=D, dshift=dmod[:, :D], dscale=dmod[:, D:], dmod_batch_stride=2 * D)
    ops.adaln_bwd(dh, x.detach(), B, Ntok, dx, accumulate=True, **kw)
    assert rel_l2(dx - base, x.grad) < 2e-5
    if mode == "per_batch":
        assert rel_l2(dmod[:, :D], shift.grad) < 1e-5 and rel_l2(dmod[:, D:], scale.grad) < 1e-5
    if norm_kind:
        assert rel_l2(dw, w.grad) < 1e-5
    dx2 = torch.full_like(base, 7.0)
    ops.adaln_bwd(dh, x.detach(), B, Ntok, dx2, accumulate=False, **{**kw, **(dict(dshift=torch.zeros(B, D, device=dev()),
                  dscale=torch.zeros(B, D, device=dev()), dmod_batch_stride=D) if mode == "per_batch" else {})})
    assert rel_l2(dx2, x.grad) < 2e-5


@pytest.mark.parametrize("with_bias", [False, True])
def test_gate_bwd_matches_autograd(ops, with_bias):
    torch.manual_seed(22)
    B, Ntok, D = 4, 91, 1280
    M = B * Ntok
    dx = torch.randn(M, D, device=dev())
    y = torch.randn(M, D, device=dev()).to(torch.bfloat16)
    gate = torch.randn(B, 3 * D, device=dev())[:, D:2 * D]      # a strided view like mod[:, 2D:3D]
    dgate = torch.zeros(B, D, device=dev())
    dbias = torch.ones(D, device=dev()) if with_bias else None
    dy = ops.gate_bwd(dx, y, gate, B, Ntok, dgate, mod_batch_stride=3 * D, dmod_batch_stride=D, dbias=dbias)
    g = gate.repeat_interleave(Ntok, 0)
    assert rel_l2(dy.float(), g * dx) < 4e-3
    assert rel_l2(dgate, (dx * y.float()).view(B, Ntok, D).sum(1)) < 1e-5
    if with_bias:
        assert rel_l2(dbias - 1, (g * dx).sum(0)) < 1e-5


@pytest.mark.parametrize("norm_kind", [0, 1])
# D % 8 == 0: the shared-memory staged kernel; D = 1004: the register version (rows not 16-byte granular in bf16)
@pytest.mark.parametrize("D,B,Ntok", [(1280, 28, 345), (1280, 3, 77), (128, 4, 22), (1000, 2, 5), (1004, 2, 9), (2048, 2, 40),
                                      (1280, 300, 7)])
@pytest.mark.parametrize("gate_mode", ["none", "gate", "gate_bias_drop"])
def test_adaln_gate_bwd_fused_matches_autograd(ops, L, norm_kind, D, B, Ntok, gate_mode):
    """Fused norm + modulate backward (+ the gate backward of the branch below, on the updated dx row) against torch
    autograd of   h = norm(x) (1 + scale_b) + shift_b   and   x_out = x_in + rs_b gate_b mask (y)."""
    torch.manual_seed(41)
    M = B * Ntok
    x = (torch.randn(M, D, device=dev()) * 1.3 + 0.2).requires_grad_(True)
    w = (1 + 0.1 * torch.randn(D, device=dev())).requires_grad_(True)
    mod = (0.3 * torch.randn(B, 3 * D, device=dev())).requires_grad_(True)     # shift | scale | gate, row stride 3D
    shift, scale, gate = mod[:, :D], mod[:, D:2 * D], mod[:, 2 * D:]
    dh = torch.randn(M, D, device=dev()).to(torch.bfloat16)
    if norm_kind == 0:
        yn = torch.nn.functional.layer_norm(x, (D,), eps=1e-6)
        mean, rstd = x.mean(-1), torch.rsqrt(x.var(-1, unbiased=False) + 1e-6)
    else:
        rstd = torch.rsqrt((x * x).mean(-1) + 1e-6)
        mean = torch.zeros_like(rstd)
        yn = x * rstd[:, None] * w
    h = yn * (1 + scale.repeat_interleave(Ntok, 0)) + shift.repeat_interleave(Ntok, 0)
    rowstats = torch.stack([mean, rstd], -1).detach().contiguous()
    base = torch.randn(M, D, device=dev())                                     # gradient already on the residual stream
    (gx,) = torch.autograd.grad(h, x, dh.float(), retain_graph=True)
    dx_want = base + gx
    dmod = torch.zeros(B, 3 * D, device=dev())
    dw = torch.zeros(D, device=dev())
    kw = dict(scale=scale.detach(), mod_batch_stride=3 * D, dshift=dmod[:, :D], dscale=dmod[:, D:2 * D], dmod_batch_stride=3 * D,
              norm_kind=norm_kind, weight=w.detach() if norm_kind else None, dweight=dw if norm_kind else None)
    dx = base.clone()
    if gate_mode == "none":
        ops.adaln_gate_bwd(dh, x.detach(), rowstats, B, Ntok, dx, **kw)
    else:
        drop = gate_mode == "gate_bias_drop"
        p, seed = (0.1, 77) if drop else (0.0, 0)
        yb = torch.randn(M, D, device=dev()).to(torch.bfloat16)
        rs = (torch.rand(B, device=dev()) > 0.3).float() / 0.7 if drop else None
        dbias = torch.ones(D, device=dev()) if drop else None
        dy = ops.adaln_gate_bwd(dh, x.detach(), rowstats, B, Ntok, dx, y=yb, gate=gate.detach(), dgate=dmod[:, 2 * D:], dbias=dbias,
                                drop_p=p, drop_seed=seed, gate_rowscale=rs, **kw)
        mask = ops.dropout_scale_mask(M, D, p, seed, dev()) if drop else torch.ones(M, D, device=dev())
        g_eff = (gate.detach() * (rs[:, None] if rs is not None else 1.0)).repeat_interleave(Ntok, 0)
        assert rel_l2(dy.float(), dx_want * mask * g_eff) < 4e-3               # bf16 output
        want_dgate = (dx_want * yb.float()).view(B, Ntok, D).sum(1) * (rs[:, None] if rs is not None else 1.0)
        assert rel_l2(dmod[:, 2 * D:], want_dgate) < 2e-5
        if drop:
            assert rel_l2(dbias - 1, (dx_want * mask * g_eff).sum(0)) < 2e-5
    assert rel_l2(dx, dx_want) < 2e-5
    h.backward(dh.float())
    assert rel_l2(dmod[:, :D], mod.grad[:, :D]) < 1e-5 and rel_l2(dmod[:, D:2 * D], mod.grad[:, D:2 * D]) < 1e-5
    if norm_kind:
        assert rel_l2(dw, w.grad) < 1e-5


@pytest.mark.parametrize("D,B,Ntok", [(1280, 28, 345), (128, 3, 22), (2048, 2, 40)])
def test_adaln_gate_bwd_rows_are_bitwise_reproducible(ops, D, B, Ntok):
    """dx and dy are written once per element (no atomics): two launches on the same inputs must agree bit for bit.  A race
    in the shared-memory staging (a stage refilled while still being read, a barrier phase off by one) would show up here."""
    torch.manual_seed(3)
    M = B * Ntok
    x = torch.randn(M, D, device=dev()) * 1.3 + 0.2
    rowstats = torch.stack([x.mean(-1), torch.rsqrt(x.var(-1, unbiased=False) + 1e-6)], -1).contiguous()
    mod = 0.3 * torch.randn(B, 3 * D, device=dev())
    dh = torch.randn(M, D, device=dev()).to(torch.bfloat16)
    yb = torch.randn(M, D, device=dev()).to(torch.bfloat16)
    base = torch.randn(M, D, device=dev())
    outs = []
    for _ in range(3):
        dmod = torch.zeros(B, 3 * D, device=dev())
        dx = base.clone()
        dy = ops.adaln_gate_bwd(dh, x, rowstats, B, Ntok, dx, scale=mod[:, D:2 * D], mod_batch_stride=3 * D, dshift=dmod[:, :D],
                                dscale=dmod[:, D:2 * D], dmod_batch_stride=3 * D, y=yb, gate=mod[:, 2 * D:], dgate=dmod[:, 2 * D:],
                                drop_p=0.1, drop_seed=11)
        outs.append((dx, dy))
    for dx, dy in outs[1:]:
        assert torch.equal(dx, outs[0][0]) and torch.equal(dy, outs[0][1])


def test_colsum_and_cast(ops):
    torch.manual_seed(23)
    a = torch.randn(9660, 512, device=dev()).to(torch.bfloat16)
    out = torch.ones(512, device=dev())
    ops.colsum_bf16(a, out)
    assert rel_l2(out - 1, a.float().sum(0)) < 1e-5
    x = torch.randn(28 * 7680 + 3, device=dev())
    assert torch.equal(ops.cast_f32_bf16(x), x.to(torch.bfloat16))


# ------------------------------------------------------------------------------------------------ attention backward
def _bwd_rope_tables(npos):
    inv = 1.0 / (10000 ** (torch.arange(0, 64, 2, device=dev()).float() / 64))
    f = torch.outer(torch.arange(npos, device=dev()).float(), inv)
    e = torch.cat([f, f], -1)
    return e.cos().contiguous(), e.sin().contiguous()


def _bwd_rope(x, cos, sin):  # x [B, N, H, 64]
    x1, x2 = x[..., :32], x[..., 32:]
    return x * cos[None, :, None, :] + torch.cat([-x2, x1], -1) * sin[None, :, None, :]


@pytest.mark.parametrize("B,N,Hq,Hkv", [(2, 345, 20, 4), (1, 22, 8, 4), (3, 129, 4, 2), (2, 256, 2, 2), (1, 300, 5, 1)])
def test_gqa_attention_bwd_matches_autograd(ops, B, N, Hq, Hkv):
    """dqkv (w.r.t. the PRE-RoPE projections) vs torch autograd through RoPE + repeat_interleave + softmax attention."""
    torch.manual_seed(31)
    G = Hq // Hkv
    cos, sin = _bwd_rope_tables(N)
    raw = (torch.randn(B, N, Hq + 2 * Hkv, 64, device=dev())).to(torch.bfloat16).float().requires_grad_(True)
    q = _bwd_rope(raw[:, :, :Hq], cos, sin)
    k = _bwd_rope(raw[:, :, Hq:Hq + Hkv], cos, sin)
    v = raw[:, :, Hq + Hkv:]
    # the kernels see bf16 RoPE'd projections (what the QKV GEMM epilogue stores)
    qkv = torch.cat([q, k, v], 2).detach().reshape(B * N, -1).to(torch.bfloat16)
    lse = torch.empty(B, Hq, N, device=dev())
    out = ops.gqa_attention_fwd(qkv, B, N, Hq, Hkv, lse=lse)
    d_out = torch.randn(B * N, Hq * 64, device=dev()).to(torch.bfloat16)
    got = ops.gqa_attention_bwd(qkv, d_out, out, lse, cos, sin, B, N, Hq, Hkv).float().view(B, N, Hq + 2 * Hkv, 64)

    kk = k.repeat_interleave(G, dim=2)
    vv = v.repeat_interleave(G, dim=2)
    s = torch.einsum("bnhd,bmhd->bhnm", q, kk) / 8.0
    want_lse = torch.logsumexp(s, -1) * math.log2(math.e)
    assert (lse - want_lse.detach()).abs().max() < 2e-2
    o = torch.einsum("bhnm,bmhd->bnhd", torch.softmax(s, -1), vv).reshape(B * N, Hq * 64)
    o.backward(d_out.float())
    want = raw.grad
    for name, sl in (("dq", slice(0, Hq)), ("dk", slice(Hq, Hq + Hkv)), ("dv", slice(Hq + Hkv, Hq + 2 * Hkv))):
        err = rel_l2(got[:, :, sl], want[:, :, sl])
        assert err < 1.5e-2, (name, err)
