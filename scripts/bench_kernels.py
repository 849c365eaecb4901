"""Per-kernel micro-benchmarks at the headline shapes (v3mod2, B_eff=56, N=345 -> M=19320).
CUDA-event timing on the launching stream, L2 flushed between iterations.  Prints one JSON line per case.
Usage: python scripts/bench_kernels.py [--only gemm|attn|elem] [--iters 20]"""
import argparse
import json
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jat_b200 import _lib as L  # noqa: E402
from jat_b200 import ops  # noqa: E402

PEAKS = {"hbm_gbs": 6530.0, "bf16_tflops": 1665.4, "bf16_tflops_sustained": 1401.0}
try:
    PEAKS.update(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json"))))
except Exception:
    pass

dev = torch.device("cuda", 0)
flush_buf = None


def flush_l2():
    global flush_buf
    if flush_buf is None:
        flush_buf = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    flush_buf.zero_()


def timeit(fn, iters, warmup=3, flush=True):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    times = []
    for _ in range(iters):
        if flush:
            flush_l2()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        times.append(s.elapsed_time(e))
    times.sort()
    return times[len(times) // 2], times[0]


def report(name, ms_med, ms_min, flops=None, bytes_=None, **extra):
    d = {"case": name, "ms_median": round(ms_med, 4), "ms_min": round(ms_min, 4)}
    if flops:
        d["tflops"] = round(flops / ms_med / 1e9, 1)
        d["frac_burst"] = round(flops / ms_med / 1e9 / PEAKS["bf16_tflops"], 3)
    if bytes_:
        d["gbs"] = round(bytes_ / ms_med / 1e6, 1)
        d["frac_hbm"] = round(bytes_ / ms_med / 1e6 / PEAKS["hbm_gbs"], 3)
    d.update(extra)
    print(json.dumps(d), flush=True)


def bench_gemm(iters):
    M, D, F, QKV, Ntok, B = 19320, 1280, 5120, 1792, 345, 56
    g = torch.Generator(device="cpu").manual_seed(0)
    mod = torch.randn(B, 6 * D, device=dev)
    inv_freq = 1.0 / (10000 ** (torch.arange(0, 64, 2).float() / 64))
    emb = torch.outer(torch.arange(4096).float(), inv_freq)
    emb = torch.cat([emb, emb], -1)
    cos, sin = emb.cos().to(dev), emb.sin().to(dev)
    cases = [
        ("qkv_rope", M, QKV, D, dict(kind=L.EPI_QKV_ROPE, tokens_per_batch=Ntok, rope_cos=cos, rope_sin=sin, rope_cols=1536)),
        ("out_proj_gate_res", M, D, D, dict(kind=L.EPI_GATE_RESIDUAL, gate=mod[:, 2 * D:], gate_batch_stride=6 * D, tokens_per_batch=Ntok)),
        ("fc1_bias_gelu", M, F, D, dict(kind=L.EPI_BIAS_ACT, act=L.ACT_GELU_ERF, out_dtype=L.DTYPE_BF16)),
        ("fc2_bias_gate_res", M, D, F, dict(kind=L.EPI_GATE_RESIDUAL, gate=mod[:, 5 * D:], gate_batch_stride=6 * D, tokens_per_batch=Ntok)),
        ("patch1_bias_gelu", M, 512, 8192, dict(kind=L.EPI_BIAS_ACT, act=L.ACT_GELU_ERF, out_dtype=L.DTYPE_BF16)),
        ("final_unpatchify", M, 4096, D, dict(kind=L.EPI_UNPATCHIFY, tokens_per_batch=Ntok, patch_len=4, t_out=1378)),
        ("plain_bf16_8192", 8192, 8192, 8192, dict(kind=L.EPI_BIAS_ACT, act=L.ACT_NONE, out_dtype=L.DTYPE_BF16)),
    ]
    for name, m, n, k, kw in cases:
        A = (torch.randn(m, k, generator=g) * 0.5).to(dev).to(torch.bfloat16)
        W = (torch.randn(n, k, generator=g) / math.sqrt(k)).to(dev).to(torch.bfloat16)
        bias = torch.randn(n, device=dev)
        if kw["kind"] == L.EPI_GATE_RESIDUAL:
            out = torch.randn(m, n, device=dev)
        elif kw["kind"] == L.EPI_UNPATCHIFY:
            out = torch.empty(B, 1024, 1378, device=dev)
        else:
            out = torch.empty(m, n, dtype=torch.bfloat16, device=dev)
        for cta_pair in (0, 1, 2):
            for bn in (256, 128):
                fn = lambda: ops.gemm(A, W, bias=bias if kw["kind"] != L.EPI_QKV_ROPE else None, out=out,
                                      cta_pair=cta_pair, block_n=bn, **kw)
                med, mn = timeit(fn, iters)
                report(f"gemm_{name}", med, mn, flops=2.0 * m * n * k, M=m, N=n, K=k, cta_pair=cta_pair, block_n=bn)
        if name == "plain_bf16_8192":
            C = torch.empty(m, n, dtype=torch.bfloat16, device=dev)
            med, mn = timeit(lambda: torch.matmul(A, W.t(), out=C), iters)
            report("cublas_bf16_8192", med, mn, flops=2.0 * m * n * k)
        if name == "fc1_bias_gelu":
            C = torch.empty(m, n, dtype=torch.bfloat16, device=dev)
            med, mn = timeit(lambda: torch.matmul(A, W.t(), out=C), iters)
            report("cublas_fc1_shape_noepi", med, mn, flops=2.0 * m * n * k)


def bench_bwd(iters):
    """Backward-pass kernels at the training shapes (B = 28, M = 9660)."""
    B, Ntok, D, F, QKV, Hq, Hkv = 28, 345, 1280, 5120, 1792, 20, 4
    M = B * Ntok
    bf = lambda *s: torch.randn(*s, device=dev).to(torch.bfloat16)
    dh, x, dx = bf(M, D), torch.randn(M, D, device=dev), torch.randn(M, D, device=dev)
    mod = torch.randn(B, 28 * 6 * D, device=dev)
    dmod = torch.zeros(28, B, 6 * D, device=dev)
    fn = lambda: ops.adaln_bwd(dh, x, B, Ntok, dx, scale=mod[:, D:2 * D], mod_batch_stride=mod.stride(0), dshift=dmod[0, :, :D],
                               dscale=dmod[0, :, D:2 * D], dmod_batch_stride=6 * D)
    med, mn = timeit(fn, iters)
    report("adaln_bwd_layernorm", med, mn, bytes_=M * D * 14.0)
    y, dy = bf(M, D), torch.empty(M, D, dtype=torch.bfloat16, device=dev)
    rowstats = torch.stack([x.mean(-1), torch.rsqrt(x.var(-1, unbiased=False) + 1e-6)], -1).contiguous()
    dbias = torch.zeros(D, device=dev)
    for name, kw in (("adaln_gate_bwd_nogate", {}),
                     ("adaln_gate_bwd_gate_drop", dict(y=y, gate=mod[:, 2 * D:3 * D], dgate=dmod[0, :, 2 * D:3 * D], dbias=dbias,
                                                       drop_p=0.1, drop_seed=5, dy=dy))):
        fn = lambda: ops.adaln_gate_bwd(dh, x, rowstats, B, Ntok, dx, scale=mod[:, D:2 * D], mod_batch_stride=mod.stride(0),
                                        dshift=dmod[0, :, :D], dscale=dmod[0, :, D:2 * D], dmod_batch_stride=6 * D, **kw)
        med, mn = timeit(fn, iters)
        report(name, med, mn, bytes_=M * D * (18.0 if kw else 14.0))
    fn = lambda: ops.gate_bwd(dx, y, mod[:, 2 * D:3 * D], B, Ntok, dmod[0, :, 2 * D:3 * D], mod_batch_stride=mod.stride(0),
                              dmod_batch_stride=6 * D, dy=dy)
    med, mn = timeit(fn, iters)
    report("gate_bwd", med, mn, bytes_=M * D * 8.0)
    du = bf(M, F)
    out = torch.zeros(F, device=dev)
    med, mn = timeit(lambda: ops.colsum_bf16(du, out), iters)
    report("colsum_bf16_MxF", med, mn, bytes_=M * F * 2.0)
    # GEMMs: dgrad fc2 (+GELU'), dgrad fc1, wgrad fc1 / fc2 / qkv / out_proj
    W1, W2 = bf(F, D), bf(D, F)
    u, h2 = bf(M, F), bf(M, D)
    o_du, o_dh = torch.empty(M, F, dtype=torch.bfloat16, device=dev), torch.empty(M, D, dtype=torch.bfloat16, device=dev)
    med, mn = timeit(lambda: ops.gemm(dh, W2, kind=L.EPI_DACT, act=L.ACT_GELU_ERF, aux=u, w_transposed=True, out=o_du), iters)
    report("dgrad_fc2_dgelu", med, mn, flops=2.0 * M * F * D)
    med, mn = timeit(lambda: ops.gemm(du, W1, w_transposed=True, out=o_dh), iters)
    report("dgrad_fc1", med, mn, flops=2.0 * M * F * D)
    for name, dY, X in (("wgrad_fc1", du, h2), ("wgrad_fc2", dh, u), ("wgrad_out_proj", dh, h2), ("wgrad_qkv", bf(M, QKV), h2)):
        g = torch.zeros(dY.shape[1], X.shape[1], device=dev)
        for ks in (1, 2, 4, 8):
            med, mn = timeit(lambda: ops.gemm(dY, X, kind=L.EPI_ACCUM, out=g, a_transposed=True, w_transposed=True, k_splits=ks), iters)
            report(name, med, mn, flops=2.0 * M * dY.shape[1] * X.shape[1], k_splits=ks)
    # attention backward
    qkv = bf(M, QKV)
    lse = torch.empty(B, Hq, Ntok, device=dev)
    o = ops.gqa_attention_fwd(qkv, B, Ntok, Hq, Hkv, lse=lse)
    inv_freq = 1.0 / (10000 ** (torch.arange(0, 64, 2).float() / 64))
    emb = torch.outer(torch.arange(4096).float(), inv_freq)
    emb = torch.cat([emb, emb], -1)
    cos, sin = emb.cos().to(dev), emb.sin().to(dev)
    dqkv = torch.empty_like(qkv)
    med, mn = timeit(lambda: ops.gqa_attention_bwd(qkv, dh, o, lse, cos, sin, B, Ntok, Hq, Hkv, dqkv=dqkv), iters)
    report("gqa_attention_bwd", med, mn, flops=2.5 * 4.0 * B * Hq * Ntok * Ntok * 64)


def bench_attn(iters):
    B, N, Hq, Hkv = 56, 345, 20, 4
    qkv = torch.randn(B * N, (Hq + 2 * Hkv) * 64, device=dev).to(torch.bfloat16)
    out = torch.empty(B * N, Hq * 64, dtype=torch.bfloat16, device=dev)
    med, mn = timeit(lambda: ops.gqa_attention_fwd(qkv, B, N, Hq, Hkv, out=out), iters)
    report("gqa_attention", med, mn, flops=4.0 * B * Hq * N * N * 64, B=B, N=N)
    q = qkv.view(B, N, Hq + 2 * Hkv, 64)
    Q = q[:, :, :Hq].transpose(1, 2).contiguous()
    K = q[:, :, Hq:Hq + Hkv].transpose(1, 2).repeat_interleave(Hq // Hkv, 1).contiguous()
    V = q[:, :, Hq + Hkv:].transpose(1, 2).repeat_interleave(Hq // Hkv, 1).contiguous()
    med, mn = timeit(lambda: torch.nn.functional.scaled_dot_product_attention(Q, K, V), iters)
    report("torch_sdpa_same_shape", med, mn, flops=4.0 * B * Hq * N * N * 64)


def bench_elem(iters):
    M, D, B, Ntok = 19320, 1280, 56, 345
    x = torch.randn(M, D, device=dev)
    mod = torch.randn(B, 6 * D, device=dev)
    out = torch.empty(M, D, dtype=torch.bfloat16, device=dev)
    med, mn = timeit(lambda: ops.adaln_norm_modulate(x, mod[:, :D], mod[:, D:], 6 * D, None, 0, 1e-6, Ntok, out=out), iters)
    report("adaln_layernorm", med, mn, bytes_=M * D * 6.0)
    w = torch.ones(D, device=dev)
    med, mn = timeit(lambda: ops.adaln_norm_modulate(x, mod[:, :D], mod[:, D:], 6 * D, w, 1, 1e-6, Ntok, out=out), iters)
    report("adaln_rmsnorm", med, mn, bytes_=M * D * 6.0)
    z = torch.randn(28, 1024, 1378, device=dev)
    xc, xu = torch.randn_like(z), torch.randn_like(z)
    t_dt = torch.tensor([[0.1, 0.02]], device=dev)
    med, mn = timeit(lambda: ops.cfg_euler_update(z, xc, xu, 3.0, t_dt, 0), iters)
    report("cfg_euler_update", med, mn, bytes_=z.numel() * 16.0)
    xt = torch.randn(28, 1024, 1378, device=dev)
    cond = torch.randn(28, 1024, 1378, device=dev)
    po = torch.empty(56 * 345, 8192, dtype=torch.bfloat16, device=dev)
    med, mn = timeit(lambda: ops.patchify_cast(xt, cond, 56, out=po), iters)
    report("patchify_cast", med, mn, bytes_=2 * xt.numel() * 4.0 + po.numel() * 2.0)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    ap.add_argument("--iters", type=int, default=20)
    a = ap.parse_args()
    if a.only in ("", "gemm"):
        bench_gemm(a.iters)
    if a.only in ("bwd",):
        bench_bwd(a.iters)
    if a.only in ("", "attn"):
        bench_attn(a.iters)
    if a.only in ("", "elem"):
        bench_elem(a.iters)
