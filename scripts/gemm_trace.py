"""Decode the GEMM kernel's clock64 trace (jat_debug_set_gemm_trace) for a few headline shapes."""
import math, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jat_b200 import _lib as L, ops
dev = torch.device("cuda", 0)
M, Ntok, B = 19320, 345, 56
g = torch.Generator(device="cpu").manual_seed(0)
trace = torch.zeros(64 * 8, dtype=torch.int64, device=dev)
ctx = L.context(0)
for (N, K, epi) in ((5120, 1280, "bias_bf16"), (5120, 1280, "gelu_bf16"), (1280, 1280, "gate_res"), (1280, 5120, "gate_res"), (1792, 1280, "qkv")):
    A = (torch.randn(M, K, generator=g) * 0.5).to(dev).to(torch.bfloat16)
    W = (torch.randn(N, K, generator=g) / math.sqrt(K)).to(dev).to(torch.bfloat16)
    bias = torch.randn(N, device=dev); gate = torch.randn(B, N, device=dev); xf = torch.randn(M, N, device=dev)
    ob = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
    inv = 1.0 / (10000 ** (torch.arange(0, 64, 2).float() / 64)); e = torch.outer(torch.arange(4096).float(), inv); e = torch.cat([e, e], -1)
    cos, sin = e.cos().to(dev), e.sin().to(dev)
    fn = {"bias_bf16": lambda: ops.gemm(A, W, bias=bias, out=ob),
          "gelu_bf16": lambda: ops.gemm(A, W, bias=bias, out=ob, act=L.ACT_GELU_ERF),
          "gate_res": lambda: ops.gemm(A, W, kind=L.EPI_GATE_RESIDUAL, out=xf, bias=bias, gate=gate, gate_batch_stride=N, tokens_per_batch=Ntok),
          "qkv": lambda: ops.gemm(A, W, kind=L.EPI_QKV_ROPE, out=ob, tokens_per_batch=Ntok, rope_cos=cos, rope_sin=sin, rope_cols=1536)}[epi]
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    trace.zero_()
    L.check(L.load().jat_debug_set_gemm_trace(ctx, trace.data_ptr()))
    fn()
    torch.cuda.synchronize()
    L.check(L.load().jat_debug_set_gemm_trace(ctx, None))
    t = trace.cpu().view(64, 8)
    n = int((t[:, 3] != 0).sum())
    t0 = int(t[0, 0])
    print(f"== N={N} K={K} {epi}: {n} work items on cluster 0; per item (clk rel. to start): prod_start mma_wait mma_go mma_commit | epi_wait epi_go epi_done | mma_busy epi_busy")
    for i in range(min(n, 8)):
        r = [int(v) - t0 for v in t[i, :7]]
        print(f"   item {i}: {r[0]:7d} {r[1]:7d} {r[2]:7d} {r[3]:7d} | {r[4]:7d} {r[5]:7d} {r[6]:7d} | {r[3]-r[2]:6d} {r[6]-r[5]:6d}")
    if n > 2:
        per = (int(t[n - 1, 3]) - int(t[1, 3])) / (n - 2)
        print(f"   steady-state period {per:.0f} clk per item (commit to commit)")
