"""Experiment: which epilogue limits the short-K (K = 1280) GEMMs?  Same M x N x K, different epilogues, L2 warm / cold,
with and without operand loads (JAT_DBG_GEMM_SKIP=3 -> pure MMA + epilogue)."""
import json, math, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jat_b200 import _lib as L, ops
from scripts.bench_kernels import timeit, dev

M, Ntok, B = 19320, 345, 56
g = torch.Generator(device="cpu").manual_seed(0)
for (N, K) in ((1280, 1280), (5120, 1280), (1792, 1280)):
    A = (torch.randn(M, K, generator=g) * 0.5).to(dev).to(torch.bfloat16)
    W = (torch.randn(N, K, generator=g) / math.sqrt(K)).to(dev).to(torch.bfloat16)
    bias = torch.randn(N, device=dev)
    gate = torch.randn(B, N, device=dev)
    xf = torch.randn(M, N, device=dev)
    ob = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
    cases = {
        "gate_residual_f32_redadd": lambda: ops.gemm(A, W, kind=L.EPI_GATE_RESIDUAL, out=xf, bias=bias, gate=gate, gate_batch_stride=N, tokens_per_batch=Ntok),
        "bias_f32_store": lambda: ops.gemm(A, W, bias=bias, out=xf, out_dtype=L.DTYPE_F32),
        "bias_bf16_store": lambda: ops.gemm(A, W, bias=bias, out=ob),
        "bias_gelu_bf16_store": lambda: ops.gemm(A, W, bias=bias, out=ob, act=L.ACT_GELU_ERF),
    }
    for name, fn in cases.items():
        for flush in (True, False):
            med, mn = timeit(fn, 10, flush=flush)
            print(json.dumps({"N": N, "K": K, "epi": name, "l2": "cold" if flush else "warm", "ms": round(med, 4),
                              "tflops": round(2.0 * M * N * K / med / 1e9, 1)}), flush=True)
