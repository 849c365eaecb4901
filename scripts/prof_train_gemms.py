"""Driver for ncu: the training-step GEMM shapes of one v3mod2 block (M = 9660 token rows), each launched a few times:
fc1 forward with the pre-activation copy + dropout, GELU' dgrad, out_proj / fc2 forward with the pre-gate copy, wgrad W1."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jat_b200 import _lib as L, ops
dev = torch.device("cuda", 0)
M, D, F, Ntok = 9660, 1280, 5120, 345
g = torch.Generator(device=dev).manual_seed(0)
bf = lambda *s: (torch.randn(*s, generator=g, device=dev) * 0.5).to(torch.bfloat16)
h, W1, W2, dy = bf(M, D), bf(F, D), bf(D, F), bf(M, D)
b1, b2 = torch.randn(F, device=dev), torch.randn(D, device=dev)
u, y = torch.empty(M, F, dtype=torch.bfloat16, device=dev), torch.empty(M, D, dtype=torch.bfloat16, device=dev)
x = torch.zeros(M, D, device=dev)
gate = torch.randn(M // Ntok, D, device=dev)
gW1 = torch.zeros(F, D, device=dev)
reps = int(os.environ.get("REPS", "3"))
ev = lambda: torch.cuda.Event(enable_timing=True)
def timed(name, fn):
    for _ in range(2): fn()
    a, b = ev(), ev(); a.record()
    for _ in range(reps): out = fn()
    b.record(); torch.cuda.synchronize()
    print(f"{name:28s} {a.elapsed_time(b) / reps * 1e3:8.1f} us")
    return out
act = timed("fc1 fwd (aux + drop)", lambda: ops.gemm(h, W1, bias=b1, act=L.ACT_GELU_ERF, aux=u, drop_p=0.1, drop_seed=3))
timed("fc1 fwd (plain)", lambda: ops.gemm(h, W1, bias=b1, act=L.ACT_GELU_ERF))
timed("fc1 fwd (aux only)", lambda: ops.gemm(h, W1, bias=b1, act=L.ACT_GELU_ERF, aux=u))
timed("fc1 fwd (drop only)", lambda: ops.gemm(h, W1, bias=b1, act=L.ACT_GELU_ERF, drop_p=0.1, drop_seed=3))
timed("fc2 fwd (aux + drop)", lambda: ops.gemm(act, W2, kind=L.EPI_GATE_RESIDUAL, bias=b2, out=x, gate=gate, gate_batch_stride=D,
                                                tokens_per_batch=Ntok, aux=y, drop_p=0.1, drop_seed=4))
timed("fc2 fwd (plain)", lambda: ops.gemm(act, W2, kind=L.EPI_GATE_RESIDUAL, bias=b2, out=x, gate=gate, gate_batch_stride=D,
                                           tokens_per_batch=Ntok))
du = timed("gelu' dgrad (aux + drop)", lambda: ops.gemm(dy, W2, kind=L.EPI_DACT, act=L.ACT_GELU_ERF, aux=u, w_transposed=True, drop_p=0.1, drop_seed=3))
timed("dgrad plain bf16", lambda: ops.gemm(du, W1, w_transposed=True))
for ks in (1, 2, 3, 4, 6, 8):
    timed(f"wgrad W1 k_splits={ks}", lambda: ops.gemm(du, h, kind=L.EPI_ACCUM, out=gW1, a_transposed=True, w_transposed=True, k_splits=ks))
