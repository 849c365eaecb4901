"""Diagnostics of the fused parameter update at full model size: isolated timing of each pass (CUDA events) next to
torch's clip_grad_norm_ + fused AdamW + re-cast, and the first steps' losses under both optimizers."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import jat_b200  # noqa: E402
from jat_b200 import _lib as L  # noqa: E402

dev = torch.device("cuda", 0)
shapes = []
for _ in range(28):
    shapes += [(7680, 1280), (7680,), (1280, 1280), (256, 1280), (256, 1280), (1280, 1280), (5120, 1280), (5120,), (1280, 5120), (1280,)]
shapes += [(512, 8192), (512,), (1280, 512), (1280,), (4096, 1280), (4096,)]
g = torch.Generator(device=dev).manual_seed(0)


def make():
    return [torch.nn.Parameter(torch.randn(s, generator=g, device=dev) * 0.02) for s in shapes]


def timeit(fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


ps = make()
numel = sum(p.numel() for p in ps)
flat = torch.randn(numel, generator=g, device=dev) * 1e-3
for p, v in zip(ps, flat.split([p.numel() for p in ps])):
    p.grad = v.view(p.shape)
bf = [torch.empty(p.shape, dtype=torch.bfloat16, device=dev) for p in ps]
print("elements", numel / 1e6, "M")
topt = torch.optim.AdamW(ps, lr=5e-5, weight_decay=0.1, fused=True)


def torch_step():
    torch.nn.utils.clip_grad_norm_(ps, 1.0)
    topt.step()
    torch._foreach_copy_(bf, [p.detach() for p in ps])


print("torch clip + fused AdamW + recast: %.3f ms" % timeit(torch_step))
print("  torch clip only: %.3f ms" % timeit(lambda: torch.nn.utils.clip_grad_norm_(ps, 1.0)))
print("  torch fused AdamW only: %.3f ms" % timeit(topt.step))
print("  torch recast only: %.3f ms" % timeit(lambda: torch._foreach_copy_(bf, [p.detach() for p in ps])))
fopt = jat_b200.FusedAdamW(ps, lr=5e-5, weight_decay=0.1, max_grad_norm=1.0)
print("FusedAdamW (clip, no packed copy): %.3f ms" % timeit(fopt.step))
L.profile_begin(0)
for _ in range(3):
    fopt.step()
print("  per class:", {k: (round(v[0] / 3, 3), v[1] // 3) for k, v in L.profile_end(0).items()})
fopt2 = jat_b200.FusedAdamW(ps, lr=5e-5, weight_decay=0.1)
print("FusedAdamW (no clip): %.3f ms" % timeit(fopt2.step))
t0 = time.perf_counter()
for _ in range(5):
    fopt.step()
print("host time per step() call: %.3f ms" % ((time.perf_counter() - t0) / 5 * 1e3))
torch.cuda.synchronize()
