import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, jat_b200
from jat_b200 import chunked
dev = torch.device("cuda", 0)
model = bench.build_model(dev, "layernorm")
C, frames = 1024, 51679
track = (torch.randn(C, frames) * 2.0 + 0.3).pin_memory()
mean, std = torch.full((C,), 0.3), torch.full((C,), 2.0)
import jat_b200.sampler as S
orig = S.flow_matching_sample
def timed_sample(*a, **k):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    r = orig(*a, **k)
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"   flow_matching_sample B={a[1].shape[0]} T={a[1].shape[2]}: host {1e3*(t1-t0):.0f} ms, +gpu drain {1e3*(t2-t1):.0f} ms")
    return r
chunked.flow_matching_sample = timed_sample
for i in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out = chunked.sample_long(model, track.to(dev, non_blocking=True), mean, std, mean, std, num_steps=50, cfg_scale=3.0, device=dev)
    torch.cuda.synchronize(); print(f"run {i}: {1e3*(time.perf_counter()-t0):.0f} ms")
