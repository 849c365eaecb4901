"""torch.profiler timeline of the DDP training step on rank 0: where the compute stream idles and what NCCL does meanwhile.
torchrun --nproc-per-node N scripts/ddp_timeline.py"""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from torch.profiler import profile, ProfilerActivity

class A: norm = "layernorm"
D_ = bench.Dist()
import jat_b200
from jat_b200 import training
dev, local, world, rank = D_.dev, D_.local, D_.world, D_.rank
model = bench.build_model(dev, "layernorm").train()
model.grad_handoff = "view"
net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local], find_unused_parameters=False,
                                                bucket_cap_mb=int(os.environ.get("JAT_DDP_BUCKET_MB", "200")), gradient_as_bucket_view=True) if world > 1 else model
if os.environ.get("JAT_DDP_GRAD_DTYPE") == "bf16" and world > 1:
    jat_b200.ddp.register_bf16_allreduce(net)
opt = jat_b200.FusedAdamW(model.parameters(), lr=5e-5, weight_decay=0.1, max_grad_norm=1.0, model=model)
B, C, T = bench.B, bench.C, bench.T
gd = torch.Generator(device=dev).manual_seed(100 + rank)
hr = torch.randn(B, C, T, generator=gd, device=dev); lr = torch.randn(B, C, T, generator=gd, device=dev)
mean = torch.zeros(1, C, 1, device=dev); std = torch.ones(1, C, 1, device=dev)
def step():
    t = torch.rand(B, generator=gd, device=dev)
    noise = torch.randn(B, C, T, generator=gd, device=dev)
    hr_n, lr_c, z_t = training.prepare_inputs(hr, lr, mean, std, mean, std, t, noise)
    opt.zero_grad(set_to_none=True)
    loss = training.mse_loss(net(z_t, t, lr_c), hr_n)
    loss.backward()
    opt.step()
for _ in range(4): step()
D_.barrier()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(2): step()
    torch.cuda.synchronize()
D_.barrier()
if rank == 0:
    path = "gpurun_out/ddp_trace.json"
    prof.export_chrome_trace(path)
    ev = [e for e in json.load(open(path))["traceEvents"] if e.get("ph") == "X" and e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
    ev.sort(key=lambda e: e["ts"])
    streams = {}
    for e in ev: streams.setdefault(e["args"].get("stream"), []).append(e)
    for s, es in streams.items():
        busy = sum(e["dur"] for e in es)
        print(f"stream {s}: {len(es)} events, busy {busy/1e3:.2f} ms, span {(es[-1]['ts']+es[-1]['dur']-es[0]['ts'])/1e3:.2f} ms; e.g. {es[0]['name'][:50]}")
    main = max(streams.values(), key=len)
    gaps = []
    for a, b in zip(main, main[1:]):
        g = b["ts"] - (a["ts"] + a["dur"])
        if g > 80: gaps.append((g, a["name"][:60], b["name"][:60], a["ts"] - main[0]["ts"]))
    print("gaps > 80 us on the compute stream (us, after, before, at):")
    for g in gaps: print(f"  {g[0]:8.0f}  at {g[3]/1e3:8.2f} ms   {g[1]}  ->  {g[2]}")
    print("total gap time", sum(g[0] for g in gaps) / 1e3, "ms over 2 steps")
    os.remove(path)
D_.close()
