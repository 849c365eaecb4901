"""torch.profiler timeline of the sampling step: idle time on the stream between consecutive kernels, by (previous, next) kernel class."""
import collections, json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from torch.profiler import profile, ProfilerActivity
from jat_b200.sampler import _Plan
dev = torch.device("cuda", 0)
model = bench.build_model(dev, "layernorm")
B, C, T = bench.B, bench.C, bench.T
plan = _Plan(model, B, C, T, 12, bench.CFG_SCALE, dev)
plan.z.normal_(); plan.lr.normal_()
plan.mod = model._engine.modulation(plan.ws, plan.t_curr)
bench.run_steps(model, plan, 0, 4)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    bench.run_steps(model, plan, 4, 4)
    torch.cuda.synchronize()
path = "gpurun_out/sample_trace.json"
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("ph") == "X" and e.get("cat") == "kernel"]
os.remove(path)
ev.sort(key=lambda e: e["ts"])
def cls(n):
    import re
    m = re.search(r"gemm_tcgen05_kernel<(\d+), (\d+), (\d+)", n)
    if m: return "gemm_" + ["bias_act", "qkv_rope", "gate_res", "unpatch", "accum", "dact"][int(m.group(3))]
    for k in ("gqa_attention_fwd", "adaln_norm", "patchify", "cfg_euler"):
        if k in n: return k
    return n[:20]
gaps = collections.defaultdict(list)
for a, b in zip(ev, ev[1:]):
    gaps[(cls(a["name"]), cls(b["name"]))].append(b["ts"] - (a["ts"] + a["dur"]))
busy = sum(e["dur"] for e in ev); span = ev[-1]["ts"] + ev[-1]["dur"] - ev[0]["ts"]
print(f"{len(ev)} kernels, busy {busy/1e3:.2f} ms, span {span/1e3:.2f} ms, idle {100*(1-busy/span):.1f} %  (PDL={os.environ.get('JAT_PDL','0')})")
for k, v in sorted(gaps.items(), key=lambda kv: -sum(kv[1])):
    print(f"  {k[0]:18s} -> {k[1]:18s} n={len(v):4d} mean gap {sum(v)/len(v):6.2f} us  total {sum(v)/1e3:6.3f} ms")
