"""Turn the ncu artefacts of scripts/gpu_profile.sh (gpurun_out/) into the tracked summaries under profiles/:
  profiles/<tag>_launches.csv            the launch list (per-launch gpu__time_duration, cold cache, serialised)
  profiles/<tag>_launch_shares.json      per-kernel-class share of one step from that list
  profiles/<tag>_ncu_full_summary.json   selected metrics of the `--set full` captures (per launch)
Usage: python scripts/summarize_ncu.py r1b"""
import csv, io, json, os, re, subprocess, sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = os.path.join(root, "profiles")
os.makedirs(out, exist_ok=True)
KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum", "lts__t_sectors_srcunit_tex_op_read.sum",
        "lts__t_sectors_srcunit_tex_op_red.sum", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second"]

# ---- launch list
src = os.path.join(root, "gpurun_out", "launches.csv")
lines = [l for l in open(src) if l.startswith('"')]
open(os.path.join(out, f"{tag}_launches.csv"), "w").writelines(lines)
rows = list(csv.DictReader(io.StringIO("".join(lines))))
def cls(name):
    m = re.search(r"gemm_tcgen05_kernel<(\d+), (\d+), (\d+)", name)
    if m:
        return "gemm_" + ["bias_act", "qkv_rope", "gate_residual", "unpatchify", "accum", "dact"][int(m.group(3))]
    for k in ("gqa_attention_fwd", "adaln_norm_modulate", "patchify_cast", "cfg_euler_update", "timestep_features"):
        if k in name:
            return k
    return name.split("(")[0]
tot = {}
for r in rows:
    c = cls(r["Kernel Name"])
    t = tot.setdefault(c, [0.0, 0])
    t[0] += float(r["Metric Value"]) / 1e6
    t[1] += 1
allms = sum(v[0] for v in tot.values())
shares = {k: {"ms_total": round(v[0], 3), "launches": v[1], "share": round(v[0] / allms, 4)} for k, v in sorted(tot.items(), key=lambda kv: -kv[1][0])}
json.dump({"note": "ncu --metrics gpu__time_duration.sum --clock-control none over `bench.py --steps 2 --warmup 1` (3 forward passes + "
                   "the once-per-run timestep path); cold-cache, serialised launches: compare SHARES with bench.py's live event timing",
           "total_ms": round(allms, 3), "classes": shares}, open(os.path.join(out, f"{tag}_launch_shares.json"), "w"), indent=1)
print(json.dumps(shares, indent=1))

# ---- full captures
summary = {}
for name in ("gemm", "attn", "adaln"):
    rep = os.path.join(root, "gpurun_out", f"prof_{name}.ncu-rep")
    if not os.path.exists(rep):
        continue
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rd = list(csv.reader(io.StringIO(txt)))
    hdr, units = rd[0], rd[1]
    items = []
    for r in rd[2:]:
        d = {"Kernel Name": r[hdr.index("Kernel Name")]}
        for k in KEEP:
            if k in hdr:
                i = hdr.index(k)
                d[k] = f"{r[i]} {units[i]}".strip()
        items.append(d)
    summary[name] = items
json.dump(summary, open(os.path.join(out, f"{tag}_ncu_full_summary.json"), "w"), indent=1)
for name, items in summary.items():
    for d in items:
        print(name, d["Kernel Name"][:70], "|", d.get("gpu__time_duration.sum"), "| tensor active", d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
              "| dram r/w", d.get("dram__bytes_read.sum"), d.get("dram__bytes_write.sum"), "| tma_ld", d.get("l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum"))
