"""Turn the ncu artefacts of scripts/gpu_profile_train.sh (gpurun_out/) into tracked summaries under profiles/:
  profiles/<tag>_train_launches.csv.gz       launch list of the whole `bench.py --mode train --steps 1 --warmup 1` run
                                             (model construction + 3 training steps; cold-cache, serialised launches)
  profiles/<tag>_train_launch_shares.json    per-kernel totals / launch counts / average duration from that list
  profiles/<tag>_train_ncu_full_summary.json selected metrics of the `--set full` captures of the backward / optimizer kernels
Usage: python scripts/summarize_ncu_train.py r1c"""
import collections, csv, gzip, io, json, os, re, subprocess, sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r1c"
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = os.path.join(root, "profiles")
KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second"]

src = os.path.join(root, "gpurun_out", "train_launches.csv")
lines = [l for l in open(src) if l.startswith('"')]
with gzip.open(os.path.join(out, f"{tag}_train_launches.csv.gz"), "wt") as f:
    f.writelines(lines)
rows = list(csv.DictReader(io.StringIO("".join(lines))))
tot = collections.OrderedDict()
for r in rows:
    n = re.sub(r"^void ", "", r["Kernel Name"])
    n = re.sub(r"\(.*", "", n)[:90]
    t = tot.setdefault(n, [0.0, 0])
    t[0] += float(r["Metric Value"]) / 1e6
    t[1] += 1
mine = {k: v for k, v in tot.items() if k.startswith("jat::")}
all_ms, mine_ms = sum(v[0] for v in tot.values()), sum(v[0] for v in mine.values())
js = {"note": "ncu --metrics gpu__time_duration.sum --clock-control none over the whole `bench.py --mode train --steps 1 --warmup 1` "
              "process: model construction + 3 training steps (warm-up, timed, profiled).  Library kernels (jat::*) run only "
              "inside the steps, so ms_total / 3 is per step; torch kernels include the one-off initialisation.  Cold-cache, "
              "serialised launches: compare SHARES with bench.py's live event timing.",
      "total_ms": round(all_ms, 3), "library_ms": round(mine_ms, 3),
      "kernels": {k: {"ms_total": round(v[0], 3), "launches": v[1], "avg_us": round(v[0] / v[1] * 1e3, 2),
                      "share_of_library_time": round(v[0] / mine_ms, 4) if k in mine else None}
                  for k, v in sorted(tot.items(), key=lambda kv: -kv[1][0]) if v[0] > 0.05}}
json.dump(js, open(os.path.join(out, f"{tag}_train_launch_shares.json"), "w"), indent=1)
for k, v in list(js["kernels"].items())[:25]:
    print(f"{v['ms_total'] / 3:8.3f} ms/step  x{v['launches'] / 3:6.1f}  avg {v['avg_us']:8.1f} us  {k}")

summary = {}
for name in ("train_bwd", "train_opt"):
    rep = os.path.join(root, "gpurun_out", f"prof_{name}.ncu-rep")
    if not os.path.exists(rep):
        continue
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rd = list(csv.reader(io.StringIO(txt)))
    hdr, units = rd[0], rd[1]
    items = []
    for r in rd[2:]:
        d = {"Kernel Name": re.sub(r"\(.*", "", r[hdr.index("Kernel Name")])}
        for k in KEEP:
            if k in hdr:
                i = hdr.index(k)
                d[k] = f"{r[i]} {units[i]}".strip()
        items.append(d)
    summary[name] = items
json.dump(summary, open(os.path.join(out, f"{tag}_train_ncu_full_summary.json"), "w"), indent=1)
for name, items in summary.items():
    for d in items:
        print(name, d["Kernel Name"][:60], "|", d.get("gpu__time_duration.sum"), "| dram r/w", d.get("dram__bytes_read.sum"),
              d.get("dram__bytes_write.sum"), "| dram %", d.get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
              "| regs", d.get("launch__registers_per_thread"), "| warps %", d.get("sm__warps_active.avg.pct_of_peak_sustained_active"))
