"""BASELINE configs[1] only (v2 288 M, B = 1, 25 steps): the `v2` sub-record of bench.py on its own."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
class A: norm = "layernorm"
D_ = bench.Dist()
r = bench.v2_record(A, D_)
k = r.pop("kernels"); r.pop("config")
print(json.dumps(r))
for n, e in sorted(k.items(), key=lambda x: -x[1]["ms_per_step"]): print(f"   {n:22s} {e['ms_per_step']:.4f} ms/step  {e['us_per_launch']:6.2f} us x {e['launches_per_step']}")
