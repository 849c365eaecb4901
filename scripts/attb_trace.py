"""Debug: timeline (clock64) of the attention BACKWARD kernel (v2), CTA (0,0,0): compute thread 0 and the MMA issuer.
compute slots it*8+e: 0 top, 1 chunk0 in regs, 2 math0 done, 3 math1 done, 4 P/dS tiles free, 5 stored + released MMA2,
6 next chunk0 requested, 7 dQ epilogue done;  MMA slots 128+it*4+e: 0 S/dP free, 1 next scores issued, 2 P/dS full, 3 MMA2 issued."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jat_b200 import _lib as L, ops
lib = L.load(); ctx = L.context(0)
B, N, Hq, Hkv = 28, 345, 20, 4
p = float(os.environ.get("DROP", "0.1"))
torch.manual_seed(0)
qkv = torch.randn(B * N, (Hq + 2 * Hkv) * 64, device="cuda").to(torch.bfloat16)
lse = torch.empty(B, Hq, N, device="cuda")
out = ops.gqa_attention_fwd(qkv, B, N, Hq, Hkv, lse=lse, drop_p=p, drop_seed=7)
d_out = torch.randn(B * N, Hq * 64, device="cuda").to(torch.bfloat16)
rope = torch.zeros(4096, 64, device="cuda"); cos = rope + 1.0
buf = torch.zeros(256, dtype=torch.int64, device="cuda")
for _ in range(3): ops.gqa_attention_bwd(qkv, d_out, out, lse, cos, rope, B, N, Hq, Hkv, drop_p=p, drop_seed=7)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): ops.gqa_attention_bwd(qkv, d_out, out, lse, cos, rope, B, N, Hq, Hkv, drop_p=p, drop_seed=7)
e1.record(); torch.cuda.synchronize()
print(f"bwd (rowdot + main + finalize + memset) {e0.elapsed_time(e1) / 10 * 1e3:.1f} us per call, dropout {p}")
L.check(lib.jat_debug_set_attention_trace(ctx, buf.data_ptr()))
ops.gqa_attention_bwd(qkv, d_out, out, lse, cos, rope, B, N, Hq, Hkv, drop_p=p, drop_seed=7); torch.cuda.synchronize()
L.check(lib.jat_debug_set_attention_trace(ctx, None))
t = buf.cpu().tolist(); t0 = t[0]
for it in range(15):
    c = [t[it * 8 + e] - t0 for e in range(8)]
    m = [t[128 + it * 4 + e] - t0 if t[128 + it * 4 + e] else -1 for e in range(4)]
    print(f"it {it:2d} top@{c[0]:6d} | ld0 {c[1]-c[0]:5d} math0 {c[2]-c[1]:5d} ld1+math1 {c[3]-c[2]:5d} wait_mma2 {c[4]-c[3]:5d} store {c[5]-c[4]:5d} "
          f"next_ld {c[6]-c[5]:5d} dq_epi {c[7]-c[6]:5d} | MMA: sp_free@{m[0]} scores_issued@{m[1]} pds_full@{m[2]} mma2_issued@{m[3]}")
