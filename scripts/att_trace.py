"""Debug: dump the attention kernel's pipeline timeline (clock64) for CTA (0,0,0)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jat_b200 import _lib as L, ops
lib = L.load(); ctx = L.context(0)
B, N, Hq, Hkv = 56, 345, 20, 4
qkv = torch.randn(B * N, (Hq + 2 * Hkv) * 64, device="cuda").to(torch.bfloat16)
buf = torch.zeros(128, dtype=torch.int64, device="cuda")
for _ in range(3): ops.gqa_attention_fwd(qkv, B, N, Hq, Hkv)
L.check(lib.jat_debug_set_attention_trace(ctx, buf.data_ptr()))
ops.gqa_attention_fwd(qkv, B, N, Hq, Hkv); torch.cuda.synchronize()
L.check(lib.jat_debug_set_attention_trace(ctx, None))
t = buf.cpu().tolist(); t0 = t[127]
print("kernel start -> end:", t[126] - t0)
for h in range(5):
    m = [t[4*h+i]-t0 for i in range(4)]
    a = [t[32+4*h+i]-t0 for i in range(4)]; bb = [t[64+4*h+i]-t0 for i in range(4)]
    print(f"head {h}: MMA SB(h)@{m[0]} SA(h+1)@{m[1]} PV_A@{m[2]} PV_B@{m[3]} | WG0 s_full@{a[0]} turn@{a[1]} exp@{a[2]} p_full@{a[3]}"
          f" | WG1 s_full@{bb[0]} turn@{bb[1]} exp@{bb[2]} p_full@{bb[3]} | EPI done@{t[96+h]-t0}")
