import sys, os, torch
sys.path.insert(0, "/root/repo")
from tests.test_model_gpu import build, dev
cfg = dict(hidden_size=1280, depth=int(os.environ.get("DEPTH", "2")), num_q_heads=20, num_kv_heads=4, bottleneck_dim=512)
model = build("JaT_AudioSR_V2", cfg, seed=77, bf16_exact=False).to(dev())
g = torch.Generator().manual_seed(11)
B = 28
W = int(os.environ.get("W", "2"))
z = torch.randn(B, 1024, 1378, generator=g).to(dev())
lr = torch.randn(B, 1024, 1378, generator=g).to(dev())
t = torch.full((2 * B,), 0.37, device=dev())
big, bb = model.forward_with_blocks(torch.cat([z, z]), t, torch.cat([lr, torch.zeros_like(lr)]))
big = big.clone(); bb = bb.clone()
for lo in (5, 26, 20):
    for uncond in (False, True):
        sm, sb = model.forward_with_blocks(z[lo:lo + W], t[:W], torch.zeros_like(lr[lo:lo+W]) if uncond else lr[lo:lo + W])
        o = lo + (B if uncond else 0)
        d = (big[o:o + W] - sm).abs()
        N = 345
        blk = [(bb[i, o * N:(o + W) * N] - sb[i]).abs().max().item() for i in range(bb.shape[0])]
        print(f"item {o}: out max diff {d.max().item():.3e} nonzero {int((d>0).sum())}; per-block residual max diff {blk}")
        if d.max() > 0:
            idx = (d[0] > 0).nonzero()
            print("   first diffs (c, t):", idx[:5].tolist(), " last:", idx[-3:].tolist())
            r = (bb[0, o * N:(o + W) * N] - sb[0]).abs()
            rows = (r.max(1).values > 0).nonzero().flatten()
            print("   block-0 differing token rows:", rows[:10].tolist(), "...", rows[-5:].tolist(), "count", len(rows), " cols:", (r.max(0).values > 0).nonzero().flatten()[:8].tolist())
