"""Small end-to-end exercise of every kernel family for compute-sanitizer (memcheck / racecheck / initcheck):
tiny model forward (both norm kinds), 2-step CFG sampler, training forward + backward with Dropout / DropPath, long
sequence attention, chunk kernels, training glue.  Exits non-zero on any mismatch against loose sanity bounds."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import jat_b200
from jat_b200 import ops, training
from tests._util import rerandomise_zero_init

dev = torch.device("cuda", 0)
cfg = dict(input_channels=32, cond_channels=32, patch_len=4, hidden_size=128, depth=2, num_q_heads=2, num_kv_heads=1,
           bottleneck_dim=128, mlp_ratio=2.0, dropout=0.1, drop_path_rate=0.1)
for cls in (jat_b200.JaT_AudioSR_V2, jat_b200.JaT_AudioSR_V3):
    torch.manual_seed(0)
    m = rerandomise_zero_init(cls(**cfg)).to(dev).eval()
    x, c, t = torch.randn(2, 32, 86, device=dev), torch.randn(2, 32, 86, device=dev), torch.rand(2, device=dev)
    with torch.no_grad():
        y = m(x, t, c)
        z = jat_b200.flow_matching_sample(m, c, num_steps=2, cfg_scale=3.0, device=dev, verbose=False, use_graph=False)
        yl = m(torch.randn(1, 32, 1600, device=dev), t[:1], torch.randn(1, 32, 1600, device=dev))   # N = 400 > 352
    assert torch.isfinite(y).all() and torch.isfinite(z).all() and torch.isfinite(yl).all()
    m.train()
    hr_n, lr_c, z_t = training.prepare_inputs(x, c, torch.zeros(32, device=dev), torch.ones(32, device=dev), torch.zeros(32, device=dev),
                                              torch.ones(32, device=dev), t, torch.randn_like(x), cond_noise=torch.randn_like(x),
                                              cond_scale=0.05)
    loss = training.mse_loss(m(z_t, t, lr_c), hr_n)
    loss.backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in m.parameters())
# round 2: 16-epilogue-warp GEMM epilogues (256-wide CTA-pair tiles with the training extras), attention backward v2 at the
# headline sequence length with dropout, Charbonnier loss, bf16 gradient exchange passes, cond_channels != input_channels
from jat_b200 import _lib as L
A = torch.randn(300, 128, device=dev).to(torch.bfloat16)
W = torch.randn(512, 128, device=dev).to(torch.bfloat16) * 0.1
u = torch.empty(300, 512, dtype=torch.bfloat16, device=dev)
act = ops.gemm(A, W, bias=torch.randn(512, device=dev), act=L.ACT_GELU_ERF, aux=u, drop_p=0.1, drop_seed=3, cta_pair=1, block_n=256)
W2 = torch.randn(256, 512, device=dev).to(torch.bfloat16) * 0.1
du = ops.gemm(torch.randn(300, 256, device=dev).to(torch.bfloat16), W2, kind=L.EPI_DACT, act=L.ACT_GELU_ERF, aux=u, w_transposed=True,
              drop_p=0.1, drop_seed=3, cta_pair=1, block_n=256)
assert torch.isfinite(act.float()).all() and torch.isfinite(du.float()).all() and torch.isfinite(u.float()).all()
Bq, Nq, Hq, Hkv = 1, 345, 5, 1
qkv = torch.randn(Bq * Nq, (Hq + 2 * Hkv) * 64, device=dev).to(torch.bfloat16)
lse = torch.empty(Bq, Hq, Nq, device=dev)
o = ops.gqa_attention_fwd(qkv, Bq, Nq, Hq, Hkv, lse=lse, drop_p=0.1, drop_seed=5)
rope = torch.zeros(4096, 64, device=dev)
dqkv = ops.gqa_attention_bwd(qkv, torch.randn_like(o), o, lse, rope + 1.0, rope, Bq, Nq, Hq, Hkv, drop_p=0.1, drop_seed=5)
assert torch.isfinite(dqkv.float()).all()
pr = torch.randn(3, 8, 87, device=dev, requires_grad=True)
training.charbonnier_loss(pr, torch.randn(3, 8, 87, device=dev)).backward()
assert torch.isfinite(pr.grad).all()
gbuf, pay = torch.randn(100003, device=dev)[:100000], torch.empty(100000, dtype=torch.bfloat16, device=dev)
L.check(L.load().jat_grad_compress(L.context(0), gbuf.data_ptr(), pay.data_ptr(), 100000, 0.5, torch.cuda.current_stream().cuda_stream))
L.check(L.load().jat_grad_decompress(L.context(0), pay.data_ptr(), gbuf.data_ptr(), 100000, torch.cuda.current_stream().cuda_stream))
assert torch.isfinite(gbuf).all()
mc = rerandomise_zero_init(jat_b200.JaT_AudioSR_V2(**dict(cfg, cond_channels=96))).to(dev).eval()
with torch.no_grad():
    assert torch.isfinite(mc(torch.randn(2, 32, 86, device=dev), torch.rand(2, device=dev), torch.randn(2, 96, 86, device=dev))).all()
lat = torch.randn(32, 3000, device=dev)
ch = ops.chunk_normalize(lat, 2, 1378, 1206, torch.zeros(32, device=dev), torch.ones(32, device=dev))
fi, fo = torch.linspace(0, 1, 172, device=dev), torch.linspace(1, 0, 172, device=dev)
out = ops.crossfade_denorm(ch, 172, 2584, fi, fo, torch.zeros(32, device=dev), torch.ones(32, device=dev))
assert torch.isfinite(out).all()
torch.cuda.synchronize()
print("sanitize_small: ok")
