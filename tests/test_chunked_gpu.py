"""GPU parity of the long-audio chunk path (config C5): chunk extraction + normalisation, de-normalise + crossfade
(bit-exact against the oracle / torch expressions), and the batched chunk scheduler against the reference's
chunk-by-chunk loop."""
import numpy as np
import pytest
import torch

from oracle import dit_oracle as O  # checker only
from tests._util import rerandomise_zero_init

pytestmark = pytest.mark.gpu


def dev():
    return torch.device("cuda", 0)


@pytest.mark.parametrize("total,Tc,ov", [(5000, 1378, 172), (1379, 1378, 172), (230, 50, 10), (50, 50, 10), (37, 50, 10)])
def test_chunk_normalize_bit_exact(total, Tc, ov):
    from jat_b200 import ops
    from jat_b200.chunked import plan_chunks
    torch.manual_seed(0)
    C = 24
    track = torch.randn(C, total, device=dev()) * 3 + 1
    mean, std = torch.randn(C, device=dev()), torch.rand(C, device=dev()) + 0.5
    plan = plan_chunks(total, Tc, ov)
    got = ops.chunk_normalize(track, len(plan), Tc, Tc - ov, mean, std)
    for i, (s, e) in enumerate(plan):
        want = (track[:, s:e] - mean[:, None]) / std[:, None]          # infer_test_v3m2.py:381-382
        assert torch.equal(got[i, :, : e - s], want)
        assert (got[i, :, e - s:] == 0).all()
    # strided shard (rank 1 of 2) and no-normalisation mode
    if len(plan) >= 3:
        sub = ops.chunk_normalize(track, (len(plan) - 1 + 1) // 2, Tc, Tc - ov, mean, std, first_chunk=1, chunk_step=2)
        assert torch.equal(sub[0], got[1])
    raw = ops.chunk_normalize(track, len(plan), Tc, Tc - ov)
    assert torch.equal(raw[0, :, : plan[0][1]], track[:, : plan[0][1]])


@pytest.mark.parametrize("lens,ov", [((1378, 1378, 1378, 500), 172), ((40, 40, 40, 23), 8), ((30,), 8), ((20, 20), 0),
                                      ((1378,) * 6 + (1100,), 172)])
@pytest.mark.parametrize("denorm", [False, True])
def test_crossfade_bit_exact(lens, ov, denorm):
    from jat_b200.chunked import crossfade_chunks
    g = torch.Generator().manual_seed(5)
    C = 16
    chunks = [torch.randn(1, C, n, generator=g) for n in lens]
    mean, std = torch.randn(C, generator=g), torch.rand(C, generator=g) + 0.5
    if denorm:
        ref_chunks = [(c * std.view(1, C, 1) + mean.view(1, C, 1)).numpy() for c in chunks]   # :394
        got = crossfade_chunks([c.to(dev()) for c in chunks], ov, mean.to(dev()), std.to(dev()))
    else:
        ref_chunks = [c.numpy() for c in chunks]
        got = crossfade_chunks([c.to(dev()) for c in chunks], ov)
    want = O.crossfade_chunks(ref_chunks, ov)
    assert tuple(got.shape) == want.shape
    assert np.array_equal(got.cpu().numpy(), want)


def test_crossfade_rejects_bad_geometry():
    from jat_b200 import _lib as L, ops
    chunks = torch.zeros(3, 4, 20, device=dev())
    fi = torch.linspace(0, 1, 12, device=dev())
    with pytest.raises(L.JatError):
        ops.crossfade_denorm(chunks, 12, 40, fi, fi)      # chunk_frames < 2 * overlap
    with pytest.raises(L.JatError):
        ops.crossfade_denorm(chunks, 4, 500, fi, fi)      # total does not match 3 chunks


def test_sample_long_matches_chunk_by_chunk_loop():
    """Batched scheduler (all full chunks as one batch) == the reference's serial loop: per chunk normalise, sample
    with the same per-chunk noise, de-normalise, left-fold crossfade (oracle.sample_long around OUR sampler at B=1)."""
    import jat_b200
    from jat_b200.chunked import plan_chunks, sample_long
    cfg = dict(input_channels=32, cond_channels=32, patch_len=4, hidden_size=128, depth=2, num_q_heads=2, num_kv_heads=1,
               bottleneck_dim=128, mlp_ratio=2.0)
    torch.manual_seed(0)
    model = rerandomise_zero_init(jat_b200.JaT_AudioSR_V2(**cfg)).to(dev()).eval()
    C, total, Tc, ov = 32, 330, 120, 20
    g = torch.Generator().manual_seed(11)
    track = torch.randn(C, total, generator=g) * 2 + 0.5
    stats = [torch.randn(C, generator=g), torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g),
             torch.rand(C, generator=g) + 0.5]
    lr_mean, lr_std, hr_mean, hr_std = [s.view(1, C, 1) for s in stats]
    plan = plan_chunks(total, Tc, ov)
    assert len(plan) == 4 and plan[-1][1] - plan[-1][0] < Tc

    torch.manual_seed(123)
    got = sample_long(model, track.to(dev()), lr_mean, lr_std, hr_mean, hr_std, num_steps=4, cfg_scale=3.0,
                      chunk_frames=Tc, overlap_frames=ov).cpu().numpy()

    torch.manual_seed(123)   # the serial loop draws the same per-chunk noise in the same order
    def per_chunk(lr_norm, i):
        return jat_b200.flow_matching_sample(model, torch.from_numpy(lr_norm).to(dev()), num_steps=4, cfg_scale=3.0,
                                             device=dev(), verbose=False).cpu().numpy()
    want = O.sample_long(per_chunk, track.numpy(), lr_mean.numpy(), lr_std.numpy(), hr_mean.numpy(), hr_std.numpy(),
                         chunk_frames=Tc, overlap_frames=ov)
    assert got.shape == want.shape == (1, C, total)
    rel = float(np.linalg.norm(got - want) / np.linalg.norm(want))
    assert rel < 1e-5, rel   # same kernels, same noise; only the batch composition differs
