"""CPU tests of the long-audio chunk scheduler (host logic) and of its oracle."""
import os
import sys

import numpy as np
import pytest
import torch

from tests._util import have_reference

from oracle import dit_oracle as O  # checker only


def test_plan_matches_reference_arithmetic():
    from jat_b200.chunked import CHUNK_FRAMES, OVERLAP_FRAMES, plan_chunks
    assert (CHUNK_FRAMES, OVERLAP_FRAMES) == (1378, 172)          # infer_test_v3m2.py:345-346
    for total in (173, 1000, 1378, 1379, 2584, 2585, 51679, 100000):
        assert plan_chunks(total) == O.plan_chunks(total)
    assert len(plan_chunks(51679)) == 43                          # 10-minute track
    for total in (173, 1379, 51679):                              # the last chunk is always longer than the overlap
        s, e = plan_chunks(total)[-1]
        assert e - s > OVERLAP_FRAMES and e == total


def test_assign_is_a_partition():
    from jat_b200.chunked import assign_chunks
    for n, w in ((43, 8), (5, 8), (1, 2), (16, 4)):
        got = sorted(c for r in range(w) for c in assign_chunks(n, r, w))
        assert got == list(range(n))
        assert max(len(assign_chunks(n, r, w)) for r in range(w)) - min(len(assign_chunks(n, r, w)) for r in range(w)) <= 1


def test_oracle_linspace_matches_torch():
    for a, b, n in ((0.0, 1.0, 172), (1.0, 0.0, 172), (0.0, 1.0, 7), (1.0, 0.0, 2)):
        assert np.array_equal(O.linspace_f32(a, b, n), torch.linspace(a, b, n).numpy())


@pytest.mark.skipif(not have_reference(), reason="/root/reference not present (GPU box)")
def test_oracle_crossfade_matches_reference():
    from tests._util import import_reference
    import_reference()
    import infer_test_v3m2 as ref
    g = torch.Generator().manual_seed(3)
    for lens, ov in (((40, 40, 40, 23), 8), ((1378, 1378, 500), 172), ((30,), 8), ((20, 20), 0)):
        chunks = [torch.randn(1, 6, n, generator=g) for n in lens]
        want = ref.crossfade_chunks(chunks, ov).numpy()
        got = O.crossfade_chunks([c.numpy() for c in chunks], ov)
        assert got.shape == want.shape and np.array_equal(got, want)


def _worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from jat_b200.chunked import assign_chunks, gather_chunks, plan_chunks
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        total, Tc, ov, C = 230, 50, 10, 3
        plan = plan_chunks(total, Tc, ov)
        track = torch.arange(C * total, dtype=torch.float32).reshape(C, total)
        # fake per-chunk "sampler": out = 2 * chunk + chunk id  (deterministic, chunk-local)
        local = {cid: 2.0 * track[:, plan[cid][0]:plan[cid][1]] + cid for cid in assign_chunks(len(plan), rank, world)}
        stacked = gather_chunks(local, len(plan), (C, Tc), torch.device("cpu"))
        q.put((rank, stacked.numpy(), plan))
    finally:
        dist.destroy_process_group()


def test_gather_chunks_world2_gloo():
    """Round-robin shard over 2 gloo ranks -> every rank holds all chunks in chunk order; stitched result equals
    the single-process oracle of the reference chunk loop."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 1000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=120) for _ in procs), key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    total, Tc, ov, C = 230, 50, 10, 3
    track = np.arange(C * total, dtype=np.float32).reshape(C, total)
    plan = res[0][2]
    assert np.array_equal(res[0][1], res[1][1])
    chunks = [res[0][1][i:i + 1, :, : e - s] for i, (s, e) in enumerate(plan)]
    got = O.crossfade_chunks(chunks, ov)
    want = O.sample_long(lambda lr, i: 2.0 * lr + i, track, np.float32(0), np.float32(1), np.float32(0), np.float32(1),
                         chunk_frames=Tc, overlap_frames=ov)
    assert got.shape == (1, C, total) and np.array_equal(got, want)
