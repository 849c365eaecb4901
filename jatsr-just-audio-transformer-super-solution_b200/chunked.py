"""Long-audio chunk scheduler around the sampler -- the caller of the hot path for BASELINE config C5.

Mirrors the chunk loop of the reference inference script (infer_test_v3m2.py:340-406) and its
`crossfade_chunks` (:188-233):
  * 16 s chunks (1378 latent frames) every 1206 frames (172 frames = 2 s overlap), the last one shorter;
  * each chunk: (x - mean) / std -> flow_matching_sample -> * std + mean;
  * results stitched by a left-fold linear crossfade over the overlaps.
The reference runs the chunks one at a time at batch 1.  Here all full-length chunks of a rank form ONE
batch [n, C, 1378] built by a single kernel (`jat_chunk_normalize`), are denoised together, and one kernel
(`jat_crossfade_denorm`) de-normalises and stitches them.  Across GPUs the chunks are dealt round-robin
(no collective on the data path; finished chunks are all-gathered once at the end).

Randomness: the reference draws `torch.randn(1, C, T)` per chunk from the global generator, in chunk order.
With one rank we draw exactly the same way, so the same seed gives the same noise per chunk.
"""
from __future__ import annotations

import torch

from . import ops
from .sampler import flow_matching_sample

DAC_SAMPLE_RATE, DAC_HOP = 44100, 512
CHUNK_FRAMES = int(16.0 * DAC_SAMPLE_RATE / DAC_HOP)    # 1378 (infer_test_v3m2.py:345)
OVERLAP_FRAMES = int(2.0 * DAC_SAMPLE_RATE / DAC_HOP)   # 172  (infer_test_v3m2.py:346)


def plan_chunks(total_frames, chunk_frames=CHUNK_FRAMES, overlap_frames=OVERLAP_FRAMES):
    """[(start, end)] of every chunk, same arithmetic as infer_test_v3m2.py:358-372."""
    if total_frames <= 0:
        return []
    if chunk_frames <= overlap_frames:
        raise ValueError("chunk_frames must exceed overlap_frames")
    stride = chunk_frames - overlap_frames
    num_chunks = max((total_frames - overlap_frames + stride - 1) // stride, 1)
    return [(i * stride, min(i * stride + chunk_frames, total_frames)) for i in range(num_chunks)]


def assign_chunks(num_chunks, rank, world_size):
    """Round-robin shard: the chunk ids rank `rank` denoises."""
    return list(range(rank, num_chunks, world_size))


def gather_chunks(local, num_chunks, chunk_shape, device, group=None):
    """All ranks contribute {chunk id: tensor[C, <=Tc]}; returns the stacked [num_chunks, C, Tc] tensor (short
    chunks zero-padded) on every rank.  world_size 1 needs no process group."""
    import torch.distributed as dist
    Cc, Tc = chunk_shape
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank(group) if world > 1 else 0
    per_rank = (num_chunks + world - 1) // world
    mine = torch.zeros(per_rank, Cc, Tc, dtype=torch.float32, device=device)
    for slot, cid in enumerate(assign_chunks(num_chunks, rank, world)):
        t = local[cid]
        mine[slot, :, : t.shape[-1]] = t
    if world == 1:
        parts = [mine]
    else:
        parts = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine, group=group)
    out = torch.zeros(num_chunks, Cc, Tc, dtype=torch.float32, device=device)
    for r in range(world):
        for slot, cid in enumerate(assign_chunks(num_chunks, r, world)):
            out[cid] = parts[r][slot]
    return out


def crossfade_chunks(chunks, overlap_frames, mean=None, std=None):
    """Drop-in for the reference `crossfade_chunks(chunks, overlap_frames)` (:188-233): `chunks` is a list of
    [1, C, T_i] CUDA tensors (all T_i equal except the last, which may be shorter); returns [1, C, total].
    Optional per-channel mean/std ([C] or [1, C, 1]) are applied as chunk * std + mean before blending."""
    if len(chunks) == 0:
        return None
    if len(chunks) == 1 and mean is None:
        return chunks[0]
    Cc, Tc = chunks[0].shape[-2], chunks[0].shape[-1]
    dev = chunks[0].device
    if any(c.shape[-1] != Tc for c in chunks[:-1]) or chunks[-1].shape[-1] > Tc:
        raise ValueError("all chunks but the last must have the same length")
    if len(chunks) > 1 and chunks[-1].shape[-1] < overlap_frames:
        raise ValueError("last chunk shorter than the overlap")
    stacked = torch.zeros(len(chunks), Cc, Tc, dtype=torch.float32, device=dev)
    for i, c in enumerate(chunks):
        stacked[i, :, : c.shape[-1]] = c.reshape(Cc, -1)
    total = (len(chunks) - 1) * (Tc - overlap_frames) + chunks[-1].shape[-1]
    return _stitch(stacked, overlap_frames, total, mean, std).unsqueeze(0)


def _stitch(stacked, overlap_frames, total_frames, mean, std):
    dev = stacked.device
    fi = torch.linspace(0.0, 1.0, overlap_frames, device=dev) if overlap_frames > 0 else None
    fo = torch.linspace(1.0, 0.0, overlap_frames, device=dev) if overlap_frames > 0 else None
    flat = lambda v: None if v is None else v.to(device=dev, dtype=torch.float32).reshape(-1).contiguous()
    return ops.crossfade_denorm(stacked, overlap_frames, total_frames, fi, fo, flat(mean), flat(std))


@torch.no_grad()
def sample_long(model, lr_latent, lr_mean=None, lr_std=None, hr_mean=None, hr_std=None, num_steps=50, cfg_scale=1.0,
                device="cuda", total_frames=None, chunk_frames=CHUNK_FRAMES, overlap_frames=OVERLAP_FRAMES,
                group=None, sample_fn=None, verbose=False):
    """lr_latent [C, total] (un-normalised LR latent track) -> generated HR latent [1, C, total] on every rank.

    Equivalent to infer_test_v3m2.py:358-402 (chunk loop + crossfade).  `sample_fn(model, lr_batch, z0)` defaults
    to the fused CFG sampler; under torch.distributed the chunks are dealt round-robin over the ranks of `group`."""
    import torch.distributed as dist
    device = torch.device(device)
    lr_latent = lr_latent.to(device=device, dtype=torch.float32)
    Cc = getattr(model, "input_channels", lr_latent.shape[0])   # channels of the GENERATED latent (= the condition's in the reference's configs)
    total = lr_latent.shape[-1] if total_frames is None else min(total_frames, lr_latent.shape[-1])
    plan = plan_chunks(total, chunk_frames, overlap_frames)
    stride = chunk_frames - overlap_frames
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank(group) if world > 1 else 0
    mine = assign_chunks(len(plan), rank, world)
    flat = lambda v: None if v is None else v.to(device=device, dtype=torch.float32).reshape(-1).contiguous()
    lr_m, lr_s = flat(lr_mean), flat(lr_std)
    if sample_fn is None:
        sample_fn = lambda m, lr, z0: flow_matching_sample(m, lr, num_steps=num_steps, cfg_scale=cfg_scale, device=device,
                                                           verbose=verbose, z0=z0)
    local = {}
    if mine:
        # noise per chunk, drawn in chunk order exactly like the reference (:133 inside the chunk loop)
        noise = {cid: torch.randn(1, Cc, plan[cid][1] - plan[cid][0], device=device) for cid in mine}
        full = [cid for cid in mine if plan[cid][1] - plan[cid][0] == chunk_frames]
        if full:
            step = world if len(full) > 1 else 1
            assert all(full[i] == full[0] + i * step for i in range(len(full)))
            batch = ops.chunk_normalize(lr_latent[:, :total], len(full), chunk_frames, stride, lr_m, lr_s,
                                        first_chunk=full[0], chunk_step=step)
            gen = sample_fn(model, batch, torch.cat([noise[cid] for cid in full], 0))
            for i, cid in enumerate(full):
                local[cid] = gen[i]
        for cid in mine:  # the short last chunk has its own token count -> its own batch-1 problem
            if cid in local:
                continue
            s, e = plan[cid]
            part = ops.chunk_normalize(lr_latent[:, :total], 1, chunk_frames, stride, lr_m, lr_s, first_chunk=cid)
            local[cid] = sample_fn(model, part[:, :, : e - s].contiguous(), noise[cid])[0]
    stacked = gather_chunks(local, len(plan), (Cc, chunk_frames), device, group)
    return _stitch(stacked, overlap_frames if len(plan) > 1 else 0, total, hr_mean, hr_std).unsqueeze(0)
