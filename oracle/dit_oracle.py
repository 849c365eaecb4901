"""CPU oracle for the JaT-AudioSR DiT denoiser hot path -- TEST INFRASTRUCTURE ONLY.

A plain numpy restatement of the reference algorithm (HUSRCF/JaTSR-Just-audio-transformer-super-solution):
  * model forward   src/models/jat_audiosr_v2.py:399-448 (LayerNorm class) and
                    src/models/jat_audiosr_v3.py:422-471 (RMSNorm class; differs only at :261,264,384)
  * Euler/CFG sampler  infer_test_v3m2.py:108-185

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / --impl reference legs may import
this module, and only as the checker / reported baseline -- never on the product path.

Parity pinning: the reference ships no golden vectors or known-answer tests for this path (SURVEY.md
8c), so this oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF: `tests/golden/make_golden.py`
imports the reference modules from /root/reference on CPU, runs them in fp32 and stores inputs,
weights and outputs in `tests/golden/*.npz`; `tests/test_oracle.py` checks this file against those
fixtures (and directly against the reference modules when /root/reference is present).

Weights are passed as a dict {state_dict key: ndarray} using the reference's key layout.
All math is done in `dtype` (float32 to mirror the reference's inference path, float64 for a
higher-precision truth).
"""
from __future__ import annotations

import math

import numpy as np

try:  # scipy is in the image; keep a dependency-free fallback
    from scipy.special import erf as _erf
except Exception:  # pragma: no cover
    _erf = np.vectorize(math.erf)


# ------------------------------------------------------------------------------------------- pieces
def rope_tables(max_seq_len=4096, dim=64, base=10000.0, dtype=np.float32):
    """RoPE.__init__ (jat_audiosr_v2.py:53-68): cos/sin [max_seq_len, dim] of cat(freqs, freqs)."""
    inv_freq = (1.0 / (np.float32(base) ** (np.arange(0, dim, 2, dtype=np.float32) / np.float32(dim)))).astype(np.float32)
    t = np.arange(max_seq_len, dtype=np.float32)
    freqs = np.outer(t, inv_freq).astype(np.float32)
    emb = np.concatenate([freqs, freqs], axis=-1)
    return np.cos(emb).astype(dtype), np.sin(emb).astype(dtype), inv_freq


def apply_rope(x, cos, sin):
    """RoPE.forward / rotate_half (jat_audiosr_v2.py:70-91).  x [B, N, H, hd]; position = token index."""
    n = x.shape[1]
    hd = x.shape[-1]
    x1, x2 = x[..., : hd // 2], x[..., hd // 2:]
    rot = np.concatenate([-x2, x1], axis=-1)
    return x * cos[None, :n, None, :] + rot * sin[None, :n, None, :]


def timestep_embedding(t, dim, dtype=np.float32):
    """TimeEmbedding.forward (jat_audiosr_v2.py:177-190): t in [0,1], no x1000 scaling."""
    half = dim // 2
    k = np.float32(math.log(10000) / (half - 1))
    f = np.exp(np.arange(half, dtype=np.float32) * -k).astype(np.float32)
    e = t.astype(np.float32)[:, None] * f[None, :]
    return np.concatenate([np.sin(e), np.cos(e)], axis=-1).astype(dtype)


def gelu_erf(x):
    """nn.GELU() default = exact erf form (jat_audiosr_v2.py:206,249)."""
    return (0.5 * x * (1.0 + _erf(x / math.sqrt(2.0)))).astype(x.dtype)


def silu(x):
    return (x / (1.0 + np.exp(-x))).astype(x.dtype)


def linear(x, w, b=None):
    y = x @ w.T
    if b is not None:
        y = y + b
    return y


def layer_norm(x, eps=1e-6):
    """nn.LayerNorm(elementwise_affine=False, eps=1e-6): biased variance (jat_audiosr_v2.py:242)."""
    mu = x.mean(axis=-1, keepdims=True)
    var = ((x - mu) ** 2).mean(axis=-1, keepdims=True)
    return (x - mu) / np.sqrt(var + eps)


def rms_norm(x, weight, eps=1e-6):
    """nn.RMSNorm(D, eps=1e-6) (jat_audiosr_v3.py:261)."""
    return x / np.sqrt((x * x).mean(axis=-1, keepdims=True) + eps) * weight


def patchify(x_in, patch_len):
    """BottleneckPatchEmbed1D.forward reshape (jat_audiosr_v2.py:225-227): feature index = c*P + p."""
    B, C, T = x_in.shape
    assert T % patch_len == 0
    return x_in.reshape(B, C, T // patch_len, patch_len).transpose(0, 2, 1, 3).reshape(B, T // patch_len, C * patch_len)


def unpatchify(x, B, C, patch_len, T_orig):
    """JaT_AudioSR_V2.unpatchify (jat_audiosr_v2.py:383-397)."""
    N = x.shape[1]
    y = x.reshape(B, N, C, patch_len).transpose(0, 2, 1, 3).reshape(B, C, N * patch_len)
    return y[:, :, :T_orig]


def gqa_attention(x, p, prefix, num_q_heads, num_kv_heads, cos, sin):
    """GroupedQueryAttention.forward, eval mode (jat_audiosr_v2.py:127-167)."""
    B, N, D = x.shape
    hd = D // num_q_heads
    G = num_q_heads // num_kv_heads
    q = linear(x, p[prefix + "q_proj.weight"]).reshape(B, N, num_q_heads, hd)
    k = linear(x, p[prefix + "k_proj.weight"]).reshape(B, N, num_kv_heads, hd)
    v = linear(x, p[prefix + "v_proj.weight"]).reshape(B, N, num_kv_heads, hd)
    q = apply_rope(q, cos, sin)
    k = apply_rope(k, cos, sin)
    k = np.repeat(k, G, axis=2)  # repeat_interleave: q head h uses kv head h // G
    v = np.repeat(v, G, axis=2)
    q, k, v = (a.transpose(0, 2, 1, 3) for a in (q, k, v))  # [B, H, N, hd]
    s = (q @ k.transpose(0, 1, 3, 2)) / math.sqrt(hd)
    s = s - s.max(axis=-1, keepdims=True)
    w = np.exp(s)
    w = w / w.sum(axis=-1, keepdims=True)
    o = (w @ v).transpose(0, 2, 1, 3).reshape(B, N, D)
    return linear(o, p[prefix + "out_proj.weight"])


# ------------------------------------------------------------------------------------------- model
def model_dims(p):
    """Recover the constructor hyper-parameters from state_dict shapes."""
    hidden = p["t_embedder.1.weight"].shape[0]
    depth = 0
    while f"blocks.{depth}.attn.q_proj.weight" in p:
        depth += 1
    kv = p["blocks.0.attn.k_proj.weight"].shape[0]
    return dict(hidden=hidden, depth=depth, bottleneck=p["patch_embed.proj.0.weight"].shape[0],
                flat_in=p["patch_embed.proj.0.weight"].shape[1], out_dim=p["final_layer.1.weight"].shape[0],
                kv_hidden=kv, rms=("blocks.0.norm1.weight" in p))


def dit_forward(p, x_t, t, x_cond, *, num_q_heads, num_kv_heads, patch_len=4, dtype=np.float32,
                return_blocks=False, max_len=2048):
    """JaT_AudioSR_V2/V3.forward in eval mode.  `p` = state_dict as numpy arrays.  The norm kind is
    inferred from the presence of `blocks.0.norm1.weight` (RMSNorm class)."""
    p = {k: np.asarray(v, dtype=dtype) for k, v in p.items() if not k.endswith(("inv_freq", "cos_cached", "sin_cached"))}
    x_t = np.asarray(x_t, dtype=dtype)
    x_cond = np.asarray(x_cond, dtype=dtype)
    dims = model_dims(p)
    D, depth, rms = dims["hidden"], dims["depth"], dims["rms"]
    B, C, T_orig = x_t.shape
    P = patch_len
    pad = (P - T_orig % P) % P
    if pad:  # jat_audiosr_v2.py:411-416
        x_t = np.pad(x_t, ((0, 0), (0, 0), (0, pad)))
        x_cond = np.pad(x_cond, ((0, 0), (0, 0), (0, pad)))
    x_in = np.concatenate([x_t, x_cond], axis=1)  # :421
    tok = patchify(x_in, P)
    x = linear(gelu_erf(linear(tok, p["patch_embed.proj.0.weight"], p["patch_embed.proj.0.bias"])),
               p["patch_embed.proj.2.weight"], p["patch_embed.proj.2.bias"])  # :204-208
    N = x.shape[1]
    if N > max_len:
        raise ValueError(f"Sequence length {N} exceeds max_len {max_len}")  # :428-429
    te = timestep_embedding(np.asarray(t), D, dtype)
    t_emb = linear(silu(linear(te, p["t_embedder.1.weight"], p["t_embedder.1.bias"])),
                   p["t_embedder.3.weight"], p["t_embedder.3.bias"])  # :341-346
    hd = D // num_q_heads
    cos, sin, _ = rope_tables(4096, hd, dtype=dtype)
    blocks = []
    for i in range(depth):  # DiTBlock_GQA.forward :265-289
        pre = f"blocks.{i}."
        mod = linear(silu(t_emb), p[pre + "adaLN_modulation.1.weight"], p[pre + "adaLN_modulation.1.bias"])
        sh_a, sc_a, g_a, sh_m, sc_m, g_m = np.split(mod, 6, axis=1)
        h = rms_norm(x, p[pre + "norm1.weight"]) if rms else layer_norm(x)
        h = h * (1 + sc_a[:, None, :]) + sh_a[:, None, :]
        x = x + g_a[:, None, :] * gqa_attention(h, p, pre + "attn.", num_q_heads, num_kv_heads, cos, sin)
        h = rms_norm(x, p[pre + "norm2.weight"]) if rms else layer_norm(x)
        h = h * (1 + sc_m[:, None, :]) + sh_m[:, None, :]
        m = linear(gelu_erf(linear(h, p[pre + "mlp.0.weight"], p[pre + "mlp.0.bias"])),
                   p[pre + "mlp.3.weight"], p[pre + "mlp.3.bias"])
        x = x + g_m[:, None, :] * m
        if return_blocks:
            blocks.append(x.copy())
    h = rms_norm(x, p["final_layer.0.weight"]) if rms else layer_norm(x)  # no modulation on the final layer, :361
    y = linear(h, p["final_layer.1.weight"], p["final_layer.1.bias"])
    out = unpatchify(y, B, C, P, T_orig)  # crop removes the padding, :442-446
    return (out, blocks) if return_blocks else out


# ------------------------------------------------------------------------------------------- sampler
def sampler_schedule(num_steps):
    """timesteps = linspace(0, 1, steps+1) in fp32; per step (t_i, dt_i = t_{i+1} - t_i)
    (infer_test_v3m2.py:136,145-147)."""
    # torch.linspace(0., 1., steps+1) in fp32: ATen fills symmetrically from both ends with a fused
    # multiply-add (RangeFactories; same on CPU-vectorised and CUDA), emulated here in float64.
    n = num_steps + 1
    step = np.float64(np.float32(1.0) / np.float32(n - 1))
    i = np.arange(n)
    lo = (step * i).astype(np.float32)
    hi = (1.0 - step * (n - 1 - i)).astype(np.float32)
    return np.where(i < n // 2, lo, hi).astype(np.float32)


def euler_cfg_update(z, x_c, x_u, cfg_scale, t, dt):
    """infer_test_v3m2.py:164,173-179 in fp32, operation by operation."""
    f = np.float32
    x = x_c if x_u is None else (x_u + f(cfg_scale) * (x_c - x_u)).astype(np.float32)
    if f(t) < f(0.999):
        den = f(f(1.0) - f(t)) + f(1e-5)
        return (z + ((x - z) / den).astype(np.float32) * f(dt)).astype(np.float32)
    return x.astype(np.float32)


def flow_matching_sample(p, lr_latent, z0, *, num_steps=50, cfg_scale=1.0, num_q_heads, num_kv_heads,
                         patch_len=4, dtype=np.float32, timesteps=None):
    """flow_matching_sample (infer_test_v3m2.py:108-185) with the initial noise z0 injected
    (the reference draws it with torch.randn from the global generator, :133)."""
    lr = np.asarray(lr_latent, dtype=np.float32)
    z = np.asarray(z0, dtype=np.float32).copy()
    B = lr.shape[0]
    ts = sampler_schedule(num_steps) if timesteps is None else np.asarray(timesteps, dtype=np.float32)
    use_cfg = cfg_scale != 1.0
    kw = dict(num_q_heads=num_q_heads, num_kv_heads=num_kv_heads, patch_len=patch_len, dtype=dtype)
    for i in range(num_steps):
        t, dt = ts[i], np.float32(ts[i + 1] - ts[i])
        tb = np.full((B,), t, dtype=np.float32)
        if use_cfg:
            both = dit_forward(p, np.concatenate([z, z]), np.concatenate([tb, tb]),
                               np.concatenate([lr, np.zeros_like(lr)]), **kw).astype(np.float32)
            x_c, x_u = both[:B], both[B:]
        else:
            x_c, x_u = dit_forward(p, z, tb, lr, **kw).astype(np.float32), None
        z = euler_cfg_update(z, x_c, x_u, cfg_scale, t, dt)
    return z


# ------------------------------------------------------------------------------------------- long-audio chunk loop
def plan_chunks(total_frames, chunk_frames=1378, overlap_frames=172):
    """Chunk boundaries of infer_test_v3m2.py:358-372: stride = chunk - overlap,
    num_chunks = ceil((total - overlap) / stride), chunk i = [i*stride, min(i*stride + chunk, total))."""
    stride = chunk_frames - overlap_frames
    n = (total_frames - overlap_frames + stride - 1) // stride
    return [(i * stride, min(i * stride + chunk_frames, total_frames)) for i in range(n)]


def linspace_f32(start, end, steps):
    """torch.linspace(start, end, steps) in fp32: ATen fills symmetrically from both ends with a fused multiply-add
    (RangeFactories; same on CPU-vectorised and CUDA) -- emulated in float64 with the fp32 step (cf. sampler_schedule)."""
    if steps == 1:
        return np.array([start], dtype=np.float32)
    step = np.float64(np.float32((np.float32(end) - np.float32(start)) / np.float32(steps - 1)))
    i = np.arange(steps)
    lo = (np.float64(np.float32(start)) + step * i).astype(np.float32)
    hi = (np.float64(np.float32(end)) - step * (steps - 1 - i)).astype(np.float32)
    return np.where(i < steps // 2, lo, hi).astype(np.float32)


def crossfade_chunks(chunks, overlap_frames):
    """crossfade_chunks (infer_test_v3m2.py:188-233): left fold over [1, C, T_i] chunks; in every overlap the running
    result fades out with linspace(1, 0, overlap) and the new chunk fades in with linspace(0, 1, overlap)."""
    if len(chunks) == 0:
        return None
    result = np.asarray(chunks[0], dtype=np.float32)
    for cur in chunks[1:]:
        cur = np.asarray(cur, dtype=np.float32)
        if overlap_frames > 0 and result.shape[-1] >= overlap_frames:
            fo = linspace_f32(1.0, 0.0, overlap_frames).reshape(1, 1, -1)
            fi = linspace_f32(0.0, 1.0, overlap_frames).reshape(1, 1, -1)
            blended = ((result[..., -overlap_frames:] * fo).astype(np.float32)
                       + (cur[..., :overlap_frames] * fi).astype(np.float32)).astype(np.float32)
            result = np.concatenate([result[..., :-overlap_frames], blended, cur[..., overlap_frames:]], axis=-1)
        else:
            result = np.concatenate([result, cur], axis=-1)
    return result


def sample_long(sample_fn, lr_latent, lr_mean, lr_std, hr_mean, hr_std, chunk_frames=1378, overlap_frames=172,
                total_frames=None):
    """The chunk loop of infer_test_v3m2.py:370-402 around an arbitrary per-chunk sampler `sample_fn(lr_norm[1,C,T],
    chunk_index) -> [1,C,T]`: normalise (:381-382), sample (:385), de-normalise (:394), crossfade (:402)."""
    lr_latent = np.asarray(lr_latent, dtype=np.float32)
    total = lr_latent.shape[-1] if total_frames is None else min(total_frames, lr_latent.shape[-1])
    outs = []
    for i, (s, e) in enumerate(plan_chunks(total, chunk_frames, overlap_frames)):
        lr = lr_latent[None, :, s:e]
        lr_norm = ((lr - lr_mean) / lr_std).astype(np.float32)
        gen = np.asarray(sample_fn(lr_norm, i), dtype=np.float32)
        outs.append(((gen * hr_std).astype(np.float32) + hr_mean).astype(np.float32))
    return crossfade_chunks(outs, overlap_frames)
