"""Torch-tensor wrappers around the C-ABI kernels (one function per `jat_*` entry point).

PyTorch is used for device memory and streams only; every computation below happens inside
``libjat_b200.so``.  CPU tensors are rejected: there is no fallback path.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib as L


def _stream(dev) -> int:
    return torch.cuda.current_stream(dev).cuda_stream


def _ctx(t: torch.Tensor) -> int:
    if not t.is_cuda:
        raise RuntimeError("jat_b200 kernels run on CUDA tensors only (no CPU fallback)")
    return L.context(t.device.index if t.device.index is not None else torch.cuda.current_device())


def _p(t):
    return 0 if t is None else t.data_ptr()


def _chk(t, dtype, name):
    if t.dtype != dtype or not t.is_contiguous():
        raise ValueError(f"{name}: expected contiguous {dtype}, got {t.dtype} contiguous={t.is_contiguous()}")


def adaln_norm_modulate(x, shift=None, scale=None, mod_batch_stride=0, weight=None, norm_kind=L.NORM_LAYERNORM,
                        eps=1e-6, tokens_per_batch=0, out=None):
    """x f32 [M, D] -> bf16 [M, D]; shift/scale are f32 tensors (or views) whose row b starts at
    data_ptr + b * mod_batch_stride elements."""
    _chk(x, torch.float32, "x")
    M, D = x.shape
    if out is None:
        out = torch.empty(M, D, dtype=torch.bfloat16, device=x.device)
    L.check(L.load().jat_adaln_norm_modulate(_ctx(x), x.data_ptr(), out.data_ptr(), _p(shift), _p(scale),
                                             mod_batch_stride, _p(weight), norm_kind, eps, M, D, tokens_per_batch,
                                             _stream(x.device)))
    return out


def patchify_cast(x_t, x_cond, B, cond_batch=None, out=None, patch_len=4):
    _chk(x_t, torch.float32, "x_t")
    xb, Cc, T = x_t.shape
    N = (T + patch_len - 1) // patch_len
    if x_cond is not None:
        _chk(x_cond, torch.float32, "x_cond")
        if cond_batch is None:
            cond_batch = x_cond.shape[0]
    else:
        cond_batch = 0
    if out is None:
        out = torch.empty(B * N, 2 * Cc * patch_len, dtype=torch.bfloat16, device=x_t.device)
    L.check(L.load().jat_patchify_cast(_ctx(x_t), x_t.data_ptr(), xb, _p(x_cond), cond_batch, out.data_ptr(), B, Cc,
                                       T, patch_len, _stream(x_t.device)))
    return out


def timestep_features(t, D, out=None):
    _chk(t, torch.float32, "t")
    B = t.shape[0]
    if out is None:
        out = torch.empty(B, D, dtype=torch.bfloat16, device=t.device)
    L.check(L.load().jat_timestep_features(_ctx(t), t.data_ptr(), out.data_ptr(), B, D, _stream(t.device)))
    return out


def gemm(A, W, *, kind=L.EPI_BIAS_ACT, act=L.ACT_NONE, out_dtype=L.DTYPE_BF16, bias=None, out=None,
         gate=None, gate_batch_stride=0, tokens_per_batch=0, rope_cos=None, rope_sin=None, rope_cols=0,
         patch_len=4, t_out=0, cta_pair=-1, block_n=0, aux=None, k_splits=0, a_transposed=False, w_transposed=False,
         drop_p=0.0, drop_seed=0, gate_rowscale=None):
    """acc = A[M,K] @ W[N,K]^T (bf16 in, f32 accumulate) + fused epilogue; returns `out`.
    a_transposed / w_transposed: the tensor passed holds A^T [K, M] / W^T [K, N] (backward GEMMs, no copies)."""
    _chk(A, torch.bfloat16, "A")
    _chk(W, torch.bfloat16, "W")
    (K, M) = A.shape if a_transposed else A.shape[::-1]
    (K2, N) = W.shape if w_transposed else W.shape[::-1]
    assert K == K2, (A.shape, W.shape)
    if out is None:
        if kind in (L.EPI_BIAS_ACT, L.EPI_DACT):
            bf = out_dtype == L.DTYPE_BF16 or kind == L.EPI_DACT
            out = torch.empty(M, N, dtype=torch.bfloat16 if bf else torch.float32, device=A.device)
        elif kind == L.EPI_QKV_ROPE:
            out = torch.empty(M, N, dtype=torch.bfloat16, device=A.device)
        else:
            raise ValueError("this epilogue needs an explicit `out`")
    e = L.GemmEpilogue()
    e.kind, e.act, e.out_dtype, e.tokens_per_batch = kind, act, out_dtype, tokens_per_batch
    e.bias, e.out = _p(bias), out.data_ptr()
    e.ldo = out.stride(0) if kind != L.EPI_UNPATCHIFY else 0
    e.gate, e.gate_batch_stride = _p(gate), gate_batch_stride
    e.rope_cos, e.rope_sin, e.rope_cols = _p(rope_cos), _p(rope_sin), rope_cols
    e.patch_len, e.t_out, e.k_splits = patch_len, t_out, k_splits
    e.aux, e.ld_aux = _p(aux), (aux.stride(0) if aux is not None else 0)
    e.a_transposed, e.w_transposed = int(a_transposed), int(w_transposed)
    e.drop_p, e.drop_seed, e.gate_rowscale = float(drop_p), int(drop_seed), _p(gate_rowscale)
    L.check(L.load().jat_gemm_bf16(_ctx(A), A.data_ptr(), A.stride(0), W.data_ptr(), W.stride(0), M, N, K,
                                   C.byref(e), cta_pair, block_n, _stream(A.device)))
    return out


def set_gemm_tail_split(device, mode):
    """GEMM tail split on a device's context: 0 / False = off, 1 / True = deterministic split-K fix-up of a partial last
    wave (any epilogue), 2 = reduce-add epilogues only, parts added straight into the output (no workspace round trip)."""
    dev = torch.device(device)
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    L.check(L.load().jat_set_gemm_tail_split(L.context(idx), int(mode)))


def set_gemm_sm_reserve(device, reserve):
    """Keep `reserve` SMs out of the persistent GEMM grids of a device's context (room for concurrent NCCL kernels)."""
    dev = torch.device(device)
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    L.check(L.load().jat_set_gemm_sm_reserve(L.context(idx), int(reserve)))


def gqa_attention_fwd(qkv, B, N, Hq, Hkv, head_dim=64, out=None, lse=None, drop_p=0.0, drop_seed=0):
    """lse: optional f32 [B, Hq, N] output (log2-domain log-sum-exp per query row, for the backward pass).
    drop_p / drop_seed: train-mode dropout on the probabilities (mask row = (b*Hq + h)*N + query, col = key)."""
    _chk(qkv, torch.bfloat16, "qkv")
    assert qkv.shape == (B * N, (Hq + 2 * Hkv) * head_dim)
    if out is None:
        out = torch.empty(B * N, Hq * head_dim, dtype=torch.bfloat16, device=qkv.device)
    lib = L.load()
    passes = lib.jat_attention_passes(N)
    if passes <= 1:
        L.check(lib.jat_gqa_attention_fwd_dropout(_ctx(qkv), qkv.data_ptr(), out.data_ptr(), _p(lse), B, N, Hq, Hkv,
                                                  head_dim, float(drop_p), int(drop_seed), _stream(qkv.device)))
        return out
    part_o = torch.empty(passes, B * N, Hq * head_dim, dtype=torch.bfloat16, device=qkv.device)
    part_lse = torch.empty(passes, B, Hq, N, dtype=torch.float32, device=qkv.device)
    L.check(lib.jat_gqa_attention_fwd_long(_ctx(qkv), qkv.data_ptr(), out.data_ptr(), _p(lse), part_o.data_ptr(),
                                           part_lse.data_ptr(), B, N, Hq, Hkv, head_dim, float(drop_p), int(drop_seed),
                                           _stream(qkv.device)))
    return out


def dropout_site_seed(seed, block, site):
    return int(L.load().jat_dropout_site_seed(int(seed), int(block), int(site)))


def dropout_scale_mask(rows, cols, p, site_seed, device):
    """f32 [rows, cols]: the multiplier (0 or 1/(1-p)) the fused kernels apply at (row, col) of a dropout site."""
    out = torch.empty(rows, cols, dtype=torch.float32, device=device)
    L.check(L.load().jat_dropout_scale_mask(_ctx(out), out.data_ptr(), rows, cols, float(p), int(site_seed),
                                            _stream(out.device)))
    return out


def drop_path_scales(rates, B, seed):
    """rates f32 [depth] (device) -> f32 [depth, 2, B] per-sample DropPath factors (branch 0 = attention, 1 = MLP)."""
    _chk(rates, torch.float32, "rates")
    out = torch.empty(rates.shape[0], 2, B, dtype=torch.float32, device=rates.device)
    L.check(L.load().jat_drop_path_scales(_ctx(rates), out.data_ptr(), rates.data_ptr(), rates.shape[0], B, int(seed),
                                          _stream(rates.device)))
    return out


def cfg_euler_update(z, x_c, x_u, cfg_scale, t_dt, step):
    """In-place fused CFG combine + x-pred->velocity + Euler update of z (all f32, same numel)."""
    _chk(z, torch.float32, "z")
    _chk(x_c, torch.float32, "x_c")
    if x_u is not None:
        _chk(x_u, torch.float32, "x_u")
    _chk(t_dt, torch.float32, "t_dt")
    L.check(L.load().jat_cfg_euler_update(_ctx(z), z.data_ptr(), x_c.data_ptr(), _p(x_u), float(cfg_scale),
                                          t_dt.data_ptr(), step, z.numel(), _stream(z.device)))
    return z


def chunk_normalize(latent, n_chunks, chunk_frames, stride, mean=None, std=None, first_chunk=0, chunk_step=1, out=None):
    """latent f32 [C, total] (row pitch = stride(0)) -> f32 [n_chunks, C, chunk_frames] normalised chunk batch."""
    if latent.dtype != torch.float32 or latent.dim() != 2 or latent.stride(1) != 1:
        raise ValueError("latent: expected f32 [C, total] with unit inner stride")
    Cc, total = latent.shape
    if out is None:
        out = torch.empty(n_chunks, Cc, chunk_frames, dtype=torch.float32, device=latent.device)
    for v, n in ((mean, "mean"), (std, "std")):
        if v is not None:
            _chk(v, torch.float32, n)
    L.check(L.load().jat_chunk_normalize(_ctx(latent), latent.data_ptr(), total, latent.stride(0), _p(mean), _p(std),
                                         out.data_ptr(), n_chunks, first_chunk, chunk_step, Cc, chunk_frames, stride,
                                         _stream(latent.device)))
    return out


def crossfade_denorm(chunks, overlap, total_frames, fade_in=None, fade_out=None, mean=None, std=None, out=None):
    """chunks f32 [n, C, Tc] -> f32 [C, total_frames]: (chunk * std + mean) stitched with the linear crossfade."""
    _chk(chunks, torch.float32, "chunks")
    n, Cc, Tc = chunks.shape
    if out is None:
        out = torch.empty(Cc, total_frames, dtype=torch.float32, device=chunks.device)
    L.check(L.load().jat_crossfade_denorm(_ctx(chunks), chunks.data_ptr(), n, Cc, Tc, overlap, _p(fade_in), _p(fade_out),
                                          _p(mean), _p(std), out.data_ptr(), total_frames, out.stride(0),
                                          _stream(chunks.device)))
    return out


# ------------------------------------------------------------------------------------------------ backward pieces
def adaln_bwd(dh, x, B, tokens_per_batch, dx, *, scale=None, mod_batch_stride=0, weight=None, norm_kind=L.NORM_LAYERNORM,
              eps=1e-6, accumulate=True, dshift=None, dscale=None, dmod_batch_stride=0, dweight=None):
    _chk(dh, torch.bfloat16, "dh")
    _chk(x, torch.float32, "x")
    _chk(dx, torch.float32, "dx")
    D = x.shape[1]
    rowstats = torch.empty(x.shape[0], 2, dtype=torch.float32, device=x.device)
    L.check(L.load().jat_adaln_bwd(_ctx(x), dh.data_ptr(), x.data_ptr(), _p(scale), mod_batch_stride, _p(weight), norm_kind,
                                   eps, dx.data_ptr(), int(accumulate), _p(dshift), _p(dscale), dmod_batch_stride,
                                   _p(dweight), rowstats.data_ptr(), B, tokens_per_batch, D, _stream(x.device)))
    return dx


def gate_bwd(dx, y, gate, B, tokens_per_batch, dgate, *, mod_batch_stride=0, dmod_batch_stride=0, dbias=None, dy=None,
             drop_p=0.0, drop_seed=0, gate_rowscale=None):
    _chk(dx, torch.float32, "dx")
    _chk(y, torch.bfloat16, "y")
    D = dx.shape[1]
    if dy is None:
        dy = torch.empty_like(y)
    scratch = torch.empty(B, D, dtype=torch.float32, device=dx.device) if dbias is not None else None
    L.check(L.load().jat_gate_bwd_dropout(_ctx(dx), dx.data_ptr(), y.data_ptr(), gate.data_ptr(), mod_batch_stride,
                                          dy.data_ptr(), dgate.data_ptr(), dmod_batch_stride, _p(scratch), _p(dbias), B,
                                          tokens_per_batch, D, float(drop_p), int(drop_seed), _p(gate_rowscale),
                                          _stream(dx.device)))
    return dy


def adaln_gate_bwd(dh, x, rowstats, B, tokens_per_batch, dx, *, scale, mod_batch_stride, dshift, dscale, dmod_batch_stride,
                   weight=None, norm_kind=L.NORM_LAYERNORM, dweight=None, y=None, gate=None, dgate=None, dy=None, dbias=None,
                   drop_p=0.0, drop_seed=0, gate_rowscale=None):
    """Fused norm + modulate backward (dx accumulated in place, dshift / dscale [/ dweight] atomically added) and, when `y`
    is given, the gate backward of the branch below on the updated dx row (-> dy bf16, dgate [, dbias])."""
    _chk(dh, torch.bfloat16, "dh")
    _chk(x, torch.float32, "x")
    _chk(dx, torch.float32, "dx")
    _chk(rowstats, torch.float32, "rowstats")
    D = x.shape[1]
    if y is not None and dy is None:
        dy = torch.empty_like(y)
    scratch = torch.empty(B, D, dtype=torch.float32, device=dx.device) if dbias is not None else None
    L.check(L.load().jat_adaln_gate_bwd(_ctx(x), dh.data_ptr(), x.data_ptr(), rowstats.data_ptr(), scale.data_ptr(),
                                        mod_batch_stride, _p(weight), norm_kind, dx.data_ptr(), dshift.data_ptr(),
                                        dscale.data_ptr(), dmod_batch_stride, _p(dweight), _p(y), _p(gate), _p(dy), _p(dgate),
                                        _p(scratch), _p(dbias), B, tokens_per_batch, D, float(drop_p), int(drop_seed),
                                        _p(gate_rowscale), _stream(x.device)))
    return dy


def colsum_bf16(a, out):
    _chk(a, torch.bfloat16, "a")
    L.check(L.load().jat_colsum_bf16(_ctx(a), a.data_ptr(), a.stride(0), a.shape[0], a.shape[1], out.data_ptr(),
                                     _stream(a.device)))
    return out


def cast_f32_bf16(x, out=None):
    _chk(x, torch.float32, "x")
    if out is None:
        out = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    L.check(L.load().jat_cast_f32_bf16(_ctx(x), x.data_ptr(), out.data_ptr(), x.numel(), _stream(x.device)))
    return out


def gqa_attention_bwd(qkv, d_out, out, lse, rope_cos, rope_sin, B, N, Hq, Hkv, head_dim=64, dqkv=None, drop_p=0.0,
                      drop_seed=0):
    """Gradient w.r.t. the pre-RoPE packed q|k|v projections, bf16 [B*N, (Hq+2Hkv)*64]."""
    for t, n in ((qkv, "qkv"), (d_out, "d_out"), (out, "out")):
        _chk(t, torch.bfloat16, n)
    _chk(lse, torch.float32, "lse")
    if dqkv is None:
        dqkv = torch.empty_like(qkv)
    dsum = torch.empty(B, Hq, N, dtype=torch.float32, device=qkv.device)
    dq_acc = torch.empty(B * N, Hq * head_dim, dtype=torch.float32, device=qkv.device)
    L.check(L.load().jat_gqa_attention_bwd_dropout(_ctx(qkv), qkv.data_ptr(), d_out.data_ptr(), out.data_ptr(),
                                                   lse.data_ptr(), dsum.data_ptr(), dq_acc.data_ptr(), dqkv.data_ptr(),
                                                   rope_cos.data_ptr(), rope_sin.data_ptr(), B, N, Hq, Hkv, head_dim,
                                                   float(drop_p), int(drop_seed), _stream(qkv.device)))
    return dqkv
