"""ctypes binding of ``libjat_b200.so`` (C ABI declared in ``include/jat_b200.h``).

This is the only place Python touches native code.  There is NO fallback: if the shared library is
missing or the device is not an sm_100 GPU, every entry point raises.  ``build()`` compiles the
library in-tree with nvcc for sm_100a (cross-compiles without a GPU).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.environ.get("JAT_B200_LIB") or os.path.join(_HERE, "libjat_b200.so")  # env override: kernel experiments
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "jat_b200.h")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-shared", "-Xcompiler", "-fPIC",
]

# ---- enums mirrored from include/jat_b200.h
NORM_LAYERNORM, NORM_RMSNORM = 0, 1
EPI_BIAS_ACT, EPI_QKV_ROPE, EPI_GATE_RESIDUAL, EPI_UNPATCHIFY, EPI_ACCUM, EPI_DACT = 0, 1, 2, 3, 4, 5
ACT_NONE, ACT_GELU_ERF, ACT_SILU = 0, 1, 2
DTYPE_F32, DTYPE_BF16 = 0, 1
ERR_BAD_ARG = -1
ERR_SEQ_TOO_LONG = -5
DROP_SITE_ATTN, DROP_SITE_MLP_HIDDEN, DROP_SITE_MLP_OUT, DROP_SITE_PATH = 0, 1, 2, 3


def _sources():
    return sorted(os.path.join(_CSRC, f) for f in os.listdir(_CSRC) if f.endswith((".cu", ".cuh")))


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/*.cu into libjat_b200.so (in-tree) if it is missing or stale."""
    srcs = _sources() + [HEADER_PATH]
    if not force and os.path.exists(LIB_PATH):
        if os.path.getmtime(LIB_PATH) >= max(os.path.getmtime(s) for s in srcs):
            return LIB_PATH
    nvcc = os.environ.get("NVCC") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        nvcc = "nvcc"
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + [
        "-o", LIB_PATH, os.path.join(_CSRC, "api.cu")]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return LIB_PATH


class GemmEpilogue(C.Structure):
    _fields_ = [
        ("kind", C.c_int32), ("act", C.c_int32), ("out_dtype", C.c_int32), ("tokens_per_batch", C.c_int32),
        ("bias", C.c_void_p), ("out", C.c_void_p), ("ldo", C.c_int64),
        ("gate", C.c_void_p), ("gate_batch_stride", C.c_int64),
        ("rope_cos", C.c_void_p), ("rope_sin", C.c_void_p),
        ("rope_cols", C.c_int32), ("patch_len", C.c_int32), ("t_out", C.c_int32), ("k_splits", C.c_int32),
        ("aux", C.c_void_p), ("ld_aux", C.c_int64), ("a_transposed", C.c_int32), ("w_transposed", C.c_int32),
        ("drop_p", C.c_float), ("drop_seed", C.c_uint32), ("gate_rowscale", C.c_void_p),
    ]


class DitWeights(C.Structure):
    _fields_ = [
        ("hidden", C.c_int32), ("depth", C.c_int32), ("n_q_heads", C.c_int32), ("n_kv_heads", C.c_int32),
        ("head_dim", C.c_int32), ("mlp_hidden", C.c_int32), ("bottleneck", C.c_int32), ("channels", C.c_int32),
        ("patch_len", C.c_int32), ("norm_kind", C.c_int32), ("max_len", C.c_int32), ("rope_max_pos", C.c_int32),
        ("norm_eps", C.c_float), ("cond_channels", C.c_int32),
        ("pe_w1", C.c_void_p), ("pe_b1", C.c_void_p), ("pe_w2", C.c_void_p), ("pe_b2", C.c_void_p),
        ("te_w1", C.c_void_p), ("te_b1", C.c_void_p), ("te_w2", C.c_void_p), ("te_b2", C.c_void_p),
        ("ada_w", C.c_void_p), ("ada_b", C.c_void_p),
        ("wqkv", C.POINTER(C.c_void_p)), ("wo", C.POINTER(C.c_void_p)),
        ("w1", C.POINTER(C.c_void_p)), ("b1", C.POINTER(C.c_void_p)),
        ("w2", C.POINTER(C.c_void_p)), ("b2", C.POINTER(C.c_void_p)),
        ("norm1_w", C.POINTER(C.c_void_p)), ("norm2_w", C.POINTER(C.c_void_p)),
        ("final_norm_w", C.c_void_p), ("final_w", C.c_void_p), ("final_b", C.c_void_p),
        ("rope_cos", C.c_void_p), ("rope_sin", C.c_void_p),
    ]


class DitWorkspace(C.Structure):
    _fields_ = [
        ("patches", C.c_void_p), ("pe_hid", C.c_void_p), ("x", C.c_void_p), ("h", C.c_void_p),
        ("qkv", C.c_void_p), ("attn", C.c_void_p), ("mlp_hid", C.c_void_p),
        ("t_feat", C.c_void_p), ("t_hid", C.c_void_p), ("t_act", C.c_void_p), ("mod", C.c_void_p),
        ("block_out", C.c_void_p), ("attn_part", C.c_void_p), ("lse_part", C.c_void_p),
    ]


class DitSaved(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("x_in", "x_mid", "h1", "qkv", "attn", "lse", "y1", "h2", "u", "mact", "y2",
                                          "pe_u", "t_u1", "t_u2")] + [
        ("dropout_p", C.c_float), ("reserved", C.c_int32), ("seed", C.c_uint64),
        ("drop_path_rates", C.c_void_p), ("dp_scale", C.c_void_p), ("rs1", C.c_void_p), ("rs2", C.c_void_p)]


class DitBwdScratch(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("dx", "dy", "dh", "da", "du", "dqkv", "dsum", "dq_acc", "dmod", "dmod_bf16",
                                          "dxsum", "dout_p", "dpe", "dt_a", "dt_b", "dt_acc", "rowstats")]


# name -> (restype, argtypes); every symbol include/jat_b200.h declares
_vp, _i, _i64, _f, _u32, _u64 = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_uint32, C.c_uint64
SIGNATURES = {
    "jat_dropout_site_seed": (_u32, [_u64, _i, _i]),
    "jat_dropout_scale_mask": (_i, [_vp, _vp, _i64, _i, _f, _u32, _vp]),
    "jat_drop_path_scales": (_i, [_vp, _vp, _vp, _i, _i, _u64, _vp]),
    "jat_attention_passes": (_i, [_i]),
    "jat_gqa_attention_fwd_long": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _f, _u32, _vp]),
    "jat_gqa_attention_fwd_dropout": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _f, _u32, _vp]),
    "jat_gqa_attention_bwd_dropout": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _f, _u32,
                                           _vp]),
    "jat_gate_bwd_dropout": (_i, [_vp, _vp, _vp, _vp, _i64, _vp, _vp, _i64, _vp, _vp, _i, _i, _i, _f, _u32, _vp, _vp]),
    "jat_adamw_chunk_elems": (_i, []),
    "jat_grad_sumsq": (_i, [_vp, _vp, _vp, _i, _i, _vp, _vp, _i, _vp]),
    "jat_adamw_step": (_i, [_vp, _vp, _vp, _i, _i, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, _f, _vp, _vp]),
    "jat_abi_version": (_i, []),
    "jat_last_error": (C.c_char_p, []),
    "jat_create": (_i, [_i, C.POINTER(_vp)]),
    "jat_destroy": (None, [_vp]),
    "jat_sm_count": (_i, [_vp]),
    "jat_launch_count": (_i64, [_vp]),
    "jat_set_gemm_config": (_i, [_vp, _i, _i]),
    "jat_set_gemm_tail_split": (_i, [_vp, _i]),
    "jat_set_gemm_sm_reserve": (_i, [_vp, _i]),
    "jat_debug_set_attention_trace": (_i, [_vp, _vp]),
    "jat_debug_set_gemm_trace": (_i, [_vp, _vp]),
    "jat_profile_begin": (_i, [_vp]),
    "jat_profile_end": (_i, [_vp, _i, C.POINTER(C.c_char_p), C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    "jat_adaln_norm_modulate": (_i, [_vp, _vp, _vp, _vp, _vp, _i64, _vp, _i, _f, _i, _i, _i, _vp]),
    "jat_patchify_cast": (_i, [_vp, _vp, _i, _vp, _i, _vp, _i, _i, _i, _i, _vp]),
    "jat_patchify_cast2": (_i, [_vp, _vp, _i, _vp, _i, _vp, _i, _i, _i, _i, _i, _vp]),
    "jat_timestep_features": (_i, [_vp, _vp, _vp, _i, _i, _vp]),
    "jat_gemm_bf16": (_i, [_vp, _vp, _i64, _vp, _i64, _i, _i, _i, C.POINTER(GemmEpilogue), _i, _i, _vp]),
    "jat_gqa_attention_fwd": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "jat_gqa_attention_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "jat_cfg_euler_update": (_i, [_vp, _vp, _vp, _vp, _f, _vp, _i, _i64, _vp]),
    "jat_adaln_gate_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _i64, _vp, _i, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                _i, _i, _i, _f, _u32, _vp, _vp]),
    "jat_adaln_bwd": (_i, [_vp, _vp, _vp, _vp, _i64, _vp, _i, _f, _vp, _i, _vp, _vp, _i64, _vp, _vp, _i, _i, _i, _vp]),
    "jat_gate_bwd": (_i, [_vp, _vp, _vp, _vp, _i64, _vp, _vp, _i64, _vp, _vp, _i, _i, _i, _vp]),
    "jat_colsum_bf16": (_i, [_vp, _vp, _i64, _i, _i, _vp, _vp]),
    "jat_cast_f32_bf16": (_i, [_vp, _vp, _vp, _i64, _vp]),
    "jat_grad_compress": (_i, [_vp, _vp, _vp, _i64, C.c_float, _vp]),
    "jat_grad_decompress": (_i, [_vp, _vp, _vp, _i64, _vp]),
    "jat_train_inputs": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _f, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "jat_mse_loss": (_i, [_vp, _vp, _vp, _vp, _vp, _i64, _vp]),
    "jat_charbonnier_loss": (_i, [_vp, _vp, _vp, _vp, _vp, _i64, C.c_float, _vp]),
    "jat_chunk_normalize": (_i, [_vp, _vp, _i64, _i64, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "jat_crossfade_denorm": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _vp]),
    "jat_patchify_single": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "jat_dit_forward_train": (_i, [_vp, C.POINTER(DitWeights), C.POINTER(DitWorkspace), C.POINTER(DitSaved), _vp, _vp, _vp,
                                    _vp, _i, _i, _vp]),
    "jat_dit_backward": (_i, [_vp, C.POINTER(DitWeights), C.POINTER(DitWorkspace), C.POINTER(DitSaved),
                               C.POINTER(DitBwdScratch), C.POINTER(DitWeights), _vp, _i, _i, _vp]),
    "jat_dit_backward_begin": (_i, [_vp, C.POINTER(DitWeights), C.POINTER(DitWorkspace), C.POINTER(DitSaved),
                                     C.POINTER(DitBwdScratch), C.POINTER(DitWeights), _vp, _i, _i, _vp]),
    "jat_dit_backward_block": (_i, [_vp, C.POINTER(DitWeights), C.POINTER(DitWorkspace), C.POINTER(DitSaved),
                                     C.POINTER(DitBwdScratch), C.POINTER(DitWeights), _i, _i, _i, _vp]),
    "jat_dit_backward_end": (_i, [_vp, C.POINTER(DitWeights), C.POINTER(DitWorkspace), C.POINTER(DitSaved),
                                   C.POINTER(DitBwdScratch), C.POINTER(DitWeights), _i, _i, _vp]),
    "jat_dit_modulation": (_i, [_vp, C.POINTER(DitWeights), C.POINTER(DitWorkspace), _vp, _i, _vp]),
    "jat_dit_forward_tokens": (_i, [_vp, C.POINTER(DitWeights), C.POINTER(DitWorkspace), _vp, _i, _vp, _i, _vp,
                                     _i64, _vp, _i, _i, _vp]),
}

_lib = None
_lock = threading.Lock()
_ctx = {}


def load() -> C.CDLL:
    """dlopen the C-ABI library and bind every declared symbol.  Never builds implicitly."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise RuntimeError(
                    f"{LIB_PATH} is missing: the jat_b200 CUDA extension has not been built. "
                    "Run `python -c 'import __graft_entry__ as g; g.build()'` (needs nvcc). There is no CPU fallback.")
            lib = C.CDLL(LIB_PATH)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
                fn.restype = res
                fn.argtypes = args
            if lib.jat_abi_version() != 1:
                raise RuntimeError("libjat_b200.so ABI version mismatch; rebuild")
            _lib = lib
    return _lib


class JatError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"jat_b200 error {code}: {msg}")
        self.code = code


def check(code: int) -> None:
    if code != 0:
        raise JatError(code, load().jat_last_error().decode(errors="replace"))


def context(device_index: int) -> int:
    """Opaque jat_ctx* for a CUDA device (created once per process and device)."""
    lib = load()
    with _lock:
        h = _ctx.get(device_index)
        if h is None:
            out = C.c_void_p()
            code = lib.jat_create(device_index, C.byref(out))
            if code != 0:
                raise JatError(code, lib.jat_last_error().decode(errors="replace"))
            h = out.value
            _ctx[device_index] = h
    return h


def profile_begin(device_index: int) -> None:
    check(load().jat_profile_begin(context(device_index)))


def profile_end(device_index: int) -> dict:
    """{kernel class: (total_ms, launches)} since profile_begin (synchronises the device)."""
    n = 32
    names = (C.c_char_p * n)()
    ms = (C.c_double * n)()
    cnt = (C.c_int64 * n)()
    got = load().jat_profile_end(context(device_index), n, names, ms, cnt)
    if got < 0:
        check(got)
    return {names[i].decode(): (ms[i], cnt[i]) for i in range(got) if cnt[i] > 0}
