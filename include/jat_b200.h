/*
 * jat_b200.h -- C ABI of libjat_b200.so: hand-written sm_100a kernels for the JaT-AudioSR DiT
 * denoiser hot path (DiT forward inside the flow-matching Euler/CFG sampler).
 *
 * The reference (HUSRCF/JaTSR-Just-audio-transformer-super-solution) is pure PyTorch and has no
 * FFI / plugin layer of its own; its boundary for this path is the nn.Module
 * `JaT_AudioSR_V2/V3.forward(x_t, t, x_cond)` (src/models/jat_audiosr_v2.py:399-448,
 * src/models/jat_audiosr_v3.py:422-471) plus the free function `flow_matching_sample`
 * (infer_test_v3m2.py:108-185). The Python mirror of that surface lives in
 * `jatsr-just-audio-transformer-super-solution_b200/{models,sampler}.py`; every device operation it
 * performs goes through the entry points below (ctypes binding in `_lib.py`).
 *
 * Conventions
 *   - plain C types only; every pointer is a DEVICE pointer unless the name ends in `_host`;
 *   - `stream` is a `cudaStream_t` passed as `void*`; nothing here synchronises the device or
 *     allocates device memory; all buffers are owned by the caller;
 *   - return value: 0 = ok; > 0 = a `cudaError_t`; < 0 = JAT_ERR_* (argument / shape errors);
 *     `jat_last_error()` returns a human-readable message for the calling thread;
 *   - row-major everywhere; "bf16" = __nv_bfloat16 bits (uint16_t), "f32" = float;
 *   - token rows: m = b * tokens_per_batch + n  (b = batch item, n = token index = RoPE position).
 */
#ifndef JAT_B200_H
#define JAT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define JAT_ABI_VERSION 1

#define JAT_ERR_BAD_ARG (-1)      /* null pointer, negative size, unsupported enum */
#define JAT_ERR_BAD_SHAPE (-2)    /* shape not supported by the kernels (see each function) */
#define JAT_ERR_NO_DRIVER (-3)    /* cuTensorMapEncodeTiled could not be resolved */
#define JAT_ERR_TENSORMAP (-4)    /* cuTensorMapEncodeTiled failed */
#define JAT_ERR_SEQ_TOO_LONG (-5) /* token count > max_len (reference raises ValueError, jat_audiosr_v2.py:428) */

typedef struct jat_ctx jat_ctx; /* opaque: device id, SM count, TMA descriptor cache */

int jat_abi_version(void);
const char* jat_last_error(void);
/* Binds to CUDA device `device`, resolves the driver entry points and allocates the library's only device memory: a
 * fixed ~19 MB scratch for the GEMM tail split (below).  Calls on one ctx must be issued to one stream at a time. */
int jat_create(int device, jat_ctx** out);
void jat_destroy(jat_ctx* ctx);
int jat_sm_count(const jat_ctx* ctx);

/* ----------------------------------------------------------------------------------------------
 * Normalisation kinds (jat_audiosr_v2.py:242,245,361 vs jat_audiosr_v3.py:261,264,384)
 * -------------------------------------------------------------------------------------------- */
#define JAT_NORM_LAYERNORM 0 /* nn.LayerNorm(D, elementwise_affine=False, eps), biased variance */
#define JAT_NORM_RMSNORM 1   /* nn.RMSNorm(D, eps) with learnable weight[D] */

/* Fused AdaLN: out[m,:] = norm(x[m,:]) * (1 + scale[b,:]) + shift[b,:], b = m / tokens_per_batch.
 * Replaces norm1/norm2 + modulate (jat_audiosr_v2.py:278-279,284-285) and, with shift == scale ==
 * NULL, the un-modulated final norm (jat_audiosr_v2.py:361).
 *   x      f32 [M, D]      out  bf16 [M, D]      weight f32 [D] (RMSNorm only, else NULL)
 *   shift/scale f32, element (b, d) at  ptr[b * mod_batch_stride + d]  (stride 0 = one t for all b)
 * D % 4 == 0, D <= 4096. */
int jat_adaln_norm_modulate(jat_ctx* ctx, const float* x, void* out_bf16, const float* shift, const float* scale,
                            int64_t mod_batch_stride, const float* weight, int norm_kind, float eps, int M, int D,
                            int tokens_per_batch, void* stream);

/* Patchify + concat + cast: builds the bf16 A operand of patch_embed.proj.0
 * (jat_audiosr_v2.py:411-421 pad+cat, :225-227 reshape/permute).
 *   x_t, x_cond f32 [B, C, T]  ->  out bf16 [B * N, 2*C*P],  N = ceil(T / P)
 *   out[b*N + n, c*P + p]       = x_t  [b, c, n*P + p]   (0 when n*P+p >= T)
 *   out[b*N + n, (C + c)*P + p] = x_cond[b, c, n*P + p]
 * x_cond == NULL means an all-zero condition (the CFG unconditional half, infer_test_v3m2.py:142).
 * cond_batch: number of batch items present in x_cond; items b >= cond_batch read zeros
 * (lets the sampler pass [z ; z] x [cond ; 0] without materialising the cats of :154-156).
 * xt_batch: number of batch items present in x_t; item b reads x_t[b % xt_batch].
 * P must be 4. */
int jat_patchify_cast(jat_ctx* ctx, const float* x_t, int xt_batch, const float* x_cond, int cond_batch,
                      void* out_bf16, int B, int C, int T, int P, void* stream);
/* The same with a condition latent of its own channel count (reference ctor `cond_channels` != `input_channels`):
 * x_t [xt_batch, C, T], x_cond [cond_batch, Cc, T] -> out bf16 [B*N, (C + Cc) * P]; C and Cc multiples of 32. */
int jat_patchify_cast2(jat_ctx* ctx, const float* x_t, int xt_batch, const float* x_cond, int cond_batch, void* out_bf16, int B, int C,
                       int Cc, int T, int P, void* stream);

/* Sinusoidal timestep features (TimeEmbedding.forward, jat_audiosr_v2.py:177-190), bf16 output
 * (the A operand of t_embedder.1):  out[b, i] = sin(t[b] * f_i), out[b, half + i] = cos(t[b] * f_i),
 * f_i = exp(-i * ln(1e4) / (half - 1)), half = D / 2.   t f32 [B] -> out bf16 [B, D]. */
int jat_timestep_features(jat_ctx* ctx, const float* t, void* out_bf16, int B, int D, void* stream);

/* ----------------------------------------------------------------------------------------------
 * tcgen05 / TMEM GEMM fed by TMA:  acc[M, N] = A[M, K] (bf16, row pitch lda) * W[N, K]^T (bf16,
 * nn.Linear layout, row pitch ldw), fp32 accumulation in tensor memory, fused epilogue.
 * Requirements: N % 128 == 0, lda/ldw % 8 == 0, 16-byte aligned pointers; K % 64 == 0 unless both operands are
 * given transposed (then the TMA unit zero-fills the reduction rows past K).
 * -------------------------------------------------------------------------------------------- */
#define JAT_EPI_BIAS_ACT 0      /* out = act(acc + bias)                  (Linear [+GELU|SiLU])      */
#define JAT_EPI_QKV_ROPE 1      /* out = RoPE(acc) on columns < rope_cols (q_proj/k_proj/v_proj+RoPE) */
#define JAT_EPI_GATE_RESIDUAL 2 /* out(f32, in place) += gate[b,:] * (acc + bias)   (adaLN-Zero gate) */
#define JAT_EPI_UNPATCHIFY 3    /* out[b, c, n*P+p] = acc[m, c*P+p] + bias  (final Linear+unpatchify) */
#define JAT_EPI_ACCUM 4         /* out(f32, in place) += acc (+ bias)      (weight gradients, split-K partial sums) */
#define JAT_EPI_DACT 5          /* out(bf16) = (acc + bias) * act'(aux)    (dgrad through GELU / SiLU; aux = pre-activation) */

#define JAT_ACT_NONE 0
#define JAT_ACT_GELU_ERF 1 /* nn.GELU() default (exact erf form), jat_audiosr_v2.py:206,249 */
#define JAT_ACT_SILU 2     /* nn.SiLU(), jat_audiosr_v2.py:344,257 */

#define JAT_DTYPE_F32 0
#define JAT_DTYPE_BF16 1

typedef struct jat_gemm_epilogue {
    int32_t kind;      /* JAT_EPI_* */
    int32_t act;       /* JAT_ACT_*   (BIAS_ACT only) */
    int32_t out_dtype; /* JAT_DTYPE_* (BIAS_ACT only; QKV_ROPE is bf16; GATE_RESIDUAL/UNPATCHIFY f32) */
    int32_t tokens_per_batch; /* N tokens per batch item: b = m / N, RoPE position = m % N */
    const float* bias;        /* f32 [N] or NULL */
    void* out;                /* [M, ldo] (UNPATCHIFY: f32 [B, C, T_out]) */
    int64_t ldo;              /* output row pitch in elements */
    const float* gate;        /* GATE_RESIDUAL: gate[b * gate_batch_stride + n] */
    int64_t gate_batch_stride;
    const float* rope_cos; /* QKV_ROPE: cos_cached / sin_cached f32 [max_pos, 64] (jat_audiosr_v2.py:60-68) */
    const float* rope_sin;
    int32_t rope_cols; /* QKV_ROPE: columns [0, rope_cols) are 64-wide heads to rotate (Q and K) */
    int32_t patch_len; /* UNPATCHIFY: P (must be 4) */
    int32_t t_out;     /* UNPATCHIFY: cropped length T (<= tokens_per_batch * P) */
    int32_t k_splits;  /* GATE_RESIDUAL / ACCUM: split the reduction over this many work items per tile (0 or 1 = off) */
    void* aux;         /* DACT: pre-activation u bf16 [M, ld_aux] (read).  BIAS_ACT/bf16: if non-NULL, a bf16 copy of
                          acc + bias (the pre-activation) is written here for the backward pass */
    int64_t ld_aux;
    int32_t a_transposed; /* 1: A is given as A^T [K, M] row-major with pitch lda (reduction index = row) */
    int32_t w_transposed; /* 1: W is given as W^T [K, N] row-major with pitch ldw.  Backward GEMMs without copies:
                             dgrad dX = dY W      -> A = dY, W = weight with w_transposed = 1;
                             wgrad dW = dY^T X    -> A = dY with a_transposed = 1, W = X with w_transposed = 1 */
    float drop_p;         /* train-mode nn.Dropout(p) fused into the epilogue (0 = off), element (m, n) of site drop_seed:
                             BIAS_ACT/bf16: out = drop(act(acc + bias));  GATE_RESIDUAL: x += gate * drop(acc + bias) (aux
                             keeps the dropped value);  DACT: out = (acc + bias) * mask * act'(aux)  (same mask as the
                             forward BIAS_ACT call with the same seed) */
    uint32_t drop_seed;   /* jat_dropout_site_seed(...) */
    const float* gate_rowscale; /* GATE_RESIDUAL: optional f32 [B] factor on the gate per batch item (DropPath) or NULL */
} jat_gemm_epilogue;

/* cta_pair: 0 = one CTA per 128-row tile (tcgen05 cta_group::1), 1 = CTA pair per 256-row tile
 * (cta_group::2, halves the per-SM B-operand traffic). block_n: 128 or 256 (0 = library picks). */
int jat_gemm_bf16(jat_ctx* ctx, const void* A, int64_t lda, const void* W, int64_t ldw, int M, int N, int K,
                  const jat_gemm_epilogue* epi, int cta_pair, int block_n, void* stream);

/* ----------------------------------------------------------------------------------------------
 * GQA attention (GroupedQueryAttention.forward, jat_audiosr_v2.py:141-164, eval mode):
 *   out[b, n, h*64 : h*64+64] = softmax(Q_h K_g^T / sqrt(64)) V_g,  g = h / (Hq / Hkv)
 * on the packed projection buffer written by the QKV_ROPE GEMM:
 *   qkv bf16 [B * N, (Hq + 2*Hkv) * 64] = [ Q heads | K heads | V heads ],   out bf16 [B * N, Hq*64].
 * No mask, bidirectional. K/V of one KV head are staged in shared memory once per CTA and reused
 * by the Hq/Hkv query heads of the group. head_dim must be 64.  One launch covers up to 352 keys (the S row of a query tile
 * lives in one TMEM accumulator, so the softmax is exact); longer sequences -- the reference allows up to 2048 tokens,
 * jat_audiosr_v2.py:428 -- go through jat_gqa_attention_fwd_long: one launch per 352-key chunk into caller scratch, then a
 * merge by the chunks' log-sum-exps (jat_attention_passes(N) chunks).
 * lse_or_null: optional f32 [B, Hq, N] output, log2(sum_j 2^(s_ij * log2(e) / 8)) per query row, kept by the training
 * forward for jat_gqa_attention_bwd.
 * -------------------------------------------------------------------------------------------- */
int jat_gqa_attention_fwd(jat_ctx* ctx, const void* qkv_bf16, void* out_bf16, float* lse_or_null, int B, int N, int Hq,
                          int Hkv, int head_dim, void* stream);
int jat_attention_passes(int N); /* ceil(N / 352) */
/* Any N: part_o bf16 [passes, B*N, Hq*64] and part_lse f32 [passes, B, Hq, N] are scratch (may be NULL when N <= 352);
 * drop_p / drop_seed as in jat_gqa_attention_fwd_dropout (mask column = global key index). */
int jat_gqa_attention_fwd_long(jat_ctx* ctx, const void* qkv_bf16, void* out_bf16, float* lse_or_null, void* part_o_bf16,
                               float* part_lse, int B, int N, int Hq, int Hkv, int head_dim, float drop_p, uint32_t drop_seed,
                               void* stream);

/* Backward of jat_gqa_attention_fwd.  d_out / out bf16 [B*N, Hq*64] (gradient of, and the saved, forward
 * output), lse from the forward call.  Writes dqkv bf16 [B*N, (Hq+2Hkv)*64] = gradient w.r.t. the PRE-RoPE q | k | v
 * projections (the RoPE rotation of the QKV epilogue is transposed inside).  Scratch (caller-owned):
 * dsum_scratch f32 [B, Hq, N], dq_acc_scratch f32 [B*N, Hq*64].  rope_cos/sin: the forward's tables. */
int jat_gqa_attention_bwd(jat_ctx* ctx, const void* qkv_bf16, const void* d_out_bf16, const void* out_bf16, const float* lse,
                          float* dsum_scratch, float* dq_acc_scratch, void* dqkv_bf16, const float* rope_cos,
                          const float* rope_sin, int B, int N, int Hq, int Hkv, int head_dim, void* stream);

/* ----------------------------------------------------------------------------------------------
 * Train-mode stochastic regularisers.  The reference draws nn.Dropout(p) masks on the attention probabilities
 * (jat_audiosr_v2.py:158), after the MLP's GELU (:250) and after mlp.3 (:252), and a per-sample DropPath factor on
 * each gated branch (:21-34, :281, :287) from torch's global Philox stream.  Here every mask element is a pure function
 * of (site seed, row, col): one 32-bit hash of (row, col >> 1, site_seed) gives the two 16-bit lanes of columns col & ~1
 * and col | 1; KEEP iff lane >= round(p * 2^16), kept values scaled by 1 / (1 - round(p 2^16) / 2^16) -- so the backward
 * kernels regenerate the forward's masks instead of storing them, with Bernoulli(1-p) statistics (p exact to 2^-17).
 * Site seeds: jat_dropout_site_seed(seed, block, site).  Mask coordinates: attention (row = (b*Hq + h)*N + query,
 * col = key); MLP sites (row = token row m, col = feature).
 * -------------------------------------------------------------------------------------------- */
#define JAT_DROP_SITE_ATTN 0
#define JAT_DROP_SITE_MLP_HIDDEN 1
#define JAT_DROP_SITE_MLP_OUT 2
#define JAT_DROP_SITE_PATH 3
uint32_t jat_dropout_site_seed(uint64_t seed, int block, int site);
/* out f32 [rows, cols] = the multiplier (0 or 1/(1-p)) the fused kernels apply at (row, col) of a site (parity tests). */
int jat_dropout_scale_mask(jat_ctx* ctx, float* out, int64_t rows, int cols, float p, uint32_t site_seed, void* stream);
/* DropPath factors: out f32 [depth, 2, B], out[i][branch][b] = floor(keep_i + U) / keep_i with keep_i = 1 - rates[i]
 * (rates: DEVICE f32 [depth]; 1 where rates[i] == 0); branch 0 = attention, 1 = MLP. */
int jat_drop_path_scales(jat_ctx* ctx, float* out, const float* rates, int depth, int B, uint64_t seed, void* stream);
int jat_gqa_attention_fwd_dropout(jat_ctx* ctx, const void* qkv_bf16, void* out_bf16, float* lse_or_null, int B, int N,
                                  int Hq, int Hkv, int head_dim, float drop_p, uint32_t drop_seed, void* stream);
int jat_gqa_attention_bwd_dropout(jat_ctx* ctx, const void* qkv_bf16, const void* d_out_bf16, const void* out_bf16,
                                  const float* lse, float* dsum_scratch, float* dq_acc_scratch, void* dqkv_bf16,
                                  const float* rope_cos, const float* rope_sin, int B, int N, int Hq, int Hkv, int head_dim,
                                  float drop_p, uint32_t drop_seed, void* stream);

/* ----------------------------------------------------------------------------------------------
 * Fused sampler update (infer_test_v3m2.py:161-179): CFG combine + x-prediction -> velocity + Euler.
 *   x   = x_u + cfg_scale * (x_c - x_u)            (x = x_c when x_u == NULL, i.e. cfg_scale == 1)
 *   z  <- z + (x - z) / (1 - t + 1e-5) * dt         if t < 0.999, else z <- x
 * t and dt are read from DEVICE memory (t_dt[2*step], t_dt[2*step+1]) so a captured CUDA graph can be
 * replayed for every step; all tensors f32 with `numel` elements, updated in place on z.
 * -------------------------------------------------------------------------------------------- */
int jat_cfg_euler_update(jat_ctx* ctx, float* z, const float* x_c, const float* x_u, float cfg_scale,
                         const float* t_dt, int step, int64_t numel, void* stream);

/* ----------------------------------------------------------------------------------------------
 * Whole DiT forward (JaT_AudioSR_V2/V3.forward, eval mode) as one host call that enqueues every
 * kernel on `stream`. All weights are the caller's packed bf16 copies (see engine.py: pack order).
 * -------------------------------------------------------------------------------------------- */
typedef struct jat_dit_weights {
    int32_t hidden, depth, n_q_heads, n_kv_heads, head_dim, mlp_hidden, bottleneck, channels, patch_len,
        norm_kind, max_len, rope_max_pos;
    float norm_eps;
    int32_t cond_channels; /* channels of x_cond (reference ctor `cond_channels`); 0 = the same as `channels` */
    const void* pe_w1;   /* bf16 [bottleneck, (C+Cc)*P]   patch_embed.proj.0.weight */
    const float* pe_b1;  /* f32  [bottleneck] */
    const void* pe_w2;   /* bf16 [hidden, bottleneck]  patch_embed.proj.2.weight */
    const float* pe_b2;  /* f32  [hidden] */
    const void* te_w1;   /* bf16 [hidden, hidden]      t_embedder.1 */
    const float* te_b1;
    const void* te_w2;   /* bf16 [hidden, hidden]      t_embedder.3 */
    const float* te_b2;
    const void* ada_w;   /* bf16 [depth * 6 * hidden, hidden]  blocks.i.adaLN_modulation.1.weight stacked */
    const float* ada_b;  /* f32  [depth * 6 * hidden] */
    const void* const* wqkv; /* host array [depth] of bf16 [(Hq+2Hkv)*64, hidden]  (q_proj|k_proj|v_proj rows) */
    const void* const* wo;   /* host array [depth] of bf16 [hidden, hidden]        out_proj */
    const void* const* w1;   /* host array [depth] of bf16 [mlp_hidden, hidden]    mlp.0 */
    const float* const* b1;  /* host array [depth] of f32  [mlp_hidden] */
    const void* const* w2;   /* host array [depth] of bf16 [hidden, mlp_hidden]    mlp.3 */
    const float* const* b2;  /* host array [depth] of f32  [hidden] */
    const float* const* norm1_w; /* host array [depth] of f32 [hidden] (RMSNorm) or NULL */
    const float* const* norm2_w;
    const float* final_norm_w; /* f32 [hidden] (RMSNorm) or NULL */
    const void* final_w;       /* bf16 [C*P, hidden]  final_layer.1.weight */
    const float* final_b;      /* f32  [C*P] */
    const float* rope_cos;     /* f32 [rope_max_pos, 64] */
    const float* rope_sin;
} jat_dit_weights;

typedef struct jat_dit_workspace {
    /* all device buffers, sized for M = B * N token rows (N = ceil(T / P)) */
    void* patches;  /* bf16 [M, 2*C*P] */
    void* pe_hid;   /* bf16 [M, bottleneck] */
    float* x;       /* f32  [M, hidden]       residual stream */
    void* h;        /* bf16 [M, hidden]       normalised+modulated GEMM operand */
    void* qkv;      /* bf16 [M, (Hq+2Hkv)*64] */
    void* attn;     /* bf16 [M, hidden] */
    void* mlp_hid;  /* bf16 [M, mlp_hidden] */
    void* t_feat;   /* bf16 [Bt, hidden]      sinusoid features */
    void* t_hid;    /* bf16 [Bt, hidden] */
    void* t_act;    /* bf16 [Bt, hidden]      SiLU(t_emb) */
    float* mod;     /* f32  [Bt, depth*6*hidden]  all blocks' shift/scale/gate */
    float* block_out; /* optional f32 [depth, M, hidden] per-block residual snapshots for parity tests, or NULL */
    void* attn_part;  /* N > 352 tokens only: bf16 [jat_attention_passes(N), M, hidden] per-key-chunk attention outputs */
    float* lse_part;  /* N > 352 tokens only: f32 [jat_attention_passes(N), B, Hq, N] */
} jat_dit_workspace;

/* Timestep path only: t f32 [Bt] -> ws->mod [Bt, depth*6*hidden] (t_embedder + every block's
 * adaLN_modulation, jat_audiosr_v2.py:341-346,256-259,274-275). */
int jat_dit_modulation(jat_ctx* ctx, const jat_dit_weights* w, const jat_dit_workspace* ws, const float* t, int Bt,
                       void* stream);

/* Token path: x_t/x_cond f32 -> out f32 [B, C, T]; modulation rows taken from `mod`
 * (row b * mod_batch_stride; stride 0 = shared t). Semantics of xt_batch / cond_batch as in
 * jat_patchify_cast. */
int jat_dit_forward_tokens(jat_ctx* ctx, const jat_dit_weights* w, const jat_dit_workspace* ws, const float* x_t,
                           int xt_batch, const float* x_cond, int cond_batch, const float* mod,
                           int64_t mod_batch_stride, float* out, int B, int T, void* stream);

/* ----------------------------------------------------------------------------------------------
 * Training step (train_ddp_v3mod2.py:886 forward, :922 backward).  `jat_dit_forward_train` = JaT_AudioSR_V2/V3.forward with
 * per-sample t, keeping in `saved` what the backward needs; `jat_dit_backward` = autograd of that forward: given
 * d_out = dL/d(out) f32 [B, C, T] it ACCUMULATES (+=) f32 parameter gradients into `grads`, a jat_dit_weights whose
 * pointers address f32 buffers with exactly the shapes of the (bf16 / f32) weights they mirror.  No gradient is
 * produced for x_t / x_cond / t (the reference never differentiates them).
 * -------------------------------------------------------------------------------------------- */
typedef struct jat_dit_saved {
    /* per-block slabs: [depth, ...] */
    float* x_in;  /* f32  [depth, M, hidden]  residual stream entering the block (norm1 input) */
    float* x_mid; /* f32  [depth, M, hidden]  after the attention branch (norm2 input) */
    void* h1;     /* bf16 [depth, M, hidden] */
    void* qkv;    /* bf16 [depth, M, (Hq+2Hkv)*64] */
    void* attn;   /* bf16 [depth, M, hidden] */
    float* lse;   /* f32  [depth, B, Hq, N] */
    void* y1;     /* bf16 [depth, M, hidden]  out_proj output (before the gate) */
    void* h2;     /* bf16 [depth, M, hidden] */
    void* u;      /* bf16 [depth, M, mlp_hidden]  mlp.0 pre-activation */
    void* mact;   /* bf16 [depth, M, mlp_hidden]  gelu(u) */
    void* y2;     /* bf16 [depth, M, hidden]  mlp.3 output (before the gate) */
    void* pe_u;   /* bf16 [M, bottleneck]  patch_embed.proj.0 pre-activation */
    void* t_u1;   /* bf16 [B, hidden]  t_embedder.1 pre-activation */
    void* t_u2;   /* bf16 [B, hidden]  t_emb (input of the adaLN SiLU) */
    /* train-mode regularisers of this step (the backward regenerates the forward's masks from them) */
    float dropout_p;               /* nn.Dropout p of the attention / MLP sites, 0 = off */
    int32_t reserved;
    uint64_t seed;                 /* step seed: every (block, site) mask derives from it */
    const float* drop_path_rates;  /* DEVICE f32 [depth] DropPath.drop_prob per block, or NULL = no DropPath */
    float* dp_scale;               /* f32 [depth, 2, B] per-sample DropPath factors (written by the forward) */
    /* row statistics (mean, rstd) of norm1 / norm2 per block, f32 [depth, M, 2], written by the forward's norm kernels;
     * with both given, the block backward runs the fused norm + gate backward kernel (jat_adaln_gate_bwd); NULL = the
     * backward recomputes them (jat_adaln_bwd + jat_gate_bwd) */
    float* rs1;
    float* rs2;
} jat_dit_saved;

typedef struct jat_dit_bwd_scratch {
    float* dx;       /* f32  [M, hidden]  residual-stream gradient */
    void* dy;        /* bf16 [M, hidden] */
    void* dh;        /* bf16 [M, hidden] */
    void* da;        /* bf16 [M, hidden] */
    void* du;        /* bf16 [M, mlp_hidden] */
    void* dqkv;      /* bf16 [M, (Hq+2Hkv)*64] */
    float* dsum;     /* f32  [B, Hq, N] */
    float* dq_acc;   /* f32  [M, Hq*64] */
    float* dmod;     /* f32  [depth, B, 6*hidden]  (block-major) */
    void* dmod_bf16; /* bf16 [depth, B, 6*hidden] */
    float* dxsum;    /* f32  [B, hidden] */
    void* dout_p;    /* bf16 [M, C*P] */
    void* dpe;       /* bf16 [M, bottleneck] */
    void* dt_a;      /* bf16 [B, hidden] */
    void* dt_b;      /* bf16 [B, hidden] */
    float* dt_acc;   /* f32  [B, hidden]  gradient w.r.t. silu(t_emb), summed over the blocks */
    float* rowstats; /* f32  [M, 2]  row mean / rstd scratch of jat_adaln_bwd */
} jat_dit_bwd_scratch;

int jat_dit_forward_train(jat_ctx* ctx, const jat_dit_weights* w, const jat_dit_workspace* ws, const jat_dit_saved* saved,
                          const float* x_t, const float* x_cond, const float* t, float* out, int B, int T, void* stream);
int jat_dit_backward(jat_ctx* ctx, const jat_dit_weights* w, const jat_dit_workspace* ws, const jat_dit_saved* saved,
                     const jat_dit_bwd_scratch* scratch, const jat_dit_weights* grads, const float* d_out, int B, int T,
                     void* stream);
/* The same backward pass in stages, so that the caller can start the gradient all-reduce of a stage (DDP buckets,
 * train_ddp_v3mod2.py:822) while the next stage runs.  Order: begin, block depth-1, ..., block 0, end.
 *   begin: final_layer.{0,1} gradients;   block i: every parameter of blocks.i (incl. its adaLN_modulation);
 *   end:   patch_embed.* and t_embedder.* gradients. */
int jat_dit_backward_begin(jat_ctx* ctx, const jat_dit_weights* w, const jat_dit_workspace* ws, const jat_dit_saved* saved,
                           const jat_dit_bwd_scratch* scratch, const jat_dit_weights* grads, const float* d_out, int B, int T,
                           void* stream);
int jat_dit_backward_block(jat_ctx* ctx, const jat_dit_weights* w, const jat_dit_workspace* ws, const jat_dit_saved* saved,
                           const jat_dit_bwd_scratch* scratch, const jat_dit_weights* grads, int block, int B, int T,
                           void* stream);
int jat_dit_backward_end(jat_ctx* ctx, const jat_dit_weights* w, const jat_dit_workspace* ws, const jat_dit_saved* saved,
                         const jat_dit_bwd_scratch* scratch, const jat_dit_weights* grads, int B, int T, void* stream);

/* Single-source patchify + cast (the transpose of the un-patchify epilogue, used on d_out):
 * out bf16 [B*N, C*P], out[b*N + n, c*P + p] = x[b, c, n*P + p] (0 past T). */
int jat_patchify_single(jat_ctx* ctx, const float* x, void* out_bf16, int B, int C, int T, int P, void* stream);

/* ---- parameter update of the training step (replaces torch.nn.utils.clip_grad_norm_ + torch.optim.AdamW.step of
 * train_ddp_v3mod2.py:926-928 and the bf16 re-pack of the updated weights; same arithmetic as ATen's fused AdamW) ----
 * One entry per parameter tensor, all f32 and contiguous; the table lives in DEVICE memory.  `packed` (optional) receives
 * the updated values as well, in packed_dtype (JAT_DTYPE_BF16 / JAT_DTYPE_F32): the copy the forward GEMMs read.
 * packed_dtype bit 1 (JAT_ADAMW_GRAD_BF16 = 2) marks `grad` as pointing at bf16 values (the all-reduced payload of the bf16
 * gradient exchange): the update then consumes the payload directly and the f32 bucket is never re-expanded.
 * vec_ok = 1 when every pointer is 16-byte aligned (8-byte for a bf16 grad) and numel % 4 == 0 (128-bit accesses), else 0.
 * chunk_first [n_tensors] (device, int32): index of each tensor's first chunk, a chunk being jat_adamw_chunk_elems()
 * consecutive elements (the last chunk of a tensor may be short); total_chunks = their sum. */
#define JAT_ADAMW_GRAD_BF16 2
typedef struct jat_adamw_tensor {
    float* param;
    const float* grad;
    float* exp_avg;
    float* exp_avg_sq;
    void* packed;
    int64_t numel;
    int32_t packed_dtype;
    int32_t vec_ok;
    float bias_corr1;      /* 1 - beta1^step of this tensor (step counts from 1, after the increment of this update) */
    float bias_corr2_sqrt; /* sqrt(1 - beta2^step) */
} jat_adamw_tensor;
int jat_adamw_chunk_elems(void);
/* sumsq_dev[0] (+)= sum over all table entries of grad^2 (f32 per-chunk partials in partials_dev [total_chunks], summed in
 * chunk order in f64: bit-reproducible).  accumulate != 0 adds to the existing value (several parameter groups). */
int jat_grad_sumsq(jat_ctx* ctx, const jat_adamw_tensor* table_dev, const int32_t* chunk_first_dev, int n_tensors,
                   int total_chunks, float* partials_dev, double* sumsq_dev, int accumulate, void* stream);
/* AdamW (decoupled weight decay, no amsgrad) on every table entry, bias corrections per entry.  max_norm > 0: the gradients are
 * scaled by min(1, max_norm / (sqrt(*sumsq_dev) + 1e-6)) on the fly (clip_grad_norm_; `grad` itself is not modified). */
int jat_adamw_step(jat_ctx* ctx, const jat_adamw_tensor* table_dev, const int32_t* chunk_first_dev, int n_tensors,
                   int total_chunks, double lr, double beta1, double beta2, double eps, double weight_decay, float max_norm,
                   const double* sumsq_dev, void* stream);

/* Number of kernels the library has launched on this context since creation (bench `gpu_launches`). */
int64_t jat_launch_count(const jat_ctx* ctx);

/* Leave `reserve` SMs out of the persistent GEMM grids (default 0; env JAT_SM_RESERVE).  For DDP training: the GEMMs are
 * persistent one-CTA-per-SM kernels with a static tile schedule, so an NCCL all-reduce kernel that holds a few SMs would
 * make every GEMM launched meanwhile wait for it; with a reserve the two run side by side. */
int jat_set_gemm_sm_reserve(jat_ctx* ctx, int reserve);
/* Default GEMM tile configuration used when a call passes cta_pair < 0 / block_n == 0. */
int jat_set_gemm_config(jat_ctx* ctx, int cta_pair, int block_n);
/* Tail split (default off): when the persistent tile schedule ends in a partial wave, the tiles of that wave are cut
 * along K into parts that fill the machine; the parts park their f32 accumulators in the ctx scratch and the part that
 * arrives last sums them IN PART ORDER and runs the fused epilogue, so results stay bit-reproducible run to run.
 * enable == 2: only the reduce-add epilogues (GATE_RESIDUAL / ACCUM without a pre-gate copy) are split and every part
 * reduce-adds its partial sum straight into the output (bias with part 0): no scratch round trip, but the f32 adds of
 * one tile's parts land in arrival order, so those tiles may differ in the last bit run to run. */
int jat_set_gemm_tail_split(jat_ctx* ctx, int enable);

/* Per-launch timing with CUDA events on the launching stream (used by bench.py for the roofline of
 * each kernel class inside a real step).  Between begin and end every launch is bracketed by an event
 * pair; `jat_profile_end` synchronises the device and returns, per kernel class, the summed duration
 * and the launch count.  Not for use during CUDA-graph capture. Returns the number of classes filled. */
int jat_profile_begin(jat_ctx* ctx);
int jat_profile_end(jat_ctx* ctx, int max_tags, const char** names, double* total_ms, int64_t* counts);

/* Fused form of jat_adaln_bwd (accumulate = 1, modulated) and the jat_gate_bwd[_dropout] that follows it in the block
 * backward, for callers that kept the row statistics of the forward norm: rowstats f32 [M, 2] = (mean, rstd) per token row
 * (what jat_dit_forward_train writes to jat_dit_saved.rs1 / rs2).  One pass over dh, x and dx: dx += norm backward,
 * dshift / dscale (/ dweight) += column sums; and, when y_bf16 != NULL, on the updated dx row: dy = dropout-mask(dx) *
 * gate_b * rowscale_b (bf16), dgate_b += rowscale_b * sum_n dx * y, dbias[:] += rowscale_b gate_b sum_n mask(dx) (optional; dxsum_scratch is unused, may be NULL).
 * scale and gate share mod_batch_stride; dshift / dscale / dgate share dmod_batch_stride. */
int jat_adaln_gate_bwd(jat_ctx* ctx, const void* dh_bf16, const float* x, const float* rowstats, const float* scale,
                       int64_t mod_batch_stride, const float* weight, int norm_kind, float* dx, float* dshift, float* dscale,
                       int64_t dmod_batch_stride, float* dweight, const void* y_bf16, const float* gate, void* dy_bf16,
                       float* dgate, float* dxsum_scratch, float* dbias, int B, int tokens_per_batch, int D, float drop_p,
                       uint32_t drop_seed, const float* gate_rowscale, void* stream);

/* ----------------------------------------------------------------------------------------------
 * Backward pass of the DiT block (training step: train_ddp_v3mod2.py:886-922 differentiates
 * jat_audiosr_v2.py:265-289 with autograd).  The weight / input gradients of every Linear are jat_gemm_bf16 calls
 * with transposed operands (dgrad: w_transposed; wgrad: a_transposed + w_transposed + JAT_EPI_ACCUM); the
 * functions below are the HBM-bound pieces in between.  All gradient outputs are ACCUMULATED (+=) except dy/dx-
 * overwrite modes stated below; f32 unless the name says bf16.
 *
 * jat_adaln_bwd: backward of jat_adaln_norm_modulate.  dh bf16 [M, D] (M = B * tokens_per_batch), x the saved f32 input;
 *   dx (+)= d norm/dx (accumulate != 0 adds to the residual-stream gradient already in dx);
 *   dshift[b,:] += sum_n dh, dscale[b,:] += sum_n dh * norm(x)[*w]   (rows b * dmod_batch_stride; skipped if scale == NULL);
 *   dweight[D] += RMSNorm weight gradient (RMSNorm only, may be NULL).  rowstats_scratch: f32 [M, 2] (row mean / rstd
 *   handed from the row-wise dx kernel to the column-sum kernel; may be NULL when scale == dweight == NULL).
 * jat_gate_bwd: backward of x += gate_b * y:  dy bf16 = gate_b * dx;  dgate[b,:] += sum_n dx * y;
 *   if dbias != NULL: dbias[:] += sum_b gate_b * sum_n dx (f32 atomics; dxsum_scratch is unused and may be NULL).
 * jat_colsum_bf16: out[c] += sum_m a[m, c] (bias gradients).    jat_cast_f32_bf16: elementwise cast.
 * -------------------------------------------------------------------------------------------- */
int jat_adaln_bwd(jat_ctx* ctx, const void* dh_bf16, const float* x, const float* scale, int64_t mod_batch_stride,
                  const float* weight, int norm_kind, float eps, float* dx, int accumulate, float* dshift, float* dscale,
                  int64_t dmod_batch_stride, float* dweight, float* rowstats_scratch, int B, int tokens_per_batch, int D,
                  void* stream);
int jat_gate_bwd(jat_ctx* ctx, const float* dx, const void* y_bf16, const float* gate, int64_t mod_batch_stride,
                 void* dy_bf16, float* dgate, int64_t dmod_batch_stride, float* dxsum_scratch, float* dbias, int B,
                 int tokens_per_batch, int D, void* stream);
/* jat_gate_bwd with the forward's regularisers: y was dropout(acc + bias) with (drop_p, drop_seed) -- dy and dbias get the
 * same mask -- and entered x as gate_rowscale[b] * gate_b * y (DropPath; NULL = 1). */
int jat_gate_bwd_dropout(jat_ctx* ctx, const float* dx, const void* y_bf16, const float* gate, int64_t mod_batch_stride,
                         void* dy_bf16, float* dgate, int64_t dmod_batch_stride, float* dxsum_scratch, float* dbias, int B,
                         int tokens_per_batch, int D, float drop_p, uint32_t drop_seed, const float* gate_rowscale,
                         void* stream);
int jat_colsum_bf16(jat_ctx* ctx, const void* a_bf16, int64_t lda, int M, int cols, float* out, void* stream);
int jat_cast_f32_bf16(jat_ctx* ctx, const float* in, void* out_bf16, int64_t n, void* stream);

/* Gradient exchange in bf16 (replaces the f32 payload of DDP's bucketed all-reduce, reference train_ddp_v3mod2.py:822, 922):
 *   jat_grad_compress:   out_bf16[i] = bf16(in[i] * scale)   -- scale = 1 / world size, so that a SUM all-reduce yields the mean
 *   jat_grad_decompress: out[i] = float(in_bf16[i])           -- back into the f32 gradient bucket
 * Both buffers 16-byte aligned, n elements; one pass each (6 bytes per element).  Used by jat_b200.ddp.bf16_allreduce_hook. */
int jat_grad_compress(jat_ctx* ctx, const float* in, void* out_bf16, int64_t n, float scale, void* stream);
int jat_grad_decompress(jat_ctx* ctx, const void* in_bf16, float* out, int64_t n, void* stream);

/* ----------------------------------------------------------------------------------------------
 * The elementwise work either side of the model call in the training step (train_ddp_v3mod2.py:856-889,
 * train_ddp_v3m2.py:547-585), one pass each instead of ~10 torch kernels and 8 `.item()` syncs.
 *
 * jat_train_inputs: all tensors f32 [B, C, T]; hr_mean / hr_std / lr_mean / lr_std f32 [C]; t, keep f32 [B].
 *   hr_norm = (hr - hr_mean) / hr_std                                   (the regression target)
 *   lr_cond = ((lr - lr_mean) / lr_std + cond_noise * s) * keep[b]      s = cond_scale (* *cond_scale_dev if non-NULL:
 *             the adaptive variant's batch std stays on the device); cond_noise == NULL: no augmentation;
 *             keep == NULL: no CFG condition dropout (keep[b] = 0 zeroes the condition of sample b)
 *   z_t     = t[b] * hr_norm + (1 - t[b]) * noise
 *   Unfused round-to-nearest fp32 in the reference's order: bit-identical to the torch expressions.
 * jat_mse_loss: stats4 (DEVICE double[4], overwritten) = { sum (pred-target)^2, sum pred, sum pred^2, sum target^2 } over
 *   n elements -- the MSE and the monitoring figures of :900-911 -- and, if d_pred != NULL, d_pred = (pred - target) * 2/n,
 *   the gradient of mean((pred - target)^2) that seeds jat_dit_backward.
 * -------------------------------------------------------------------------------------------- */
int jat_train_inputs(jat_ctx* ctx, const float* hr, const float* lr, const float* hr_mean, const float* hr_std,
                     const float* lr_mean, const float* lr_std, const float* noise, const float* cond_noise,
                     const float* cond_scale_dev, float cond_scale, const float* keep, const float* t, float* hr_norm,
                     float* lr_cond, float* z_t, int B, int C, int T, void* stream);
int jat_mse_loss(jat_ctx* ctx, const float* pred, const float* target, float* d_pred, double* stats4, int64_t n, void* stream);
/* Charbonnier reconstruction loss of the MOD3 training script (train_ddp_v3mod3.py:57-85): mean(sqrt((pred - target)^2 + eps)).
 * stats5 (DEVICE double[5], overwritten) = { sum sqrt(d^2 + eps), sum pred, sum pred^2, sum target^2, sum d^2 };
 * d_pred (optional) = d / sqrt(d^2 + eps) / n, the gradient that seeds jat_dit_backward. */
int jat_charbonnier_loss(jat_ctx* ctx, const float* pred, const float* target, float* d_pred, double* stats5, int64_t n, float eps,
                         void* stream);

/* ----------------------------------------------------------------------------------------------
 * Long-audio chunk plumbing (infer_test_v3m2.py:340-406 chunk loop, :188-233 crossfade_chunks).
 * A track latent[C, total_frames] (row pitch ld) is cut into chunks of `chunk_frames` frames starting every
 * `stride` = chunk_frames - overlap frames.
 *
 * jat_chunk_normalize: out[k, c, t] = (latent[c, s_k + t] - mean[c]) / std[c] with
 * s_k = (first_chunk + k * chunk_step) * stride, k < n_chunks; frames past the end of the track are 0
 * (:377-382 for every chunk of a rank at once; chunk_step = world size gives the round-robin shard).
 * mean == std == NULL copies without normalising.
 *
 * jat_crossfade_denorm: out[c, :] = left-fold linear crossfade (:188-233) of the chunks
 * chunks[i, c, :] * std[c] + mean[c] (:394); fade_in / fade_out are the reference's
 * torch.linspace(0, 1, overlap) / torch.linspace(1, 0, overlap) tables (f32 [overlap]).
 * Requires chunk_frames >= 2 * overlap and (n-1)*stride + overlap < total_frames <= (n-1)*stride + chunk_frames.
 * All arithmetic is unfused round-to-nearest fp32 in the reference's order: results are bit-identical.
 * -------------------------------------------------------------------------------------------- */
int jat_chunk_normalize(jat_ctx* ctx, const float* latent, int64_t total_frames, int64_t ld, const float* mean,
                        const float* std, float* out, int n_chunks, int first_chunk, int chunk_step, int C,
                        int chunk_frames, int stride, void* stream);
int jat_crossfade_denorm(jat_ctx* ctx, const float* chunks, int n_chunks, int C, int chunk_frames, int overlap,
                         const float* fade_in, const float* fade_out, const float* mean, const float* std, float* out,
                         int64_t total_frames, int64_t ldo, void* stream);

/* Debug aid: when `buf` (DEVICE, 128 x int64) is non-NULL, CTA (0,0,0) of every following attention launch
 * stores clock64() timestamps of its pipeline events there (scripts/att_trace.py decodes them). NULL = off. */
int jat_debug_set_attention_trace(jat_ctx* ctx, void* buf);
/* The same for the GEMM kernel: 64 work items x 8 slots of int64 (CTA 0): 0 producer starts the item, 1/2 MMA issuer
 * before/after the accumulator-stage wait, 3 MMA issuer committed the item, 4/5 epilogue warp before/after the
 * accumulator-full wait, 6 epilogue done. */
int jat_debug_set_gemm_trace(jat_ctx* ctx, void* buf);

#ifdef __cplusplus
}
#endif
#endif /* JAT_B200_H */
