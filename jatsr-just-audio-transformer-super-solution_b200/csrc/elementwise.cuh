// HBM-bound kernels of the DiT denoiser hot path: fused AdaLN norm+modulate, patchify+concat+cast,
// sinusoidal timestep features, fused CFG-combine + x-pred->velocity + Euler update.
// All are single-pass, vectorised and coalesced; no shared-memory reuse is needed except for the
// patchify transpose.
#pragma once
#include "common.cuh"

namespace jat {

// ------------------------------------------------------------------------------------------------
// Fused AdaLN:  out = norm(x) * (1 + scale_b) + shift_b      (jat_audiosr_v2.py:278-279, 284-285, 361;
// RMSNorm variant jat_audiosr_v3.py:261,264,384).  One warp per token row: the row lives in
// registers (NV float4 per lane), mean / variance by warp-shuffle reduction, one read of x (f32)
// and one write of the bf16 GEMM operand -> algorithmic bytes = M*D*(4+2).
// ------------------------------------------------------------------------------------------------
constexpr int ADALN_WARPS = 8;

template <int NV, int NORM_KIND>
__global__ void __launch_bounds__(ADALN_WARPS * 32)
adaln_norm_modulate_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out,
                           const float* __restrict__ shift, const float* __restrict__ scale, long long mod_bstride,
                           const float* __restrict__ weight, float eps, int M, int D, int tokens_per_batch,
                           float2* __restrict__ rowstats, float* __restrict__ x_copy) {
    // rowstats / x_copy (training forward, optional): the row's (mean, rstd) and an f32 copy of the row, kept for the backward
    pdl_wait();  // (no early launch_dependents: the successor's CTAs would take occupancy from this grid's later waves)
    // Rows are visited from the LAST to the first: the GEMM that produced x wrote its row blocks in ascending order, so the
    // highest rows are the ones still in L2; and the GEMM that consumes `out` starts at row block 0, i.e. at the rows this
    // grid writes last.
    const int row = M - 1 - (int)(blockIdx.x * ADALN_WARPS + (threadIdx.x >> 5));
    if (row < 0) return;
    const int lane = threadIdx.x & 31;
    const int nvec = D >> 2;
    const float4* xr = reinterpret_cast<const float4*>(x + (long long)row * D);

    float4 v[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int idx = lane + 32 * i;
        v[i] = idx < nvec ? xr[idx] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (x_copy != nullptr) {
        float4* cr = reinterpret_cast<float4*>(x_copy + (long long)row * D);
#pragma unroll
        for (int i = 0; i < NV; ++i)
            if (lane + 32 * i < nvec) __stcs(cr + lane + 32 * i, v[i]);
    }
    const float inv_d = 1.0f / (float)D;
    float mean = 0.f, rstd;
    if constexpr (NORM_KIND == 0) {  // LayerNorm, biased variance, two-pass in registers
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
        mean = warp_sum(s) * inv_d;
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            if (lane + 32 * i < nvec) {
                const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
                q += (a * a + b * b) + (c * c + d * d);
            }
        }
        rstd = rsqrtf(warp_sum(q) * inv_d + eps);
    } else {  // RMSNorm
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) q += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
        rstd = rsqrtf(warp_sum(q) * inv_d + eps);
    }

    if (rowstats != nullptr && lane == 0) rowstats[row] = make_float2(mean, rstd);
    const bool has_mod = shift != nullptr;
    const long long boff = has_mod ? (long long)(row / tokens_per_batch) * mod_bstride : 0;
    const float4* sh = has_mod ? reinterpret_cast<const float4*>(shift + boff) : nullptr;
    const float4* sc = has_mod ? reinterpret_cast<const float4*>(scale + boff) : nullptr;
    const float4* wv = reinterpret_cast<const float4*>(weight);
    uint2* orow = reinterpret_cast<uint2*>(out + (long long)row * D);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int idx = lane + 32 * i;
        if (idx < nvec) {
            float4 y;
            y.x = (v[i].x - mean) * rstd;
            y.y = (v[i].y - mean) * rstd;
            y.z = (v[i].z - mean) * rstd;
            y.w = (v[i].w - mean) * rstd;
            if constexpr (NORM_KIND == 1) {
                const float4 w4 = __ldg(wv + idx);
                y.x *= w4.x; y.y *= w4.y; y.z *= w4.z; y.w *= w4.w;
            }
            if (has_mod) {
                const float4 s4 = __ldg(sc + idx), h4 = __ldg(sh + idx);
                y.x = y.x * (1.0f + s4.x) + h4.x;
                y.y = y.y * (1.0f + s4.y) + h4.y;
                y.z = y.z * (1.0f + s4.z) + h4.z;
                y.w = y.w * (1.0f + s4.w) + h4.w;
            }
            orow[idx] = make_uint2(pack_bf16(y.x, y.y), pack_bf16(y.z, y.w));
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Patchify + channel concat + zero pad + f32->bf16 cast (jat_audiosr_v2.py:411-421, 225-227).
//   out[b*N + n, c*4 + p] = src[b', c, 4n + p]     src = x_t for c < C, x_cond (Cc channels) for C <= c < C + Cc
// Tile = 32 channels x 32 tokens: coalesced f32 reads along T into shared memory, then each warp
// writes 32 channels x 4 = 128 contiguous bf16 (256 B) of one token row.
// ------------------------------------------------------------------------------------------------
constexpr int PATCH_TC = 32;   // channels per tile
constexpr int PATCH_TN = 32;   // tokens per tile
constexpr int PATCH_ROW = PATCH_TN * 4 + 4;  // padded smem row (floats)

__global__ void __launch_bounds__(256)
patchify_cast_kernel(const float* __restrict__ x_t, int xt_batch, const float* __restrict__ x_cond, int cond_batch,
                     __nv_bfloat16* __restrict__ out, int C, int Cc, int T, int N, int K) {
    __shared__ __align__(16) float tile[PATCH_TC][PATCH_ROW];
    pdl_wait();  // (no early launch_dependents: the successor's CTAs would take occupancy from this grid's later waves)
    const int n0 = blockIdx.x * PATCH_TN;
    const int c0 = blockIdx.y * PATCH_TC;  // in [0, C + Cc): x_t channels first, then the Cc condition channels
    const int b = blockIdx.z;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    const bool is_cond = c0 >= C;
    const float* src = nullptr;
    if (!is_cond) src = x_t + ((long long)(b % xt_batch) * C + c0) * T;
    else if (x_cond != nullptr && b < cond_batch) src = x_cond + ((long long)b * Cc + (c0 - C)) * T;

    const int t0 = n0 * 4;
#pragma unroll
    for (int cc = warp; cc < PATCH_TC; cc += 8) {
        const float* row = src ? src + (long long)cc * T : nullptr;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int t = t0 + q * 32 + lane;
            tile[cc][q * 32 + lane] = (row != nullptr && t < T) ? __ldg(row + t) : 0.0f;
        }
    }
    __syncthreads();
#pragma unroll
    for (int nn = warp; nn < PATCH_TN; nn += 8) {
        const int n = n0 + nn;
        if (n < N) {
            const float4 f = *reinterpret_cast<const float4*>(&tile[lane][nn * 4]);
            uint2* o = reinterpret_cast<uint2*>(out + ((long long)b * N + n) * K + (long long)(c0 + lane) * 4);
            *o = make_uint2(pack_bf16(f.x, f.y), pack_bf16(f.z, f.w));
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Sinusoidal timestep features (TimeEmbedding.forward, jat_audiosr_v2.py:177-190); t in [0,1] is NOT
// scaled by 1000.  Uses accurate sinf/cosf/expf: arguments are < 1 rad * 1, tiny kernel.
// ------------------------------------------------------------------------------------------------
__global__ void timestep_features_kernel(const float* __restrict__ t, __nv_bfloat16* __restrict__ out, int B, int D) {
    pdl_wait();  // (no early launch_dependents: the successor's CTAs would take occupancy from this grid's later waves)
    const int half = D >> 1;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (i >= half || b >= B) return;
    const float k = logf(10000.0f) / (float)(half - 1);
    const float f = expf((float)i * -k);
    const float a = t[b] * f;
    out[(long long)b * D + i] = __float2bfloat16(sinf(a));
    out[(long long)b * D + half + i] = __float2bfloat16(cosf(a));
}

// ------------------------------------------------------------------------------------------------
// Fused sampler update (infer_test_v3m2.py:161-179): CFG combine + x-pred -> velocity + Euler step,
// one read of x_c, x_u, z and one write of z (4 * numel * 4 bytes).  Every operation is an explicit
// round-to-nearest fp32 op in the reference's order (no FMA contraction) so that, given identical
// model outputs, the update is bit-identical to the PyTorch expression.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float euler_one(float z, float xc, float xu, bool has_u, float s, float den, float dt,
                                           bool direct) {
    float x = xc;
    if (has_u) x = __fadd_rn(xu, __fmul_rn(s, __fsub_rn(xc, xu)));
    if (direct) return x;
    const float vel = __fdiv_rn(__fsub_rn(x, z), den);
    return __fadd_rn(z, __fmul_rn(vel, dt));
}

__global__ void __launch_bounds__(256)
cfg_euler_update_kernel(float* __restrict__ z, const float* __restrict__ x_c, const float* __restrict__ x_u,
                        float cfg_scale, const float* __restrict__ t_dt, int step, long long numel) {
    pdl_wait();  // (no early launch_dependents: the successor's CTAs would take occupancy from this grid's later waves)
    const float t = t_dt[2 * step], dt = t_dt[2 * step + 1];
    const bool direct = !(t < 0.999f);
    const float den = __fadd_rn(__fsub_rn(1.0f, t), 1e-5f);
    const bool has_u = x_u != nullptr;
    const long long nvec = numel >> 2;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const bool aligned = ((reinterpret_cast<uintptr_t>(z) | reinterpret_cast<uintptr_t>(x_c) |
                           reinterpret_cast<uintptr_t>(x_u)) & 15) == 0;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (aligned) {
        float4* z4 = reinterpret_cast<float4*>(z);
        const float4* c4 = reinterpret_cast<const float4*>(x_c);
        const float4* u4 = reinterpret_cast<const float4*>(x_u);
        for (long long v = i; v < nvec; v += stride) {
            float4 zz = z4[v];
            const float4 cc = __ldcs(c4 + v);
            const float4 uu = has_u ? __ldcs(u4 + v) : cc;
            zz.x = euler_one(zz.x, cc.x, uu.x, has_u, cfg_scale, den, dt, direct);
            zz.y = euler_one(zz.y, cc.y, uu.y, has_u, cfg_scale, den, dt, direct);
            zz.z = euler_one(zz.z, cc.z, uu.z, has_u, cfg_scale, den, dt, direct);
            zz.w = euler_one(zz.w, cc.w, uu.w, has_u, cfg_scale, den, dt, direct);
            z4[v] = zz;
        }
        for (long long e = (nvec << 2) + i; e < numel; e += stride)
            z[e] = euler_one(z[e], x_c[e], has_u ? x_u[e] : 0.f, has_u, cfg_scale, den, dt, direct);
    } else {
        for (long long e = i; e < numel; e += stride)
            z[e] = euler_one(z[e], x_c[e], has_u ? x_u[e] : 0.f, has_u, cfg_scale, den, dt, direct);
    }
}

// ------------------------------------------------------------------------------------------------
// Train-mode stochastic regularisers (nn.Dropout jat_audiosr_v2.py:158,250,252; DropPath :21-34).
// dropout_scale_mask: materialises the multiplier matrix (0 or 1/keep) the fused kernels apply at element (row, col)
// of a site -- used by the parity tests to rebuild the forward's masks in torch.
// drop_path_scales: out[(i*2 + branch)*B + b] = floor(keep_i + U) / keep_i, the per-sample factor the reference
// multiplies onto gate * branch of block i (1 when rates[i] == 0).
// ------------------------------------------------------------------------------------------------
__global__ void dropout_scale_mask_kernel(float* __restrict__ out, long long n, int cols, DropCfg d) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[i] = d.thresh != 0u ? drop_scale(d, (uint32_t)(i / cols), (uint32_t)(i % cols)) : 1.0f;
}

__global__ void drop_path_scales_kernel(float* __restrict__ out, const float* __restrict__ rates, int depth, int B,
                                        uint32_t seed) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= depth * 2 * B) return;
    const int site = i / B, b = i - site * B;
    const float r = rates[site >> 1];
    float v = 1.0f;
    if (r > 0.0f) {
        const double th = (double)r * 4294967296.0;
        const uint32_t thresh = th >= 4294967295.0 ? 4294967295u : (uint32_t)(th + 0.5);
        v = jat_hash3((uint32_t)b, (uint32_t)site, seed) >= thresh ? 1.0f / (1.0f - r) : 0.0f;
    }
    out[i] = v;
}

}  // namespace jat
