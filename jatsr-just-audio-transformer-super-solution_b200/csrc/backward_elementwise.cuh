// HBM-bound kernels of the DiT block BACKWARD pass (training step, train_ddp_v3mod2.py:886-922 runs autograd
// through jat_audiosr_v2.py:265-289).  All single-pass, vectorised; per-batch-item column reductions (the gradients
// of the adaLN shift / scale / gate vectors, which are broadcast over the N tokens of a batch item) are
// accumulated in registers over the rows a CTA owns, reduced across its warps in shared memory and flushed with
// one f32 atomicAdd per column per CTA.
#pragma once
#include "common.cuh"

namespace jat {

constexpr int BWD_WARPS = 8;
constexpr int BWD_ROWS_PER_CTA = 32;  // rows of ONE batch item per CTA (4 per warp)

__device__ __forceinline__ float4 bf16x4_to_f32(uint2 v) {
    return make_float4(__uint_as_float(v.x << 16), __uint_as_float(v.x & 0xffff0000u), __uint_as_float(v.y << 16),
                       __uint_as_float(v.y & 0xffff0000u));
}

// cross-warp reduction of per-lane column partials acc[NV] (float4 each; lane owns vec4 columns lane + 32 i) and
// atomicAdd into dst[0 .. D)
template <int NV>
__device__ __forceinline__ void cta_colsum_flush(const float4 (&acc)[NV], float* smem /* [D] */, float* dst, int nvec) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float4* s4 = reinterpret_cast<float4*>(smem);
    __syncthreads();
    for (int i = threadIdx.x; i < nvec; i += blockDim.x) s4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
    for (int w = 0; w < BWD_WARPS; ++w) {  // warps take turns: no shared-memory atomics needed
        if (warp == w) {
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                const int idx = lane + 32 * i;
                if (idx < nvec) {
                    float4 t = s4[idx];
                    t.x += acc[i].x; t.y += acc[i].y; t.z += acc[i].z; t.w += acc[i].w;
                    s4[idx] = t;
                }
            }
        }
        __syncthreads();
    }
    for (int i = threadIdx.x; i < nvec * 4; i += blockDim.x) atomicAdd(dst + i, smem[i]);
}

// ------------------------------------------------------------------------------------------------
// AdaLN backward.  Forward (jat_audiosr_v2.py:278-279 / jat_audiosr_v3.py RMSNorm):
//     y = norm(x) [* w]          h = y * (1 + scale_b) + shift_b
// Given dh (bf16) and the saved x (f32):
//     dshift_b += sum_n dh        dscale_b += sum_n dh * y        [dw += sum_rows dh (1+scale) * xhat   (RMSNorm)]
//     LayerNorm: g = dh (1+scale);            dx (+)= rstd (g - mean(g) - xhat mean(g xhat))
//     RMSNorm:   g = dh (1+scale) w;          dx (+)= rstd (g - xhat mean(g xhat)),   xhat = x rstd
// grid = (ceil(N / 32), B): a CTA owns 32 consecutive tokens of one batch item.  dx is accumulated in place
// (accumulate = 1: the residual-stream gradient already holds the skip-path gradient) or overwritten.
// ------------------------------------------------------------------------------------------------
template <int NV, int NORM_KIND>
__global__ void __launch_bounds__(BWD_WARPS * 32)
adaln_bwd_kernel(const __nv_bfloat16* __restrict__ dh, const float* __restrict__ x, const float* __restrict__ scale,
                 long long mod_bstride, const float* __restrict__ weight, float eps, float* __restrict__ dx, int accumulate,
                 float* __restrict__ dshift, float* __restrict__ dscale, long long dmod_bstride, float* __restrict__ dweight,
                 int D, int tokens_per_batch) {
    extern __shared__ float red_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.y;
    const int nvec = D >> 2;
    const bool has_mod = scale != nullptr;
    const float4* sc = has_mod ? reinterpret_cast<const float4*>(scale + (long long)b * mod_bstride) : nullptr;
    const float4* wv = reinterpret_cast<const float4*>(weight);
    const float inv_d = 1.0f / (float)D;

    float4 a_shift[NV], a_scale[NV], a_w[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) a_shift[i] = a_scale[i] = a_w[i] = make_float4(0.f, 0.f, 0.f, 0.f);

    for (int rr = warp; rr < BWD_ROWS_PER_CTA; rr += BWD_WARPS) {
        const int n = blockIdx.x * BWD_ROWS_PER_CTA + rr;
        if (n >= tokens_per_batch) break;
        const long long row = (long long)b * tokens_per_batch + n;
        const float4* xr = reinterpret_cast<const float4*>(x + row * D);
        const uint2* gr = reinterpret_cast<const uint2*>(dh + row * D);
        float4 xv[NV], gv[NV];
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int idx = lane + 32 * i;
            xv[i] = idx < nvec ? xr[idx] : make_float4(0.f, 0.f, 0.f, 0.f);
            gv[i] = idx < nvec ? bf16x4_to_f32(__ldg(gr + idx)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        float mean = 0.f, rstd;
        if constexpr (NORM_KIND == 0) {
            float s = 0.f;
#pragma unroll
            for (int i = 0; i < NV; ++i) s += (xv[i].x + xv[i].y) + (xv[i].z + xv[i].w);
            mean = warp_sum(s) * inv_d;
            float q = 0.f;
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                if (lane + 32 * i < nvec) {
                    const float a = xv[i].x - mean, bb = xv[i].y - mean, c = xv[i].z - mean, d = xv[i].w - mean;
                    q += (a * a + bb * bb) + (c * c + d * d);
                }
            }
            rstd = rsqrtf(warp_sum(q) * inv_d + eps);
        } else {
            float q = 0.f;
#pragma unroll
            for (int i = 0; i < NV; ++i) q += (xv[i].x * xv[i].x + xv[i].y * xv[i].y) + (xv[i].z * xv[i].z + xv[i].w * xv[i].w);
            rstd = rsqrtf(warp_sum(q) * inv_d + eps);
        }
        // xhat in xv, g (gradient w.r.t. xhat) in gv; column partials on the way
        float sg = 0.f, sgx = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int idx = lane + 32 * i;
            if (idx < nvec) {
                float4 xh = make_float4((xv[i].x - mean) * rstd, (xv[i].y - mean) * rstd, (xv[i].z - mean) * rstd,
                                        (xv[i].w - mean) * rstd);
                float4 g = gv[i];
                float4 y = xh;
                float4 w4 = make_float4(1.f, 1.f, 1.f, 1.f);
                if constexpr (NORM_KIND == 1) {
                    w4 = __ldg(wv + idx);
                    y.x *= w4.x; y.y *= w4.y; y.z *= w4.z; y.w *= w4.w;
                }
                if (has_mod) {
                    a_shift[i].x += g.x; a_shift[i].y += g.y; a_shift[i].z += g.z; a_shift[i].w += g.w;
                    a_scale[i].x += g.x * y.x; a_scale[i].y += g.y * y.y; a_scale[i].z += g.z * y.z; a_scale[i].w += g.w * y.w;
                    const float4 s4 = __ldg(sc + idx);
                    g.x *= 1.0f + s4.x; g.y *= 1.0f + s4.y; g.z *= 1.0f + s4.z; g.w *= 1.0f + s4.w;
                }
                if constexpr (NORM_KIND == 1) {
                    a_w[i].x += g.x * xh.x; a_w[i].y += g.y * xh.y; a_w[i].z += g.z * xh.z; a_w[i].w += g.w * xh.w;
                    g.x *= w4.x; g.y *= w4.y; g.z *= w4.z; g.w *= w4.w;
                }
                sg += (g.x + g.y) + (g.z + g.w);
                sgx += (g.x * xh.x + g.y * xh.y) + (g.z * xh.z + g.w * xh.w);
                xv[i] = xh;
                gv[i] = g;
            }
        }
        const float mg = NORM_KIND == 0 ? warp_sum(sg) * inv_d : 0.0f;
        const float mgx = warp_sum(sgx) * inv_d;
        float4* dxr = reinterpret_cast<float4*>(dx + row * D);
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int idx = lane + 32 * i;
            if (idx < nvec) {
                float4 o;
                o.x = rstd * (gv[i].x - mg - xv[i].x * mgx);
                o.y = rstd * (gv[i].y - mg - xv[i].y * mgx);
                o.z = rstd * (gv[i].z - mg - xv[i].z * mgx);
                o.w = rstd * (gv[i].w - mg - xv[i].w * mgx);
                if (accumulate) {
                    const float4 old = dxr[idx];
                    o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
                }
                dxr[idx] = o;
            }
        }
    }
    if (has_mod) {
        cta_colsum_flush<NV>(a_shift, red_smem, dshift + (long long)b * dmod_bstride, nvec);
        cta_colsum_flush<NV>(a_scale, red_smem, dscale + (long long)b * dmod_bstride, nvec);
    }
    if constexpr (NORM_KIND == 1) {
        if (dweight != nullptr) cta_colsum_flush<NV>(a_w, red_smem, dweight, nvec);
    }
}

// ------------------------------------------------------------------------------------------------
// Gate backward.  Forward (jat_audiosr_v2.py:281,287): x += gate_b * y.   Given the residual-stream gradient dx (f32)
// and the saved y (bf16):   dy = gate_b * dx (bf16, the A operand of the following dgrad / wgrad GEMMs)
//     dgate_b += sum_n dx * y          dxsum_b += sum_n dx   (db = sum_b gate_b * dxsum_b, finished by gate_bias_grad_kernel)
// ------------------------------------------------------------------------------------------------
template <int NV>
__global__ void __launch_bounds__(BWD_WARPS * 32)
gate_bwd_kernel(const float* __restrict__ dx, const __nv_bfloat16* __restrict__ y, const float* __restrict__ gate,
                long long mod_bstride, __nv_bfloat16* __restrict__ dy, float* __restrict__ dgate, long long dmod_bstride,
                float* __restrict__ dxsum, int D, int tokens_per_batch) {
    extern __shared__ float red_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.y;
    const int nvec = D >> 2;
    const float4* gt = reinterpret_cast<const float4*>(gate + (long long)b * mod_bstride);
    float4 a_gate[NV], a_sum[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) a_gate[i] = a_sum[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int rr = warp; rr < BWD_ROWS_PER_CTA; rr += BWD_WARPS) {
        const int n = blockIdx.x * BWD_ROWS_PER_CTA + rr;
        if (n >= tokens_per_batch) break;
        const long long row = (long long)b * tokens_per_batch + n;
        const float4* dr = reinterpret_cast<const float4*>(dx + row * D);
        const uint2* yr = reinterpret_cast<const uint2*>(y + row * D);
        uint2* or_ = reinterpret_cast<uint2*>(dy + row * D);
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int idx = lane + 32 * i;
            if (idx < nvec) {
                const float4 d = dr[idx];
                const float4 yv = bf16x4_to_f32(__ldg(yr + idx));
                const float4 g4 = __ldg(gt + idx);
                a_gate[i].x += d.x * yv.x; a_gate[i].y += d.y * yv.y; a_gate[i].z += d.z * yv.z; a_gate[i].w += d.w * yv.w;
                a_sum[i].x += d.x; a_sum[i].y += d.y; a_sum[i].z += d.z; a_sum[i].w += d.w;
                or_[idx] = make_uint2(pack_bf16(d.x * g4.x, d.y * g4.y), pack_bf16(d.z * g4.z, d.w * g4.w));
            }
        }
    }
    cta_colsum_flush<NV>(a_gate, red_smem, dgate + (long long)b * dmod_bstride, nvec);
    if (dxsum != nullptr) cta_colsum_flush<NV>(a_sum, red_smem, dxsum + (long long)b * D, nvec);
}

// db[d] += sum_b gate[b, d] * dxsum[b, d]      (bias of mlp.3: y = acc + bias enters x through the gate)
__global__ void gate_bias_grad_kernel(const float* __restrict__ gate, long long mod_bstride, const float* __restrict__ dxsum,
                                      float* __restrict__ dbias, int B, int D) {
    const int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= D) return;
    float s = 0.f;
    for (int b = 0; b < B; ++b) s += gate[(long long)b * mod_bstride + d] * dxsum[(long long)b * D + d];
    dbias[d] += s;
}

// ------------------------------------------------------------------------------------------------
// Column sums of a bf16 matrix: out[c] += sum_m a[m, c]   (bias gradients of mlp.0, patch_embed, final_layer, ...).
// grid = (ceil(cols / 256), row chunks); thread = 2 adjacent columns, rows strided by the chunk count.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
colsum_bf16_kernel(const __nv_bfloat16* __restrict__ a, long long lda, int M, int cols, float* __restrict__ out) {
    const int c = (blockIdx.x * 128 + threadIdx.x) * 2;
    if (c >= cols) return;
    float s0 = 0.f, s1 = 0.f;
    for (int m = blockIdx.y; m < M; m += gridDim.y) {
        const uint32_t v = __ldg(reinterpret_cast<const uint32_t*>(a + (long long)m * lda + c));
        s0 += __uint_as_float(v << 16);
        s1 += __uint_as_float(v & 0xffff0000u);
    }
    atomicAdd(out + c, s0);
    if (c + 1 < cols) atomicAdd(out + c + 1, s1);
}

// f32 -> bf16 cast of a matrix (gradients that become GEMM operands, e.g. dmod)
__global__ void cast_f32_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, long long n) {
    const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i + 3 < n) {
        const float4 v = *reinterpret_cast<const float4*>(in + i);
        *reinterpret_cast<uint2*>(out + i) = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
    } else {
        for (long long j = i; j < n; ++j) out[j] = __float2bfloat16(in[j]);
    }
}

}  // namespace jat
