// The parameter update of the training step (train_ddp_v3mod2.py:926-928: clip_grad_norm_(1.0) + AdamW.step(), followed
// in this library by the bf16 re-pack of the updated weights for the next forward) as TWO multi-tensor passes:
//   grad_sumsq_kernel   sum of squares of every gradient element -> per-chunk f32 partials; grad_sumsq_final_kernel
//                       adds them in chunk order in f64 (deterministic) -> ||g||^2 on the device, no host sync
//   adamw_kernel        reads p, g, m, v once, applies the clip coefficient min(1, max_norm / (||g|| + 1e-6)) on the fly
//                       (clip_grad_norm_'s formula), the decoupled weight decay and the Adam update in torch's fused
//                       operation order, writes p, m, v and the packed bf16 / f32 copy the GEMMs read
// torch runs the same step as foreach-norm + stack/norm + foreach-mul (read g, read+write g), fused AdamW (read 4, write 3)
// and a re-cast pass (read p, write bf16): 15 parameter-sized f32 streams; here 8.5.  HBM-bound, 128-bit accesses.
// Tensors are described by a device table (one entry per parameter tensor); a chunk = OPT_CHUNK consecutive elements of
// one tensor, chunk_first[i] = index of tensor i's first chunk (binary search per CTA).
#pragma once
#include "common.cuh"

namespace jat {

constexpr int OPT_THREADS = 256;
constexpr int OPT_CHUNK = 4096;  // elements per CTA: 4 x float4 per thread

struct OptTensor {  // mirrors jat_adamw_tensor (include/jat_b200.h)
    float* param;
    const float* grad;
    float* exp_avg;
    float* exp_avg_sq;
    void* packed;       // optional copy of the updated value in the layout the forward reads (NULL: none)
    long long numel;
    int packed_dtype;   // bit 0: dtype of `packed` (JAT_DTYPE_BF16 = 1 / JAT_DTYPE_F32 = 0); bit 1 (JAT_ADAMW_GRAD_BF16): `grad`
                        // points at bf16 values (the all-reduced payload of the bf16 gradient exchange, jat_b200.ddp)
    int vec_ok;         // every pointer 16-byte aligned and numel % 4 == 0
    float bias_corr1;       // 1 - beta1^step of THIS tensor (parameters may have been frozen for some steps)
    float bias_corr2_sqrt;  // sqrt(1 - beta2^step)
};

// 4 consecutive gradient elements, f32 or bf16 source
__device__ __forceinline__ float4 opt_load_grad4(const float* g, long long q, bool bf16) {
    if (!bf16) return __ldcs(reinterpret_cast<const float4*>(g) + q);
    const uint2 w = __ldcs(reinterpret_cast<const uint2*>(g) + q);
    return make_float4(__uint_as_float(w.x << 16), __uint_as_float(w.x & 0xffff0000u), __uint_as_float(w.y << 16),
                       __uint_as_float(w.y & 0xffff0000u));
}
__device__ __forceinline__ float opt_load_grad1(const float* g, long long i, bool bf16) {
    return bf16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(g)[i]) : g[i];
}

__device__ __forceinline__ int opt_find_tensor(const int* __restrict__ chunk_first, int n_tensors, int chunk) {
    int lo = 0, hi = n_tensors - 1;  // largest i with chunk_first[i] <= chunk
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (__ldg(chunk_first + mid) <= chunk) lo = mid; else hi = mid - 1;
    }
    return lo;
}

__global__ void __launch_bounds__(OPT_THREADS)
grad_sumsq_kernel(const OptTensor* __restrict__ table, const int* __restrict__ chunk_first, int n_tensors,
                  float* __restrict__ partials) {
    __shared__ float red[OPT_THREADS / 32];
    const int chunk = blockIdx.x;
    const int ti = opt_find_tensor(chunk_first, n_tensors, chunk);
    const OptTensor t = table[ti];
    const long long e0 = (long long)(chunk - __ldg(chunk_first + ti)) * OPT_CHUNK;
    const long long n = t.numel - e0 < OPT_CHUNK ? t.numel - e0 : OPT_CHUNK;
    const bool gbf = (t.packed_dtype & 2) != 0;
    const float* g = gbf ? reinterpret_cast<const float*>(reinterpret_cast<const __nv_bfloat16*>(t.grad) + e0) : t.grad + e0;
    float acc = 0.f;
    if (t.vec_ok) {
        float4 v[OPT_CHUNK / 4 / OPT_THREADS];
#pragma unroll
        for (int i = 0; i < OPT_CHUNK / 4 / OPT_THREADS; ++i) {
            const int q = i * OPT_THREADS + threadIdx.x;
            v[i] = q * 4 < n ? opt_load_grad4(g, q, gbf) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int i = 0; i < OPT_CHUNK / 4 / OPT_THREADS; ++i)
            acc += v[i].x * v[i].x + v[i].y * v[i].y + v[i].z * v[i].z + v[i].w * v[i].w;
    } else {
        for (long long i = threadIdx.x; i < n; i += OPT_THREADS) { const float x = opt_load_grad1(g, i, gbf); acc += x * x; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < OPT_THREADS / 32; ++w) s += red[w];
        partials[chunk] = s;
    }
}

// one CTA: out[0] (+)= sum of the partials, fixed order, f64
__global__ void __launch_bounds__(1024)
grad_sumsq_final_kernel(const float* __restrict__ partials, int n, double* __restrict__ out, int accumulate) {
    __shared__ double red[32];
    double acc = 0.0;
    for (int i = threadIdx.x; i < n; i += 1024) acc += (double)partials[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        double s = red[threadIdx.x];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (threadIdx.x == 0) out[0] = accumulate ? out[0] + s : s;
    }
}

struct AdamWArgs {
    double lr, beta1, beta2, eps, weight_decay;
    float max_norm;         // <= 0: no clipping
    const double* sumsq;    // ||g||^2 over ALL tensors that are clipped together (device), or NULL
    // f32 images of the coefficients, each rounded once from its f64 value on the host
    float lr_wd, b1, one_minus_b1, b2, one_minus_b2, eps_f;
};

// AdamW (ATen fused_adam_utils.cuh adam_math, ADAMW mode, no amsgrad / maximize / grad scaler).
//   F64 = 0 (default): f32 arithmetic with FMA contraction, coefficients pre-rounded from f64 -- within 1-2 ulp of ATen.
//   F64 = 1 (JAT_ADAMW_F64=1): ATen's operand types verbatim (hyper-parameters are doubles there, so the decay and the
//           moment updates round once from f64).  Measured on B200 at 763 M parameters: 6.8 ms vs 3.5 ms -- the f64 /
//           conversion pipes, not HBM, bound that form (torch's own fused AdamW: 4.7 ms).
template <int F64>
__device__ __forceinline__ void adamw_one(float& p, float g, float& m, float& v, const AdamWArgs& a, float clip, float step_size,
                                          float bc2_sqrt) {
    g *= clip;
    if constexpr (F64) {
        p = (float)((double)p - a.lr * a.weight_decay * (double)p);
        m = (float)(a.beta1 * (double)m + (1.0 - a.beta1) * (double)g);
        v = (float)(a.beta2 * (double)v + (1.0 - a.beta2) * (double)g * (double)g);
        const float denom = (float)((double)(sqrtf(v) / bc2_sqrt) + a.eps);
        p -= step_size * m / denom;
    } else {
        p -= a.lr_wd * p;
        m = a.b1 * m + a.one_minus_b1 * g;
        v = a.b2 * v + a.one_minus_b2 * g * g;
        const float denom = sqrtf(v) / bc2_sqrt + a.eps_f;
        p -= step_size * m / denom;
    }
}

template <int F64>
__global__ void __launch_bounds__(OPT_THREADS)
adamw_kernel(const OptTensor* __restrict__ table, const int* __restrict__ chunk_first, int n_tensors, const AdamWArgs a) {
    const int chunk = blockIdx.x;
    const int ti = opt_find_tensor(chunk_first, n_tensors, chunk);
    const OptTensor t = table[ti];
    const long long e0 = (long long)(chunk - __ldg(chunk_first + ti)) * OPT_CHUNK;
    const long long n = t.numel - e0 < OPT_CHUNK ? t.numel - e0 : OPT_CHUNK;
    float clip = 1.0f;
    if (a.max_norm > 0.f && a.sumsq != nullptr) {
        const float total = (float)sqrt(*a.sumsq);
        const float c = a.max_norm / (total + 1e-6f);    // torch.nn.utils.clip_grad_norm_
        clip = c < 1.0f ? c : 1.0f;
    }
    const float step_size = (float)(a.lr / (double)t.bias_corr1);
    const bool gbf = (t.packed_dtype & 2) != 0;
    const int pdt = t.packed_dtype & 1;
    const float* gsrc = gbf ? reinterpret_cast<const float*>(reinterpret_cast<const __nv_bfloat16*>(t.grad) + e0) : t.grad + e0;
    if (t.vec_ok) {
        constexpr int NV = OPT_CHUNK / 4 / OPT_THREADS;
        float4* p4 = reinterpret_cast<float4*>(t.param + e0);
        float4* m4 = reinterpret_cast<float4*>(t.exp_avg + e0);
        float4* v4 = reinterpret_cast<float4*>(t.exp_avg_sq + e0);
        float4 p[NV], g[NV], m[NV], v[NV];
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int q = i * OPT_THREADS + threadIdx.x;
            if (q * 4 < n) { p[i] = __ldcs(p4 + q); g[i] = opt_load_grad4(gsrc, q, gbf); m[i] = __ldcs(m4 + q); v[i] = __ldcs(v4 + q); }
        }
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int q = i * OPT_THREADS + threadIdx.x;
            if (q * 4 >= n) continue;
            adamw_one<F64>(p[i].x, g[i].x, m[i].x, v[i].x, a, clip, step_size, t.bias_corr2_sqrt);
            adamw_one<F64>(p[i].y, g[i].y, m[i].y, v[i].y, a, clip, step_size, t.bias_corr2_sqrt);
            adamw_one<F64>(p[i].z, g[i].z, m[i].z, v[i].z, a, clip, step_size, t.bias_corr2_sqrt);
            adamw_one<F64>(p[i].w, g[i].w, m[i].w, v[i].w, a, clip, step_size, t.bias_corr2_sqrt);
            __stcs(p4 + q, p[i]); __stcs(m4 + q, m[i]); __stcs(v4 + q, v[i]);
            if (t.packed != nullptr) {
                if (pdt == 1)   // JAT_DTYPE_BF16: stays in L2 for the next forward where it fits
                    reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(t.packed) + e0)[q] =
                        make_uint2(pack_bf16(p[i].x, p[i].y), pack_bf16(p[i].z, p[i].w));
                else
                    reinterpret_cast<float4*>(reinterpret_cast<float*>(t.packed) + e0)[q] = p[i];
            }
        }
    } else {
        for (long long i = threadIdx.x; i < n; i += OPT_THREADS) {
            float p = t.param[e0 + i], m = t.exp_avg[e0 + i], v = t.exp_avg_sq[e0 + i];
            adamw_one<F64>(p, opt_load_grad1(gsrc, i, gbf), m, v, a, clip, step_size, t.bias_corr2_sqrt);
            t.param[e0 + i] = p; t.exp_avg[e0 + i] = m; t.exp_avg_sq[e0 + i] = v;
            if (t.packed != nullptr) {
                if (pdt == 1) reinterpret_cast<__nv_bfloat16*>(t.packed)[e0 + i] = __float2bfloat16_rn(p);
                else reinterpret_cast<float*>(t.packed)[e0 + i] = p;
            }
        }
    }
}

}  // namespace jat
