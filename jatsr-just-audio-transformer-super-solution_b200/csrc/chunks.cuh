// Long-audio chunk plumbing around the sampler (infer_test_v3m2.py:340-406, 188-233): the reference cuts a
// latent track [C, total] into 16 s chunks (1378 frames, 172 frames overlap), normalises each chunk per
// channel, samples it, de-normalises and stitches the results with a linear crossfade in latent space, one
// chunk at a time with ~10 small torch kernels per chunk.  Here: one kernel builds the whole normalised
// chunk batch [n, C, Tc] (so the chunks can be denoised as ONE batch), and one kernel de-normalises +
// crossfades all chunks into the output track.  Both are single-pass, coalesced along time, HBM-bound.
// Arithmetic is explicit round-to-nearest fp32 in the reference's order (no FMA contraction): bit-exact.
#pragma once
#include "common.cuh"

namespace jat {

// out[k, c, t] = (latent[c, s_k + t] - mean[c]) / std[c],  s_k = (first_chunk + k * chunk_step) * stride,
// zero where s_k + t >= total (the short last chunk).  mean == NULL -> plain copy.
__global__ void __launch_bounds__(256)
chunk_normalize_kernel(const float* __restrict__ latent, long long total, long long ld, const float* __restrict__ mean,
                       const float* __restrict__ std, float* __restrict__ out, int C, int Tc, int stride, int first_chunk,
                       int chunk_step) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int c = blockIdx.y, k = blockIdx.z;
    if (t >= Tc) return;
    const long long src = (long long)(first_chunk + k * chunk_step) * stride + t;
    float v = 0.0f;
    if (src < total) {
        v = __ldg(latent + (long long)c * ld + src);
        if (mean != nullptr) v = __fdiv_rn(__fsub_rn(v, __ldg(mean + c)), __ldg(std + c));
    }
    out[((long long)k * C + c) * Tc + t] = v;
}

// out[c, p] = crossfade over chunks of (chunks[i, c, :] * std[c] + mean[c])   (infer_test_v3m2.py:394, 188-233)
// chunk i covers frames [i * stride, i * stride + Tc); in the first `overlap` frames of chunk i >= 1 the previous
// chunk's tail fades out and chunk i fades in with the reference's torch.linspace weights (passed as tables).
// Requires Tc >= 2 * overlap (so at most two chunks meet at any frame), checked by the host.
__global__ void __launch_bounds__(256)
crossfade_denorm_kernel(const float* __restrict__ chunks, int n, int C, int Tc, int overlap,
                        const float* __restrict__ fade_in, const float* __restrict__ fade_out,
                        const float* __restrict__ mean, const float* __restrict__ std, float* __restrict__ out,
                        long long total, long long ldo) {
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int c = blockIdx.y;
    if (p >= total) return;
    const int stride = Tc - overlap;
    long long i0 = p / stride;
    if (i0 > n - 1) i0 = n - 1;
    const int k = (int)(p - i0 * stride);
    const bool dn = mean != nullptr;
    const float m = dn ? __ldg(mean + c) : 0.0f, s = dn ? __ldg(std + c) : 1.0f;
    float cur = __ldg(chunks + ((long long)i0 * C + c) * Tc + k);
    if (dn) cur = __fadd_rn(__fmul_rn(cur, s), m);
    if (i0 >= 1 && k < overlap) {
        float prev = __ldg(chunks + ((long long)(i0 - 1) * C + c) * Tc + k + stride);
        if (dn) prev = __fadd_rn(__fmul_rn(prev, s), m);
        cur = __fadd_rn(__fmul_rn(prev, __ldg(fade_out + k)), __fmul_rn(cur, __ldg(fade_in + k)));
    }
    out[(long long)c * ldo + p] = cur;
}

}  // namespace jat
