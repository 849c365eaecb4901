// The elementwise work either side of the model call in the training step:
//   train_inputs_kernel   normalise + conditional-noise augmentation + CFG condition dropout + flow-matching mix
//                         (train_ddp_v3mod2.py:856-883, train_ddp_v3m2.py:547-580) in ONE pass over the batch
//   recon_loss_kernel     x-prediction MSE / Charbonnier (train_ddp_v3mod2.py:889) with its gradient seed and the monitoring sums
//                         (:900-911: pred mean / std, signal and noise power for the SNR) in ONE pass
// HBM-bound, 128-bit accesses.  All arithmetic is unfused round-to-nearest fp32 in the reference's order, so the three
// outputs of train_inputs are bit-identical to the torch expressions.
#pragma once
#include "common.cuh"

namespace jat {

__global__ void __launch_bounds__(256)
train_inputs_kernel(const float* __restrict__ hr, const float* __restrict__ lr, const float* __restrict__ hr_mean,
                    const float* __restrict__ hr_std, const float* __restrict__ lr_mean, const float* __restrict__ lr_std,
                    const float* __restrict__ noise, const float* __restrict__ cond_noise, const float* __restrict__ cond_scale_dev,
                    float cond_scale, const float* __restrict__ keep, const float* __restrict__ t, float* __restrict__ hr_norm,
                    float* __restrict__ lr_cond, float* __restrict__ z_t, int C, int T) {
    const int c = blockIdx.y, b = blockIdx.z;
    const long long row = ((long long)b * C + c) * T;
    const float hm = hr_mean[c], hs = hr_std[c], lm = lr_mean[c], ls = lr_std[c];
    const float tb = t[b], omt = __fsub_rn(1.0f, tb);
    const float cs = cond_scale_dev != nullptr ? __fmul_rn(cond_scale, *cond_scale_dev) : cond_scale;
    const float kp = keep != nullptr ? keep[b] : 1.0f;
    auto one = [&](float h, float l, float n, float cn, float& o_h, float& o_l, float& o_z) {
        o_h = __fdiv_rn(__fsub_rn(h, hm), hs);
        float ln = __fdiv_rn(__fsub_rn(l, lm), ls);
        if (cond_noise != nullptr) ln = __fadd_rn(ln, __fmul_rn(cn, cs));
        if (keep != nullptr) ln = __fmul_rn(ln, kp);
        o_l = ln;
        o_z = __fadd_rn(__fmul_rn(tb, o_h), __fmul_rn(omt, n));
    };
    const bool vec = ((T & 3) == 0);
    if (vec) {
        for (int i = (blockIdx.x * blockDim.x + threadIdx.x) * 4; i < T; i += gridDim.x * blockDim.x * 4) {
            const float4 h = __ldcs(reinterpret_cast<const float4*>(hr + row + i));
            const float4 l = __ldcs(reinterpret_cast<const float4*>(lr + row + i));
            const float4 n = __ldcs(reinterpret_cast<const float4*>(noise + row + i));
            float4 cn = make_float4(0.f, 0.f, 0.f, 0.f);
            if (cond_noise != nullptr) cn = __ldcs(reinterpret_cast<const float4*>(cond_noise + row + i));
            float4 oh, ol, oz;
            one(h.x, l.x, n.x, cn.x, oh.x, ol.x, oz.x);
            one(h.y, l.y, n.y, cn.y, oh.y, ol.y, oz.y);
            one(h.z, l.z, n.z, cn.z, oh.z, ol.z, oz.z);
            one(h.w, l.w, n.w, cn.w, oh.w, ol.w, oz.w);
            *reinterpret_cast<float4*>(hr_norm + row + i) = oh;
            *reinterpret_cast<float4*>(lr_cond + row + i) = ol;
            *reinterpret_cast<float4*>(z_t + row + i) = oz;
        }
    } else {
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < T; i += gridDim.x * blockDim.x)
            one(hr[row + i], lr[row + i], noise[row + i], cond_noise != nullptr ? cond_noise[row + i] : 0.f, hr_norm[row + i],
                lr_cond[row + i], z_t[row + i]);
    }
}

// LOSS = 0 (MSE, train_ddp_v3mod2.py:889):
//   stats[0] += sum (pred - target)^2   stats[1] += sum pred   stats[2] += sum pred^2   stats[3] += sum target^2   (double)
//   d_pred = (pred - target) * scale  with scale = 2 / n  (the gradient of mean((pred - target)^2)), if d_pred != NULL
// LOSS = 1 (Charbonnier, train_ddp_v3mod3.py:57-85: mean(sqrt((pred - target)^2 + eps))):
//   stats[0] += sum sqrt(d^2 + eps), stats[1..3] as above, stats[4] += sum d^2 (the SNR monitor still needs it);
//   d_pred = d / sqrt(d^2 + eps) * scale  with scale = 1 / n
template <int LOSS>
__global__ void __launch_bounds__(256)
recon_loss_kernel(const float* __restrict__ pred, const float* __restrict__ target, float* __restrict__ d_pred, double* __restrict__ stats,
                  long long n, float scale, float eps) {
    float sl = 0.f, sp = 0.f, spp = 0.f, stt = 0.f, se = 0.f;
    auto one = [&](float p, float q) -> float {
        const float d = p - q;
        sp += p; spp += p * p; stt += q * q;
        if constexpr (LOSS == 0) { sl += d * d; return d * scale; }
        else { const float r = sqrtf(d * d + eps); sl += r; se += d * d; return d / r * scale; }
    };
    const long long stride = (long long)gridDim.x * blockDim.x * 4;
    long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    for (; i + 3 < n; i += stride) {
        const float4 p = __ldcs(reinterpret_cast<const float4*>(pred + i));
        const float4 q = __ldcs(reinterpret_cast<const float4*>(target + i));
        const float4 g = make_float4(one(p.x, q.x), one(p.y, q.y), one(p.z, q.z), one(p.w, q.w));
        if (d_pred != nullptr) *reinterpret_cast<float4*>(d_pred + i) = g;
    }
    if (i < n && i + 3 >= n) {  // ragged tail (n % 4 != 0): handled by the one thread that lands on it
        for (long long j = i; j < n; ++j) {
            const float g = one(pred[j], target[j]);
            if (d_pred != nullptr) d_pred[j] = g;
        }
    }
    sl = warp_sum(sl); sp = warp_sum(sp); spp = warp_sum(spp); stt = warp_sum(stt); se = warp_sum(se);
    __shared__ float red[5][8];
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) { red[0][w] = sl; red[1][w] = sp; red[2][w] = spp; red[3][w] = stt; red[4][w] = se; }
    __syncthreads();
    if (threadIdx.x < (LOSS == 0 ? 4 : 5)) {
        double s = 0.0;
        for (int k = 0; k < 8; ++k) s += (double)red[threadIdx.x][k];
        atomicAdd(stats + threadIdx.x, s);
    }
}

}  // namespace jat
