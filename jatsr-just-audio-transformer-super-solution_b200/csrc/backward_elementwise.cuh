// HBM-bound kernels of the DiT block BACKWARD pass (training step, train_ddp_v3mod2.py:886-922 runs autograd
// through jat_audiosr_v2.py:265-289).  All single-pass, vectorised.  Thread = 4 fixed columns and a CTA walks over
// token rows of ONE batch item, so the per-batch-item column reductions (the gradients of the adaLN shift / scale /
// gate vectors, which are broadcast over the N tokens of a batch item) accumulate in a few registers and are
// flushed with one f32 atomicAdd per column per CTA.
#pragma once
#include "common.cuh"

namespace jat {

constexpr int BWD_WARPS = 8;
constexpr int BWD_ROWS_PER_CTA = 32;  // rows of ONE batch item per CTA (4 per warp)

__device__ __forceinline__ float4 bf16x4_to_f32(uint2 v) {
    return make_float4(__uint_as_float(v.x << 16), __uint_as_float(v.x & 0xffff0000u), __uint_as_float(v.y << 16),
                       __uint_as_float(v.y & 0xffff0000u));
}

// ------------------------------------------------------------------------------------------------
// AdaLN backward.  Forward (jat_audiosr_v2.py:278-279 / jat_audiosr_v3.py RMSNorm):
//     y = norm(x) [* w]          h = y * (1 + scale_b) + shift_b
// Given dh (bf16) and the saved x (f32):
//     dshift_b += sum_n dh        dscale_b += sum_n dh * y        [dw += sum_rows dh (1+scale) * xhat   (RMSNorm)]
//     LayerNorm: g = dh (1+scale);            dx (+)= rstd (g - mean(g) - xhat mean(g xhat))
//     RMSNorm:   g = dh (1+scale) w;          dx (+)= rstd (g - xhat mean(g xhat)),   xhat = x rstd
// Two kernels, each with the thread mapping its reductions want:
//   adaln_bwd_dx_kernel      one WARP per token row (row statistics are warp-shuffle reductions, like the forward
//                            kernel); writes dx and the row's (mean, rstd);
//   adaln_bwd_colsum_kernel  one THREAD per 4 columns, a CTA walks 32 rows of one batch item (column sums in
//                            registers, coalesced 5 KB row reads, one atomicAdd per column per CTA).
// dx is accumulated in place (accumulate = 1: the residual-stream gradient already holds the skip-path gradient) or
// overwritten.
// ------------------------------------------------------------------------------------------------
template <int NV, int NORM_KIND>
__global__ void __launch_bounds__(BWD_WARPS * 32)
adaln_bwd_dx_kernel(const __nv_bfloat16* __restrict__ dh, const float* __restrict__ x, const float* __restrict__ scale,
                    long long mod_bstride, const float* __restrict__ weight, float eps, float* __restrict__ dx, int accumulate,
                    float2* __restrict__ rowstats, int M, int D, int tokens_per_batch) {
    const int row = blockIdx.x * BWD_WARPS + (threadIdx.x >> 5);
    if (row >= M) return;
    const int lane = threadIdx.x & 31;
    const int nvec = D >> 2;
    const float inv_d = 1.0f / (float)D;
    const float4* xr = reinterpret_cast<const float4*>(x + (long long)row * D);
    const uint2* gr = reinterpret_cast<const uint2*>(dh + (long long)row * D);
    float4 xv[NV], gv[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int idx = lane + 32 * i;
        xv[i] = idx < nvec ? __ldcs(xr + idx) : make_float4(0.f, 0.f, 0.f, 0.f);
        gv[i] = idx < nvec ? bf16x4_to_f32(__ldcs(gr + idx)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (accumulate) {  // the old dx row is only needed after two warp reductions: pull it towards L2 now
        const float4* dxp = reinterpret_cast<const float4*>(dx + (long long)row * D);
#pragma unroll
        for (int i = 0; i < NV; ++i)
            if (lane + 32 * i < nvec) asm volatile("prefetch.global.L2 [%0];" ::"l"(dxp + lane + 32 * i));
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
    if (lane == 0 && rowstats != nullptr) rowstats[row] = make_float2(mean, rstd);
    const bool has_mod = scale != nullptr;
    const float4* sc = has_mod ? reinterpret_cast<const float4*>(scale + (long long)(row / tokens_per_batch) * mod_bstride) : nullptr;
    const float4* wv = reinterpret_cast<const float4*>(weight);
    float sg = 0.f, sgx = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int idx = lane + 32 * i;
        if (idx < nvec) {
            const float4 xh = make_float4((xv[i].x - mean) * rstd, (xv[i].y - mean) * rstd, (xv[i].z - mean) * rstd,
                                          (xv[i].w - mean) * rstd);
            float4 g = gv[i];
            if (has_mod) {
                const float4 s4 = __ldg(sc + idx);
                g.x *= 1.0f + s4.x; g.y *= 1.0f + s4.y; g.z *= 1.0f + s4.z; g.w *= 1.0f + s4.w;
            }
            if constexpr (NORM_KIND == 1) {
                const float4 w4 = __ldg(wv + idx);
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
    float4* dxr = reinterpret_cast<float4*>(dx + (long long)row * D);
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

template <int NORM_KIND>
__global__ void __launch_bounds__(512)
adaln_bwd_colsum_kernel(const __nv_bfloat16* __restrict__ dh, const float* __restrict__ x, const float2* __restrict__ rowstats,
                        const float* __restrict__ scale, long long mod_bstride, const float* __restrict__ weight,
                        float* __restrict__ dshift, float* __restrict__ dscale, long long dmod_bstride,
                        float* __restrict__ dweight, int D, int tokens_per_batch) {
    const int c4 = threadIdx.x;
    if (c4 >= (D >> 2)) return;
    const int b = blockIdx.y;
    const bool has_mod = scale != nullptr;
    float4 sc4 = make_float4(1.f, 1.f, 1.f, 1.f), w4 = sc4;
    if (has_mod) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(scale + (long long)b * mod_bstride) + c4);
        sc4 = make_float4(1.0f + t.x, 1.0f + t.y, 1.0f + t.z, 1.0f + t.w);
    }
    if (NORM_KIND == 1) w4 = __ldg(reinterpret_cast<const float4*>(weight) + c4);
    float4 a_shift = make_float4(0.f, 0.f, 0.f, 0.f), a_scale = a_shift, a_w = a_shift;
    const int n0 = blockIdx.x * BWD_ROWS_PER_CTA;
    const int n1 = min(n0 + BWD_ROWS_PER_CTA, tokens_per_batch);
    const long long row0 = (long long)b * tokens_per_batch;
#pragma unroll 4
    for (int n = n0; n < n1; ++n) {
        const long long row = row0 + n;
        const float2 st = __ldg(rowstats + row);
        const float4 xv = __ldcs(reinterpret_cast<const float4*>(x + row * D) + c4);
        const float4 g = bf16x4_to_f32(__ldcs(reinterpret_cast<const uint2*>(dh + row * D) + c4));
        const float4 xh = make_float4((xv.x - st.x) * st.y, (xv.y - st.x) * st.y, (xv.z - st.x) * st.y, (xv.w - st.x) * st.y);
        a_shift.x += g.x; a_shift.y += g.y; a_shift.z += g.z; a_shift.w += g.w;
        a_scale.x += g.x * xh.x * w4.x; a_scale.y += g.y * xh.y * w4.y; a_scale.z += g.z * xh.z * w4.z; a_scale.w += g.w * xh.w * w4.w;
        if (NORM_KIND == 1) {
            a_w.x += g.x * sc4.x * xh.x; a_w.y += g.y * sc4.y * xh.y; a_w.z += g.z * sc4.z * xh.z; a_w.w += g.w * sc4.w * xh.w;
        }
    }
    if (has_mod) {
        float* p1 = dshift + (long long)b * dmod_bstride + c4 * 4;
        float* p2 = dscale + (long long)b * dmod_bstride + c4 * 4;
        atomicAdd(p1, a_shift.x); atomicAdd(p1 + 1, a_shift.y); atomicAdd(p1 + 2, a_shift.z); atomicAdd(p1 + 3, a_shift.w);
        atomicAdd(p2, a_scale.x); atomicAdd(p2 + 1, a_scale.y); atomicAdd(p2 + 2, a_scale.z); atomicAdd(p2 + 3, a_scale.w);
    }
    if (NORM_KIND == 1 && dweight != nullptr) {
        float* p3 = dweight + c4 * 4;
        atomicAdd(p3, a_w.x); atomicAdd(p3 + 1, a_w.y); atomicAdd(p3 + 2, a_w.z); atomicAdd(p3 + 3, a_w.w);
    }
}

// ------------------------------------------------------------------------------------------------
// Gate backward.  Forward (jat_audiosr_v2.py:281,287): x += gate_b * y.   Given the residual-stream gradient dx (f32)
// and the saved y (bf16):   dy = gate_b * dx (bf16, the A operand of the following dgrad / wgrad GEMMs)
//     dgate_b += sum_n dx * y          db += gate_b * sum_n dx   (bias of mlp.3; f32 atomics)
// Thread = 4 fixed columns; a CTA (D/4 threads) walks over 32 token rows of one batch item, so every row access is one
// fully coalesced 5 KB read and the column partials live in 8 registers -- no cross-thread reduction at all.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512)
gate_bwd_kernel(const float* __restrict__ dx, const __nv_bfloat16* __restrict__ y, const float* __restrict__ gate,
                long long mod_bstride, __nv_bfloat16* __restrict__ dy, float* __restrict__ dgate, long long dmod_bstride,
                float* __restrict__ dbias, int D, int tokens_per_batch, DropCfg drop, const float* __restrict__ rowscale) {
    const int c4 = threadIdx.x;  // vec4 column
    if (c4 >= (D >> 2)) return;
    const int b = blockIdx.y;
    // DropPath (jat_audiosr_v2.py:281,287): the branch entered x as (rowscale_b * gate_b) * y
    const float rs = rowscale != nullptr ? __ldg(rowscale + b) : 1.0f;
    float4 g4 = __ldg(reinterpret_cast<const float4*>(gate + (long long)b * mod_bstride) + c4);
    g4.x *= rs; g4.y *= rs; g4.z *= rs; g4.w *= rs;
    float4 a_gate = make_float4(0.f, 0.f, 0.f, 0.f), a_sum = a_gate;
    const int n0 = blockIdx.x * BWD_ROWS_PER_CTA;
    const int n1 = min(n0 + BWD_ROWS_PER_CTA, tokens_per_batch);
    const long long row0 = (long long)b * tokens_per_batch;
#pragma unroll 4
    for (int n = n0; n < n1; ++n) {
        const long long off = (row0 + n) * D;
        const float4 d = __ldcs(reinterpret_cast<const float4*>(dx + off) + c4);
        const float4 yv = bf16x4_to_f32(__ldcs(reinterpret_cast<const uint2*>(y + off) + c4));
        a_gate.x += d.x * yv.x; a_gate.y += d.y * yv.y; a_gate.z += d.z * yv.z; a_gate.w += d.w * yv.w;
        float4 dm = d;
        if (drop.thresh != 0u) {  // forward: y = dropout(acc + bias) entered x through the gate
            const uint32_t rr = (uint32_t)(row0 + n), cc = (uint32_t)(c4 * 4);
            float m0, m1, m2, m3;
            drop_scale2(drop, rr, cc, m0, m1);
            drop_scale2(drop, rr, cc + 2, m2, m3);
            dm.x *= m0; dm.y *= m1; dm.z *= m2; dm.w *= m3;
        }
        a_sum.x += dm.x; a_sum.y += dm.y; a_sum.z += dm.z; a_sum.w += dm.w;
        const float4 o = make_float4(dm.x * g4.x, dm.y * g4.y, dm.z * g4.z, dm.w * g4.w);
        reinterpret_cast<uint2*>(dy + off)[c4] = make_uint2(pack_bf16(o.x, o.y), pack_bf16(o.z, o.w));
    }
    float* dg = dgate + (long long)b * dmod_bstride + c4 * 4;
    atomicAdd(dg, a_gate.x * rs); atomicAdd(dg + 1, a_gate.y * rs); atomicAdd(dg + 2, a_gate.z * rs); atomicAdd(dg + 3, a_gate.w * rs);
    if (dbias != nullptr) {  // db += rs_b gate_b sum_n mask(dx)   (bias of mlp.3: y = acc + bias enters x through the gate)
        float* ds = dbias + c4 * 4;
        atomicAdd(ds, a_sum.x * g4.x); atomicAdd(ds + 1, a_sum.y * g4.y); atomicAdd(ds + 2, a_sum.z * g4.z); atomicAdd(ds + 3, a_sum.w * g4.w);
    }
}

// ------------------------------------------------------------------------------------------------
// Fused AdaLN backward (+ the gate backward that follows it).  In the block backward every adaln_bwd is followed by the
// gate_bwd of the branch below it, which re-reads the dx row adaln_bwd has just written; and adaln_bwd itself is two
// kernels (row reductions / column reductions) that both read dh and x.  With the row statistics (mean, rstd) kept by the
// forward's norm kernel, ONE kernel in the column mapping does all of it: thread = 4 fixed columns, a CTA walks over rows of
// one batch item in groups of 4; the only cross-thread step is the block reduction of (sum g, sum g xhat) per row
// (warp shuffles + one __syncthreads per group, double-buffered partials in shared memory).
//     xhat = (x - mean) rstd      g = dh (1 + scale_b) [w]        dshift_b += dh      dscale_b += dh xhat [w]    [dw += dh (1 + scale_b) xhat]
//     dx  += rstd (g - mean(g) - xhat mean(g xhat))                (RMSNorm: no mean(g) term)
//     HAS_GATE:   dy = mask (dx) gate_b rs_b (bf16)     dgate_b += rs_b sum dx y     db += rs_b gate_b sum mask(dx)
// Traffic per row element: dh 2 + x 4 + dx 4 + 4 (+ y 2 + dy 2) bytes = 14 (18) vs 14 + 6 (+ 8) for the separate kernels.
// ------------------------------------------------------------------------------------------------
constexpr int AGB_ROWS = 4;

template <int NORM_KIND, int HAS_GATE, int MAXT>  // MAXT = 320: two CTAs per SM (<= 102 registers), else 512 threads, one CTA
__global__ void __launch_bounds__(MAXT, MAXT <= 320 ? 2 : 1)
adaln_gate_bwd_kernel(const __nv_bfloat16* __restrict__ dh, const float* __restrict__ x, const float2* __restrict__ rowstats,
                      const float* __restrict__ scale, long long mod_bstride, const float* __restrict__ weight,
                      float* __restrict__ dx, float* __restrict__ dshift, float* __restrict__ dscale, long long dmod_bstride,
                      float* __restrict__ dweight, const __nv_bfloat16* __restrict__ y, const float* __restrict__ gate,
                      __nv_bfloat16* __restrict__ dy, float* __restrict__ dgate, float* __restrict__ dbias, DropCfg drop,
                      const float* __restrict__ rowscale, int D, int tokens_per_batch, int rows_per_cta) {
    constexpr int R = AGB_ROWS;
    __shared__ float4 red[2][16][2];  // [buffer][warp][sum g | sum g xhat] for the R rows of a group
    const int c4 = threadIdx.x, lane = c4 & 31, warp = c4 >> 5, nw = (int)(blockDim.x >> 5);
    const bool act = c4 < (D >> 2);
    const int b = (int)(gridDim.y - 1 - blockIdx.y);  // last rows first: dh / dx were written in ascending row order, dy is consumed from row 0
    const float inv_d = 1.0f / (float)D;
    float4 sc4 = make_float4(1.f, 1.f, 1.f, 1.f), w4 = sc4, g4 = make_float4(0.f, 0.f, 0.f, 0.f);
    float rs = 1.0f;
    if (act) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(scale + (long long)b * mod_bstride) + c4);
        sc4 = make_float4(1.0f + t.x, 1.0f + t.y, 1.0f + t.z, 1.0f + t.w);
        if (NORM_KIND == 1) w4 = __ldg(reinterpret_cast<const float4*>(weight) + c4);
        if (HAS_GATE) {
            rs = rowscale != nullptr ? __ldg(rowscale + b) : 1.0f;
            g4 = __ldg(reinterpret_cast<const float4*>(gate + (long long)b * mod_bstride) + c4);
            g4.x *= rs; g4.y *= rs; g4.z *= rs; g4.w *= rs;
        }
    }
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 a_shift = zero4, a_scale = zero4, a_w = zero4, a_gate = zero4, a_sum = zero4;
    const int n0 = (int)(gridDim.x - 1 - blockIdx.x) * rows_per_cta;
    const int n1 = min(n0 + rows_per_cta, tokens_per_batch);
    const long long row0 = (long long)b * tokens_per_batch;
    int buf = 0;
    for (int n = n0; n < n1; n += R) {
        float4 xv[R], gv[R], ov[R];
        uint2 yv[R];
        float rstd[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const bool ok = act && n + r < n1;
            const long long row = row0 + n + r;
            const float2 st = ok ? __ldg(rowstats + row) : make_float2(0.f, 0.f);
            const float4 xr = ok ? __ldcs(reinterpret_cast<const float4*>(x + row * D) + c4) : zero4;
            gv[r] = ok ? bf16x4_to_f32(__ldcs(reinterpret_cast<const uint2*>(dh + row * D) + c4)) : zero4;
            ov[r] = ok ? __ldcs(reinterpret_cast<const float4*>(dx + row * D) + c4) : zero4;
            if (HAS_GATE) yv[r] = ok ? __ldcs(reinterpret_cast<const uint2*>(y + row * D) + c4) : make_uint2(0u, 0u);
            rstd[r] = st.y;
            xv[r] = make_float4((xr.x - st.x) * st.y, (xr.y - st.x) * st.y, (xr.z - st.x) * st.y, (xr.w - st.x) * st.y);
        }
        float sg[R], sgx[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const float4 d = gv[r], xh = xv[r];
            a_shift.x += d.x; a_shift.y += d.y; a_shift.z += d.z; a_shift.w += d.w;
            a_scale.x += d.x * xh.x * w4.x; a_scale.y += d.y * xh.y * w4.y; a_scale.z += d.z * xh.z * w4.z; a_scale.w += d.w * xh.w * w4.w;
            float4 g = make_float4(d.x * sc4.x, d.y * sc4.y, d.z * sc4.z, d.w * sc4.w);
            if (NORM_KIND == 1) {
                a_w.x += g.x * xh.x; a_w.y += g.y * xh.y; a_w.z += g.z * xh.z; a_w.w += g.w * xh.w;
                g.x *= w4.x; g.y *= w4.y; g.z *= w4.z; g.w *= w4.w;
            }
            gv[r] = g;
            sg[r] = (g.x + g.y) + (g.z + g.w);
            sgx[r] = (g.x * xh.x + g.y * xh.y) + (g.z * xh.z + g.w * xh.w);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
            for (int r = 0; r < R; ++r) {
                if (NORM_KIND == 0) sg[r] += __shfl_xor_sync(0xffffffffu, sg[r], o);
                sgx[r] += __shfl_xor_sync(0xffffffffu, sgx[r], o);
            }
        }
        if (lane == 0) {
            red[buf][warp][0] = make_float4(sg[0], sg[1], sg[2], sg[3]);
            red[buf][warp][1] = make_float4(sgx[0], sgx[1], sgx[2], sgx[3]);
        }
        __syncthreads();
        float4 tg = zero4, tgx = zero4;
        for (int w = 0; w < nw; ++w) {
            const float4 p0 = red[buf][w][0], p1 = red[buf][w][1];
            tg.x += p0.x; tg.y += p0.y; tg.z += p0.z; tg.w += p0.w;
            tgx.x += p1.x; tgx.y += p1.y; tgx.z += p1.z; tgx.w += p1.w;
        }
        buf ^= 1;
        const float mgs[R] = {tg.x * inv_d, tg.y * inv_d, tg.z * inv_d, tg.w * inv_d};
        const float mgxs[R] = {tgx.x * inv_d, tgx.y * inv_d, tgx.z * inv_d, tgx.w * inv_d};
#pragma unroll
        for (int r = 0; r < R; ++r) {
            if (!(act && n + r < n1)) continue;
            const long long row = row0 + n + r;
            const float mg = NORM_KIND == 0 ? mgs[r] : 0.0f, mgx = mgxs[r], rd = rstd[r];
            float4 o;
            o.x = rd * (gv[r].x - mg - xv[r].x * mgx) + ov[r].x;
            o.y = rd * (gv[r].y - mg - xv[r].y * mgx) + ov[r].y;
            o.z = rd * (gv[r].z - mg - xv[r].z * mgx) + ov[r].z;
            o.w = rd * (gv[r].w - mg - xv[r].w * mgx) + ov[r].w;
            reinterpret_cast<float4*>(dx + row * D)[c4] = o;
            if (HAS_GATE) {
                const float4 yf = bf16x4_to_f32(yv[r]);
                a_gate.x += o.x * yf.x; a_gate.y += o.y * yf.y; a_gate.z += o.z * yf.z; a_gate.w += o.w * yf.w;
                float4 dm = o;
                if (drop.thresh != 0u) {
                    const uint32_t rr = (uint32_t)row, cc = (uint32_t)(c4 * 4);
                    float m0, m1, m2, m3;
                    drop_scale2(drop, rr, cc, m0, m1);
                    drop_scale2(drop, rr, cc + 2, m2, m3);
                    dm.x *= m0; dm.y *= m1; dm.z *= m2; dm.w *= m3;
                }
                a_sum.x += dm.x; a_sum.y += dm.y; a_sum.z += dm.z; a_sum.w += dm.w;
                reinterpret_cast<uint2*>(dy + row * D)[c4] =
                    make_uint2(pack_bf16(dm.x * g4.x, dm.y * g4.y), pack_bf16(dm.z * g4.z, dm.w * g4.w));
            }
        }
    }
    if (!act) return;
    float* p1 = dshift + (long long)b * dmod_bstride + c4 * 4;
    float* p2 = dscale + (long long)b * dmod_bstride + c4 * 4;
    atomicAdd(p1, a_shift.x); atomicAdd(p1 + 1, a_shift.y); atomicAdd(p1 + 2, a_shift.z); atomicAdd(p1 + 3, a_shift.w);
    atomicAdd(p2, a_scale.x); atomicAdd(p2 + 1, a_scale.y); atomicAdd(p2 + 2, a_scale.z); atomicAdd(p2 + 3, a_scale.w);
    if (NORM_KIND == 1 && dweight != nullptr) {
        float* p3 = dweight + c4 * 4;
        atomicAdd(p3, a_w.x); atomicAdd(p3 + 1, a_w.y); atomicAdd(p3 + 2, a_w.z); atomicAdd(p3 + 3, a_w.w);
    }
    if (HAS_GATE) {
        float* dg = dgate + (long long)b * dmod_bstride + c4 * 4;
        atomicAdd(dg, a_gate.x * rs); atomicAdd(dg + 1, a_gate.y * rs); atomicAdd(dg + 2, a_gate.z * rs); atomicAdd(dg + 3, a_gate.w * rs);
        if (dbias != nullptr) {
            float* ds = dbias + c4 * 4;
            atomicAdd(ds, a_sum.x * g4.x); atomicAdd(ds + 1, a_sum.y * g4.y); atomicAdd(ds + 2, a_sum.z * g4.z); atomicAdd(ds + 3, a_sum.w * g4.w);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// The same kernel with its four input streams (x, dx, dh, y) staged through shared memory by the bulk-copy engine.  The
// register version above alternates between a load phase and a reduce / write phase (one __syncthreads per row group), so
// its loads are in flight only part of the time: 46 % of the HBM peak with the issue slots 52 % busy.  Here thread 0 keeps
// AGS_STAGES groups of AGS_ROWS rows in flight (one bulk copy per stream and group: the rows of a CTA are contiguous), the
// column threads read a landed stage into registers, and the stage is refilled right after the group's __syncthreads.
// Stage = AGS_ROWS x D x (4 + 4 + 2 [+ 2]) bytes; 3 stages of 2 rows at D = 1280: 90 KB, two CTAs per SM.
// Needs D % 8 == 0 and 16-byte aligned bases (the launcher falls back to the register version otherwise).
// ------------------------------------------------------------------------------------------------
constexpr int AGS_ROWS = 2;
constexpr int AGS_STAGES = 3;
constexpr int AGS_PAD = 512;  // bytes behind the last stage

template <int NORM_KIND, int HAS_GATE, int MAXT>
__global__ void __launch_bounds__(MAXT, MAXT <= 320 ? 2 : 1)
adaln_gate_bwd_staged_kernel(const __nv_bfloat16* __restrict__ dh, const float* __restrict__ x, const float2* __restrict__ rowstats,
                             const float* __restrict__ scale, long long mod_bstride, const float* __restrict__ weight,
                             float* __restrict__ dx, float* __restrict__ dshift, float* __restrict__ dscale, long long dmod_bstride,
                             float* __restrict__ dweight, const __nv_bfloat16* __restrict__ y, const float* __restrict__ gate,
                             __nv_bfloat16* __restrict__ dy, float* __restrict__ dgate, float* __restrict__ dbias, DropCfg drop,
                             const float* __restrict__ rowscale, int D, int tokens_per_batch, int rows_per_cta) {
    constexpr int R = AGS_ROWS, S = AGS_STAGES, NW = MAXT / 32;
    static_assert(R == 2, "the partial sums of a group travel as one float4");
    extern __shared__ __align__(128) uint8_t ags_smem[];  // S stages + AGS_PAD bytes (threads past D / 4 read, and discard, a few bytes past the last row)
    __shared__ uint64_t full[S];
    __shared__ float4 red[2][NW];  // [buffer][warp] = (sum g, sum g xhat) of the two rows of a group
    const int c4 = threadIdx.x, lane = c4 & 31, warp = c4 >> 5;
    const bool act = c4 < (D >> 2);
    const int b = (int)(gridDim.y - 1 - blockIdx.y);  // last rows first, as above
    const float inv_d = 1.0f / (float)D;
    const uint32_t off_dx = (uint32_t)(R * D) * 4u, off_dh = off_dx * 2u, off_y = off_dh + (uint32_t)(R * D) * 2u;
    const uint32_t stage_bytes = (uint32_t)(R * D) * (HAS_GATE ? 12u : 10u);
    const int n0 = (int)(gridDim.x - 1 - blockIdx.x) * rows_per_cta;
    const int n1 = min(n0 + rows_per_cta, tokens_per_batch);
    const long long row0 = (long long)b * tokens_per_batch;
    const int groups = (n1 - n0 + R - 1) / R;

    auto issue = [&](int g) {  // thread 0: the four streams of group g -> stage g % S
        uint8_t* st = ags_smem + (size_t)(g % S) * stage_bytes;
        const int rows = min(R, n1 - (n0 + g * R));
        const long long e0 = (row0 + n0 + g * R) * D;
        const uint32_t bx = (uint32_t)(rows * D) * 4u, bh = (uint32_t)(rows * D) * 2u;
        mbar_expect_tx(&full[g % S], 2u * bx + (HAS_GATE ? 2u : 1u) * bh);
        bulk_load_1d(st, x + e0, bx, &full[g % S]);
        bulk_load_1d(st + off_dx, dx + e0, bx, &full[g % S]);
        bulk_load_1d(st + off_dh, dh + e0, bh, &full[g % S]);
        if (HAS_GATE) bulk_load_1d(st + off_y, y + e0, bh, &full[g % S]);
    };
    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < S; ++s) mbar_init(&full[s], 1);
        fence_barrier_init();
        for (int g = 0; g < S && g < groups; ++g) issue(g);
    }
    if (c4 < 2 * NW) red[c4 / NW][c4 % NW] = make_float4(0.f, 0.f, 0.f, 0.f);  // entries of warps this launch does not have
    float4 sc4 = make_float4(1.f, 1.f, 1.f, 1.f), w4 = sc4, g4 = make_float4(0.f, 0.f, 0.f, 0.f);
    float rs = 1.0f;
    if (act) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(scale + (long long)b * mod_bstride) + c4);
        sc4 = make_float4(1.0f + t.x, 1.0f + t.y, 1.0f + t.z, 1.0f + t.w);
        if (NORM_KIND == 1) w4 = __ldg(reinterpret_cast<const float4*>(weight) + c4);
        if (HAS_GATE) {
            rs = rowscale != nullptr ? __ldg(rowscale + b) : 1.0f;
            g4 = __ldg(reinterpret_cast<const float4*>(gate + (long long)b * mod_bstride) + c4);
            g4.x *= rs; g4.y *= rs; g4.z *= rs; g4.w *= rs;
        }
    }
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 a_shift = zero4, a_scale = zero4, a_w = zero4, a_gate = zero4, a_sum = zero4;
    float2 st_next[R];
#pragma unroll
    for (int r = 0; r < R; ++r) st_next[r] = n0 + r < n1 ? __ldg(rowstats + row0 + n0 + r) : make_float2(0.f, 0.f);
    __syncthreads();  // barrier initialisation visible before anybody waits
    // smem offsets of this thread's 4 columns inside a stage (f32 rows: 16 bytes per thread, bf16 rows: 8)
    const uint32_t cx = (uint32_t)c4 * 16u, ch = (uint32_t)c4 * 8u;
    int buf = 0;
    for (int g = 0; g < groups; ++g) {
        const int n = n0 + g * R;
        float2 stt[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            stt[r] = st_next[r];
            st_next[r] = n + R + r < n1 ? __ldg(rowstats + row0 + n + R + r) : make_float2(0.f, 0.f);
        }
        mbar_wait(&full[g % S], (uint32_t)((g / S) & 1));
        const uint8_t* st = ags_smem + (size_t)(g % S) * stage_bytes;
        // loads are unconditional (a row past the CTA's range holds stale or uninitialised bytes, a thread past D / 4 reads its
        // neighbours' columns); what is invalid is replaced by zeros with selects, so the loop body has no branches
        float4 xv[R], gv[R], ov[R];
        uint2 yv[R];
        bool ok[R];
        float sg[R], sgx[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            ok[r] = act && n + r < n1;
            float4 xr = *reinterpret_cast<const float4*>(st + (uint32_t)(r * D) * 4u + cx);
            float4 o4 = *reinterpret_cast<const float4*>(st + off_dx + (uint32_t)(r * D) * 4u + cx);
            uint2 dr = *reinterpret_cast<const uint2*>(st + off_dh + (uint32_t)(r * D) * 2u + ch);
            if (HAS_GATE) {
                const uint2 yr = *reinterpret_cast<const uint2*>(st + off_y + (uint32_t)(r * D) * 2u + ch);
                yv[r] = ok[r] ? yr : make_uint2(0u, 0u);
            }
            dr = ok[r] ? dr : make_uint2(0u, 0u);
            ov[r] = ok[r] ? o4 : zero4;
            const float4 d = bf16x4_to_f32(dr);
            const float mean = stt[r].x, rstd = stt[r].y;
            float4 xh = make_float4((xr.x - mean) * rstd, (xr.y - mean) * rstd, (xr.z - mean) * rstd, (xr.w - mean) * rstd);
            xh = ok[r] ? xh : zero4;
            xv[r] = xh;
            a_shift.x += d.x; a_shift.y += d.y; a_shift.z += d.z; a_shift.w += d.w;
            if (NORM_KIND == 1) {
                a_scale.x += d.x * xh.x * w4.x; a_scale.y += d.y * xh.y * w4.y; a_scale.z += d.z * xh.z * w4.z; a_scale.w += d.w * xh.w * w4.w;
            } else {
                a_scale.x += d.x * xh.x; a_scale.y += d.y * xh.y; a_scale.z += d.z * xh.z; a_scale.w += d.w * xh.w;
            }
            float4 gg = make_float4(d.x * sc4.x, d.y * sc4.y, d.z * sc4.z, d.w * sc4.w);
            if (NORM_KIND == 1) {
                a_w.x += gg.x * xh.x; a_w.y += gg.y * xh.y; a_w.z += gg.z * xh.z; a_w.w += gg.w * xh.w;
                gg.x *= w4.x; gg.y *= w4.y; gg.z *= w4.z; gg.w *= w4.w;
            }
            gv[r] = gg;
            sg[r] = (gg.x + gg.y) + (gg.z + gg.w);
            sgx[r] = (gg.x * xh.x + gg.y * xh.y) + (gg.z * xh.z + gg.w * xh.w);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
            for (int r = 0; r < R; ++r) {
                if (NORM_KIND == 0) sg[r] += __shfl_xor_sync(0xffffffffu, sg[r], o);
                sgx[r] += __shfl_xor_sync(0xffffffffu, sgx[r], o);
            }
        }
        if (lane == 0) red[buf][warp] = make_float4(sg[0], sgx[0], sg[1], sgx[1]);
        __syncthreads();  // every thread holds its part of the stage in registers: the stage can be refilled
        if (threadIdx.x == 0 && g + S < groups) issue(g + S);
        float4 tot = zero4;
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            const float4 p0 = red[buf][w];
            tot.x += p0.x; tot.y += p0.y; tot.z += p0.z; tot.w += p0.w;
        }
        buf ^= 1;
        const float mgs[R] = {tot.x * inv_d, tot.z * inv_d};
        const float mgxs[R] = {tot.y * inv_d, tot.w * inv_d};
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const long long row = row0 + n + r;
            const float mg = NORM_KIND == 0 ? mgs[r] : 0.0f, mgx = mgxs[r], rd = stt[r].y;  // rd = 0 for a row past the range
            float4 o;
            o.x = rd * (gv[r].x - mg - xv[r].x * mgx) + ov[r].x;
            o.y = rd * (gv[r].y - mg - xv[r].y * mgx) + ov[r].y;
            o.z = rd * (gv[r].z - mg - xv[r].z * mgx) + ov[r].z;
            o.w = rd * (gv[r].w - mg - xv[r].w * mgx) + ov[r].w;
            if (ok[r]) reinterpret_cast<float4*>(dx + row * D)[c4] = o;
            if (HAS_GATE) {
                const float4 yf = bf16x4_to_f32(yv[r]);
                a_gate.x += o.x * yf.x; a_gate.y += o.y * yf.y; a_gate.z += o.z * yf.z; a_gate.w += o.w * yf.w;
                float4 dm = ok[r] ? o : zero4;
                if (drop.thresh != 0u) {
                    const uint32_t rr = (uint32_t)row, cc = (uint32_t)(c4 * 4);
                    float m0, m1, m2, m3;
                    drop_scale2(drop, rr, cc, m0, m1);
                    drop_scale2(drop, rr, cc + 2, m2, m3);
                    dm.x *= m0; dm.y *= m1; dm.z *= m2; dm.w *= m3;
                }
                a_sum.x += dm.x; a_sum.y += dm.y; a_sum.z += dm.z; a_sum.w += dm.w;
                if (ok[r])
                    reinterpret_cast<uint2*>(dy + row * D)[c4] =
                        make_uint2(pack_bf16(dm.x * g4.x, dm.y * g4.y), pack_bf16(dm.z * g4.z, dm.w * g4.w));
            }
        }
    }
    if (!act) return;
    float* p1 = dshift + (long long)b * dmod_bstride + c4 * 4;
    float* p2 = dscale + (long long)b * dmod_bstride + c4 * 4;
    atomicAdd(p1, a_shift.x); atomicAdd(p1 + 1, a_shift.y); atomicAdd(p1 + 2, a_shift.z); atomicAdd(p1 + 3, a_shift.w);
    atomicAdd(p2, a_scale.x); atomicAdd(p2 + 1, a_scale.y); atomicAdd(p2 + 2, a_scale.z); atomicAdd(p2 + 3, a_scale.w);
    if (NORM_KIND == 1 && dweight != nullptr) {
        float* p3 = dweight + c4 * 4;
        atomicAdd(p3, a_w.x); atomicAdd(p3 + 1, a_w.y); atomicAdd(p3 + 2, a_w.z); atomicAdd(p3 + 3, a_w.w);
    }
    if (HAS_GATE) {
        float* dg = dgate + (long long)b * dmod_bstride + c4 * 4;
        atomicAdd(dg, a_gate.x * rs); atomicAdd(dg + 1, a_gate.y * rs); atomicAdd(dg + 2, a_gate.z * rs); atomicAdd(dg + 3, a_gate.w * rs);
        if (dbias != nullptr) {
            float* ds = dbias + c4 * 4;
            atomicAdd(ds, a_sum.x * g4.x); atomicAdd(ds + 1, a_sum.y * g4.y); atomicAdd(ds + 2, a_sum.z * g4.z); atomicAdd(ds + 3, a_sum.w * g4.w);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Column sums of a bf16 matrix: out[c] += sum_m a[m, c]   (bias gradients of mlp.0, patch_embed, final_layer, ...).
// grid = (ceil(cols / 512), row chunks); thread = 4 adjacent columns (8-byte loads), rows strided by the chunk count,
// 8 loads in flight per thread.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
colsum_bf16_kernel(const __nv_bfloat16* __restrict__ a, long long lda, int M, int cols, float* __restrict__ out) {
    const int c = (blockIdx.x * 128 + threadIdx.x) * 4;
    if (c >= cols) return;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    const int step = gridDim.y;
    int m = blockIdx.y;
    for (; m + 7 * step < M; m += 8 * step) {
        uint2 v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = __ldcs(reinterpret_cast<const uint2*>(a + (long long)(m + j * step) * lda + c));
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float4 f = bf16x4_to_f32(v[j]);
            s.x += f.x; s.y += f.y; s.z += f.z; s.w += f.w;
        }
    }
    for (; m < M; m += step) {
        const float4 f = bf16x4_to_f32(__ldcs(reinterpret_cast<const uint2*>(a + (long long)m * lda + c)));
        s.x += f.x; s.y += f.y; s.z += f.z; s.w += f.w;
    }
    atomicAdd(out + c, s.x); atomicAdd(out + c + 1, s.y); atomicAdd(out + c + 2, s.z); atomicAdd(out + c + 3, s.w);
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

// Gradient exchange in bf16 (DDP bucket -> NCCL payload and back; jat_b200.ddp): grid-stride, 128-bit accesses.
//   compress:    out_bf16[i] = bf16(in_f32[i] * scale)      (scale = 1 / world size: the all-reduce then yields the mean)
//   decompress:  out_f32[i]  = float(in_bf16[i])
// One read of 4 B + one write of 2 B per element (resp. 2 + 4): HBM-bound, 6 B per element.
__global__ void __launch_bounds__(256)
grad_compress_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, long long n, float scale) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long nvec = n >> 3;
    for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += stride) {
        const float4 a = __ldcs(reinterpret_cast<const float4*>(in) + 2 * v), b = __ldcs(reinterpret_cast<const float4*>(in) + 2 * v + 1);
        reinterpret_cast<uint4*>(out)[v] = make_uint4(pack_bf16(a.x * scale, a.y * scale), pack_bf16(a.z * scale, a.w * scale),
                                                      pack_bf16(b.x * scale, b.y * scale), pack_bf16(b.z * scale, b.w * scale));
    }
    for (long long j = (nvec << 3) + (long long)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride)
        out[j] = __float2bfloat16(in[j] * scale);
}
__global__ void __launch_bounds__(256)
grad_decompress_kernel(const __nv_bfloat16* __restrict__ in, float* __restrict__ out, long long n) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long nvec = n >> 3;
    for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += stride) {
        const uint4 w = __ldcs(reinterpret_cast<const uint4*>(in) + v);
        reinterpret_cast<float4*>(out)[2 * v] = make_float4(__uint_as_float(w.x << 16), __uint_as_float(w.x & 0xffff0000u),
                                                            __uint_as_float(w.y << 16), __uint_as_float(w.y & 0xffff0000u));
        reinterpret_cast<float4*>(out)[2 * v + 1] = make_float4(__uint_as_float(w.z << 16), __uint_as_float(w.z & 0xffff0000u),
                                                                __uint_as_float(w.w << 16), __uint_as_float(w.w & 0xffff0000u));
    }
    for (long long j = (nvec << 3) + (long long)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride)
        out[j] = __bfloat162float(in[j]);
}

// out (bf16) = g (f32) * act'(u)   (tiny: the activation backward of the timestep path)
template <int ACT>
__global__ void dact_mul_kernel(const float* __restrict__ g, const __nv_bfloat16* __restrict__ u, __nv_bfloat16* __restrict__ out,
                                long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float uu = __bfloat162float(u[i]);
    float d = 1.0f;
    if (ACT == 2) {  // SiLU
        const float sg = 1.0f / (1.0f + __expf(-uu));
        d = sg * fmaf(uu, 1.0f - sg, 1.0f);
    }
    out[i] = __float2bfloat16(g[i] * d);
}

}  // namespace jat
