// GQA attention forward on tcgen05 / TMEM (GroupedQueryAttention.forward, jat_audiosr_v2.py:141-164,
// eval mode: no mask, bidirectional, dropout = identity).
//
//   out[b, n, h, :] = softmax(Q[b,n,h,:] . K[b,:,g,:]^T / 8) V[b,:,g,:],   g = h / G,  G = Hq / Hkv
//
// The reference expands K/V G-fold with repeat_interleave (:147-148) and materialises the
// [B, Hq, N, N] score tensor in HBM several times per block; here one CTA owns (q-tile of 128 rows,
// KV head g, batch b): K_g and V_g ([N, 64] each) are staged in shared memory ONCE by TMA and reused
// by the G query heads of the group; scores and probabilities live only in tensor memory and registers.
//
// At head_dim 64 the tensor work is small; what the kernel has to schedule around is the softmax: one ex2 per
// score on the MUFU (4 lanes/clk per SM sub-partition measured: 128 x 352 scores = 2816 clk per head-tile), the
// dependent chain S MMA -> TMEM read-out -> max -> exp -> P write-back -> P.V MMA, and shared-memory bandwidth
// (operand fetch of the MMAs).  Organisation:
//   * the key range [0, 2 NKH) is split into two halves owned by two softmax warpgroups that never
//     talk to each other: each half uses ITS OWN row maximum and feeds ITS OWN P.V accumulator
//     (O_A, O_B in TMEM); the exact softmax is recovered in the output epilogue with the usual
//     split-KV combine  out = (wA O_A + wB O_B) / (wA sumA + wB sumB),  w = 2^(m_half - max(mA, mB)).
//     No cross-half barrier, so the two warps sharing an SM sub-partition run out of phase and one
//     warp's TMEM read-out / max / P write-back overlaps the other's exponentials;
//   * the S half-row (NKH fp32) is pulled into registers in one burst and the accumulator handed back at
//     once; the single S region of TMEM is time-multiplexed S_A(h), S_B(h), S_A(h+1), ... which also
//     staggers the two warpgroups by one MMA + one read-out;
//   * P goes back to TENSOR MEMORY as packed bf16 (tcgen05.st) and is the A operand of the P.V MMA
//     straight from there: no shared-memory round trip for the [128 x 352] probability tile;
//   * the output epilogue (TMEM -> combine -> bf16 -> global) has its own warpgroup.
//
// TMEM columns (NKH = 176):  S [0,176)  P_A [176,264)  P_B [264,352)  O_A [352,416)  O_B [416,480)
//
// Warp roles (512 threads; setmaxnreg moves registers from warpgroups 2-3 to the softmax warpgroups):
//   warps 0-3   softmax, key half A = keys [0, NKH); thread = query row r = TMEM lane
//   warps 4-7   softmax, key half B = keys [NKH, 2 NKH)
//   warps 8-11  output epilogue; thread = query row
//   warp 12     lane 0: issues S_half = Q K_half^T (M=128, N=NKH, K=64), alternating halves
//   warp 13     lane 0: issues O_half = P_half V_half (M=128, N=64, K=NKH; A from TMEM, V consumed MN-major
//               straight from its row-major [key, 64] shared-memory tile)
//   warp 14     lane 0: Q loads (double-buffered over heads); K, V are loaded once in the prologue
#pragma once
#include <cuda.h>
#include "common.cuh"

namespace jat {

constexpr int ATT_BQ = 128;       // query rows per tile
constexpr int ATT_HD = 64;        // head dim
constexpr int ATT_THREADS = 512;
constexpr int ATT_Q_BYTES = ATT_BQ * 128;  // 16384
constexpr int ATT_MAX_NK = 352;
constexpr int ATT_OA_COL = 352;
constexpr int ATT_OB_COL = 416;

template <int NKH>
struct AttCfg {
    static constexpr int NK = 2 * NKH;
    static constexpr int KV_BYTES = NK * 128;
    static constexpr int STAT_BYTES = 2 * 2 * ATT_BQ * 8;  // [parity][half][row] {m * scale * log2e, sum}
    static constexpr int SMEM_BYTES = 2 * KV_BYTES + 2 * ATT_Q_BYTES + STAT_BYTES + 256 + 1024;
    static constexpr int P_COL = NKH;  // P_half at [NKH + half * NKH/2, +NKH/2)
    static_assert(2 * NKH <= ATT_OA_COL, "S + P regions overlap the O accumulators");
};

struct AttnParams {
    int B, N, Hq, Hkv, G;
    int Gs;             // query heads per CTA: a KV group is cut into ceil(G / Gs) parts (the last one may be shorter), each its own
                        // CTA staging K / V itself.  Gs = G stages K / V once per group; the host picks the Gs whose CTA list has
                        // the shortest makespan on the SMs (attention_heads_per_cta in api.cu)
    int key0, NKeys;    // this launch attends to keys [key0, key0 + NKeys) of every batch item (NKeys <= 2 * NKH); longer
                        // sequences run one launch per key chunk and are merged by attention_combine_kernel
    __nv_bfloat16* out;
    float* lse;         // optional f32 [B, Hq, N]: log2-domain log-sum-exp of the scaled scores (kept for the backward pass)
    float scale_log2e;  // (1/sqrt(64)) * log2(e)
    long long* trace;   // debug: per-event clock64 timestamps of CTA (0,0,0), or NULL
    DropCfg drop;       // train-mode dropout on the probabilities (jat_audiosr_v2.py:158): P.V uses P * mask / keep, the
                        // softmax normaliser the undropped P; mask element = ((b*Hq + h)*N + query, key)
};
#define ATT_TRACE(slot)                                                                                    \
    do {                                                                                                   \
        if (p.trace && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) p.trace[(slot)] = clock64(); \
    } while (0)

// Opaque copy: stops the compiler from hoisting per-instruction descriptor words out of the head loop
// (it would then keep dozens of them live and spill them in the 40-register MMA warps).
__device__ __forceinline__ uint32_t launder_u32(uint32_t x) {
    uint32_t y;
    asm volatile("mov.u32 %0, %1;" : "=r"(y) : "r"(x));
    return y;
}
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

template <int NKH, bool DROP = false>
__global__ void __launch_bounds__(ATT_THREADS, 1)
gqa_attention_fwd_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_kv,
                         const AttnParams p) {
    using Cfg = AttCfg<NKH>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
    uint8_t* sK = smem;
    uint8_t* sV = sK + Cfg::KV_BYTES;
    uint8_t* sQ = sV + Cfg::KV_BYTES;
    float2* stats = reinterpret_cast<float2*>(sQ + 2 * ATT_Q_BYTES);  // [parity][half][row]
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(stats) + Cfg::STAT_BYTES);
    uint64_t* bar_k = bars;             // K landed
    uint64_t* bar_v = bars + 1;         // V landed
    uint64_t* bar_q = bars + 2;         // [2] Q stage landed
    uint64_t* bar_s_full = bars + 4;    // [2] S_half = Q K_half^T retired
    uint64_t* bar_s_free = bars + 6;    // [2] S_half copied to registers by the 128 threads of its warpgroup
    uint64_t* bar_p_full = bars + 8;    // [2] P_half (+ its row statistics) written by its warpgroup
    uint64_t* bar_p_free = bars + 10;   // [2] P_half V_half retired: the P_half columns may be overwritten
    uint64_t* bar_o_full = bars + 12;   // both P.V chains of the head retired
    uint64_t* bar_o_free = bars + 13;   // O_A / O_B read out by the 128 epilogue threads
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);

    // (warp-uniform values go through a full-mask shuffle so that the compiler keeps the MMA issue paths on the
    //  uniform datapath -- see elect_one())
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    // grid = (query tile, KV head, part * B + batch item): the parts of a KV group are the slowest index, so the CTAs of
    // the full parts (Gs heads) are all handed out before those of the ragged last part (G - (parts - 1) Gs heads)
    const int part = (int)blockIdx.z / p.B;
    const int qt = blockIdx.x, g = blockIdx.y, b = (int)blockIdx.z % p.B;
    const int h0 = g * p.G + part * p.Gs;             // first query head of this CTA
    const int Gs = min(p.Gs, p.G - part * p.Gs);      // its number of heads
    const int q_row0 = b * p.N + qt * ATT_BQ;  // global token row of this tile's first query
    const int kv_row0 = b * p.N + p.key0;
    const int k_col = (p.Hq + g) * ATT_HD;
    const int v_col = (p.Hq + p.Hkv + g) * ATT_HD;

    pdl_wait();  // (no early launch_dependents: the successor's CTAs would take occupancy from this grid's later waves)
    if (threadIdx.x == 0) ATT_TRACE(127);
    if (warp == 12) {
        if (lane == 0) {
            tma_prefetch_desc(&tmap_q);
            tma_prefetch_desc(&tmap_kv);
            mbar_init(bar_k, 1);
            mbar_init(bar_v, 1);
            mbar_init(&bar_q[0], 1);
            mbar_init(&bar_q[1], 1);
            for (int i = 0; i < 2; ++i) {
                mbar_init(&bar_s_full[i], 1);
                mbar_init(&bar_s_free[i], 128);
                mbar_init(&bar_p_full[i], 128);
                mbar_init(&bar_p_free[i], 1);
            }
            mbar_init(bar_o_full, 1);
            mbar_init(bar_o_free, 128);
            fence_barrier_init();
            // the loads do not depend on TMEM: issue them before the CTA-wide sync
            mbar_expect_tx(bar_k, Cfg::KV_BYTES);
            for (int i = 0; i < 2; ++i) tma_load_2d(sK + i * NKH * 128, &tmap_kv, bar_k, k_col, kv_row0 + i * NKH);
            mbar_expect_tx(&bar_q[0], ATT_Q_BYTES);
            tma_load_2d(sQ, &tmap_q, &bar_q[0], h0 * ATT_HD, q_row0);
            mbar_expect_tx(bar_v, Cfg::KV_BYTES);
            for (int i = 0; i < 2; ++i) tma_load_2d(sV + i * NKH * 128, &tmap_kv, bar_v, v_col, kv_row0 + i * NKH);
            if (Gs > 1) {
                mbar_expect_tx(&bar_q[1], ATT_Q_BYTES);
                tma_load_2d(sQ + ATT_Q_BYTES, &tmap_q, &bar_q[1], (h0 + 1) * ATT_HD, q_row0);
            }
        }
        __syncwarp();
        tmem_alloc<1>(tmem_slot, 512);
        tmem_relinquish<1>();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

    if (warp >= 12) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
        // descriptors as (lo, hi) 32-bit halves: hi is the same layout constant for every operand
        const uint32_t desc_hi = (uint32_t)(umma_smem_desc_sw128(0) >> 32);
        if (warp == 12) {
            // -------------------------------------------------------------- S MMA issue
            // The whole warp runs the (uniform) control flow; one elected lane issues.  Under elect.sync the operands stay
            // in uniform registers, so the MMAs of a chain issue back to back (under `lane == 0` every tcgen05.mma was
            // wrapped in an R2UR waterfall that cost more than the MMA itself for these short N = 64 / 176 shapes).
            constexpr uint32_t idesc_s = umma_idesc_bf16(ATT_BQ, NKH);
            const uint32_t q_lo0 = (uint32_t)umma_smem_desc_sw128(smem_u32(sQ));
            const uint32_t k_lo0 = (uint32_t)umma_smem_desc_sw128(smem_u32(sK));
            auto issue_s = [&](int h, int half) {  // S_half(h) = Q(h) K_half^T into the shared S region
                const uint32_t q_lo = q_lo0 + (uint32_t)((h & 1) * (ATT_Q_BYTES >> 4));
                const uint32_t k_lo = k_lo0 + (uint32_t)(half * ((NKH * 128) >> 4));
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < ATT_HD / 16; ++k)
                        umma_bf16_ss_lohi(tmem_base, q_lo + 2 * k, desc_hi, k_lo + 2 * k, desc_hi, idesc_s, (uint32_t)(k != 0));
                    umma_commit(&bar_s_full[half]);
                }
                __syncwarp();
            };
            mbar_wait_backoff(bar_k, 0);
            mbar_wait_backoff(&bar_q[0], 0);
            tc_fence_after();
            issue_s(0, 0);
#pragma unroll 1
            for (int h = 0; h < Gs; ++h) {
                const uint32_t ph = (uint32_t)(h & 1);
                mbar_wait_backoff(&bar_s_free[0], ph);  // S_A(h) is in registers -> the S region is free
                tc_fence_after();
                issue_s(h, 1);
                if (lane == 0) ATT_TRACE(4 * h + 0);
                if (h + 1 < Gs) {
                    mbar_wait_backoff(&bar_q[(h + 1) & 1], (uint32_t)(((h + 1) >> 1) & 1));
                    mbar_wait_backoff(&bar_s_free[1], ph);  // S_B(h) is in registers
                    tc_fence_after();
                    issue_s(h + 1, 0);
                    if (lane == 0) ATT_TRACE(4 * h + 1);
                }
            }
        } else if (warp == 13) {
            // -------------------------------------------------------------- P.V MMA issue
            constexpr uint32_t idesc_pv = umma_idesc_bf16(ATT_BQ, ATT_HD, /*a_major=*/0, /*b_major=*/1);
            const uint32_t v_lo0 = (uint32_t)umma_smem_desc_sw128(smem_u32(sV));
            auto issue_pv = [&](int half, bool last) {  // O_half = P_half V_half
                const uint32_t d = tmem_base + (uint32_t)(half ? ATT_OB_COL : ATT_OA_COL);
                // A: P_half in TMEM, 8 packed columns per 16-key step;
                // B: V rows [16 ks, 16 ks + 16): MN-major, 2 groups of 8 key rows (+2 KB per step)
                const uint32_t a = tmem_base + (uint32_t)(Cfg::P_COL + half * (NKH / 2));
                const uint32_t v_lo = v_lo0 + (uint32_t)(half * (NKH / 16) * (2048 >> 4));
                if (elect_one()) {
#pragma unroll
                    for (int i = 0; i < NKH / 16; ++i)
                        umma_bf16_ts_lohi(d, a + (uint32_t)(i * 8), v_lo + (uint32_t)(i * (2048 >> 4)), desc_hi, idesc_pv,
                                          (uint32_t)(i != 0));
                    umma_commit(&bar_p_free[half]);
                    if (last) umma_commit(bar_o_full);
                }
                __syncwarp();
            };
            mbar_wait_backoff(bar_v, 0);
#pragma unroll 1
            for (int h = 0; h < Gs; ++h) {
                const uint32_t ph = (uint32_t)(h & 1);
                if (h > 0) mbar_wait_backoff(bar_o_free, (uint32_t)((h - 1) & 1));  // O(h-1) has left TMEM
                mbar_wait_backoff(&bar_p_full[0], ph);
                tc_fence_after();
                issue_pv(0, false);
                if (lane == 0) ATT_TRACE(4 * h + 2);
                mbar_wait_backoff(&bar_p_full[1], ph);
                tc_fence_after();
                issue_pv(1, true);
                if (lane == 0) ATT_TRACE(4 * h + 3);
            }
        } else if (warp == 14 && lane == 0) {
            // Q(h+2) into stage (h & 1) once S_B(h), the last reader of that stage, has retired
#pragma unroll 1
            for (int h = 0; h + 2 < Gs; ++h) {
                mbar_wait_backoff(&bar_s_full[1], (uint32_t)(h & 1));
                mbar_expect_tx(&bar_q[h & 1], ATT_Q_BYTES);
                tma_load_2d(sQ + (h & 1) * ATT_Q_BYTES, &tmap_q, &bar_q[h & 1], (h0 + h + 2) * ATT_HD, q_row0);
            }
        }
        __syncwarp();
    } else if (warp >= 8) {
        // ------------------------------------------------------------------ output epilogue
        asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
        const int r = (warp & 3) * 32 + lane;
        const uint32_t t_row = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
        const bool row_ok = qt * ATT_BQ + r < p.N;
        __nv_bfloat16* out_row = p.out + (long long)(q_row0 + r) * (p.Hq * ATT_HD) + (long long)h0 * ATT_HD;
#pragma unroll 1
        for (int h = 0; h < Gs; ++h) {
            mbar_wait_backoff(bar_o_full, (uint32_t)(h & 1));
            tc_fence_after();
            const float2 sa = stats[(h & 1) * 2 * ATT_BQ + r];
            const float2 sb = stats[(h & 1) * 2 * ATT_BQ + ATT_BQ + r];
            const float mx = fmaxf(sa.x, sb.x);  // half A always holds at least one real key -> finite
            const float wa0 = ex2_approx(sa.x - mx), wb0 = ex2_approx(sb.x - mx);  // 2^(-inf) = 0 for an all-padding half
            const float den = wa0 * sa.y + wb0 * sb.y;
            const float inv = 1.0f / den;
            const float wa = wa0 * inv, wb = wb0 * inv;
            if (p.lse != nullptr && row_ok)
                p.lse[((long long)b * p.Hq + h0 + h) * p.N + qt * ATT_BQ + r] = mx + __log2f(den);
            __nv_bfloat16* o = out_row + (long long)h * ATT_HD;
#pragma unroll
            for (int c = 0; c < ATT_HD / 16; ++c) {
                uint32_t va[16], vb[16];
                tmem_ld_32x16(t_row + ATT_OA_COL + c * 16, va);
                tmem_ld_32x16(t_row + ATT_OB_COL + c * 16, vb);
                tmem_ld_wait();
                if (row_ok) {
                    uint32_t w[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        w[j] = pack_bf16(fmaf(__uint_as_float(va[2 * j]), wa, __uint_as_float(vb[2 * j]) * wb),
                                         fmaf(__uint_as_float(va[2 * j + 1]), wa, __uint_as_float(vb[2 * j + 1]) * wb));
                    *reinterpret_cast<uint4*>(o + c * 16) = make_uint4(w[0], w[1], w[2], w[3]);
                    *reinterpret_cast<uint4*>(o + c * 16 + 8) = make_uint4(w[4], w[5], w[6], w[7]);
                }
            }
            tc_fence_before();
            mbar_arrive(bar_o_free);
            if (threadIdx.x == 256) ATT_TRACE(96 + h);
        }
    } else {
        // ------------------------------------------------------------------ softmax warpgroups
        asm volatile("setmaxnreg.inc.sync.aligned.u32 208;");
        const int half = warp >> 2;
        const int r = (warp & 3) * 32 + lane;  // query row in tile = TMEM lane
        const uint32_t t_row = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
        const uint32_t t_p = t_row + (uint32_t)(Cfg::P_COL + half * (NKH / 2));
        const float sl2 = p.scale_log2e;
        const int col0 = half * NKH;
        const bool trace_thr = (threadIdx.x & 127) == 0;
        // dropout mask row of (b, first head of the group, this query); + h * N per head
        const uint32_t drop_row = (uint32_t)(((long long)b * p.Hq + h0) * p.N + qt * ATT_BQ + r);
#pragma unroll 1
        for (int h = 0; h < Gs; ++h) {
            mbar_wait(&bar_s_full[half], (uint32_t)(h & 1));
            if (trace_thr) ATT_TRACE(32 + 32 * half + 4 * h + 0);
            tc_fence_after();
            // ---- S half-row -> registers, then release the S region
            uint32_t s[NKH];
#pragma unroll
            for (int c = 0; c < NKH / 16; ++c)
                tmem_ld_32x16(t_row + (uint32_t)(c * 16), *reinterpret_cast<uint32_t(*)[16]>(&s[c * 16]));
            tmem_ld_wait();
            tc_fence_before();
            mbar_arrive(&bar_s_free[half]);
            {   // mask the padded keys: only the 16-column chunks that reach past the last real key (warp-uniform)
                const int nvalid = p.NKeys - col0;
#pragma unroll
                for (int c = 0; c < NKH / 16; ++c) {
                    if (nvalid < (c + 1) * 16) {
#pragma unroll
                        for (int j = c * 16; j < (c + 1) * 16; ++j)
                            if (j >= nvalid) s[j] = 0xff800000u;  // -inf
                    }
                }
            }
            // ---- maximum of this half of the row
            float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
            for (int j = 0; j < NKH; j += 4) {
                m4[0] = fmaxf(m4[0], __uint_as_float(s[j + 0]));
                m4[1] = fmaxf(m4[1], __uint_as_float(s[j + 1]));
                m4[2] = fmaxf(m4[2], __uint_as_float(s[j + 2]));
                m4[3] = fmaxf(m4[3], __uint_as_float(s[j + 3]));
            }
            const float mraw = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
            const float mstat = mraw * sl2;                      // -inf when the whole half is padding
            const float moff = mraw == -INFINITY ? 0.0f : mstat;  // then every exponential below is 2^(-inf) = 0
            if (trace_thr) ATT_TRACE(32 + 32 * half + 4 * h + 1);
            // ---- exponentials, partial row sum, bf16 packing (S registers die as P registers are born)
            float a4[4] = {0.f, 0.f, 0.f, 0.f};
            uint32_t pk[NKH / 2];
#pragma unroll
            for (int j = 0; j < NKH; j += 4) {
                const float e0 = ex2_approx(fmaf(__uint_as_float(s[j + 0]), sl2, -moff));
                const float e1 = ex2_approx(fmaf(__uint_as_float(s[j + 1]), sl2, -moff));
                const float e2 = ex2_approx(fmaf(__uint_as_float(s[j + 2]), sl2, -moff));
                const float e3 = ex2_approx(fmaf(__uint_as_float(s[j + 3]), sl2, -moff));
                a4[0] += e0; a4[1] += e1; a4[2] += e2; a4[3] += e3;
                if constexpr (DROP) {
                    const uint32_t kc = (uint32_t)(p.key0 + col0 + j);
                    float m0, m1, m2, m3;  // kc is even: two keys per hash
                    drop_scale2(p.drop, drop_row + (uint32_t)h * (uint32_t)p.N, kc, m0, m1);
                    drop_scale2(p.drop, drop_row + (uint32_t)h * (uint32_t)p.N, kc + 2, m2, m3);
                    pk[j / 2] = pack_bf16(e0 * m0, e1 * m1);
                    pk[j / 2 + 1] = pack_bf16(e2 * m2, e3 * m3);
                } else {
                    pk[j / 2] = pack_bf16(e0, e1);
                    pk[j / 2 + 1] = pack_bf16(e2, e3);
                }
            }
            if (trace_thr) ATT_TRACE(32 + 32 * half + 4 * h + 2);
            // ---- P_half(h-1) V_half must have retired before its columns are overwritten (this also orders the
            //      statistics ring: the epilogue of head h-2 finished before P.V(h-1) was issued)
            if (h > 0) {
                mbar_wait(&bar_p_free[half], (uint32_t)((h - 1) & 1));
                tc_fence_after();
            }
#pragma unroll
            for (int c = 0; c < NKH / 32; ++c) tmem_st_32x16(t_p + (uint32_t)(c * 16), &pk[c * 16]);
            if constexpr ((NKH % 32) != 0) tmem_st_32x8(t_p + (uint32_t)((NKH / 32) * 16), &pk[(NKH / 32) * 16]);
            stats[(h & 1) * 2 * ATT_BQ + half * ATT_BQ + r] = make_float2(mstat, (a4[0] + a4[1]) + (a4[2] + a4[3]));
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(&bar_p_full[half]);
            if (trace_thr) ATT_TRACE(32 + 32 * half + 4 * h + 3);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) ATT_TRACE(126);
    if (warp == 12) {
        tc_fence_after();
        tmem_dealloc<1>(tmem_base, 512);
    }
}

// ------------------------------------------------------------------------------------------------
// Merge of the per-key-chunk passes of a long sequence (N > 352 tokens): pass c produced O_c = softmax_c(S_c) V_c and
// lse_c = log2 sum_j 2^(s_ij) over its keys; then  O = sum_c 2^(lse_c - lse) O_c,  lse = log2 sum_c 2^lse_c.
// One warp per (token row, head), lane = 2 of the 64 head columns.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
attention_combine_kernel(const __nv_bfloat16* __restrict__ part_o, const float* __restrict__ part_lse, __nv_bfloat16* __restrict__ out,
                         float* __restrict__ lse, int B, int N, int Hq, int passes) {
    const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const long long M = (long long)B * N;
    if (w >= M * Hq) return;
    const int h = (int)(w % Hq);
    const long long m = w / Hq;
    const int b = (int)(m / N), n = (int)(m % N);
    const long long lse_idx = ((long long)b * Hq + h) * N + n, lse_pass = (long long)B * Hq * N;
    float mx = -INFINITY;
    for (int c = 0; c < passes; ++c) mx = fmaxf(mx, __ldg(part_lse + c * lse_pass + lse_idx));
    float den = 0.f;
    for (int c = 0; c < passes; ++c) den += exp2f(__ldg(part_lse + c * lse_pass + lse_idx) - mx);
    const float inv = 1.0f / den;
    float a0 = 0.f, a1 = 0.f;
    const long long o_idx = m * ((long long)Hq * ATT_HD) + (long long)h * ATT_HD + lane * 2, o_pass = M * Hq * ATT_HD;
    for (int c = 0; c < passes; ++c) {
        const float wgt = exp2f(__ldg(part_lse + c * lse_pass + lse_idx) - mx) * inv;
        const uint32_t v = *reinterpret_cast<const uint32_t*>(part_o + c * o_pass + o_idx);
        a0 = fmaf(__uint_as_float(v << 16), wgt, a0);
        a1 = fmaf(__uint_as_float(v & 0xffff0000u), wgt, a1);
    }
    *reinterpret_cast<uint32_t*>(out + o_idx) = pack_bf16(a0, a1);
    if (lse != nullptr && lane == 0) lse[lse_idx] = mx + log2f(den);
}

}  // namespace jat
