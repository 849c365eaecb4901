// GQA attention BACKWARD on tcgen05 / TMEM (autograd of GroupedQueryAttention.forward, jat_audiosr_v2.py:141-164,
// dropout = 0).  Inputs: the packed projections qkv (RoPE already applied to Q and K), the forward output O, its
// gradient dO and the per-row log-sum-exp kept by the forward kernel.  With P = softmax(Q K^T / 8):
//     dV = P^T dO      dP = dO V^T      dS = P * (dP - D) / 8,  D_i = sum_d dO_id O_id      dQ = dS K      dK = dS^T Q
// One CTA owns 128 KEYS of one (batch item, KV head) and walks over every query tile of the G query heads that
// share that KV head, so dK / dV accumulate in tensor memory for the whole kernel (no atomics), S^T and dP^T are
// recomputed per tile (nothing of size N x N is ever stored), and only dQ -- whose reduction runs over the key
// tiles, i.e. over CTAs -- leaves through TMA reduce-add into an f32 accumulator.
//
//   tensor memory:  S^T [0,128)   dP^T [128,256)   dV [256,320)   dK [320,384)   dQ [384,448)      (fp32 columns)
//   per (head, q-tile) iteration:
//     MMA  S^T  = K  Q^T        (M = keys 128, N = queries 128, K = 64; both operands K-major)
//     MMA  dP^T = V  dO^T
//     warps 4-7 (thread = key row): P^T = 2^(S^T log2e/8 - lse_q),  dS^T = P^T (dP^T - D_q) / 8  -> bf16 smem tiles
//     MMA  dV  += P^T  dO       (A K-major from smem; B = dO tile consumed MN-major as stored)
//     MMA  dK  += dS^T Q        (B = Q tile MN-major)
//     MMA  dQ   = dS   K        (A = the same dS^T tile read MN-major, B = K tile MN-major) -> TMA reduce-add
//   epilogue: dV -> bf16,  dK -> inverse RoPE -> bf16, written into the [M, (Hq+2Hkv)*64] gradient of qkv.
#pragma once
#include <cuda.h>
#include "attention_gqa.cuh"

namespace jat {

constexpr int ATTB_THREADS = 384;  // 4 service warps + 8 compute warps (two per SM sub-partition)
constexpr int ATTB_TILE = 128;
constexpr int ATTB_TILE_BYTES = ATTB_TILE * 128;  // [128 rows x 64 bf16]
constexpr int ATTB_SMEM_BYTES = 2 * ATTB_TILE_BYTES          // K, V
                                + 4 * ATTB_TILE_BYTES        // Q, dO double-buffered
                                + 4 * ATTB_TILE_BYTES        // P^T, dS^T ([128 x 128] bf16 = 2 blocks each)
                                + 2 * ATTB_TILE_BYTES        // dQ staging [128 x 64] f32 = 2 boxes
                                + 2 * 2 * ATTB_TILE * 4      // lse, D per query (double-buffered)
                                + 256 + 1024;

struct AttnBwdParams {
    int B, N, Hq, Hkv, G;
    const float* lse;   // [B, Hq, N] log2-domain
    const float* dsum;  // [B, Hq, N] D_i = rowsum(dO * O)
    __nv_bfloat16* dqkv;  // [B*N, (Hq + 2 Hkv) * 64]: this kernel writes the K and V column ranges
    const float* rope_cos;  // [max_pos, 64]
    const float* rope_sin;
    float scale_log2e, scale;
    DropCfg drop;  // the forward's dropout on the probabilities
    long long* trace;  // debug: clock64 timestamps of CTA (0,0,0) (v2 kernel; 256 slots), or NULL
};
#define ATTB_TRACE(slot)                                                                                   \
    do {                                                                                                   \
        if (p.trace && blockIdx.x == 0) p.trace[(slot)] = clock64(); \
    } while (0)

__device__ __forceinline__ void named_bar(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

__global__ void __launch_bounds__(ATTB_THREADS, 1)
gqa_attention_bwd_v1_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const __grid_constant__ CUtensorMap tmap_do,
                         const __grid_constant__ CUtensorMap tmap_dq, const AttnBwdParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
    uint8_t* sK = smem;
    uint8_t* sV = sK + ATTB_TILE_BYTES;
    uint8_t* sQ = sV + ATTB_TILE_BYTES;          // 2 stages
    uint8_t* sdO = sQ + 2 * ATTB_TILE_BYTES;     // 2 stages
    uint8_t* sPT = sdO + 2 * ATTB_TILE_BYTES;    // 2 blocks of 64 queries
    uint8_t* sdST = sPT + 2 * ATTB_TILE_BYTES;   // 2 blocks of 64 queries
    uint8_t* sdQ = sdST + 2 * ATTB_TILE_BYTES;   // 2 boxes of 32 f32 columns
    float* s_lse = reinterpret_cast<float*>(sdQ + 2 * ATTB_TILE_BYTES);  // [2][128]
    float* s_dsum = s_lse + 2 * ATTB_TILE;                                // [2][128]
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_dsum + 2 * ATTB_TILE);
    uint64_t* bar_kv = bars;
    uint64_t* bar_q_full = bars + 1;    // [2]
    uint64_t* bar_q_empty = bars + 3;   // [2]
    uint64_t* bar_sp_full = bars + 5;   // S^T, dP^T retired
    uint64_t* bar_pds_full = bars + 6;  // P^T, dS^T written (128 compute threads)
    uint64_t* bar_mma2_done = bars + 7; // dV, dK, dQ MMAs retired
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

    // warp-uniform values through a full-mask shuffle: keeps the MMA issue path on the uniform datapath (elect_one())
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const int kt = blockIdx.x, g = blockIdx.y, b = blockIdx.z;
    const int QT = (p.N + ATTB_TILE - 1) / ATTB_TILE;
    const int iters = p.G * QT;
    const int row_b = b * p.N;  // first token row of this batch item

    if (warp == 1 && lane == 0) {
        mbar_init(bar_kv, 1);
        for (int i = 0; i < 2; ++i) { mbar_init(&bar_q_full[i], 1); mbar_init(&bar_q_empty[i], 1); }
        mbar_init(bar_sp_full, 1);
        mbar_init(bar_pds_full, 256);
        mbar_init(bar_mma2_done, 1);
        fence_barrier_init();
    }
    if (warp == 2) {
        tmem_alloc<1>(tmem_slot, 512);
        tmem_relinquish<1>();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
    constexpr uint32_t COL_ST = 0, COL_DPT = 128, COL_DV = 256, COL_DK = 320, COL_DQ = 384;

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (lane == 0) {
            tma_prefetch_desc(&tmap_qkv);
            tma_prefetch_desc(&tmap_do);
            tma_prefetch_desc(&tmap_dq);
            mbar_expect_tx(bar_kv, 2 * ATTB_TILE_BYTES);
            tma_load_2d(sK, &tmap_qkv, bar_kv, (p.Hq + g) * ATT_HD, row_b + kt * ATTB_TILE);
            tma_load_2d(sV, &tmap_qkv, bar_kv, (p.Hq + p.Hkv + g) * ATT_HD, row_b + kt * ATTB_TILE);
            for (int it = 0; it < iters; ++it) {
                const int st = it & 1;
                mbar_wait(&bar_q_empty[st], (uint32_t)(((it >> 1) & 1) ^ 1));
                const int h = g * p.G + it / QT, qt = it % QT;
                mbar_expect_tx(&bar_q_full[st], 2 * ATTB_TILE_BYTES);
                tma_load_2d(sQ + st * ATTB_TILE_BYTES, &tmap_qkv, &bar_q_full[st], h * ATT_HD, row_b + qt * ATTB_TILE);
                tma_load_2d(sdO + st * ATTB_TILE_BYTES, &tmap_do, &bar_q_full[st], h * ATT_HD, row_b + qt * ATTB_TILE);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        // Warp-uniform control flow, one elected lane issues: the descriptors stay in uniform registers and the 32 MMAs
        // of an iteration go out back to back (no per-instruction R2UR waterfall as under `lane == 0`).
        {
            constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128, 0, 0);   // S^T, dP^T
            constexpr uint32_t idesc_kv = umma_idesc_bf16(128, 64, 0, 1);   // dV, dK: A K-major, B MN-major
            constexpr uint32_t idesc_dq = umma_idesc_bf16(128, 64, 1, 1);   // dQ: both MN-major
            const uint64_t k_desc = umma_smem_desc_sw128(smem_u32(sK));
            const uint64_t v_desc = umma_smem_desc_sw128(smem_u32(sV));
            const uint64_t k_desc_mn = umma_smem_desc_sw128_mn(smem_u32(sK), ATTB_TILE_BYTES);
            const uint64_t dst_desc_mn = umma_smem_desc_sw128_mn(smem_u32(sdST), ATTB_TILE_BYTES);
            const uint64_t pt_desc = umma_smem_desc_sw128(smem_u32(sPT));
            const uint64_t dst_desc = umma_smem_desc_sw128(smem_u32(sdST));
            auto issue_first = [&](int it) {  // S^T = K Q^T, dP^T = V dO^T
                const int st = it & 1;
                mbar_wait(&bar_q_full[st], (uint32_t)((it >> 1) & 1));
                tc_fence_after();
                const uint64_t q_desc = umma_smem_desc_sw128(smem_u32(sQ + st * ATTB_TILE_BYTES));
                const uint64_t do_desc = umma_smem_desc_sw128(smem_u32(sdO + st * ATTB_TILE_BYTES));
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_bf16_ss<1>(tmem_base + COL_ST, k_desc + 2 * k, q_desc + 2 * k, idesc_s, (uint32_t)(k != 0));
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_bf16_ss<1>(tmem_base + COL_DPT, v_desc + 2 * k, do_desc + 2 * k, idesc_s, (uint32_t)(k != 0));
                    umma_commit(bar_sp_full);
                }
                __syncwarp();
            };
            mbar_wait(bar_kv, 0);
            issue_first(0);
#pragma unroll 1
            for (int it = 0; it < iters; ++it) {
                const int st = it & 1;
                mbar_wait(bar_pds_full, (uint32_t)(it & 1));
                tc_fence_after();
                const uint64_t q_desc_mn = umma_smem_desc_sw128_mn(smem_u32(sQ + st * ATTB_TILE_BYTES), ATTB_TILE_BYTES);
                const uint64_t do_desc_mn = umma_smem_desc_sw128_mn(smem_u32(sdO + st * ATTB_TILE_BYTES), ATTB_TILE_BYTES);
                const uint32_t acc0 = (uint32_t)(it != 0);
                if (elect_one()) {
#pragma unroll
                    for (int ks = 0; ks < 8; ++ks) {  // reduction over the 128 queries of the tile, 16 per step
                        const uint64_t a_off = (uint64_t)((ks >> 2) * (ATTB_TILE_BYTES >> 4) + 2 * (ks & 3));
                        umma_bf16_ss<1>(tmem_base + COL_DV, pt_desc + a_off, do_desc_mn + (uint64_t)(ks * (2048 >> 4)), idesc_kv,
                                        ks == 0 ? acc0 : 1u);
                    }
#pragma unroll
                    for (int ks = 0; ks < 8; ++ks) {
                        const uint64_t a_off = (uint64_t)((ks >> 2) * (ATTB_TILE_BYTES >> 4) + 2 * (ks & 3));
                        umma_bf16_ss<1>(tmem_base + COL_DK, dst_desc + a_off, q_desc_mn + (uint64_t)(ks * (2048 >> 4)), idesc_kv,
                                        ks == 0 ? acc0 : 1u);
                    }
#pragma unroll
                    for (int ks = 0; ks < 8; ++ks)  // reduction over the 128 keys of this CTA
                        umma_bf16_ss<1>(tmem_base + COL_DQ, dst_desc_mn + (uint64_t)(ks * (2048 >> 4)),
                                        k_desc_mn + (uint64_t)(ks * (2048 >> 4)), idesc_dq, (uint32_t)(ks != 0));
                    umma_commit(bar_mma2_done);
                    umma_commit(&bar_q_empty[st]);
                }
                __syncwarp();
                if (it + 1 < iters) issue_first(it + 1);
            }
        }
        __syncwarp();
    } else if (warp >= 4) {
        // ------------------------------------------------------------------ compute warps: thread = key row
        // Two warps per TMEM lane quadrant (= per SM sub-partition): group 0 (warps 4-7) takes queries [0, 64) of a tile,
        // group 1 (warps 8-11) queries [64, 128); one warp's MUFU / TMEM-load latency hides behind the other's.
        const int grp = (warp - 4) >> 2;
        const int r = (warp & 3) * 32 + lane;
        const int t = threadIdx.x - 128;  // 0..255; t & 127 == r
        const uint32_t t_row = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
        const int key = kt * ATTB_TILE + r;
        const bool key_ok = key < p.N;
        const uint32_t rx = (uint32_t)(r & 7);
        const uint32_t row_off = (uint32_t)(r >> 3) * 1024u + (uint32_t)(r & 7) * 128u;
        const uint32_t sPT_row = smem_u32(sPT) + row_off, sdST_row = smem_u32(sdST) + row_off;
        const uint32_t sdQ_row = smem_u32(sdQ) + row_off;

        auto dq_epilogue = [&](int it) {  // dQ tile of iteration `it`: TMEM -> f32 smem boxes -> TMA reduce-add
            const int h = g * p.G + it / QT, qt = it % QT;
            if (t == 0) tma_store_wait_read<0>();  // the previous reduce-add has drained the staging tile
            named_bar(1, 256);
            {
                const int bx = grp;  // each group moves one 32-column box of the [128 x 64] dQ tile
                uint32_t v[32];
                tmem_ld_32x32(t_row + COL_DQ + bx * 32, v);
                tmem_ld_wait();
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const uint32_t addr = sdQ_row + (uint32_t)bx * ATTB_TILE_BYTES + ((((uint32_t)c) ^ rx) << 4);
                    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v[4 * c]), "r"(v[4 * c + 1]),
                                 "r"(v[4 * c + 2]), "r"(v[4 * c + 3]) : "memory");
                }
            }
            fence_proxy_async_smem();
            tc_fence_before();
            named_bar(1, 256);
            if (t == 0) {
                tma_reduce_add_2d(&tmap_dq, sdQ, h * ATT_HD, row_b + qt * ATTB_TILE);
                tma_reduce_add_2d(&tmap_dq, sdQ + ATTB_TILE_BYTES, h * ATT_HD + 32, row_b + qt * ATTB_TILE);
                tma_store_commit();
            }
        };

#pragma unroll 1
        for (int it = 0; it < iters; ++it) {
            const int h = g * p.G + it / QT, qt = it % QT;
            const int slot = it & 1;
            if (grp == 0) {  // per-query statistics of this tile (queries past the batch item's N tokens get lse = +inf -> P = 0)
                const int q = qt * ATTB_TILE + t;
                const bool q_ok = q < p.N;
                const long long idx = ((long long)b * p.Hq + h) * p.N + q;
                s_lse[slot * ATTB_TILE + t] = q_ok ? __ldg(p.lse + idx) : INFINITY;
                s_dsum[slot * ATTB_TILE + t] = q_ok ? __ldg(p.dsum + idx) : 0.0f;
            }
            if (it > 0) {
                mbar_wait(bar_mma2_done, (uint32_t)((it - 1) & 1));  // P^T / dS^T tiles free, dQ(it-1) complete
                tc_fence_after();
                dq_epilogue(it - 1);   // (its named barriers also publish s_lse / s_dsum)
            } else {
                named_bar(1, 256);
            }
            mbar_wait(bar_sp_full, (uint32_t)(it & 1));
            tc_fence_after();
            const float* lse_t = s_lse + slot * ATTB_TILE;
            const float* ds_t = s_dsum + slot * ATTB_TILE;
#pragma unroll 1
            for (int c = 2 * grp; c < 2 * grp + 2; ++c) {  // 32 queries at a time
                uint32_t sv[32], dv[32];
                tmem_ld_32x32(t_row + COL_ST + c * 32, sv);
                tmem_ld_32x32(t_row + COL_DPT + c * 32, dv);
                tmem_ld_wait();
                uint32_t pp[16], dd[16];
#pragma unroll
                for (int j = 0; j < 32; j += 2) {
                    float p0 = 0.f, p1 = 0.f;
                    if (key_ok) {
                        p0 = ex2_approx(fmaf(__uint_as_float(sv[j]), p.scale_log2e, -lse_t[c * 32 + j]));
                        p1 = ex2_approx(fmaf(__uint_as_float(sv[j + 1]), p.scale_log2e, -lse_t[c * 32 + j + 1]));
                    }
                    float m0 = 1.0f, m1 = 1.0f;
                    if (p.drop.thresh != 0u) {  // forward: O = (P * mask / keep) V   ->   dV uses P * m, dP = (dO V^T) * m
                        const uint32_t rq = (uint32_t)(((long long)b * p.Hq + h) * p.N + qt * ATTB_TILE + c * 32 + j);
                        m0 = drop_scale(p.drop, rq, (uint32_t)key);
                        m1 = drop_scale(p.drop, rq + 1u, (uint32_t)key);
                    }
                    const float d0 = p0 * (__uint_as_float(dv[j]) * m0 - ds_t[c * 32 + j]) * p.scale;
                    const float d1 = p1 * (__uint_as_float(dv[j + 1]) * m1 - ds_t[c * 32 + j + 1]) * p.scale;
                    p0 *= m0;
                    p1 *= m1;
                    pp[j >> 1] = pack_bf16(p0, p1);
                    dd[j >> 1] = pack_bf16(d0, d1);
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) {  // 32 queries = 64 bytes = 4 x 16 B chunks of block (c >> 1)
                    const uint32_t off = (uint32_t)(c >> 1) * ATTB_TILE_BYTES + (((uint32_t)((c & 1) * 4 + i) ^ rx) << 4);
                    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(sPT_row + off), "r"(pp[4 * i]), "r"(pp[4 * i + 1]),
                                 "r"(pp[4 * i + 2]), "r"(pp[4 * i + 3]) : "memory");
                    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(sdST_row + off), "r"(dd[4 * i]), "r"(dd[4 * i + 1]),
                                 "r"(dd[4 * i + 2]), "r"(dd[4 * i + 3]) : "memory");
                }
            }
            fence_proxy_async_smem();
            tc_fence_before();
            mbar_arrive(bar_pds_full);
        }
        mbar_wait(bar_mma2_done, (uint32_t)((iters - 1) & 1));
        tc_fence_after();
        dq_epilogue(iters - 1);
        // ---- dV, dK of this key row (dK through the transpose of the RoPE rotation, jat_audiosr_v2.py:70-91)
        {   // (tcgen05.ld is warp-collective: every lane loads, only rows that hold a real key store)
            __nv_bfloat16* orow = p.dqkv + (long long)(row_b + (key_ok ? key : 0)) * ((p.Hq + 2 * p.Hkv) * ATT_HD);
            uint32_t lo[32], hi[32];
            if (grp == 0) {  // group 0 writes dV, group 1 dK
            tmem_ld_32x32(t_row + COL_DV, lo);
            tmem_ld_32x32(t_row + COL_DV + 32, hi);
            tmem_ld_wait();
            __nv_bfloat16* ov = orow + (p.Hq + p.Hkv + g) * ATT_HD;
            if (key_ok) {
#pragma unroll
                for (int j = 0; j < 32; j += 8) {
                    *reinterpret_cast<uint4*>(ov + j) = make_uint4(
                        pack_bf16(__uint_as_float(lo[j]), __uint_as_float(lo[j + 1])), pack_bf16(__uint_as_float(lo[j + 2]), __uint_as_float(lo[j + 3])),
                        pack_bf16(__uint_as_float(lo[j + 4]), __uint_as_float(lo[j + 5])), pack_bf16(__uint_as_float(lo[j + 6]), __uint_as_float(lo[j + 7])));
                    *reinterpret_cast<uint4*>(ov + 32 + j) = make_uint4(
                        pack_bf16(__uint_as_float(hi[j]), __uint_as_float(hi[j + 1])), pack_bf16(__uint_as_float(hi[j + 2]), __uint_as_float(hi[j + 3])),
                        pack_bf16(__uint_as_float(hi[j + 4]), __uint_as_float(hi[j + 5])), pack_bf16(__uint_as_float(hi[j + 6]), __uint_as_float(hi[j + 7])));
                }
            }
            } else {
            tmem_ld_32x32(t_row + COL_DK, lo);
            tmem_ld_32x32(t_row + COL_DK + 32, hi);
            tmem_ld_wait();
            if (key_ok) {
                const float* cosr = p.rope_cos + (long long)key * 64;
                const float* sinr = p.rope_sin + (long long)key * 64;
                __nv_bfloat16* ok = orow + (p.Hq + g) * ATT_HD;
                float ra[32], rb[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) {  // forward: lo' = a c - b s, hi' = b c + a s  ->  da = glo c + ghi s, db = ghi c - glo s
                    const float c = __ldg(cosr + j), s = __ldg(sinr + j);
                    const float glo = __uint_as_float(lo[j]), ghi = __uint_as_float(hi[j]);
                    ra[j] = glo * c + ghi * s;
                    rb[j] = ghi * c - glo * s;
                }
#pragma unroll
                for (int j = 0; j < 32; j += 8) {
                    *reinterpret_cast<uint4*>(ok + j) = make_uint4(pack_bf16(ra[j], ra[j + 1]), pack_bf16(ra[j + 2], ra[j + 3]),
                                                                   pack_bf16(ra[j + 4], ra[j + 5]), pack_bf16(ra[j + 6], ra[j + 7]));
                    *reinterpret_cast<uint4*>(ok + 32 + j) = make_uint4(pack_bf16(rb[j], rb[j + 1]), pack_bf16(rb[j + 2], rb[j + 3]),
                                                                        pack_bf16(rb[j + 4], rb[j + 5]), pack_bf16(rb[j + 6], rb[j + 7]));
                }
            }
            }
        }
        if (t == 0) tma_store_wait_all<0>();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc<1>(tmem_base, 512);
    }
}


// ------------------------------------------------------------------------------------------------
// Version 2 (default): the same key-stationary schedule, re-organised around the instruction-issue bound of the
// softmax-backward math (v1: ~44 SASS instructions per score element at 14 % tensor-pipe activity):
//   * score tiles are computed as S = Q K^T / dP = dO V^T (M = queries): a compute thread owns a QUERY row, so the row's
//     log-sum-exp, D and dropout-mask row are per-thread registers (v1: two shared-memory reads per element) and one mask
//     hash serves two adjacent keys (v1: one hash per element);
//   * a thread drains its 64 S and 64 dP values into registers in one burst and hands the accumulators back at once: the
//     S / dP MMAs of iteration i+1 run under the math of iteration i, and the dV / dK / dQ MMAs of iteration i under the math
//     of iteration i+1 (dQ double-buffered in tensor memory); only the shared-memory stores of P / dS wait for them;
//   * 1/sqrt(d) is folded into the finalize pass (dQ) and the dK epilogue -- exact, a power of two;
//   * fully padded 32-key chunks / 32-query warps skip the math.
//   tensor memory:  S [0,128)  dP [128,256)  dV [256,320)  dK [320,384)  dQ_0 [384,448)  dQ_1 [448,512)
//   MMAs per iteration (operands as stored, no transposed copies):
//     S   = Q  K^T       A = Q tile K-major,               B = K tile K-major
//     dP  = dO V^T       A = dO tile K-major,              B = V tile K-major
//     dV += P^T  dO      A = P tile [query][key] MN-major, B = dO tile MN-major
//     dK += dS^T Q       A = dS tile MN-major,             B = Q tile MN-major
//     dQ  = dS   K       A = dS tile K-major,              B = K tile MN-major     -> TMA reduce-add (f32)
// ------------------------------------------------------------------------------------------------
// softmax-backward math of one 32-key chunk of a query row: P = 2^(s log2e/8 - lse), dS = P (dP m - D) [1/8 applied later],
// packed to bf16 pairs.  nv = number of real keys in the chunk (>= 32: all), warp-uniform.
template <bool DROP>
__device__ __forceinline__ void attb_chunk_math(uint32_t (&sv)[32], const uint32_t (&dv)[32], uint32_t* pp, uint32_t* dd,
                                                float sl2, float nl, float dsum, int nv, const DropCfg& drop, uint32_t rq,
                                                uint32_t key) {
    if (nv <= 0) {
#pragma unroll
        for (int j = 0; j < 16; ++j) { pp[j] = 0u; dd[j] = 0u; }
        return;
    }
    if (nv < 32) {  // the chunk that straddles the sequence end: score -inf -> P = 0 and dS = 0 * (finite) = 0
#pragma unroll
        for (int j = 0; j < 32; ++j)
            if (j >= nv) sv[j] = 0xff800000u;
    }
#pragma unroll
    for (int j = 0; j < 32; j += 2) {
        float p0 = ex2_approx(fmaf(__uint_as_float(sv[j]), sl2, nl));
        float p1 = ex2_approx(fmaf(__uint_as_float(sv[j + 1]), sl2, nl));
        float d0, d1;
        if constexpr (DROP) {  // forward: O = (P * mask / keep) V  ->  dV uses P * m, dP = (dO V^T) * m
            float m0, m1;
            drop_scale2(drop, rq, key + (uint32_t)j, m0, m1);
            d0 = p0 * fmaf(__uint_as_float(dv[j]), m0, -dsum);
            d1 = p1 * fmaf(__uint_as_float(dv[j + 1]), m1, -dsum);
            p0 *= m0;
            p1 *= m1;
        } else {
            d0 = p0 * (__uint_as_float(dv[j]) - dsum);
            d1 = p1 * (__uint_as_float(dv[j + 1]) - dsum);
        }
        pp[j >> 1] = pack_bf16(p0, p1);
        dd[j >> 1] = pack_bf16(d0, d1);
    }
}

constexpr int ATTB2_THREADS = 512;  // 4 service warps + 8 compute warps + 4 dQ read-out warps
constexpr int ATTB2_SMEM_BYTES = 2 * ATTB_TILE_BYTES      // K, V
                                 + 4 * ATTB_TILE_BYTES    // Q, dO double-buffered
                                 + 4 * ATTB_TILE_BYTES    // P, dS ([128 queries x 128 keys] bf16 = 2 key blocks each)
                                 + 2 * ATTB_TILE_BYTES    // dQ staging [128 x 64] f32 = 2 boxes
                                 + 256 + 1024;

template <bool DROP>
__global__ void __launch_bounds__(ATTB2_THREADS, 1)
gqa_attention_bwd_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const __grid_constant__ CUtensorMap tmap_do,
                         const __grid_constant__ CUtensorMap tmap_dq, const AttnBwdParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
    uint8_t* sK = smem;
    uint8_t* sV = sK + ATTB_TILE_BYTES;
    uint8_t* sQ = sV + ATTB_TILE_BYTES;          // 2 stages
    uint8_t* sdO = sQ + 2 * ATTB_TILE_BYTES;     // 2 stages
    uint8_t* sP = sdO + 2 * ATTB_TILE_BYTES;     // 2 blocks of 64 keys
    uint8_t* sdS = sP + 2 * ATTB_TILE_BYTES;     // 2 blocks of 64 keys
    uint8_t* sdQ = sdS + 2 * ATTB_TILE_BYTES;    // 2 boxes of 32 f32 columns
    uint64_t* bars = reinterpret_cast<uint64_t*>(sdQ + 2 * ATTB_TILE_BYTES);
    uint64_t* bar_kv = bars;
    uint64_t* bar_q_full = bars + 1;    // [2] Q / dO stage landed
    uint64_t* bar_q_empty = bars + 3;   // [2] the dV / dK / dQ MMAs that read the stage have retired
    uint64_t* bar_sp_full = bars + 5;   // S, dP retired
    uint64_t* bar_sp_free = bars + 6;   // S, dP copied to registers by the 256 compute threads
    uint64_t* bar_pds_full = bars + 7;  // P, dS written (256 compute threads)
    uint64_t* bar_mma2_done = bars + 8; // dV, dK, dQ MMAs retired: P / dS tiles free, dQ complete
    uint64_t* bar_dq_free = bars + 9;   // [2] dQ buffer copied to registers by the 128 read-out threads
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 11);

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    // 1-D grid, heaviest work first: every full 128-key tile of every (KV head, batch item) precedes the partial last key
    // tiles (N = 345: 89 of 128 keys, ~70 % of the work), so the short CTAs fill the last, partial wave of the launch
    const int QT = (p.N + ATTB_TILE - 1) / ATTB_TILE;
    int kt, g, b;
    {
        const int kt_full = p.N / ATTB_TILE, per = p.Hkv * p.B, lin = (int)blockIdx.x;
        int rest;
        if (lin < kt_full * per) { kt = lin % kt_full; rest = lin / kt_full; }
        else { const int l2 = lin - kt_full * per, kp = QT - kt_full; kt = kt_full + l2 % kp; rest = l2 / kp; }
        g = rest % p.Hkv;
        b = rest / p.Hkv;
    }
    const int iters = p.G * QT;
    const int row_b = b * p.N;  // first token row of this batch item

    if (warp == 1 && lane == 0) {
        mbar_init(bar_kv, 1);
        for (int i = 0; i < 2; ++i) { mbar_init(&bar_q_full[i], 1); mbar_init(&bar_q_empty[i], 1); }
        mbar_init(bar_sp_full, 1);
        mbar_init(bar_sp_free, 256);
        mbar_init(bar_pds_full, 256);
        mbar_init(bar_mma2_done, 1);
        mbar_init(&bar_dq_free[0], 128);
        mbar_init(&bar_dq_free[1], 128);
        fence_barrier_init();
    }
    if (warp == 2) {
        tmem_alloc<1>(tmem_slot, 512);
        tmem_relinquish<1>();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
    constexpr uint32_t COL_S = 0, COL_DP = 128, COL_DV = 256, COL_DK = 320, COL_DQ = 384;

    if (warp < 4) {
    // service warpgroup (TMA producer, MMA issuer, two idle warps): hands registers to the compute warpgroups
    asm volatile("setmaxnreg.dec.sync.aligned.u32 72;");
    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (lane == 0) {
            tma_prefetch_desc(&tmap_qkv);
            tma_prefetch_desc(&tmap_do);
            tma_prefetch_desc(&tmap_dq);
            mbar_expect_tx(bar_kv, 2 * ATTB_TILE_BYTES);
            tma_load_2d(sK, &tmap_qkv, bar_kv, (p.Hq + g) * ATT_HD, row_b + kt * ATTB_TILE);
            tma_load_2d(sV, &tmap_qkv, bar_kv, (p.Hq + p.Hkv + g) * ATT_HD, row_b + kt * ATTB_TILE);
            for (int it = 0; it < iters; ++it) {
                const int st = it & 1;
                mbar_wait(&bar_q_empty[st], (uint32_t)(((it >> 1) & 1) ^ 1));
                const int h = g * p.G + it / QT, qt = it % QT;
                mbar_expect_tx(&bar_q_full[st], 2 * ATTB_TILE_BYTES);
                tma_load_2d(sQ + st * ATTB_TILE_BYTES, &tmap_qkv, &bar_q_full[st], h * ATT_HD, row_b + qt * ATTB_TILE);
                tma_load_2d(sdO + st * ATTB_TILE_BYTES, &tmap_do, &bar_q_full[st], h * ATT_HD, row_b + qt * ATTB_TILE);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer (warp-uniform control, one elected lane)
        constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128, 0, 0);   // S, dP
        constexpr uint32_t idesc_kv = umma_idesc_bf16(128, 64, 1, 1);   // dV, dK: A and B MN-major
        constexpr uint32_t idesc_dq = umma_idesc_bf16(128, 64, 0, 1);   // dQ: A K-major, B MN-major
        const uint64_t k_desc = umma_smem_desc_sw128(smem_u32(sK));
        const uint64_t v_desc = umma_smem_desc_sw128(smem_u32(sV));
        const uint64_t k_desc_mn = umma_smem_desc_sw128_mn(smem_u32(sK), ATTB_TILE_BYTES);
        const uint64_t p_desc_mn = umma_smem_desc_sw128_mn(smem_u32(sP), ATTB_TILE_BYTES);
        const uint64_t ds_desc_mn = umma_smem_desc_sw128_mn(smem_u32(sdS), ATTB_TILE_BYTES);
        const uint64_t ds_desc = umma_smem_desc_sw128(smem_u32(sdS));
        auto issue_scores = [&](int it) {  // S = Q K^T, dP = dO V^T
            const int st = it & 1;
            mbar_wait(&bar_q_full[st], (uint32_t)((it >> 1) & 1));
            tc_fence_after();
            const uint64_t q_desc = umma_smem_desc_sw128(smem_u32(sQ + st * ATTB_TILE_BYTES));
            const uint64_t do_desc = umma_smem_desc_sw128(smem_u32(sdO + st * ATTB_TILE_BYTES));
            if (elect_one()) {
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_bf16_ss<1>(tmem_base + COL_S, q_desc + 2 * k, k_desc + 2 * k, idesc_s, (uint32_t)(k != 0));
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_bf16_ss<1>(tmem_base + COL_DP, do_desc + 2 * k, v_desc + 2 * k, idesc_s, (uint32_t)(k != 0));
                umma_commit(bar_sp_full);
            }
            __syncwarp();
        };
        mbar_wait(bar_kv, 0);
        issue_scores(0);
#pragma unroll 1
        for (int it = 0; it < iters; ++it) {
            const int st = it & 1;
            if (it + 1 < iters) {  // the next score tiles as soon as this iteration's are in registers
                mbar_wait(bar_sp_free, (uint32_t)(it & 1));
                tc_fence_after();
                if (lane == 0 && it < 16) ATTB_TRACE(128 + it * 4 + 0);
                issue_scores(it + 1);
                if (lane == 0 && it < 16) ATTB_TRACE(128 + it * 4 + 1);
            }
            mbar_wait(bar_pds_full, (uint32_t)(it & 1));
            if (it >= 2) mbar_wait(&bar_dq_free[st], (uint32_t)(((it - 2) >> 1) & 1));  // dQ buffer `st` was read out (iteration it - 2)
            tc_fence_after();
            if (lane == 0 && it < 16) ATTB_TRACE(128 + it * 4 + 2);
            const uint64_t q_desc_mn = umma_smem_desc_sw128_mn(smem_u32(sQ + st * ATTB_TILE_BYTES), ATTB_TILE_BYTES);
            const uint64_t do_desc_mn = umma_smem_desc_sw128_mn(smem_u32(sdO + st * ATTB_TILE_BYTES), ATTB_TILE_BYTES);
            const uint32_t acc0 = (uint32_t)(it != 0);
            const uint32_t dq_col = tmem_base + COL_DQ + (uint32_t)(st * 64);
            if (elect_one()) {
#pragma unroll
                for (int ks = 0; ks < 8; ++ks)  // reduction over the 128 queries of the tile, 16 per step
                    umma_bf16_ss<1>(tmem_base + COL_DV, p_desc_mn + (uint64_t)(ks * (2048 >> 4)), do_desc_mn + (uint64_t)(ks * (2048 >> 4)),
                                    idesc_kv, ks == 0 ? acc0 : 1u);
#pragma unroll
                for (int ks = 0; ks < 8; ++ks)
                    umma_bf16_ss<1>(tmem_base + COL_DK, ds_desc_mn + (uint64_t)(ks * (2048 >> 4)), q_desc_mn + (uint64_t)(ks * (2048 >> 4)),
                                    idesc_kv, ks == 0 ? acc0 : 1u);
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) {  // reduction over the 128 keys of this CTA: 2 key blocks x 4 steps of 16
                    const uint64_t a_off = (uint64_t)((ks >> 2) * (ATTB_TILE_BYTES >> 4) + 2 * (ks & 3));
                    umma_bf16_ss<1>(dq_col, ds_desc + a_off, k_desc_mn + (uint64_t)(ks * (2048 >> 4)), idesc_dq, (uint32_t)(ks != 0));
                }
                umma_commit(bar_mma2_done);
                umma_commit(&bar_q_empty[st]);
            }
            __syncwarp();
            if (lane == 0 && it < 16) ATTB_TRACE(128 + it * 4 + 3);
        }
        __syncwarp();
    }
    } else if (warp < 12) {
        // ------------------------------------------------------------------ compute warps: thread = QUERY row
        asm volatile("setmaxnreg.inc.sync.aligned.u32 184;");
        // Two warps per TMEM lane quadrant: group 0 (warps 4-7) takes keys [0, 64) of the tile, group 1 (warps 8-11) keys [64, 128).
        const int grp = (warp - 4) >> 2;
        const int r = (warp & 3) * 32 + lane;
        const int t = threadIdx.x - 128;  // 0..255; t & 127 == r
        const uint32_t t_row = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
        const uint32_t rx = (uint32_t)(r & 7);
        const uint32_t row_off = (uint32_t)(r >> 3) * 1024u + (uint32_t)(r & 7) * 128u;
        const uint32_t sP_row = smem_u32(sP) + (uint32_t)grp * ATTB_TILE_BYTES + row_off;
        const uint32_t sdS_row = smem_u32(sdS) + (uint32_t)grp * ATTB_TILE_BYTES + row_off;
        const int key0 = kt * ATTB_TILE + grp * 64;   // first key of this thread's 64 columns
        const int nvalid = p.N - key0;                 // columns [0, nvalid) hold real keys (may be <= 0 or >= 64)
        const float sl2 = p.scale_log2e;

        // per-query statistics of iteration `it` (queries past the batch item's N tokens: lse = +inf -> P = 0, dS = 0)
        auto load_stats = [&](int it, float& nl, float& dsum) {
            const int h = g * p.G + it / QT, q = (it % QT) * ATTB_TILE + r;
            const bool q_ok = q < p.N;
            const long long idx = ((long long)b * p.Hq + h) * p.N + (q_ok ? q : 0);
            // (volatile: issued HERE, a whole math phase before the values are used -- the compiler otherwise sinks the loads
            //  to their use and the warp waits out the full global-memory latency)
            float l = INFINITY, d = 0.0f;
            if (q_ok) {
                asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(l) : "l"(p.lse + idx));
                asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(d) : "l"(p.dsum + idx));
            }
            nl = -l;
            dsum = d;
        };
        float nl, dsum;
        load_stats(0, nl, dsum);
        // Software pipeline over the thread's two 32-key chunks: the TMEM read-out of chunk 1 runs under the math of chunk 0,
        // the accumulators go back to the MMA issuer as soon as chunk 1 is in registers (the next S / dP MMAs then run under
        // the math of chunk 1), and chunk 0 of the NEXT iteration is requested before this iteration's dQ epilogue.
        uint32_t s0[32], d0[32], s1[32], d1[32];
        const uint32_t t_s = t_row + COL_S + (uint32_t)(grp * 64), t_dp = t_row + COL_DP + (uint32_t)(grp * 64);
        mbar_wait(bar_sp_full, 0u);
        tc_fence_after();
        tmem_ld_32x32(t_s, s0);
        tmem_ld_32x32(t_dp, d0);

#pragma unroll 1
        for (int it = 0; it < iters; ++it) {
            const int h = g * p.G + it / QT, qt = it % QT;
            // rows of a warp that lie wholly past the sequence end need no math (warp-uniform)
            const bool warp_live = qt * ATTB_TILE + (warp & 3) * 32 < p.N;
            const uint32_t rq = (uint32_t)(((long long)b * p.Hq + h) * p.N + qt * ATTB_TILE + r);  // dropout-mask row
            const float my_nl = nl, my_d = dsum;
            uint32_t pp[32], dd[32];
            const bool tr = t == 0 && it < 16;
            if (tr) ATTB_TRACE(it * 8 + 0);
            tmem_ld_wait_dep32(s0);
            tmem_ld_wait_dep32(d0);
            if (tr) ATTB_TRACE(it * 8 + 1);
            tmem_ld_32x32(t_s + 32u, s1);
            tmem_ld_32x32(t_dp + 32u, d1);
            attb_chunk_math<DROP>(s0, d0, &pp[0], &dd[0], sl2, my_nl, my_d, warp_live ? nvalid : 0, p.drop, rq, (uint32_t)key0);
            if (tr) ATTB_TRACE(it * 8 + 2);
            // ---- the P / dS tiles are free once the dV / dK / dQ MMAs of the previous iteration have retired (they ran under
            //      the math above); chunk 0 goes to shared memory now, its stores overlap the math of chunk 1
            if (it > 0) {
                mbar_wait(bar_mma2_done, (uint32_t)((it - 1) & 1));
                tc_fence_after();
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const uint32_t off = (((uint32_t)i) ^ rx) << 4;
                asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(sP_row + off), "r"(pp[4 * i]), "r"(pp[4 * i + 1]),
                             "r"(pp[4 * i + 2]), "r"(pp[4 * i + 3]) : "memory");
                asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(sdS_row + off), "r"(dd[4 * i]), "r"(dd[4 * i + 1]),
                             "r"(dd[4 * i + 2]), "r"(dd[4 * i + 3]) : "memory");
            }
            tmem_ld_wait_dep32(s1);
            tmem_ld_wait_dep32(d1);
            tc_fence_before();
            mbar_arrive(bar_sp_free);                            // S / dP are in registers: the next score MMAs may overwrite them
            if (it + 1 < iters) load_stats(it + 1, nl, dsum);   // prefetch under the math below
            attb_chunk_math<DROP>(s1, d1, &pp[16], &dd[16], sl2, my_nl, my_d, warp_live ? nvalid - 32 : 0, p.drop, rq,
                                  (uint32_t)(key0 + 32));
            // ---- the P / dS tiles are free once the dV / dK / dQ MMAs of the previous iteration have retired
            if (tr) ATTB_TRACE(it * 8 + 3);
            if (tr) ATTB_TRACE(it * 8 + 4);
#pragma unroll
            for (int i = 4; i < 8; ++i) {  // 64 keys = 128 bytes = 8 x 16 B chunks of this thread's row in key block `grp`
                const uint32_t off = (((uint32_t)i) ^ rx) << 4;
                asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(sP_row + off), "r"(pp[4 * i]), "r"(pp[4 * i + 1]),
                             "r"(pp[4 * i + 2]), "r"(pp[4 * i + 3]) : "memory");
                asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(sdS_row + off), "r"(dd[4 * i]), "r"(dd[4 * i + 1]),
                             "r"(dd[4 * i + 2]), "r"(dd[4 * i + 3]) : "memory");
            }
            fence_proxy_async_smem();
            tc_fence_before();
            mbar_arrive(bar_pds_full);
            if (tr) ATTB_TRACE(it * 8 + 5);
            if (it + 1 < iters) {  // chunk 0 of the next iteration (its score MMAs ran under the math above)
                mbar_wait(bar_sp_full, (uint32_t)((it + 1) & 1));
                tc_fence_after();
                tmem_ld_32x32(t_s, s0);
                tmem_ld_32x32(t_dp, d0);
            }
            if (tr) ATTB_TRACE(it * 8 + 6);
            if (tr) ATTB_TRACE(it * 8 + 7);
        }
        mbar_wait(bar_mma2_done, (uint32_t)((iters - 1) & 1));
        tc_fence_after();
        // ---- dV, dK of key row r (TMEM lane = key): dK through the transpose of the RoPE rotation (jat_audiosr_v2.py:70-91)
        {   // (tcgen05.ld is warp-collective: every lane loads, only rows that hold a real key store)
            const int key = kt * ATTB_TILE + r;
            const bool key_ok = key < p.N;
            __nv_bfloat16* orow = p.dqkv + (long long)(row_b + (key_ok ? key : 0)) * ((p.Hq + 2 * p.Hkv) * ATT_HD);
            uint32_t lo[32], hi[32];
            if (grp == 0) {  // group 0 writes dV, group 1 dK
                tmem_ld_32x32(t_row + COL_DV, lo);
                tmem_ld_32x32(t_row + COL_DV + 32, hi);
                tmem_ld_wait();
                __nv_bfloat16* ov = orow + (p.Hq + p.Hkv + g) * ATT_HD;
                if (key_ok) {
#pragma unroll
                    for (int j = 0; j < 32; j += 8) {
                        *reinterpret_cast<uint4*>(ov + j) = make_uint4(
                            pack_bf16(__uint_as_float(lo[j]), __uint_as_float(lo[j + 1])), pack_bf16(__uint_as_float(lo[j + 2]), __uint_as_float(lo[j + 3])),
                            pack_bf16(__uint_as_float(lo[j + 4]), __uint_as_float(lo[j + 5])), pack_bf16(__uint_as_float(lo[j + 6]), __uint_as_float(lo[j + 7])));
                        *reinterpret_cast<uint4*>(ov + 32 + j) = make_uint4(
                            pack_bf16(__uint_as_float(hi[j]), __uint_as_float(hi[j + 1])), pack_bf16(__uint_as_float(hi[j + 2]), __uint_as_float(hi[j + 3])),
                            pack_bf16(__uint_as_float(hi[j + 4]), __uint_as_float(hi[j + 5])), pack_bf16(__uint_as_float(hi[j + 6]), __uint_as_float(hi[j + 7])));
                    }
                }
            } else {
                tmem_ld_32x32(t_row + COL_DK, lo);
                tmem_ld_32x32(t_row + COL_DK + 32, hi);
                tmem_ld_wait();
                if (key_ok) {
                    const float* cosr = p.rope_cos + (long long)key * 64;
                    const float* sinr = p.rope_sin + (long long)key * 64;
                    __nv_bfloat16* ok = orow + (p.Hq + g) * ATT_HD;
                    float ra[32], rb[32];
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {  // forward: lo' = a c - b s, hi' = b c + a s  ->  da = glo c + ghi s, db = ghi c - glo s
                        const float4 c4 = __ldg(reinterpret_cast<const float4*>(cosr + j)), s4 = __ldg(reinterpret_cast<const float4*>(sinr + j));
                        const float cc[4] = {c4.x, c4.y, c4.z, c4.w}, ss[4] = {s4.x, s4.y, s4.z, s4.w};
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const float c = cc[q] * p.scale, s = ss[q] * p.scale;   // (1/sqrt(d) of dS, see above)
                            const float glo = __uint_as_float(lo[j + q]), ghi = __uint_as_float(hi[j + q]);
                            ra[j + q] = glo * c + ghi * s;
                            rb[j + q] = ghi * c - glo * s;
                        }
                    }
#pragma unroll
                    for (int j = 0; j < 32; j += 8) {
                        *reinterpret_cast<uint4*>(ok + j) = make_uint4(pack_bf16(ra[j], ra[j + 1]), pack_bf16(ra[j + 2], ra[j + 3]),
                                                                       pack_bf16(ra[j + 4], ra[j + 5]), pack_bf16(ra[j + 6], ra[j + 7]));
                        *reinterpret_cast<uint4*>(ok + 32 + j) = make_uint4(pack_bf16(rb[j], rb[j + 1]), pack_bf16(rb[j + 2], rb[j + 3]),
                                                                            pack_bf16(rb[j + 4], rb[j + 5]), pack_bf16(rb[j + 6], rb[j + 7]));
                    }
                }
            }
        }
    } else {
        // ------------------------------------------------------------------ dQ read-out warpgroup (warps 12-15 = TMEM lane quadrants 0-3)
        // dQ tile of every iteration: TMEM -> registers (buffer handed back at once) -> f32 smem boxes -> TMA reduce-add into the
        // f32 dQ accumulator.  Off the compute warps' critical path: their math never waits for this read-out.
        asm volatile("setmaxnreg.dec.sync.aligned.u32 72;");
        const int r = (warp & 3) * 32 + lane;
        const int t = threadIdx.x - 384;  // 0..127 == r
        const uint32_t t_row = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
        const uint32_t rx = (uint32_t)(r & 7);
        const uint32_t sdQ_row = smem_u32(sdQ) + (uint32_t)(r >> 3) * 1024u + (uint32_t)(r & 7) * 128u;
#pragma unroll 1
        for (int it = 0; it < iters; ++it) {
            const int h = g * p.G + it / QT, qt = it % QT;
            mbar_wait_backoff(bar_mma2_done, (uint32_t)(it & 1));
            tc_fence_after();
            uint32_t v0[32], v1[32];
            tmem_ld_32x32(t_row + COL_DQ + (uint32_t)((it & 1) * 64), v0);
            tmem_ld_32x32(t_row + COL_DQ + (uint32_t)((it & 1) * 64 + 32), v1);
            tmem_ld_wait_dep32(v0);
            tmem_ld_wait_dep32(v1);
            tc_fence_before();
            mbar_arrive(&bar_dq_free[it & 1]);
            if (t == 0) tma_store_wait_read<0>();  // the previous reduce-add has drained the staging tile
            named_bar(1, 128);
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const uint32_t addr = sdQ_row + ((((uint32_t)c) ^ rx) << 4);
                asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v0[4 * c]), "r"(v0[4 * c + 1]),
                             "r"(v0[4 * c + 2]), "r"(v0[4 * c + 3]) : "memory");
                asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr + (uint32_t)ATTB_TILE_BYTES), "r"(v1[4 * c]), "r"(v1[4 * c + 1]),
                             "r"(v1[4 * c + 2]), "r"(v1[4 * c + 3]) : "memory");
            }
            fence_proxy_async_smem();
            named_bar(1, 128);
            if (t == 0) {
                tma_reduce_add_2d(&tmap_dq, sdQ, h * ATT_HD, row_b + qt * ATTB_TILE);
                tma_reduce_add_2d(&tmap_dq, sdQ + ATTB_TILE_BYTES, h * ATT_HD + 32, row_b + qt * ATTB_TILE);
                tma_store_commit();
            }
        }
        if (t == 0) tma_store_wait_all<0>();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc<1>(tmem_base, 512);
    }
}

// D[b, h, n] = sum_d dO[m, h*64 + d] * O[m, h*64 + d]   (8 lanes per (token row, head): one 16-byte load each, so a warp
// reads 4 x 128 contiguous bytes of both operands; 3 shuffles finish the dot product)
__global__ void attn_bwd_rowdot_kernel(const __nv_bfloat16* __restrict__ dO, const __nv_bfloat16* __restrict__ O,
                                       float* __restrict__ dsum, int B, int N, int Hq) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long i = t >> 3;  // (row, head); each group of 8 lanes is in or out of range together
    const int k = (int)(t & 7);
    const bool ok = i < (long long)B * N * Hq;
    float s = 0.f;
    int h = 0;
    long long m = 0;
    if (ok) {
        h = (int)(i % Hq);
        m = i / Hq;
        const uint4 x = __ldcs(reinterpret_cast<const uint4*>(dO + m * (Hq * 64) + h * 64) + k);
        const uint4 y = __ldg(reinterpret_cast<const uint4*>(O + m * (Hq * 64) + h * 64) + k);
        const uint32_t xs[4] = {x.x, x.y, x.z, x.w}, ys[4] = {y.x, y.y, y.z, y.w};
#pragma unroll
        for (int q = 0; q < 4; ++q)
            s += __uint_as_float(xs[q] << 16) * __uint_as_float(ys[q] << 16) +
                 __uint_as_float(xs[q] & 0xffff0000u) * __uint_as_float(ys[q] & 0xffff0000u);
    }
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 4);
    if (ok && k == 0) {
        const int bidx = (int)(m / N), n = (int)(m % N);
        dsum[((long long)bidx * Hq + h) * N + n] = s;
    }
}

// dq (bf16, into the Q column range of dqkv) = inverse-RoPE(dQ accumulator f32 [M, Hq*64]); one thread per
// (row, head, 4 column pairs)
__global__ void attn_bwd_dq_finalize_kernel(const float* __restrict__ dq_acc, __nv_bfloat16* __restrict__ dqkv,
                                            const float* __restrict__ rope_cos, const float* __restrict__ rope_sin, int B,
                                            int N, int Hq, int Hkv, float scale) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)B * N * Hq * 8) return;
    const int j = (int)(i & 7) * 4;
    const long long mh = i >> 3;
    const int h = (int)(mh % Hq);
    const long long m = mh / Hq;
    const int pos = (int)(m % N);
    const float4 glo = *reinterpret_cast<const float4*>(dq_acc + m * (Hq * 64) + h * 64 + j);
    const float4 ghi = *reinterpret_cast<const float4*>(dq_acc + m * (Hq * 64) + h * 64 + 32 + j);
    float4 c = __ldg(reinterpret_cast<const float4*>(rope_cos + (long long)pos * 64 + j));
    float4 s = __ldg(reinterpret_cast<const float4*>(rope_sin + (long long)pos * 64 + j));
    c.x *= scale; c.y *= scale; c.z *= scale; c.w *= scale;   // v2 kernel: 1/sqrt(d) of dS applied here (scale = 1 for v1)
    s.x *= scale; s.y *= scale; s.z *= scale; s.w *= scale;
    __nv_bfloat16* o = dqkv + m * ((Hq + 2 * Hkv) * 64) + h * 64;
    *reinterpret_cast<uint2*>(o + j) = make_uint2(pack_bf16(glo.x * c.x + ghi.x * s.x, glo.y * c.y + ghi.y * s.y),
                                                  pack_bf16(glo.z * c.z + ghi.z * s.z, glo.w * c.w + ghi.w * s.w));
    *reinterpret_cast<uint2*>(o + 32 + j) = make_uint2(pack_bf16(ghi.x * c.x - glo.x * s.x, ghi.y * c.y - glo.y * s.y),
                                                       pack_bf16(ghi.z * c.z - glo.z * s.z, ghi.w * c.w - glo.w * s.w));
}

}  // namespace jat
