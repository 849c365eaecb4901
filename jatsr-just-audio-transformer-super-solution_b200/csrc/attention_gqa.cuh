// GQA attention forward on tcgen05 / TMEM (GroupedQueryAttention.forward, jat_audiosr_v2.py:141-164,
// eval mode: no mask, bidirectional, dropout = identity).
//
//   out[b, n, h, :] = softmax(Q[b,n,h,:] . K[b,:,g,:]^T / 8) V[b,:,g,:],   g = h / G,  G = Hq / Hkv
//
// The reference expands K/V G-fold with repeat_interleave (:147-148) and materialises the
// [B, Hq, N, N] score tensor in HBM several times per block; here one CTA owns (q-tile of 128 rows,
// KV head g, batch b): K_g and V_g ([N, 64] each) are staged in shared memory ONCE by TMA and reused
// by the G query heads of the group, scores live only in tensor memory.
//
// Because the token count per chunk is small (N = 345 for a 16 s chunk) the whole key range fits
// one TMEM accumulator (NK = round_up(N,16) <= 352 fp32 columns), so the softmax is exact two-pass
// (row max, then exp / sum) with no online rescaling.
//
// Warp roles (160 threads):
//   warps 0-3  softmax + output: thread = one query row (TMEM lane); S row read with tcgen05.ld,
//              P = exp2((s - max) * log2e/8) written as bf16 into a 128B-swizzled K-major smem tile,
//              O read back from TMEM, scaled by 1/sum, stored bf16.
//   warp 4     lane 0: TMA loads (K, V once; Q double-buffered over heads) and all tcgen05.mma issue:
//              S = Q K^T (M=128, N=NK split in two instructions, K=64) and O = P V (M=128, N=64, K=NK,
//              V consumed MN-major straight from its row-major [key, 64] tile).
#pragma once
#include <cuda.h>
#include "common.cuh"

namespace jat {

constexpr int ATT_BQ = 128;            // query rows per tile
constexpr int ATT_HD = 64;             // head dim
constexpr int ATT_MAX_NK = 352;        // padded key count limit (TMEM + smem budget)
constexpr int ATT_THREADS = 160;
constexpr int ATT_KV_BYTES = ATT_MAX_NK * 128;        // 45056, multiple of 1024
constexpr int ATT_Q_BYTES = ATT_BQ * 128;             // 16384
constexpr int ATT_P_BLOCK_BYTES = ATT_BQ * 128;       // one 64-key block of P
constexpr int ATT_P_BLOCKS = (ATT_MAX_NK + 63) / 64;  // 6
constexpr int ATT_O_COL = 448;                        // TMEM column of the O accumulator (S uses [0, 352))
constexpr int ATT_SMEM_BYTES = 2 * ATT_KV_BYTES + 2 * ATT_Q_BYTES + ATT_P_BLOCKS * ATT_P_BLOCK_BYTES + 128 + 1024;

struct AttnParams {
    int B, N, NK, Hq, Hkv, G;
    int kv_box_rows, kv_boxes;  // K/V are loaded as kv_boxes TMA boxes of kv_box_rows rows
    int nA, nB;                 // S = Q K^T is issued as two MMAs with N = nA and N = nB (nB may be 0)
    __nv_bfloat16* out;
    float scale_log2e;          // (1/sqrt(64)) * log2(e)
};

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__global__ void __launch_bounds__(ATT_THREADS, 1)
gqa_attention_fwd_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_kv,
                         const AttnParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
    uint8_t* sK = smem;
    uint8_t* sV = sK + ATT_KV_BYTES;
    uint8_t* sQ = sV + ATT_KV_BYTES;
    uint8_t* sP = sQ + 2 * ATT_Q_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sP + ATT_P_BLOCKS * ATT_P_BLOCK_BYTES);
    uint64_t* bar_kv = bars;          // K and V landed
    uint64_t* bar_q = bars + 1;       // [2] Q stage landed
    uint64_t* bar_s = bars + 3;       // S = Q K^T retired
    uint64_t* bar_p = bars + 4;       // P written by all 128 softmax threads (and S fully read)
    uint64_t* bar_o = bars + 5;       // O = P V retired
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qt = blockIdx.x, g = blockIdx.y, b = blockIdx.z;
    const int q_row0 = b * p.N + qt * ATT_BQ;  // global token row of this tile's first query
    const int kv_row0 = b * p.N;
    const int k_col = (p.Hq + g) * ATT_HD;
    const int v_col = (p.Hq + p.Hkv + g) * ATT_HD;

    if (warp == 4) {
        if (lane == 0) {
            tma_prefetch_desc(&tmap_q);
            tma_prefetch_desc(&tmap_kv);
            mbar_init(bar_kv, 1);
            mbar_init(&bar_q[0], 1);
            mbar_init(&bar_q[1], 1);
            mbar_init(bar_s, 1);
            mbar_init(bar_p, 128);
            mbar_init(bar_o, 1);
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc<1>(tmem_slot, 512);
        tmem_relinquish<1>();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 4) {
        if (lane == 0) {
            // ---------------------------------------------------------- loads
            const uint32_t kv_bytes = (uint32_t)(p.kv_boxes * p.kv_box_rows * 128);
            mbar_expect_tx(bar_kv, 2 * kv_bytes);
            for (int i = 0; i < p.kv_boxes; ++i) {
                tma_load_2d(sK + i * p.kv_box_rows * 128, &tmap_kv, bar_kv, k_col, kv_row0 + i * p.kv_box_rows);
                tma_load_2d(sV + i * p.kv_box_rows * 128, &tmap_kv, bar_kv, v_col, kv_row0 + i * p.kv_box_rows);
            }
            for (int h = 0; h < 2 && h < p.G; ++h) {
                mbar_expect_tx(&bar_q[h], ATT_Q_BYTES);
                tma_load_2d(sQ + h * ATT_Q_BYTES, &tmap_q, &bar_q[h], (g * p.G + h) * ATT_HD, q_row0);
            }
            const uint32_t idesc_sA = umma_idesc_bf16(ATT_BQ, p.nA);
            const uint32_t idesc_sB = umma_idesc_bf16(ATT_BQ, p.nB > 0 ? p.nB : 16);
            const uint32_t idesc_pv = umma_idesc_bf16(ATT_BQ, ATT_HD, /*a_major=*/0, /*b_major=*/1);
            const uint32_t sK_addr = smem_u32(sK), sV_addr = smem_u32(sV), sP_addr = smem_u32(sP);
            mbar_wait(bar_kv, 0);
            for (int h = 0; h < p.G; ++h) {
                const int st = h & 1;
                // ------------------------------------------------------ S = Q K^T
                mbar_wait(&bar_q[st], (uint32_t)((h >> 1) & 1));
                tc_fence_after();
                const uint64_t q_desc = umma_smem_desc_sw128(smem_u32(sQ + st * ATT_Q_BYTES));
                const uint64_t kA_desc = umma_smem_desc_sw128(sK_addr);
                const uint64_t kB_desc = umma_smem_desc_sw128(sK_addr + (uint32_t)p.nA * 128u);
#pragma unroll
                for (int k = 0; k < ATT_HD / 16; ++k)
                    umma_bf16_ss<1>(tmem_base, q_desc + 2 * k, kA_desc + 2 * k, idesc_sA, (uint32_t)(k != 0));
                if (p.nB > 0) {
#pragma unroll
                    for (int k = 0; k < ATT_HD / 16; ++k)
                        umma_bf16_ss<1>(tmem_base + (uint32_t)p.nA, q_desc + 2 * k, kB_desc + 2 * k, idesc_sB,
                                        (uint32_t)(k != 0));
                }
                umma_commit(bar_s);
                // ------------------------------------------------------ O = P V
                mbar_wait(bar_p, (uint32_t)(h & 1));
                tc_fence_after();
                // S(h) has retired (the softmax consumed it), so Q stage `st` is free: prefetch head h+2.
                if (h + 2 < p.G) {
                    mbar_expect_tx(&bar_q[st], ATT_Q_BYTES);
                    tma_load_2d(sQ + st * ATT_Q_BYTES, &tmap_q, &bar_q[st], (g * p.G + h + 2) * ATT_HD, q_row0);
                }
                const int ksteps = p.NK / 16;
                for (int ks = 0; ks < ksteps; ++ks) {
                    // A: P block (ks/4), +32 B per 16-key step inside the swizzle atom
                    const uint64_t a_desc =
                        umma_smem_desc_sw128(sP_addr + (uint32_t)(ks >> 2) * ATT_P_BLOCK_BYTES) + 2 * (ks & 3);
                    // B: V rows [16 ks, 16 ks + 16): MN-major, 2 groups of 8 key rows, 1024 B apart
                    const uint64_t b_desc = umma_smem_desc_sw128(sV_addr + (uint32_t)ks * 2048u);
                    umma_bf16_ss<1>(tmem_base + ATT_O_COL, a_desc, b_desc, idesc_pv, (uint32_t)(ks != 0));
                }
                umma_commit(bar_o);
            }
        }
        __syncwarp();
    } else {
        // -------------------------------------------------------------- softmax + output warps
        const int r = warp * 32 + lane;                              // query row in tile = TMEM lane
        const uint32_t t_row = tmem_base + ((uint32_t)(warp * 32) << 16);
        const bool row_ok = qt * ATT_BQ + r < p.N;
        const float sl2 = p.scale_log2e;
        const uint32_t sP_row = smem_u32(sP) + (uint32_t)(r >> 3) * 1024u + (uint32_t)(r & 7) * 128u;
        const uint32_t rx = (uint32_t)(r & 7);
        const int full_chunks = p.NK / 32;
        const bool tail16 = (p.NK & 16) != 0;
        for (int h = 0; h < p.G; ++h) {
            mbar_wait(bar_s, (uint32_t)(h & 1));
            tc_fence_after();
            // ---- pass 1: row max over the N valid keys
            float mx = -INFINITY;
#pragma unroll 1
            for (int c = 0; c < full_chunks; ++c) {
                uint32_t v[32];
                tmem_ld_32x32(t_row + c * 32, v);
                tmem_ld_wait();
                if ((c + 1) * 32 <= p.N) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(v[j]));
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (c * 32 + j < p.N) mx = fmaxf(mx, __uint_as_float(v[j]));
                }
            }
            if (tail16) {
                uint32_t v[16];
                tmem_ld_32x16(t_row + full_chunks * 32, v);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    if (full_chunks * 32 + j < p.N) mx = fmaxf(mx, __uint_as_float(v[j]));
            }
            const float moff = mx * sl2;
            // ---- pass 2: p = exp2(s*sl2 - moff), row sum, bf16 P -> swizzled smem
            float sum = 0.f;
#pragma unroll 1
            for (int c = 0; c < full_chunks; ++c) {
                uint32_t v[32];
                tmem_ld_32x32(t_row + c * 32, v);
                tmem_ld_wait();
                float e[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const float x = ex2_approx(fmaf(__uint_as_float(v[j]), sl2, -moff));
                    e[j] = (c * 32 + j < p.N) ? x : 0.f;
                    sum += e[j];
                }
                // 32 keys = 4 x 16-byte chunks of key block (c/2); chunk index within block = (c&1)*4 + q
                const uint32_t blk = sP_row + (uint32_t)(c >> 1) * ATT_P_BLOCK_BYTES;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const uint32_t chunk = (uint32_t)((c & 1) * 4 + q);
                    const uint32_t addr = blk + ((chunk ^ rx) << 4);
                    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr),
                                 "r"(pack_bf16(e[q * 8 + 0], e[q * 8 + 1])), "r"(pack_bf16(e[q * 8 + 2], e[q * 8 + 3])),
                                 "r"(pack_bf16(e[q * 8 + 4], e[q * 8 + 5])), "r"(pack_bf16(e[q * 8 + 6], e[q * 8 + 7]))
                                 : "memory");
                }
            }
            if (tail16) {
                uint32_t v[16];
                tmem_ld_32x16(t_row + full_chunks * 32, v);
                tmem_ld_wait();
                float e[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const float x = ex2_approx(fmaf(__uint_as_float(v[j]), sl2, -moff));
                    e[j] = (full_chunks * 32 + j < p.N) ? x : 0.f;
                    sum += e[j];
                }
                const int c = full_chunks;
                const uint32_t blk = sP_row + (uint32_t)(c >> 1) * ATT_P_BLOCK_BYTES;
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const uint32_t chunk = (uint32_t)((c & 1) * 4 + q);
                    const uint32_t addr = blk + ((chunk ^ rx) << 4);
                    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr),
                                 "r"(pack_bf16(e[q * 8 + 0], e[q * 8 + 1])), "r"(pack_bf16(e[q * 8 + 2], e[q * 8 + 3])),
                                 "r"(pack_bf16(e[q * 8 + 4], e[q * 8 + 5])), "r"(pack_bf16(e[q * 8 + 6], e[q * 8 + 7]))
                                 : "memory");
                }
            }
            fence_proxy_async_smem();  // P (generic-proxy stores) -> visible to the tensor core's async proxy
            tc_fence_before();         // orders this thread's tcgen05.ld of S before the next S MMA
            mbar_arrive(bar_p);

            // ---- O epilogue
            mbar_wait(bar_o, (uint32_t)(h & 1));
            tc_fence_after();
            const float inv = 1.0f / sum;
            __nv_bfloat16* orow =
                p.out + (long long)(q_row0 + r) * (p.Hq * ATT_HD) + (long long)(g * p.G + h) * ATT_HD;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                uint32_t v[32];
                tmem_ld_32x32(t_row + ATT_O_COL + c * 32, v);
                tmem_ld_wait();
                if (row_ok) {
#pragma unroll
                    for (int j = 0; j < 32; j += 8) {
                        uint4 o;
                        o.x = pack_bf16(__uint_as_float(v[j + 0]) * inv, __uint_as_float(v[j + 1]) * inv);
                        o.y = pack_bf16(__uint_as_float(v[j + 2]) * inv, __uint_as_float(v[j + 3]) * inv);
                        o.z = pack_bf16(__uint_as_float(v[j + 4]) * inv, __uint_as_float(v[j + 5]) * inv);
                        o.w = pack_bf16(__uint_as_float(v[j + 6]) * inv, __uint_as_float(v[j + 7]) * inv);
                        *reinterpret_cast<uint4*>(orow + c * 32 + j) = o;
                    }
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 4) {
        tc_fence_after();
        tmem_dealloc<1>(tmem_base, 512);
    }
}

}  // namespace jat
