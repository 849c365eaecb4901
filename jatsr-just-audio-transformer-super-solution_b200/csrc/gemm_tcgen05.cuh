// tcgen05 / TMEM GEMM fed by TMA, with the DiT block's elementwise work fused into the epilogue.
//
//   acc[M, N] = A[M, K] * W[N, K]^T     A, W bf16 K-major (nn.Linear weight layout), fp32 accumulate
//
// Replaces the cuBLAS calls + separate bias / GELU / RoPE / gate / residual / permute kernels the
// reference issues for patch_embed.proj, t_embedder, adaLN_modulation, q/k/v_proj + RoPE, out_proj,
// mlp.0/mlp.3 and final_layer (jat_audiosr_v2.py:204-208, 341-346, 256-259, 137-145, 165, 247-253, 281,
// 286-287, 359-363, 383-397).
//
// Structure (one persistent CTA -- or CTA pair -- per SM, 384 threads):
//   warp 0      TMA producer: A tile [128 x 64] + W tile [BN/CG x 64] per stage, 128B swizzle
//   warp 1      MMA issuer (one lane): tcgen05.mma kind::f16, M = 128*CG, N = BN, K = 16, accumulators
//               double-buffered in TMEM (2 x BN columns) so the epilogue of tile i overlaps tile i+1
//   warp 2      TMEM allocator
//   warps 4-11  epilogue: tcgen05.ld (thread = one accumulator row) -> fused math in registers ->
//               128B-swizzled shared-memory slab (32 rows x 128 B per warp, double-buffered) -> TMA store,
//               or TMA reduce-add for the gated residual (x += gate * (acc + bias) is performed by the
//               TMA unit in L2: the epilogue never reads x).  Warpgroup 0 takes columns [0, BN/2) of the
//               tile, warpgroup 1 the rest.  The un-patchify epilogue stores directly (its [B,C,T] target
//               is already coalesced across lanes).
// Pipelines: smem full/empty mbarriers (TMA <-> MMA), TMEM full/empty mbarriers (MMA <-> epilogue),
// bulk async-groups (epilogue smem slab <-> TMA store).
#pragma once
#include <cuda.h>
#include "common.cuh"

namespace jat {

enum { EPI_BIAS_ACT = 0, EPI_QKV_ROPE = 1, EPI_GATE_RESIDUAL = 2, EPI_UNPATCHIFY = 3, EPI_ACCUM = 4, EPI_DACT = 5 };
enum { ACT_NONE = 0, ACT_GELU = 1, ACT_SILU = 2 };

struct GemmParams {
    int M, N, K;
    int num_n_blocks, num_k_blocks, num_tiles;
    const float* bias;
    void* out;
    long long ldo;
    const float* gate;
    long long gate_bstride;
    const float* gate_rowscale;  // GATE_RESIDUAL: optional per-batch-item factor on the gate (DropPath), f32 [B] or NULL
    int tokens_per_batch;
    const float* rope_cos;
    const float* rope_sin;
    int rope_cols;
    int t_out;
    int k_splits;            // split-K: every output tile is computed by k_splits work items (reduce-add epilogues only)
    DropCfg drop;            // train-mode dropout on the epilogue output (BIAS_ACT: after the activation; GATE_RESIDUAL:
                             // on acc + bias before the gate; DACT: the same mask applied to the incoming gradient)
    const void* aux;         // EPI_DACT: pre-activation u bf16 [M, ld_aux];  EPI_BIAS_ACT: optional bf16 copy of (acc + bias)
    long long ld_aux;
    // Tail split (deterministic stream-K fix-up for the last, partial wave): tiles [0, head_tiles) are whole work items
    // (x k_splits); each tile in [head_tiles, num_tiles) is cut into tail_splits reduction parts whose f32 partial
    // accumulators go to tail_ws; the part that arrives last (tail_cnt) sums them in part order and runs the epilogue.
    int head_tiles, tail_splits;
    int tail_direct;  // 1: reduce-add epilogues only -- the parts of a tail tile add their partial sums straight into the output
    float* tail_ws;
    int* tail_cnt;
    long long* trace;  // diagnostics: clock64 timestamps of cluster 0 / CTA rank 0, 8 slots per work item, or NULL
    int dbg_skip;  // diagnostics only (JAT_DBG_GEMM_SKIP): bit 0 = do not load A, bit 1 = do not load B (results are garbage)
};

#define GEMM_TRACE(item, slot)                                                                      \
    do {                                                                                            \
        if (p.trace != nullptr && blockIdx.x == 0 && (item) < 64) p.trace[(item) * 8 + (slot)] = clock64(); \
    } while (0)

struct GemmWork {
    int tile, kb0, kb1, part, tail;  // tail = index of the split tail tile, or -1
};
__device__ __forceinline__ int gemm_num_work(const GemmParams& p) {
    return p.head_tiles * p.k_splits + (p.num_tiles - p.head_tiles) * p.tail_splits;
}
__device__ __forceinline__ GemmWork gemm_decode_work(const GemmParams& p, int work) {
    GemmWork w;
    int parts;
    const int head_items = p.head_tiles * p.k_splits;
    if (work < head_items) {
        w.tile = work / p.k_splits; w.part = work - w.tile * p.k_splits; parts = p.k_splits; w.tail = -1;
    } else {
        const int w2 = work - head_items;
        const int t2 = w2 / p.tail_splits;
        w.tile = p.head_tiles + t2; w.part = w2 - t2 * p.tail_splits; parts = p.tail_splits;
        w.tail = (p.tail_splits > 1 && !p.tail_direct) ? t2 : -1;
    }
    if (parts == 1) { w.kb0 = 0; w.kb1 = p.num_k_blocks; }
    else { w.kb0 = p.num_k_blocks * w.part / parts; w.kb1 = p.num_k_blocks * (w.part + 1) / parts; }
    return w;
}

// Sum of the `parts` partial accumulators of a split tail tile, columns [col, col + 4*NV) of this thread's row, in
// part order (deterministic).  Workspace region layout: float4 unit u = col/4 of lane l at [(u * 32 + l)].
template <int NV>
__device__ __forceinline__ void gemm_ws_sum(const float* region0, long long part_stride, int parts, int col, int lane,
                                            uint32_t* v) {
    float4 acc[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int q = 0; q < parts; ++q) {
        const float4* src = reinterpret_cast<const float4*>(region0 + q * part_stride) + (col >> 2) * 32 + lane;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const float4 t = __ldcg(src + i * 32);
            acc[i].x += t.x; acc[i].y += t.y; acc[i].z += t.z; acc[i].w += t.w;
        }
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        v[4 * i + 0] = __float_as_uint(acc[i].x); v[4 * i + 1] = __float_as_uint(acc[i].y);
        v[4 * i + 2] = __float_as_uint(acc[i].z); v[4 * i + 3] = __float_as_uint(acc[i].w);
    }
}

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;
constexpr int GEMM_THREADS = 384;
constexpr int GEMM_EPI_WARPS = 8;
constexpr int GEMM_SLAB_BYTES = 32 * 128;  // one epilogue slab: 32 rows x 128 B (64 bf16 or 32 f32 columns)

template <int BN, int CG>
struct GemmCfg {
    static constexpr int B_ROWS = BN / CG;
    static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;
    static constexpr int B_BYTES = B_ROWS * GEMM_BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int STAGING_BYTES = GEMM_EPI_WARPS * 2 * GEMM_SLAB_BYTES;  // 64 KB
#ifndef JAT_GEMM_STAGE_BUDGET_KB
#define JAT_GEMM_STAGE_BUDGET_KB 160
#endif
    static constexpr int STAGES_RAW = (JAT_GEMM_STAGE_BUDGET_KB * 1024) / STAGE_BYTES;
    static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
    static constexpr int TMEM_COLS = 2 * BN;
    static constexpr int BAR_BYTES = (2 * STAGES + 4) * 8 + 16;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + STAGING_BYTES + BAR_BYTES + 1024;  // +1024: alignment slack
};

template <int ACT>
__device__ __forceinline__ float apply_act(float v) {
    if constexpr (ACT == ACT_GELU) return gelu_erf(v);
    if constexpr (ACT == ACT_SILU) return silu(v);
    return v;
}

// derivative of the activation at pre-activation u (backward of mlp.1 / patch_embed.proj.1 GELU, t_embedder.2 SiLU)
template <int ACT>
__device__ __forceinline__ float dact(float u) {
    if constexpr (ACT == ACT_GELU) {  // d/du [u Phi(u)] = Phi(u) + u phi(u), Phi from the same A&S 7.1.26 form as gelu_erf
        const float a = fabsf(u) * 0.84932180028801904f;  // |u| sqrt(log2(e) / 2)
        float t, e;
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.27273748f, a, 1.0f)));
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-a * a));  // = exp(-u^2 / 2)
        float pl = fmaf(t, 0.5f * 1.061405429f, 0.5f * -1.453152027f);
        pl = fmaf(t, pl, 0.5f * 1.421413741f);
        pl = fmaf(t, pl, 0.5f * -0.284496736f);
        pl = fmaf(t, pl, 0.5f * 0.254829592f);
        const float q = pl * t * e;                        // 1 - Phi(|u|)
        const float cdf = u >= 0.0f ? 1.0f - q : q;
        return fmaf(u * 0.3989422804014327f, e, cdf);
    }
    if constexpr (ACT == ACT_SILU) {  // d/du [u sigma(u)] = sigma(u) (1 + u (1 - sigma(u)))
        const float sg = 1.0f / (1.0f + __expf(-u));
        return sg * fmaf(u, 1.0f - sg, 1.0f);
    }
    return 1.0f;
}

// 16-byte chunk `chunk` (0..7) of row `lane` inside a 128B-swizzled 32-row slab
__device__ __forceinline__ void st_slab_chunk(uint32_t slab_row_addr, int lane, int chunk, uint32_t a, uint32_t b,
                                              uint32_t c, uint32_t d) {
    const uint32_t addr = slab_row_addr + (uint32_t)(((chunk ^ (lane & 7)) << 4));
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

__device__ __forceinline__ uint4 ld_slab_chunk(uint32_t slab_row_addr, int row, int chunk) {
    uint4 v;
    const uint32_t addr = slab_row_addr + (uint32_t)(((chunk ^ (row & 7)) << 4));
    asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
    return v;
}
// Warp-cooperative, fully coalesced copies between a 32-row epilogue slab (128B-swizzled, CH 16-byte chunks used per row)
// and a row-major global matrix: CH lanes cover one row's contiguous bytes, 32 / CH rows per instruction -- instead of
// one 16-byte access per lane with a row-pitch stride (32 separate sectors per instruction).  The training forward keeps
// pre-activations / pre-gate values this way, the GELU' dgrad reads them back this way.  Rows >= rows_ok are skipped.
template <int CH>
__device__ __forceinline__ void slab_to_global(uint32_t slab_addr, int lane, __nv_bfloat16* gbase, long long ld, int rows_ok) {
    constexpr int RPI = 32 / CH;  // rows per instruction
    const int c = lane % CH, sub = lane / CH;
#pragma unroll
    for (int i = 0; i < 32 / RPI; ++i) {
        const int row = i * RPI + sub;
        const uint4 v = ld_slab_chunk(slab_addr + (uint32_t)row * 128u, row, c);
        if (row < rows_ok) *reinterpret_cast<uint4*>(gbase + (long long)row * ld + c * 8) = v;
    }
}
template <int CH>
__device__ __forceinline__ void global_to_slab(uint32_t slab_addr, int lane, const __nv_bfloat16* gbase, long long ld, int rows_ok) {
    constexpr int RPI = 32 / CH;
    const int c = lane % CH, sub = lane / CH;
    uint4 v[32 / RPI];
#pragma unroll
    for (int i = 0; i < 32 / RPI; ++i) {  // all loads in flight before the first shared-memory store
        const int row = i * RPI + sub;
        v[i] = row < rows_ok ? __ldg(reinterpret_cast<const uint4*>(gbase + (long long)row * ld + c * 8)) : make_uint4(0u, 0u, 0u, 0u);
    }
#pragma unroll
    for (int i = 0; i < 32 / RPI; ++i) {
        const int row = i * RPI + sub;
        st_slab_chunk(slab_addr + (uint32_t)row * 128u, row, c, v[i].x, v[i].y, v[i].z, v[i].w);
    }
}

// A_MN / B_MN: the operand is stored "MN-major" = as the TRANSPOSE of the K-major layout, i.e. A^T [K, M] /
// W^T [K, N] row-major (reduction index = row).  That is what the backward GEMMs see without any transposed
// copies:  dgrad  dX[M, K'] = dY[M, N'] W[N', K']      -> B_MN (W rows = reduction index)
//          wgrad  dW[N', K'] = dY^T[N', M] X[M, K']    -> A_MN and B_MN (token index = reduction index)
// Tiles are then TMA boxes of 64 reduction rows x 64 elements (128 B), one box per 64 output rows/columns,
// consumed through MN-major UMMA descriptors (leading byte offset = one 8 KB box).
// MC = CTA pairs per cluster (CG == 2 only).  MC == 2: a cluster of 4 CTAs computes two 256-row M blocks of the same N block;
// every CTA loads HALF of its pair's share of the W tile and multicasts it to the CTA with the same pair-rank in the other
// pair, so the cluster reads W once instead of twice (the long-K GEMMs are bound by the L2 -> SM operand stream).
// EW = epilogue warps: 8 (default; two column halves, two staging slabs per warp) or 16 (bf16-output epilogues of 256-wide
// tiles: four column quarters = ONE 64-column slab per warp and tile, single-buffered -- the slab's TMA store was issued a
// whole tile ago; 640 threads at <= 102 registers).  The training-mode fc1 / GELU' dgrad epilogues run ~33 instructions per
// output element and are bound by instruction issue and latency with two epilogue warps per scheduler; four hide it.
template <int BN, int CG, int EPI, int ACT, int OUT_BF16, int A_MN = 0, int B_MN = 0, int MC = 1, int EW = GEMM_EPI_WARPS>
__global__ void __launch_bounds__(128 + EW * 32, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                    const __grid_constant__ CUtensorMap tmap_out, const GemmParams p) {
    using Cfg = GemmCfg<BN, CG>;
    constexpr int STAGES = Cfg::STAGES;

    extern __shared__ uint8_t smem_raw[];
    // 128B-swizzled tiles need a 1024-byte aligned base (the swizzle is a function of address bits 4..9).
    const uint32_t raw_addr = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
    uint8_t* staging = smem + STAGES * Cfg::STAGE_BYTES;

    uint64_t* bar_full = reinterpret_cast<uint64_t*>(staging + Cfg::STAGING_BYTES);
    uint64_t* bar_empty = bar_full + STAGES;
    uint64_t* bar_tmem_full = bar_empty + STAGES;
    uint64_t* bar_tmem_empty = bar_tmem_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_tmem_empty + 2);

    // warp-uniform values are passed through a full-mask shuffle: that is what tells the compiler they are uniform, so
    // that the MMA / TMA issue paths keep descriptors and addresses in uniform registers
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;
    static_assert(MC == 1 || (CG == 2 && !A_MN && !B_MN), "multicast clusters: CTA pairs with K-major operands only");
    const uint32_t cluster_rank = (CG == 2) ? __shfl_sync(0xffffffffu, cluster_ctarank(), 0) : 0u;
    const uint32_t cta_rank = cluster_rank & 1u;   // rank inside the CTA pair (0 = leader: issues the MMAs)
    const int pair = (int)(cluster_rank >> 1);     // which pair of the cluster (0 .. MC-1)
    const int cluster_id = blockIdx.x / (CG * MC);
    const int num_clusters = gridDim.x / (CG * MC);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_b);
        if constexpr (EPI != EPI_UNPATCHIFY) tma_prefetch_desc(&tmap_out);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&bar_full[s], 1);
            mbar_init(&bar_empty[s], MC);  // a stage is free when the MMAs of every pair that reads it have retired
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&bar_tmem_full[a], 1);
            mbar_init(&bar_tmem_empty[a], EW * CG);
        }
        fence_barrier_init();
    }
    if (warp == 2) {
        tmem_alloc<CG>(tmem_slot, Cfg::TMEM_COLS);
        tmem_relinquish<CG>();
    }
    tc_fence_before();
    if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
    pdl_launch_dependents();
    pdl_wait();  // everything above overlapped the predecessor's tail; operands / outputs are touched only from here on

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer (one elected lane issues)
        {
            int stage = 0;
            uint32_t phase = 0;
            const int num_work = gemm_num_work(p);
            for (int work = cluster_id; work < num_work; work += num_clusters) {
                const GemmWork wk = gemm_decode_work(p, work);
                const int tile = wk.tile, kb0 = wk.kb0, kb1 = wk.kb1;
                const int m_blk = tile / p.num_n_blocks, n_blk = tile % p.num_n_blocks;
                const int a_row0 = ((m_blk * MC + pair) * CG + (int)cta_rank) * GEMM_BM;
                const int b_row0 = n_blk * BN + (int)cta_rank * Cfg::B_ROWS;
                if (lane == 0) GEMM_TRACE(work / num_clusters, 0);
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(&bar_empty[stage], phase ^ 1);
                    uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
                    uint8_t* sb = sa + Cfg::A_BYTES;
                    if (elect_one()) {
                    if (p.dbg_skip != 0) {  // bandwidth diagnostics: stream only one (or neither) operand
                        const uint32_t bytes = ((p.dbg_skip & 1) ? 0u : (uint32_t)Cfg::A_BYTES) + ((p.dbg_skip & 2) ? 0u : (uint32_t)Cfg::B_BYTES);
                        if (bytes == 0) { if (CG == 1 || cta_rank == 0) mbar_arrive(&bar_full[stage]); }
                        else if (CG == 1) mbar_expect_tx(&bar_full[stage], bytes);
                        else if (cta_rank == 0) mbar_expect_tx(&bar_full[stage], bytes * 2);
                        if (!(p.dbg_skip & 1)) tma_load_tile<CG>(sa, &tmap_a, &bar_full[stage], kb * GEMM_BK, a_row0);
                        if (!(p.dbg_skip & 2)) tma_load_tile<CG>(sb, &tmap_b, &bar_full[stage], kb * GEMM_BK, b_row0);
                    } else {
                    if (CG == 1) mbar_expect_tx(&bar_full[stage], Cfg::STAGE_BYTES);
                    else if (cta_rank == 0) mbar_expect_tx(&bar_full[stage], Cfg::STAGE_BYTES * 2);  // both CTAs' bytes
                    if constexpr (A_MN) {  // boxes of 64 reduction rows x 64 output rows (8 KB each)
#pragma unroll
                        for (int j = 0; j < GEMM_BM / 64; ++j)
                            tma_load_tile<CG>(sa + j * 8192, &tmap_a, &bar_full[stage], a_row0 + j * 64, kb * GEMM_BK);
                    } else {
                        tma_load_tile<CG>(sa, &tmap_a, &bar_full[stage], kb * GEMM_BK, a_row0);
                    }
                    if constexpr (B_MN) {
#pragma unroll
                        for (int j = 0; j < Cfg::B_ROWS / 64; ++j)
                            tma_load_tile<CG>(sb + j * 8192, &tmap_b, &bar_full[stage], b_row0 + j * 64, kb * GEMM_BK);
                    } else if constexpr (MC == 2) {
                        // my half of this pair-rank's W rows, to both CTAs of the cluster with this pair-rank
                        constexpr int PART = Cfg::B_ROWS / 2;
                        tma_load_2d_2sm_mc(sb + pair * PART * 128, &tmap_b, &bar_full[stage], kb * GEMM_BK, b_row0 + pair * PART,
                                           (uint16_t)(cta_rank ? 0xAu : 0x5u));
                    } else {
                        tma_load_tile<CG>(sb, &tmap_b, &bar_full[stage], kb * GEMM_BK, b_row0);
                    }
                    }
                    }
                    __syncwarp();
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
        __syncwarp();  // reconverge before the (warp-aligned) teardown barrier
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer (lane 0 issues; warp-uniform control)
        // The whole warp runs the (uniform) control flow and ONE ELECTED lane issues: under elect.sync the descriptors stay
        // in uniform registers and the four MMAs of a k-block go out back to back -- measured 512 clk per k-block, the
        // tcgen05 rate.  (Under `if (lane == 0)` every tcgen05.mma was wrapped in an ELECT / R2UR.BROADCAST waterfall:
        // ~570-700 clk per k-block, more when the epilogue warps compete for issue slots.)  The path between the last MMA
        // of a tile and the first of the next is kept minimal: no work decoding when every item spans all k-blocks.
        if (cta_rank == 0) {
            constexpr uint32_t idesc = umma_idesc_bf16(GEMM_BM * CG, BN, A_MN, B_MN);
            // per 16-element k-step: +32 B inside the 128B swizzle atom (K-major) or +16 rows x 128 B (MN-major)
            constexpr uint32_t a_kstep = A_MN ? (2048u >> 4) : 2u, b_kstep = B_MN ? (2048u >> 4) : 2u;
            const int num_work = gemm_num_work(p);
            const bool uniform = p.k_splits == 1 && p.tail_splits == 1;  // every item spans all k-blocks
            int stage = 0, as = 0;
            uint32_t phase = 0, aphase = 0;
            for (int work = cluster_id; work < num_work; work += num_clusters) {
                int nkb = p.num_k_blocks;
                if (!uniform) { const GemmWork wk = gemm_decode_work(p, work); nkb = wk.kb1 - wk.kb0; }
                mbar_wait(&bar_tmem_empty[as], aphase ^ 1);
                const uint32_t d_tmem = tmem_base + (uint32_t)(as * BN);
                if (lane == 0) GEMM_TRACE(work / num_clusters, 2);
                for (int i = 0; i < nkb; ++i) {
                    mbar_wait(&bar_full[stage], phase);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
                        const uint64_t a_desc = A_MN ? umma_smem_desc_sw128_mn(sa, 8192) : umma_smem_desc_sw128(sa);
                        const uint64_t b_desc = B_MN ? umma_smem_desc_sw128_mn(sa + Cfg::A_BYTES, 8192)
                                                     : umma_smem_desc_sw128(sa + Cfg::A_BYTES);
#pragma unroll
                        for (int k = 0; k < GEMM_BK / 16; ++k)
                            umma_bf16_ss<CG>(d_tmem, a_desc + a_kstep * k, b_desc + b_kstep * k, idesc, (uint32_t)((i | k) != 0));
                        if constexpr (CG == 1) umma_commit(&bar_empty[stage]);
                        else umma_commit_2sm(&bar_empty[stage], MC == 2 ? 0xF : 0x3);
                        if (i + 1 == nkb) {
                            if constexpr (CG == 1) umma_commit(&bar_tmem_full[as]);
                            else umma_commit_2sm(&bar_tmem_full[as], (uint16_t)(0x3 << (2 * pair)));
                            GEMM_TRACE(work / num_clusters, 3);
                        }
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                if ((as ^= 1) == 0) aphase ^= 1;
            }
        }
        __syncwarp();
    } else if (warp >= 4) {
        // ------------------------------------------------------------------ epilogue
        const int quad = warp & 3;       // TMEM lane quadrant this warp may access
        const int wg = (warp - 4) >> 2;  // column half
        static_assert(EW == 8 || (EW == 16 && BN == 256 && OUT_BF16 && (EPI == EPI_BIAS_ACT || EPI == EPI_DACT)),
                      "16 epilogue warps: bf16-output epilogues of 256-wide tiles only");
        constexpr int HALF = BN / (EW / 4);          // columns of the tile this warp owns (a half, or a quarter with EW = 16)
        constexpr int SLABS = EW == 8 ? 2 : 1;       // staging slabs per warp
        uint8_t* my_stage = staging + (warp - 4) * SLABS * GEMM_SLAB_BYTES;
        const uint32_t slab_row[2] = {smem_u32(my_stage) + (uint32_t)lane * 128u,
                                      smem_u32(my_stage + (SLABS - 1) * GEMM_SLAB_BYTES) + (uint32_t)lane * 128u};
        int buf = 0;
        int as = 0;
        uint32_t aphase = 0;
        const int num_work = gemm_num_work(p);
        constexpr long long WS_REGION = 32LL * HALF;                         // floats per (part, CTA, warp) region
        constexpr long long WS_PART_STRIDE = (long long)CG * EW * WS_REGION;
        for (int work = cluster_id; work < num_work; work += num_clusters) {
            const GemmWork wk = gemm_decode_work(p, work);
            const int tile = wk.tile;
            // split-K (reduce-add epilogues): the bias goes in with the first partial sum; tail split: with the fix-up
            const bool add_bias = p.bias != nullptr && (wk.part == 0 || wk.tail >= 0);
            const int m_blk = tile / p.num_n_blocks, n_blk = tile % p.num_n_blocks;
            const int row0 = ((m_blk * MC + pair) * CG + (int)cta_rank) * GEMM_BM + quad * 32;  // first row of this warp's slab
            const int m = row0 + lane;
            const bool row_ok = m < p.M;
            const int n_base = n_blk * BN + wg * HALF;
            if (warp == 4 && lane == 0) GEMM_TRACE(work / num_clusters, 4);
            mbar_wait(&bar_tmem_full[as], aphase);
            tc_fence_after();
            if (warp == 4 && lane == 0) GEMM_TRACE(work / num_clusters, 5);
            const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(as * BN + wg * HALF);
            // ---- split tail tile: park this part's accumulator in the workspace; only the warp that arrives last
            //      (per tile, CTA, warp slab) goes on to the epilogue, on the in-order sum of all parts
            const float* ws_region = nullptr;
            if (wk.tail >= 0) {
                const long long slab = ((long long)cta_rank * EW + (warp - 4)) * WS_REGION;
                float* tile_ws = p.tail_ws + (long long)wk.tail * p.tail_splits * WS_PART_STRIDE + slab;
                float4* dst = reinterpret_cast<float4*>(tile_ws + wk.part * WS_PART_STRIDE) + lane;
#pragma unroll 1
                for (int c = 0; c < HALF / 32; ++c) {
                    uint32_t v[32];
                    tmem_ld_32x32(taddr + c * 32, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        __stcg(dst + (c * 8 + i) * 32, make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]),
                                                                   __uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3])));
                }
                tc_fence_before();
                __threadfence();
                __syncwarp();
                int last = 0;
                if (lane == 0) {
                    if constexpr (CG == 1) mbar_arrive(&bar_tmem_empty[as]);
                    else mbar_arrive_cluster(&bar_tmem_empty[as], cluster_rank & ~1u);
                    int* cnt = p.tail_cnt + (wk.tail * CG + (int)cta_rank) * EW + (warp - 4);
                    last = atomicAdd(cnt, 1) == p.tail_splits - 1;
                    if (last) *cnt = 0;  // every part has arrived: reset for the next launch
                }
                last = __shfl_sync(0xffffffffu, last, 0);
                if ((as ^= 1) == 0) aphase ^= 1;
                if (!last) continue;
                __threadfence();
                ws_region = tile_ws;
            }
            // accumulator columns [col, col+32) / [col, col+16) of this thread's row: from TMEM, or the summed partials
            auto load32 = [&](int col, uint32_t (&v)[32]) {
                if (ws_region == nullptr) { tmem_ld_32x32(taddr + col, v); tmem_ld_wait(); }
                else gemm_ws_sum<8>(ws_region, WS_PART_STRIDE, p.tail_splits, col, lane, v);
            };
            auto load16 = [&](int col, uint32_t (&v)[16]) {
                if (ws_region == nullptr) tmem_ld_32x16(taddr + col, v);
                else gemm_ws_sum<4>(ws_region, WS_PART_STRIDE, p.tail_splits, col, lane, v);
            };

            if constexpr ((EPI == EPI_BIAS_ACT && OUT_BF16) || EPI == EPI_DACT) {
                // BIAS_ACT: out = act(acc + bias)  [+ optional bf16 copy of the pre-activation acc + bias to p.aux,
                //           kept by the training forward for the backward of the activation]
                // DACT:     out = (acc + bias) * act'(u),  u = pre-activation read from p.aux   (dgrad through GELU)
                // The aux matrix moves through the staging slab with warp-cooperative coalesced accesses (slab_to_global /
                // global_to_slab), not one strided 16-byte access per thread.
                const bool use_aux = p.aux != nullptr;  // warp-uniform
                const int rows_ok = p.M - row0;          // rows of this warp's slab that exist (may be <= 0 or >= 32)
                __nv_bfloat16* aux_slab = reinterpret_cast<__nv_bfloat16*>(const_cast<void*>(p.aux)) + (long long)row0 * p.ld_aux;
#pragma unroll 1
                for (int sl = 0; sl < HALF / 64; ++sl) {  // slab = 64 bf16 columns
                    const int n0 = n_base + sl * 64;
                    if (lane == 0) tma_store_wait_read<SLABS - 1>();
                    __syncwarp();
                    // EPI_DACT: the 64 pre-activations of this thread's row (packed bf16 pairs), one 16-byte load per chunk, all
                    // issued ahead of the TMEM read-out (measured faster than staging the rows through the slab with coalesced
                    // loads: the extra shared-memory round trip sits on the epilogue's critical path, 3.99 vs 3.77 ms per step).
                    // With 16 epilogue warps (96 registers) the loads are issued per 32-column half instead.
                    uint32_t uw[EW == 8 ? 32 : 16];
                    const __nv_bfloat16* aux_row = aux_slab + (long long)lane * p.ld_aux + n0;
                    auto load_u = [&](int c0, int nchunks) {
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            if (c >= nchunks) break;
                            uint4 t4 = make_uint4(0u, 0u, 0u, 0u);
                            if (use_aux && row_ok) t4 = __ldg(reinterpret_cast<const uint4*>(aux_row) + c0 + c);
                            uw[4 * c] = t4.x; uw[4 * c + 1] = t4.y; uw[4 * c + 2] = t4.z; uw[4 * c + 3] = t4.w;
                        }
                    };
                    if constexpr (EPI == EPI_DACT && EW == 8) {
#pragma unroll
                        for (int c = 0; c < 8; ++c) {
                            uint4 t4 = make_uint4(0u, 0u, 0u, 0u);
                            if (use_aux && row_ok) t4 = __ldg(reinterpret_cast<const uint4*>(aux_row) + c);
                            uw[4 * c] = t4.x; uw[4 * c + 1] = t4.y; uw[4 * c + 2] = t4.z; uw[4 * c + 3] = t4.w;
                        }
                    }
                    // (the accumulator half is read from tensor memory where it is used -- twice when the pre-activation copy
                    //  is kept -- instead of holding 64 values across both passes: the 16-warp variant has 96 registers)
                    if constexpr (EPI == EPI_BIAS_ACT) {
                        if (use_aux) {  // pass 1: the pre-activation acc + bias (bf16) through the slab into p.aux
#pragma unroll
                            for (int hx = 0; hx < 2; ++hx) {
                                uint32_t v[32];
                                load32(sl * 64 + hx * 32, v);
#pragma unroll
                                for (int j = 0; j < 32; j += 8) {
                                    float f[8];
#pragma unroll
                                    for (int q = 0; q < 8; q += 4) {
                                        const float4 b4 = add_bias ? __ldg(reinterpret_cast<const float4*>(p.bias + n0 + hx * 32 + j + q))
                                                                   : make_float4(0.f, 0.f, 0.f, 0.f);
                                        f[q + 0] = __uint_as_float(v[j + q + 0]) + b4.x;
                                        f[q + 1] = __uint_as_float(v[j + q + 1]) + b4.y;
                                        f[q + 2] = __uint_as_float(v[j + q + 2]) + b4.z;
                                        f[q + 3] = __uint_as_float(v[j + q + 3]) + b4.w;
                                    }
                                    st_slab_chunk(slab_row[buf], lane, hx * 4 + j / 8, pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]),
                                                  pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
                                }
                            }
                            __syncwarp();
                            slab_to_global<8>(smem_u32(my_stage + buf * GEMM_SLAB_BYTES), lane, aux_slab + n0, p.ld_aux, rows_ok);
                            __syncwarp();
                        }
                    }
#pragma unroll
                    for (int hx = 0; hx < 2; ++hx) {
                        if constexpr (EPI == EPI_DACT && EW != 8) load_u(hx * 4, 4);
                        uint32_t v[32];
                        load32(sl * 64 + hx * 32, v);
#pragma unroll
                        for (int j = 0; j < 32; j += 8) {
                            float f[8];
#pragma unroll
                            for (int q = 0; q < 8; q += 4) {
                                const float4 b4 = add_bias ? __ldg(reinterpret_cast<const float4*>(p.bias + n0 + hx * 32 + j + q))
                                                           : make_float4(0.f, 0.f, 0.f, 0.f);
                                f[q + 0] = __uint_as_float(v[j + q + 0]) + b4.x;
                                f[q + 1] = __uint_as_float(v[j + q + 1]) + b4.y;
                                f[q + 2] = __uint_as_float(v[j + q + 2]) + b4.z;
                                f[q + 3] = __uint_as_float(v[j + q + 3]) + b4.w;
                            }
                            if constexpr (EPI == EPI_DACT) {
#pragma unroll
                                for (int q = 0; q < 4; ++q) {
                                    const uint32_t w2 = uw[(EW == 8 ? hx * 16 : 0) + j / 2 + q];
                                    f[2 * q] *= dact<ACT>(__uint_as_float(w2 << 16));
                                    f[2 * q + 1] *= dact<ACT>(__uint_as_float(w2 & 0xffff0000u));
                                }
                            } else {
#pragma unroll
                                for (int q = 0; q < 8; ++q) f[q] = apply_act<ACT>(f[q]);
                            }
                            if (p.drop.thresh != 0u) {
#pragma unroll
                                for (int q = 0; q < 8; q += 2) {
                                    float m0, m1;
                                    drop_scale2(p.drop, (uint32_t)m, (uint32_t)(n0 + hx * 32 + j + q), m0, m1);
                                    f[q] *= m0; f[q + 1] *= m1;
                                }
                            }
                            st_slab_chunk(slab_row[buf], lane, hx * 4 + j / 8, pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]),
                                          pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
                        }
                    }
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        if (row0 < p.M) tma_store_2d(&tmap_out, my_stage + buf * GEMM_SLAB_BYTES, n0, row0);
                        tma_store_commit();
                    }
                    if constexpr (SLABS == 2) buf ^= 1;
                }
            } else if constexpr (EPI == EPI_BIAS_ACT || EPI == EPI_GATE_RESIDUAL || EPI == EPI_ACCUM) {
                // f32 slabs of 32 columns.  GATE_RESIDUAL: slab = gate * (acc + bias), TMA-reduce-added into x.
                // ACCUM: slab = acc (+ bias), TMA-reduce-added into out (gradient accumulation, split-K partial sums).
                const float* grow = nullptr;
                float gsc = 1.0f;
                if constexpr (EPI == EPI_GATE_RESIDUAL) {
                    const int bi = row_ok ? (m / p.tokens_per_batch) : 0;
                    grow = p.gate + (long long)bi * p.gate_bstride;
                    if (p.gate_rowscale != nullptr) gsc = __ldg(p.gate_rowscale + bi);
                }
#pragma unroll 1
                for (int sl = 0; sl < HALF / 32; ++sl) {
                    const int n0 = n_base + sl * 32;
                    if (lane == 0) tma_store_wait_read<1>();
                    __syncwarp();
                    uint32_t v[32];
                    load32(sl * 32, v);
                    bool aux_pass = false;
                    if constexpr (EPI == EPI_GATE_RESIDUAL) aux_pass = p.aux != nullptr;  // training forward (warp-uniform)
                    if (aux_pass) {
                        // y = (acc + bias) * mask, kept in bf16 for the gate gradient: through the slab (32 bf16 = 64 B per
                        // row = 4 chunks) and out with warp-cooperative coalesced stores; then the gated f32 values as usual
                        float r[32];
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const float4 b4 = add_bias ? __ldg(reinterpret_cast<const float4*>(p.bias + n0 + j))
                                                     : make_float4(0.f, 0.f, 0.f, 0.f);
                            r[j] = __uint_as_float(v[j + 0]) + b4.x; r[j + 1] = __uint_as_float(v[j + 1]) + b4.y;
                            r[j + 2] = __uint_as_float(v[j + 2]) + b4.z; r[j + 3] = __uint_as_float(v[j + 3]) + b4.w;
                            if (p.drop.thresh != 0u) {
                                float m0, m1, m2, m3;
                                drop_scale2(p.drop, (uint32_t)m, (uint32_t)(n0 + j), m0, m1);
                                drop_scale2(p.drop, (uint32_t)m, (uint32_t)(n0 + j + 2), m2, m3);
                                r[j] *= m0; r[j + 1] *= m1; r[j + 2] *= m2; r[j + 3] *= m3;
                            }
                        }
#pragma unroll
                        for (int c = 0; c < 4; ++c)
                            st_slab_chunk(slab_row[buf], lane, c, pack_bf16(r[8 * c], r[8 * c + 1]), pack_bf16(r[8 * c + 2], r[8 * c + 3]),
                                          pack_bf16(r[8 * c + 4], r[8 * c + 5]), pack_bf16(r[8 * c + 6], r[8 * c + 7]));
                        __syncwarp();
                        slab_to_global<4>(smem_u32(my_stage + buf * GEMM_SLAB_BYTES), lane,
                                          reinterpret_cast<__nv_bfloat16*>(const_cast<void*>(p.aux)) + (long long)row0 * p.ld_aux + n0,
                                          p.ld_aux, p.M - row0);
                        __syncwarp();
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const float4 g4 = __ldg(reinterpret_cast<const float4*>(grow + n0 + j));
                            st_slab_chunk(slab_row[buf], lane, j / 4, __float_as_uint(r[j] * (g4.x * gsc)), __float_as_uint(r[j + 1] * (g4.y * gsc)),
                                          __float_as_uint(r[j + 2] * (g4.z * gsc)), __float_as_uint(r[j + 3] * (g4.w * gsc)));
                        }
                    } else {
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const float4 b4 = add_bias ? __ldg(reinterpret_cast<const float4*>(p.bias + n0 + j))
                                                 : make_float4(0.f, 0.f, 0.f, 0.f);
                        float r0 = __uint_as_float(v[j + 0]) + b4.x, r1 = __uint_as_float(v[j + 1]) + b4.y,
                              r2 = __uint_as_float(v[j + 2]) + b4.z, r3 = __uint_as_float(v[j + 3]) + b4.w;
                        if constexpr (EPI == EPI_GATE_RESIDUAL) {
                            if (p.drop.thresh != 0u) {
                                float m0, m1, m2, m3;
                                drop_scale2(p.drop, (uint32_t)m, (uint32_t)(n0 + j), m0, m1);
                                drop_scale2(p.drop, (uint32_t)m, (uint32_t)(n0 + j + 2), m2, m3);
                                r0 *= m0; r1 *= m1; r2 *= m2; r3 *= m3;
                            }
                            const float4 g4 = __ldg(reinterpret_cast<const float4*>(grow + n0 + j));
                            r0 *= g4.x * gsc; r1 *= g4.y * gsc; r2 *= g4.z * gsc; r3 *= g4.w * gsc;
                        } else if constexpr (EPI == EPI_BIAS_ACT) {
                            r0 = apply_act<ACT>(r0); r1 = apply_act<ACT>(r1); r2 = apply_act<ACT>(r2); r3 = apply_act<ACT>(r3);
                        }
                        st_slab_chunk(slab_row[buf], lane, j / 4, __float_as_uint(r0), __float_as_uint(r1),
                                      __float_as_uint(r2), __float_as_uint(r3));
                    }
                    }
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        if (row0 < p.M) {
                            if constexpr (EPI == EPI_GATE_RESIDUAL || EPI == EPI_ACCUM)
                                tma_reduce_add_2d(&tmap_out, my_stage + buf * GEMM_SLAB_BYTES, n0, row0);
                            else
                                tma_store_2d(&tmap_out, my_stage + buf * GEMM_SLAB_BYTES, n0, row0);
                        }
                        tma_store_commit();
                    }
                    buf ^= 1;
                }
            } else if constexpr (EPI == EPI_QKV_ROPE) {
                // slab = one 64-wide head; rotate_half pairs column j with j+32 (jat_audiosr_v2.py:70-91).
                const int pos = row_ok ? (m % p.tokens_per_batch) : 0;
                const float* cosr = p.rope_cos + (long long)pos * 64;
                const float* sinr = p.rope_sin + (long long)pos * 64;
#pragma unroll 1
                for (int hh = 0; hh < HALF / 64; ++hh) {
                    const int n0 = n_base + hh * 64;
                    const bool rot = n0 < p.rope_cols;
                    if (lane == 0) tma_store_wait_read<1>();
                    __syncwarp();
#pragma unroll
                    for (int s = 0; s < 2; ++s) {
                        uint32_t lo[16], hi[16];
                        load16(hh * 64 + s * 16, lo);
                        load16(hh * 64 + 32 + s * 16, hi);
                        tmem_ld_wait();
                        float ol[16], oh[16];
#pragma unroll
                        for (int j = 0; j < 16; j += 4) {
                            float4 c4 = make_float4(1.f, 1.f, 1.f, 1.f), s4 = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (rot) {
                                c4 = __ldg(reinterpret_cast<const float4*>(cosr + s * 16 + j));
                                s4 = __ldg(reinterpret_cast<const float4*>(sinr + s * 16 + j));
                            }
                            const float cc[4] = {c4.x, c4.y, c4.z, c4.w}, ss[4] = {s4.x, s4.y, s4.z, s4.w};
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                const float a = __uint_as_float(lo[j + q]), b = __uint_as_float(hi[j + q]);
                                ol[j + q] = a * cc[q] - b * ss[q];
                                oh[j + q] = b * cc[q] + a * ss[q];
                            }
                        }
#pragma unroll
                        for (int c = 0; c < 2; ++c) {
                            st_slab_chunk(slab_row[buf], lane, s * 2 + c, pack_bf16(ol[c * 8 + 0], ol[c * 8 + 1]),
                                          pack_bf16(ol[c * 8 + 2], ol[c * 8 + 3]), pack_bf16(ol[c * 8 + 4], ol[c * 8 + 5]),
                                          pack_bf16(ol[c * 8 + 6], ol[c * 8 + 7]));
                            st_slab_chunk(slab_row[buf], lane, 4 + s * 2 + c, pack_bf16(oh[c * 8 + 0], oh[c * 8 + 1]),
                                          pack_bf16(oh[c * 8 + 2], oh[c * 8 + 3]), pack_bf16(oh[c * 8 + 4], oh[c * 8 + 5]),
                                          pack_bf16(oh[c * 8 + 6], oh[c * 8 + 7]));
                        }
                    }
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        if (row0 < p.M) tma_store_2d(&tmap_out, my_stage + buf * GEMM_SLAB_BYTES, n0, row0);
                        tma_store_commit();
                    }
                    buf ^= 1;
                }
            } else {  // EPI_UNPATCHIFY, patch_len 4: column = c*4 + p -> out[b, c, n*4 + p]
                const int b = row_ok ? (m / p.tokens_per_batch) : 0;
                const int n = row_ok ? (m % p.tokens_per_batch) : 0;
                const int C = p.N >> 2;
                const int t0 = n * 4;
                float* obase = reinterpret_cast<float*>(p.out) + (long long)b * C * p.t_out + t0;
                const bool vec_ok = (p.t_out & 1) == 0;
#pragma unroll 1
                for (int c = 0; c < HALF / 32; ++c) {
                    uint32_t v[32];
                    load32(c * 32, v);
                    const int n0 = n_base + c * 32;
                    if (row_ok) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const float4 b4 = add_bias ? __ldg(reinterpret_cast<const float4*>(p.bias + n0 + j))
                                                       : make_float4(0.f, 0.f, 0.f, 0.f);
                            const float r0 = __uint_as_float(v[j + 0]) + b4.x, r1 = __uint_as_float(v[j + 1]) + b4.y,
                                        r2 = __uint_as_float(v[j + 2]) + b4.z, r3 = __uint_as_float(v[j + 3]) + b4.w;
                            float* o = obase + (long long)((n0 + j) >> 2) * p.t_out;
                            if (vec_ok && t0 + 3 < p.t_out) {
                                *reinterpret_cast<float2*>(o) = make_float2(r0, r1);
                                *reinterpret_cast<float2*>(o + 2) = make_float2(r2, r3);
                            } else {
                                if (t0 + 0 < p.t_out) o[0] = r0;
                                if (t0 + 1 < p.t_out) o[1] = r1;
                                if (t0 + 2 < p.t_out) o[2] = r2;
                                if (t0 + 3 < p.t_out) o[3] = r3;
                            }
                        }
                    }
                }
            }

            if (wk.tail >= 0) continue;  // a split tail tile handed its accumulator stage back right after parking it
            // all TMEM reads of this accumulator stage are done -> hand it back to the MMA issuer
            tc_fence_before();
            __syncwarp();
            if (warp == 4 && lane == 0) GEMM_TRACE(work / num_clusters, 6);
            if (lane == 0) {
                if constexpr (CG == 1) mbar_arrive(&bar_tmem_empty[as]);
                else mbar_arrive_cluster(&bar_tmem_empty[as], cluster_rank & ~1u);
            }
            if ((as ^= 1) == 0) aphase ^= 1;
        }
        if (lane == 0) tma_store_wait_all<0>();  // slabs must stay valid until the TMA unit has drained them
        __syncwarp();
    }

    // ---------------------------------------------------------------------- teardown
    tc_fence_before();
    if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc<CG>(tmem_base, Cfg::TMEM_COLS);
    }
}

}  // namespace jat
