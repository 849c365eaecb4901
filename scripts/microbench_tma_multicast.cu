// Micro-benchmark: how many bytes per clock can the TMA unit deliver into shared memory per SM from L2, and does
// cluster multicast lift that figure?  (The GEMM's A/B operand stream needs 64 B/clk/SM at full tcgen05 rate with
// 256x256 CTA-pair tiles; the measured full-chip L2->SM cap is ~6300 B/clk = 42.6 B/clk/SM.)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o /tmp/mbt scripts/microbench_tma_multicast.cu -lcuda
// Modes (cluster size CS in {1, 2, 4}; every CTA RECEIVES a full 16 KB box per stage in all modes):
//   0  unicast, every CTA streams its own region
//   1  unicast, the CS CTAs of a cluster stream the SAME region in lock-step (does L2 de-duplicate?)
//   2  multicast: each CTA loads 1/CS of the box rows and multicasts it to all CTAs of the cluster
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>
#include "../jatsr-just-audio-transformer-super-solution_b200/csrc/common.cuh"
using namespace jat;

constexpr int STAGES = 8;
constexpr int BOX_ROWS = 128;           // 128 rows x 64 bf16 = 16 KB
constexpr int BOX_BYTES = BOX_ROWS * 128;

__device__ __forceinline__ void tma_load_2d_mc(void* dst, const void* tmap, uint64_t* bar, int c0, int c1, uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], "
        "[%2], %5;"
        ::"r"(smem_u32(dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask)
        : "memory");
}

template <int CS>
__global__ void __launch_bounds__(128, 1)
k_stream(const __grid_constant__ CUtensorMap tm_full, const __grid_constant__ CUtensorMap tm_part, long long* out, int iters,
         int mode, int rows_per_region) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * BOX_BYTES);
    uint64_t* empty = full + STAGES;
    const uint32_t rank = CS > 1 ? cluster_ctarank() : 0u;
    const int cluster_id = blockIdx.x / CS;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], CS); }
        fence_barrier_init();
    }
    if constexpr (CS > 1) cluster_sync_all(); else __syncthreads();
    const int region = (mode == 0) ? blockIdx.x : cluster_id;
    const int row0 = region * rows_per_region;
    const int boxes = rows_per_region / BOX_ROWS;
    long long t0 = 0, t1 = 0;
    if (threadIdx.x == 0) {
        t0 = clock64();
        int stage = 0; uint32_t phase = 0;
        // producer + consumer in one thread: issue up to STAGES ahead, consume in order
        int issued = 0, consumed = 0;
        uint32_t cphase = 0; int cstage = 0;
        while (consumed < iters) {
            while (issued < iters && issued - consumed < STAGES) {
                if (issued >= STAGES) mbar_wait(&empty[stage], phase ^ 1);   // all CTAs of the cluster are done with it
                const int r = row0 + (issued % boxes) * BOX_ROWS;
                mbar_expect_tx(&full[stage], BOX_BYTES);
                if (mode == 2 && CS > 1) {
                    constexpr int PART = BOX_ROWS / CS;
                    tma_load_2d_mc(smem + stage * BOX_BYTES + rank * PART * 128, &tm_part, &full[stage], 0, r + rank * PART,
                                   (uint16_t)((1u << CS) - 1));
                } else {
                    tma_load_2d(smem + stage * BOX_BYTES, &tm_full, &full[stage], 0, r);
                }
                ++issued;
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
            mbar_wait(&full[cstage], cphase);
            if (mode == 2 && CS > 1) {
                for (int c = 0; c < CS; ++c) mbar_arrive_cluster(&empty[cstage], c);
            } else {
                for (int c = 0; c < CS; ++c) mbar_arrive(&empty[cstage]);
            }
            ++consumed;
            if (++cstage == STAGES) { cstage = 0; cphase ^= 1; }
        }
        t1 = clock64();
    }
    if constexpr (CS > 1) cluster_sync_all(); else __syncthreads();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static CUtensorMap make_map(PFN_encodeTiled enc, void* ptr, uint64_t rows, uint32_t box_rows) {
    CUtensorMap tm;
    cuuint64_t gdim[2] = {64, rows};
    cuuint64_t gstr[1] = {128};
    cuuint32_t box[2] = {64, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, ptr, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
    return tm;
}

template <int CS>
static void run(PFN_encodeTiled enc, void* buf, uint64_t rows, int mode, int sms, long long* d_out) {
    const int iters = 4096;
    const int smem = STAGES * BOX_BYTES + 2 * STAGES * 8 + 1024 + 64;
    cudaFuncSetAttribute(k_stream<CS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    CUtensorMap tf = make_map(enc, buf, rows, BOX_ROWS), tp = make_map(enc, buf, rows, BOX_ROWS / CS);
    const int grid = sms / CS * CS;
    const int regions = (mode == 0) ? grid : grid / CS;
    int rows_per_region = (int)(rows / regions) / BOX_ROWS * BOX_ROWS;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    for (int rep = 0; rep < 2; ++rep) {   // first pass warms L2
        cudaError_t e = cudaLaunchKernelEx(&cfg, k_stream<CS>, tf, tp, d_out, iters, mode, rows_per_region);
        if (e != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(e)); return; }
        e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(e)); exit(1); }
    }
    long long h[256];
    cudaMemcpy(h, d_out, grid * sizeof(long long), cudaMemcpyDeviceToHost);
    double mx = 0, sum = 0;
    for (int i = 0; i < grid; ++i) { mx = h[i] > mx ? h[i] : mx; sum += h[i]; }
    const double bytes = (double)iters * BOX_BYTES;
    printf("{\"case\": \"tma_stream\", \"cluster\": %d, \"mode\": %d, \"region_MB\": %.1f, \"B_per_clk_per_SM_avg\": %.1f, "
           "\"B_per_clk_per_SM_slowest\": %.1f, \"chip_B_per_clk\": %.0f}\n",
           CS, mode, rows_per_region * 128.0 / 1e6, bytes / (sum / grid), bytes / mx, bytes / (sum / grid) * grid);
}

int main() {
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    PFN_encodeTiled enc = (PFN_encodeTiled)fn;
    const uint64_t rows = 1ull << 19;   // 512 Ki rows x 128 B = 64 MB (L2-resident after the warm-up pass)
    void* buf;
    cudaMalloc(&buf, rows * 128);
    cudaMemset(buf, 0, rows * 128);
    long long* d_out;
    cudaMalloc(&d_out, 256 * sizeof(long long));
    const int sms = prop.multiProcessorCount;
    run<1>(enc, buf, rows, 0, sms, d_out);
    run<2>(enc, buf, rows, 0, sms, d_out);
    run<2>(enc, buf, rows, 1, sms, d_out);
    run<2>(enc, buf, rows, 2, sms, d_out);
    run<4>(enc, buf, rows, 0, sms, d_out);
    run<4>(enc, buf, rows, 1, sms, d_out);
    run<4>(enc, buf, rows, 2, sms, d_out);
    return 0;
}
