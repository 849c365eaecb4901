// Micro-benchmarks of the sm_100a resources the attention kernel is limited by (run on the GPU box):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o /tmp/mb scripts/microbench_sm100.cu && /tmp/mb
// 1. tcgen05.ld read-out bandwidth of TMEM (4 and 8 warps per SM)
// 2. MUFU.EX2 throughput (8 warps per SM)
// 3. tcgen05.mma issue/execute time of the attention's MMA chains (SS mode, garbage operands)
#include <cstdio>
#include <cuda_runtime.h>
#include "../jatsr-just-audio-transformer-super-solution_b200/csrc/common.cuh"
using namespace jat;

__global__ void __launch_bounds__(256, 1) k_ldtm(long long* out, int iters, int width) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) { tmem_alloc<1>(&slot, 512); tmem_relinquish<1>(); }
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t t_row = slot + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t acc = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (width == 32) {
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                uint32_t v[32];
                tmem_ld_32x32(t_row + (uint32_t)(((warp >> 2) * 256) + c * 32), v);
                tmem_ld_wait_dep32(v);
                acc ^= v[0] ^ v[31];
            }
        } else {
            uint32_t v[8][32];
#pragma unroll
            for (int c = 0; c < 8; ++c) tmem_ld_32x32(t_row + (uint32_t)(((warp >> 2) * 256) + c * 32), v[c]);
            tmem_ld_wait();
#pragma unroll
            for (int c = 0; c < 8; ++c) acc ^= v[c][0] ^ v[c][31];
        }
    }
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
    if (acc == 0x12345) out[1] = acc;
    tc_fence_before(); __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc<1>(slot, 512); }
}

__global__ void __launch_bounds__(256, 1) k_mufu(long long* out, float* sink, int iters) {
    float x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 1e-3f + i;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
    }
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i];
    if (s == 123.f) sink[0] = s;
}

// chain of `count` MMAs with the given N and k-stride pattern, operands = whatever is in smem
template <int N, int B_MAJOR>
__global__ void __launch_bounds__(128, 1) k_mma(long long* out, int count, int reps) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint32_t slot;
    __shared__ uint64_t bar;
    const uint32_t raw = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) { tmem_alloc<1>(&slot, 512); tmem_relinquish<1>(); }
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    fence_proxy_async_smem();
    tc_fence_before(); __syncthreads(); tc_fence_after();
    if (warp == 1 && lane == 0) {
        constexpr uint32_t idesc = umma_idesc_bf16(128, N, 0, B_MAJOR);
        const uint64_t a_desc = umma_smem_desc_sw128(smem_u32(smem));
        const uint64_t b_desc = umma_smem_desc_sw128(smem_u32(smem) + 98304u);
        long long t_issue = 0, t_total = 0;
        for (int r = 0; r < reps; ++r) {
            const long long t0 = clock64();
#pragma unroll 1
            for (int ks = 0; ks < count; ++ks)
                umma_bf16_ss<1>(slot, a_desc + (uint64_t)((ks >> 2) * 1024 + 2 * (ks & 3)),
                                b_desc + (uint64_t)(B_MAJOR ? ks * 128 : 2 * (ks & 3)), idesc, (uint32_t)(ks != 0));
            umma_commit(&bar);
            const long long t1 = clock64();
            mbar_wait(&bar, (uint32_t)(r & 1));
            const long long t2 = clock64();
            t_issue += t1 - t0; t_total += t2 - t0;
        }
        if (blockIdx.x == 0) { out[0] = t_issue / reps; out[1] = t_total / reps; }
    }
    tc_fence_before(); __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc<1>(slot, 512); }
}

__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// TS mode: A (128 x 16 bf16 = 8 packed columns per k-step) from TMEM, B from smem
template <int N, int B_MAJOR, int M>
__global__ void __launch_bounds__(128, 1) k_mma_ts(long long* out, int count, int reps, int ss) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint32_t slot;
    __shared__ uint64_t bar;
    const uint32_t raw = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) { tmem_alloc<1>(&slot, 512); tmem_relinquish<1>(); }
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    fence_proxy_async_smem();
    tc_fence_before(); __syncthreads(); tc_fence_after();
    if (warp == 1 && lane == 0) {
        constexpr uint32_t idesc = umma_idesc_bf16(M, N, 0, B_MAJOR);
        const uint64_t a_desc = umma_smem_desc_sw128(smem_u32(smem));
        const uint64_t b_desc = umma_smem_desc_sw128(smem_u32(smem) + 98304u);
        long long t_issue = 0, t_total = 0;
        for (int r = 0; r < reps; ++r) {
            const long long t0 = clock64();
#pragma unroll 1
            for (int ks = 0; ks < count; ++ks) {
                const uint64_t bd = b_desc + (uint64_t)(B_MAJOR ? ks * 128 : 2 * (ks & 3));
                if (ss) umma_bf16_ss<1>(slot, a_desc + (uint64_t)((ks >> 2) * 1024 + 2 * (ks & 3)), bd, idesc, (uint32_t)(ks != 0));
                else umma_bf16_ts(slot, slot + 256 + (uint32_t)(ks * 8), bd, idesc, (uint32_t)(ks != 0));
            }
            umma_commit(&bar);
            const long long t1 = clock64();
            mbar_wait(&bar, (uint32_t)(r & 1));
            const long long t2 = clock64();
            t_issue += t1 - t0; t_total += t2 - t0;
        }
        if (blockIdx.x == 0) { out[0] = t_issue / reps; out[1] = t_total / reps; }
    }
    tc_fence_before(); __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc<1>(slot, 512); }
}

// fully unrolled chain with compile-time descriptor offsets (what the production kernels should look like)
template <int N, int B_MAJOR, int COUNT, int TS>
__global__ void __launch_bounds__(128, 1) k_mma_unrolled(long long* out, int reps) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint32_t slot;
    __shared__ uint64_t bar;
    const uint32_t raw = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) { tmem_alloc<1>(&slot, 512); tmem_relinquish<1>(); }
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    fence_proxy_async_smem();
    tc_fence_before(); __syncthreads(); tc_fence_after();
    if (warp == 1 && lane == 0) {
        constexpr uint32_t idesc = umma_idesc_bf16(128, N, 0, B_MAJOR);
        const uint64_t a_desc = umma_smem_desc_sw128(smem_u32(smem));
        const uint64_t b_desc = umma_smem_desc_sw128(smem_u32(smem) + 98304u);
        const uint32_t tm = slot;
        long long t_issue = 0, t_total = 0;
        for (int r = 0; r < reps; ++r) {
            const long long t0 = clock64();
#pragma unroll
            for (int ks = 0; ks < COUNT; ++ks) {
                const uint64_t bd = b_desc + (uint64_t)(B_MAJOR ? ks * 128 : 2 * (ks & 3));
                if (TS) umma_bf16_ts(tm, tm + 256 + (uint32_t)(ks * 8), bd, idesc, (uint32_t)(ks != 0));
                else umma_bf16_ss<1>(tm, a_desc + (uint64_t)((ks >> 2) * 1024 + 2 * (ks & 3)), bd, idesc, (uint32_t)(ks != 0));
            }
            umma_commit(&bar);
            const long long t1 = clock64();
            mbar_wait(&bar, (uint32_t)(r & 1));
            const long long t2 = clock64();
            t_issue += t1 - t0; t_total += t2 - t0;
        }
        if (blockIdx.x == 0) { out[0] = t_issue / reps; out[1] = t_total / reps; }
    }
    tc_fence_before(); __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc<1>(slot, 512); }
}

int main() {
    long long* d; float* sink;
    cudaMalloc(&d, 64); cudaMalloc(&sink, 64);
    long long h[2];
    for (int width : {32, 256}) {
        for (int threads : {128, 256}) {
            const int iters = 200;
            k_ldtm<<<148, threads>>>(d, iters, width);
            cudaDeviceSynchronize();
            cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
            const double bytes = (double)iters * 8 * 32 * 128 * (threads / 32);
            printf("ldtm  %s warps=%d : %lld cycles, %.1f B/clk/SM\n", width == 32 ? "x32+wait each" : "8 x32 then wait",
                   threads / 32, h[0], bytes / h[0]);
        }
    }
    {
        const int iters = 2000;
        k_mufu<<<148, 256>>>(d, sink, iters);
        cudaDeviceSynchronize();
        cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        printf("mufu  ex2 8 warps: %lld cycles, %.2f lanes/clk/SM\n", h[0], (double)iters * 8 * 256 / h[0]);
    }
    cudaFuncSetAttribute(k_mma<64, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(k_mma<64, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(k_mma<176, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(k_mma<256, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(k_mma<128, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    auto run = [&](const char* name, auto kern, int count) {
        kern<<<148, 128, 200 * 1024>>>(d, count, 20);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        printf("mma   %-28s count=%2d : issue %lld cyc, issue+retire %lld cyc (%.1f / MMA) %s\n", name, count, h[0], h[1],
               (double)h[1] / count, e == cudaSuccess ? "" : cudaGetErrorString(e));
    };
    run("SS M128 N64 K16 B=MN-major", k_mma<64, 1>, 22);
    run("SS M128 N64 K16 B=K-major", k_mma<64, 0>, 22);
    run("SS M128 N128 K16", k_mma<128, 0>, 16);
    run("SS M128 N176 K16", k_mma<176, 0>, 8);
    run("SS M128 N256 K16", k_mma<256, 0>, 16);
    auto run2 = [&](const char* name, auto kern, int count, int ss) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        kern<<<148, 128, 200 * 1024>>>(d, count, 20, ss);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        printf("mma   %-34s count=%2d : issue %lld cyc, issue+retire %lld cyc (%.1f / MMA) %s\n", name, count, h[0], h[1],
               (double)h[1] / count, e == cudaSuccess ? "" : cudaGetErrorString(e));
    };
    run2("TS M128 N64 K16 B=MN-major", k_mma_ts<64, 1, 128>, 22, 0);
    run2("TS M128 N64 K16 B=K-major", k_mma_ts<64, 0, 128>, 22, 0);
    run2("TS M128 N176 K16", k_mma_ts<176, 0, 128>, 8, 0);
    run2("TS M128 N256 K16", k_mma_ts<256, 0, 128>, 16, 0);
    run2("SS M64 N128 K16", k_mma_ts<128, 0, 64>, 22, 1);
    run2("SS M64 N64 K16", k_mma_ts<64, 0, 64>, 22, 1);
    run2("SS M64 N256 K16", k_mma_ts<256, 0, 64>, 16, 1);
    auto run3 = [&](const char* name, auto kern, int count) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        kern<<<148, 128, 200 * 1024>>>(d, 20);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        printf("mma   %-34s count=%2d : issue %lld cyc, issue+retire %lld cyc (%.1f / MMA) %s\n", name, count, h[0], h[1],
               (double)h[1] / count, e == cudaSuccess ? "" : cudaGetErrorString(e));
    };
    run3("UNROLLED SS M128 N64 B=MN", k_mma_unrolled<64, 1, 22, 0>, 22);
    run3("UNROLLED TS M128 N64 B=MN", k_mma_unrolled<64, 1, 22, 1>, 22);
    run3("UNROLLED SS M128 N176", k_mma_unrolled<176, 0, 8, 0>, 8);
    run3("UNROLLED TS M128 N176", k_mma_unrolled<176, 0, 8, 1>, 8);
    run3("UNROLLED SS M128 N256", k_mma_unrolled<256, 0, 16, 0>, 16);
    run3("UNROLLED TS M128 N256", k_mma_unrolled<256, 0, 16, 1>, 16);
    return 0;
}
