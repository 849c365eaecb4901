// Micro-benchmark: how MUFU.EX2 and FMA-pipe instructions share an SM sub-partition's issue/dispatch
// bandwidth (decides whether the attention softmax is MUFU-bound or issue-bound).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o build_tmp/mbi scripts/microbench_issue.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(512, 1) k_mix(long long* out, float* sink, int iters) {
    float x[8], a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { x[i] = threadIdx.x * 1e-3f + i; a[i] = 0.f; }
    const float c1 = 0.99991f, c2 = -1e-4f;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) {  // MUFU only
                asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
            } else if (MODE == 1) {  // FFMA only (3-register form)
                asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[i]) : "f"(c1), "f"(c2));
            } else if (MODE == 2) {  // 1 MUFU + 1 FFMA + 1 FADD  (the softmax mix without the pack)
                float e;
                asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(e) : "f"(x[i]), "f"(c1), "f"(c2));
                asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(e));
                asm volatile("add.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(e));
            } else if (MODE == 3) {  // 1 MUFU + 3 FFMA
                float e;
                asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(e) : "f"(x[i]), "f"(c1), "f"(c2));
                asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(e));
                asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(a[i]) : "f"(e), "f"(c1));
                asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(a[i]) : "f"(e), "f"(c2));
            } else if (MODE == 4) {  // 2 FFMA + 2 FADD, no MUFU
                float e;
                asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(e) : "f"(x[i]), "f"(c1), "f"(c2));
                asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(e) : "f"(e), "f"(c1), "f"(c2));
                asm volatile("add.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(e));
                asm volatile("add.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(c2));
            } else if (MODE == 5) {  // FMNMX only (alu pipe)
                asm volatile("max.f32 %0, %0, %1;" : "+f"(x[i]) : "f"(a[i]));
            } else if (MODE == 6) {  // 1 MUFU + 1 FFMA + 1 FADD + 1 FMNMX + half a pack
                float e;
                asm volatile("max.f32 %0, %0, %1;" : "+f"(a[(i + 1) & 7]) : "f"(x[i]));
                asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(e) : "f"(x[i]), "f"(c1), "f"(c2));
                asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(e));
                asm volatile("add.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(e));
                if (i & 1) {
                    unsigned pk;
                    asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(pk) : "f"(e), "f"(a[i]));
                    x[i] = __uint_as_float(pk);
                }
            }
        }
    }
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i] + a[i];
    if (s == 123.f) sink[0] = s;
}

int main() {
    long long* d; float* sink;
    cudaMalloc(&d, 64); cudaMalloc(&sink, 64);
    long long h[2];
    const int iters = 2000;
    const char* names[] = {"MUFU only", "FFMA only", "MUFU+FFMA+FADD", "MUFU+3 FFMA", "2 FFMA + 2 FADD", "FMNMX only",
                           "MUFU+FFMA+FADD+FMNMX+0.5 F2FP"};
    auto run = [&](int mode, int threads) {
        switch (mode) {
            case 0: k_mix<0><<<148, threads>>>(d, sink, iters); break;
            case 1: k_mix<1><<<148, threads>>>(d, sink, iters); break;
            case 2: k_mix<2><<<148, threads>>>(d, sink, iters); break;
            case 3: k_mix<3><<<148, threads>>>(d, sink, iters); break;
            case 4: k_mix<4><<<148, threads>>>(d, sink, iters); break;
            case 5: k_mix<5><<<148, threads>>>(d, sink, iters); break;
            case 6: k_mix<6><<<148, threads>>>(d, sink, iters); break;
        }
        cudaDeviceSynchronize();
        cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        // cycles per "element" (one unrolled body) per warp, and per SMSP (warps per SMSP = threads / 128)
        const double per_elem = (double)h[0] / (iters * 8.0);
        printf("%-32s warps/SMSP=%d : %.2f clk per element per warp, %.2f clk per element per SMSP\n", names[mode],
               threads / 128, per_elem, per_elem / (threads / 128));
    };
    for (int mode = 0; mode < 7; ++mode)
        for (int threads : {128, 256, 512}) run(mode, threads);
    return 0;
}
