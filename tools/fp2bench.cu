// fp2bench.cu -- microbenchmark (development aid): issue throughput of scalar FADD/FFMA vs packed FADD2/FFMA2 on sm_100a
#include <cuda_runtime.h>
#include <stdio.h>
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) {
    unsigned long long d; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
template <int MODE>
__global__ void k(float* out, int iters, float seed) {
    float a[8], b = seed;
    unsigned long long p[8], q = __float_as_uint(seed) | ((unsigned long long)__float_as_uint(seed) << 32);
    for (int i = 0; i < 8; ++i) { a[i] = threadIdx.x * 0.001f + i; p[i] = q + i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) a[i] = a[i] + b;
            if (MODE == 1) a[i] = fmaf(a[i], b, b);
            if (MODE == 2) p[i] = add2(p[i], q);
            if (MODE == 3) p[i] = fma2(p[i], q, q);
        }
    }
    float s = 0;
    for (int i = 0; i < 8; ++i) s += a[i] + __uint_as_float((unsigned)p[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
    const char* names[] = {"FADD", "FFMA", "FADD2", "FFMA2"};
    for (int m = 0; m < 4; ++m) {
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        const int iters = 20000;
        auto launch = [&] {
            if (m == 0) k<0><<<148 * 8, 256>>>(out, iters, 1.0001f);
            if (m == 1) k<1><<<148 * 8, 256>>>(out, iters, 1.0001f);
            if (m == 2) k<2><<<148 * 8, 256>>>(out, iters, 1.0001f);
            if (m == 3) k<3><<<148 * 8, 256>>>(out, iters, 1.0001f);
        };
        launch(); cudaDeviceSynchronize();
        cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double warp_instr = 148.0 * 8 * 8 * iters * 8;   // warps * iters * 8 instr
        printf("{\"op\": \"%s\", \"ms\": %.3f, \"warp_instr_per_ns\": %.2f, \"lane_results_per_ns\": %.1f}\n", names[m], ms,
               warp_instr / (ms * 1e6), warp_instr * 32 * (m >= 2 ? 2 : 1) / (ms * 1e6));
    }
    return 0;
}
