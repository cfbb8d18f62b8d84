// intbench.cu -- microbenchmark (development aid): issue throughput of the integer instructions the Q15 / Q31 kernels are
// made of, on sm_100a: which pipe each runs on decides how the fixed-point arithmetic is best written (kf_math.h).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_build/intbench tools/intbench.cu
#include <cuda_runtime.h>
#include <stdio.h>
#define OPS(X) X(IMAD) X(IMAD_HI) X(IMAD_WIDE) X(IMAD_WIDE_ACC) X(SHF) X(LEA_HI_SX) X(PRMT) X(IADD3) X(IADD64) X(MIX_IMAD_SHF) X(MIX_IMAD_IMADHI) X(FUNNEL)
enum Op {
#define E(n) n,
    OPS(E)
#undef E
    NOPS
};
template <int MODE>
__global__ void k(int* out, int iters, int seed, int k17)
{
    int a[8], b = seed | 1;
    long long w[8];
    for (int i = 0; i < 8; ++i) { a[i] = threadIdx.x * 7 + i; w[i] = a[i]; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == IMAD) asm volatile("mad.lo.s32 %0, %0, %1, %1;" : "+r"(a[i]) : "r"(b));
            if (MODE == IMAD_HI) asm volatile("mad.hi.s32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(k17), "r"(b));
            if (MODE == IMAD_WIDE) asm volatile("mul.wide.s32 %0, %1, %2;" : "=l"(w[i]) : "r"((int)w[i]), "r"(b));
            if (MODE == IMAD_WIDE_ACC) asm volatile("mad.wide.s32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(a[i]), "r"(b));
            if (MODE == SHF) asm volatile("shr.s32 %0, %0, 3; add.s32 %0, %0, %1;" : "+r"(a[i]) : "r"(b));   // fuses to LEA.HI.SX32 or SHF+IADD
            if (MODE == LEA_HI_SX) a[i] = (a[i] >> 15) + b;
            if (MODE == PRMT) asm volatile("prmt.b32 %0, %0, %1, 0x9910;" : "+r"(a[i]) : "r"(b));
            if (MODE == IADD3) asm volatile("add.s32 %0, %0, %1;" : "+r"(a[i]) : "r"(b));
            if (MODE == IADD64) asm volatile("add.s64 %0, %0, %1;" : "+l"(w[i]) : "l"((long long)b));
            if (MODE == MIX_IMAD_SHF) { asm volatile("mad.lo.s32 %0, %0, %1, %1;" : "+r"(a[i]) : "r"(b)); asm volatile("shr.s32 %0, %0, 15;" : "+r"(a[i])); }
            if (MODE == MIX_IMAD_IMADHI) { asm volatile("mad.lo.s32 %0, %0, %1, %1;" : "+r"(a[i]) : "r"(b)); asm volatile("mul.hi.s32 %0, %0, %1;" : "+r"(a[i]) : "r"(k17)); }
            if (MODE == FUNNEL) asm volatile("shf.l.wrap.b32 %0, %0, %1, 1;" : "+r"(a[i]) : "r"(b));
        }
    }
    int s = 0;
    for (int i = 0; i < 8; ++i) s += a[i] + (int)w[i] + (int)(w[i] >> 32);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE>
static void run(const char* name, int* out, int per_iter)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int iters = 20000;
    k<MODE><<<148 * 8, 256>>>(out, iters, 12345, 1 << 17);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<MODE><<<148 * 8, 256>>>(out, iters, 12345, 1 << 17);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double warp_instr = 148.0 * 8 * 8 * iters * 8.0 * per_iter;   // warps * iters * 8 slots * PTX ops per slot
    printf("{\"op\": \"%s\", \"ms\": %.3f, \"ptx_ops_per_clk_per_sm\": %.1f}\n", name, ms, warp_instr * 32 / (ms * 1e-3) / 148 / 1.965e9);
}
int main()
{
    int* out;
    cudaMalloc(&out, 148 * 8 * 256 * 4);
    run<IMAD>("IMAD", out, 1);
    run<IMAD_HI>("IMAD.HI (mad.hi, register multiplier)", out, 1);
    run<IMAD_WIDE>("IMAD.WIDE (mul.wide)", out, 1);
    run<IMAD_WIDE_ACC>("IMAD.WIDE with 64-bit addend (mad.wide)", out, 1);
    run<SHF>("shr+add (PTX pair)", out, 1);
    run<LEA_HI_SX>("(x >> 15) + y (C)", out, 1);
    run<PRMT>("PRMT", out, 1);
    run<IADD3>("IADD", out, 1);
    run<IADD64>("64-bit add", out, 1);
    run<MIX_IMAD_SHF>("IMAD then SHF (pair)", out, 2);
    run<MIX_IMAD_IMADHI>("IMAD then IMAD.HI (pair)", out, 2);
    run<FUNNEL>("SHF.L.W funnel", out, 1);
    return 0;
}
