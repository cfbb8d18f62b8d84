// yardstick.cu -- context numbers for the roofline discussion (development aid; NEVER linked into the product):
//   * a plain device-to-device copy kernel over the same byte counts as the transforms (what "100 %" looks like)
//   * cuFFT on the same batched configurations (the only GPU FFT yardstick on the box; kissfft has no GPU code)
#include <cuda_runtime.h>
#include <cufft.h>
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA %s @%d\n", cudaGetErrorString(e_), __LINE__); exit(2);} } while (0)
#define CF(x) do { cufftResult r_ = (x); if (r_ != CUFFT_SUCCESS) { fprintf(stderr, "cuFFT %d @%d\n", (int)r_, __LINE__); exit(3);} } while (0)

template <class V>
__global__ void copy_kernel(const V* __restrict__ in, V* __restrict__ out, size_t n)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = in[i];
}

template <class F>
static float time_it(F&& f, int iters = 10)
{
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; ++i) f();
    CK(cudaDeviceSynchronize());
    std::vector<float> ms(iters);
    for (int i = 0; i < iters; ++i) {
        CK(cudaEventRecord(e0));
        f();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms[i], e0, e1));
    }
    std::sort(ms.begin(), ms.end());
    return ms[iters / 2];
}

int main()
{
    const size_t bytes = 512ull << 20;
    void *a, *b;
    CK(cudaMalloc(&a, 2 * bytes + (64 << 20)));
    CK(cudaMalloc(&b, 2 * bytes + (64 << 20)));
    CK(cudaMemset(a, 0, 2 * bytes));
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    for (int bpsm : {2, 4, 8}) {
        float m8 = time_it([&] { copy_kernel<float2><<<sms * bpsm, 256>>>((const float2*)a, (float2*)b, bytes / 8); });
        float m16 = time_it([&] { copy_kernel<float4><<<sms * bpsm, 256>>>((const float4*)a, (float4*)b, bytes / 16); });
        printf("{\"yardstick\": \"copy 512MiB->512MiB\", \"ctas_per_sm\": %d, \"float2_gbs\": %.1f, \"float4_gbs\": %.1f}\n", bpsm,
               2.0 * bytes / m8 * 1e-6, 2.0 * bytes / m16 * 1e-6);
    }
    float mm = time_it([&] { CK(cudaMemcpyAsync(b, a, bytes, cudaMemcpyDeviceToDevice)); });
    printf("{\"yardstick\": \"cudaMemcpy D2D 512MiB\", \"gbs\": %.1f}\n", 2.0 * bytes / mm * 1e-6);
    struct Cfg { const char* name; int n; int batch; cufftType type; double bytes; };
    Cfg cfgs[] = {
        {"cufft C2C f32 1024 x 65536", 1024, 65536, CUFFT_C2C, 2.0 * 1024 * 8 * 65536},
        {"cufft R2C f32 4096 x 32768", 4096, 32768, CUFFT_R2C, (4096.0 * 4 + 2049 * 8) * 32768},
        {"cufft C2R f32 4096 x 32768", 4096, 32768, CUFFT_C2R, (4096.0 * 4 + 2049 * 8) * 32768},
        {"cufft C2C f32 1000 x 100000", 1000, 100000, CUFFT_C2C, 2.0 * 1000 * 8 * 100000},
        {"cufft C2C f32 1155 x 100000", 1155, 100000, CUFFT_C2C, 2.0 * 1155 * 8 * 100000},
        {"cufft Z2Z f64 1000 x 32768", 1000, 32768, CUFFT_Z2Z, 2.0 * 1000 * 16 * 32768},
        {"cufft Z2Z f64 1155 x 32768", 1155, 32768, CUFFT_Z2Z, 2.0 * 1155 * 16 * 32768},
        {"cufft C2C f32 2048 x 65536", 2048, 65536, CUFFT_C2C, 2.0 * 2048 * 8 * 65536},
    };
    for (auto& c : cfgs) {
        cufftHandle h;
        int n[1] = {c.n};
        CF(cufftPlanMany(&h, 1, n, nullptr, 1, 0, nullptr, 1, 0, c.type, c.batch));
        float ms = time_it([&] {
            if (c.type == CUFFT_C2C) CF(cufftExecC2C(h, (cufftComplex*)a, (cufftComplex*)b, CUFFT_FORWARD));
            else if (c.type == CUFFT_R2C) CF(cufftExecR2C(h, (cufftReal*)a, (cufftComplex*)b));
            else if (c.type == CUFFT_C2R) CF(cufftExecC2R(h, (cufftComplex*)a, (cufftReal*)b));
            else CF(cufftExecZ2Z(h, (cufftDoubleComplex*)a, (cufftDoubleComplex*)b, CUFFT_FORWARD));
        });
        printf("{\"yardstick\": \"%s\", \"ms\": %.4f, \"gbs\": %.1f}\n", c.name, ms, c.bytes / ms * 1e-6);
        cufftDestroy(h);
    }
    {   // 3-D 512^3 and 1024^3 C2C
        for (int d : {256, 512, 1024}) {
            void *x, *y;
            size_t nb = (size_t)d * d * d * 8;
            if (cudaMalloc(&x, nb) != cudaSuccess || cudaMalloc(&y, nb) != cudaSuccess) { cudaGetLastError(); break; }
            cufftHandle h;
            if (cufftPlan3d(&h, d, d, d, CUFFT_C2C) != CUFFT_SUCCESS) { cudaFree(x); cudaFree(y); break; }
            float ms = time_it([&] { CF(cufftExecC2C(h, (cufftComplex*)x, (cufftComplex*)y, CUFFT_FORWARD)); }, 5);
            printf("{\"yardstick\": \"cufft C2C f32 %d^3\", \"ms\": %.4f, \"gbs_3pass\": %.1f}\n", d, ms, 6.0 * nb / ms * 1e-6);
            cufftDestroy(h);
            cudaFree(x);
            cudaFree(y);
        }
    }
    return 0;
}
