// pcie_probe.cu -- ceiling of the host<->device path on this box: pinned cudaMemcpyAsync, one direction alone and both at
// once, whole buffers and 32 MiB chunks on several streams.  The e2e number of bench.py (kiss_fftr_batch + kiss_fftri_batch on
// host buffers) is judged against this, not against the nominal PCIe rate.   nvcc -O2 -o pcie_probe pcie_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <chrono>
#include <vector>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main()
{
    const size_t n = (size_t)1 << 30, chunk = (size_t)32 << 20;
    char *h_a, *h_b, *d_a, *d_b;
    CK(cudaHostAlloc((void**)&h_a, n, cudaHostAllocDefault));
    CK(cudaHostAlloc((void**)&h_b, n, cudaHostAllocDefault));
    CK(cudaMalloc((void**)&d_a, n));
    CK(cudaMalloc((void**)&d_b, n));
    for (size_t i = 0; i < n; i += 4096) h_a[i] = h_b[i] = 1;
    cudaStream_t s[8];
    for (auto& x : s) CK(cudaStreamCreateWithFlags(&x, cudaStreamNonBlocking));
    auto run = [&](const char* name, int mode, int lanes) -> int {
        double best = 1e9;
        for (int rep = 0; rep < 4; ++rep) {
            CK(cudaDeviceSynchronize());
            const double t0 = now();
            if (lanes == 0) {
                if (mode & 1) CK(cudaMemcpyAsync(d_a, h_a, n, cudaMemcpyHostToDevice, s[0]));
                if (mode & 2) CK(cudaMemcpyAsync(h_b, d_b, n, cudaMemcpyDeviceToHost, s[1]));
            } else {
                size_t k = 0;
                for (size_t off = 0; off < n; off += chunk, ++k) {
                    if (mode & 1) CK(cudaMemcpyAsync(d_a + off, h_a + off, chunk, cudaMemcpyHostToDevice, s[k % lanes]));
                    if (mode & 2) CK(cudaMemcpyAsync(h_b + off, d_b + off, chunk, cudaMemcpyDeviceToHost, s[(k + 1) % lanes]));
                }
            }
            CK(cudaDeviceSynchronize());
            const double t = now() - t0;
            if (t < best) best = t;
        }
        printf("{\"case\": \"%s\", \"lanes\": %d, \"ms\": %.3f, \"GBps_per_direction\": %.2f}\n", name, lanes, best * 1e3, n / best / 1e9);
        return 0;
    };
    if (run("h2d alone", 1, 0) || run("d2h alone", 2, 0) || run("both, whole buffers", 3, 0) || run("both, 32 MiB chunks", 3, 2) ||
        run("both, 32 MiB chunks", 3, 4) || run("both, 32 MiB chunks", 3, 8))
        return 1;
    return 0;
}
