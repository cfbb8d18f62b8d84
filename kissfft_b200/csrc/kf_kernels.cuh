// kf_kernels.cuh -- sm_100a __global__ entry points: thin wrappers that bind the kernel bodies of kf_body.h
// to the CUDA execution environment.
#pragma once
#include <cuda_runtime.h>

#include "kf_body.h"

namespace kf {

struct DeviceEnv {
    unsigned char* smem_;
    __device__ __forceinline__ int tid() const { return (int)threadIdx.x; }
    __device__ __forceinline__ int nthreads() const { return (int)blockDim.x; }
    __device__ __forceinline__ long long bid() const { return (long long)blockIdx.x; }
    __device__ __forceinline__ long long nblocks() const { return (long long)gridDim.x; }
    __device__ __forceinline__ void sync() const { __syncthreads(); }
    __device__ __forceinline__ unsigned char* smem() const { return smem_; }
    template <class V>
    __device__ __forceinline__ V shfl(V v, int src_lane) const { return __shfl_sync(0xffffffffu, v, src_lane); }

    // ---- mbarrier + bulk asynchronous copy (TMA engine, SASS: UBLKCP / SYNCS) ----
    static __device__ __forceinline__ unsigned s32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
    __device__ __forceinline__ void mbar_init(void* bar) const
    {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(bar)) : "memory");
    }
    __device__ __forceinline__ void mbar_fence_init() const
    {
        // make the initialised barriers visible to the async proxy before the first bulk copy signals them
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    // one elected thread: arm the barrier with the byte count, then start global -> shared
    __device__ __forceinline__ void bulk_load(void* bar, void* dst, const void* src, unsigned bytes) const
    {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bar)), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(dst)),
                     "l"(src), "r"(bytes), "r"(s32(bar))
                     : "memory");
    }
    // tensor-map variant (column modes): arm the barrier once for the whole tile, then one 3-D box per call --
    // box = [ncol adjacent columns][nrows rows][1 plane] of the array described by `tmap`, landing densely as
    // [row][column].  The trailing arguments describe the same box for the CPU emulator and are unused here.
    __device__ __forceinline__ void mbar_expect(void* bar, unsigned bytes) const
    {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bar)), "r"(bytes) : "memory");
    }
    __device__ __forceinline__ void tensor_load_box(void* bar, void* dst, const void* tmap, long long col0, int row0, long long plane,
                                                    const void*, long long, int, int, int elem_bytes, bool) const
    {
        // the map counts 4-byte words along the contiguous dimension
        const int x = (int)(col0 * (elem_bytes / 4)), y = row0, z = (int)plane;
        asm volatile(
            "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(s32(dst)),
            "l"(tmap), "r"(x), "r"(y), "r"(z), "r"(s32(bar))
            : "memory");
    }
    // all threads: wait for the k-th completion of the barrier (phase parity k & 1)
    __device__ __forceinline__ void mbar_wait(void* bar, int k) const
    {
        const unsigned parity = (unsigned)k & 1u, addr = s32(bar);
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "KF_WAIT_%=:\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
            "@p bra KF_DONE_%=;\n\t"
            "bra KF_WAIT_%=;\n\t"
            "KF_DONE_%=:\n\t"
            "}" ::"r"(addr),
            "r"(parity)
            : "memory");
    }
};

// PT is a tag type carrying the plan as `static constexpr PlanDesc D` (a class-type non-type template argument
// cannot appear in a __global__ signature: nvcc's host stub cannot spell it).
template <class A, class PT, int MODE>
__global__ void __launch_bounds__(PT::D.threads(), PT::D.minblocks) kf_fused_kernel(const __grid_constant__ KParams<A> P)
{
    extern __shared__ __align__(16) unsigned char kf_smem_raw[];
    DeviceEnv env{kf_smem_raw};
    fused_body<A, PT, MODE>(P, env);
}

template <class A>
__global__ void __launch_bounds__(256) kf_generic_kernel(const __grid_constant__ GParams<A> G)
{
    extern __shared__ __align__(16) unsigned char kf_smem_raw[];
    DeviceEnv env{kf_smem_raw};
    generic_body<A>(G, env);
}

template <class A, class PT>
__global__ void __launch_bounds__(PT::D.threads(), PT::D.minblocks) kf_fastconv_kernel(const __grid_constant__ FCParams<A> P)
{
    extern __shared__ __align__(16) unsigned char kf_smem_raw[];
    DeviceEnv env{kf_smem_raw};
    fastconv_body<A, PT>(P, env);
}

template <class A>
__global__ void __launch_bounds__(256) kf_stage_kernel(const __grid_constant__ StageParams<A> S)
{
    DeviceEnv env{nullptr};
    stage_body<A>(S, env);
}

template <class A>
__global__ void __launch_bounds__(256) kf_realpass_kernel(const __grid_constant__ RealPassParams<A> S)
{
    DeviceEnv env{nullptr};
    realpass_body<A>(S, env);
}

// overlap-scrap block gather (unfused fast convolution): out[b][i] = in[b*advance + i], i < len.  T = scalar or complex.
template <class T>
__global__ void __launch_bounds__(256) kf_gather_blocks_kernel(const T* __restrict__ in, T* __restrict__ out, long long nblocks, int len,
                                                              long long advance)
{
    const long long total = nblocks * len, step = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += step) {
        const long long b = i / len;
        out[i] = in[b * advance + (i - b * len)];
    }
}

// x[b][k] = C_MUL(x[b][k], h[k]) for b < rows, k < n (fastconv1buf, tools/kiss_fastfir.c:171-176); float / double
template <class A>
__global__ void __launch_bounds__(256) kf_cmul_rows_kernel(typename A::C* __restrict__ x, const typename A::C* __restrict__ h, long long rows, int n)
{
    const long long total = rows * n, step = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += step)
        x[i] = A::store(A::cmul(A::load(x[i]), A::load(h[i % n])));
}

// tiled transpose of a rows x cols array of storage complexes (32 x 32 tiles, +1 column of padding)
template <class C>
__global__ void __launch_bounds__(256) kf_transpose_kernel(const C* __restrict__ in, C* __restrict__ out, long long rows,
                                                          long long cols)
{
    __shared__ C tile[32][33];
    const long long tiles_c = (cols + 31) / 32, tiles_r = (rows + 31) / 32;
    const int tx = threadIdx.x % 32, ty = threadIdx.x / 32;   // 32 x 8
    for (long long tI = blockIdx.x; tI < tiles_c * tiles_r; tI += gridDim.x) {
        const long long r0 = (tI / tiles_c) * 32, c0 = (tI % tiles_c) * 32;
        for (int j = ty; j < 32; j += 8)
            if (r0 + j < rows && c0 + tx < cols) tile[j][tx] = in[(r0 + j) * cols + c0 + tx];
        __syncthreads();
        for (int j = ty; j < 32; j += 8)
            if (c0 + j < cols && r0 + tx < rows) out[(c0 + j) * rows + r0 + tx] = tile[tx][j];
        __syncthreads();
    }
}

// the same tiles with the columns split into npeers blocks: block s (columns s*cols_per_peer + [0, cols_per_peer) of a
// rows x (npeers*cols_per_peer) array with row pitch in_pitch) lands transposed in peers.p[s] as
// [cols_per_peer][out_pitch], shifted by out_off elements inside every output row.  Exchange step of the 2-D slab transform
// (kf_mgpu.c): every destination row piece is `rows` contiguous elements, written straight into the owner's memory.
struct KfPeerPtrs {
    void* p[16];
};
template <class C>
__global__ void __launch_bounds__(256) kf_transpose_peers_kernel(const C* __restrict__ in, long long in_pitch, KfPeerPtrs peers, int npeers,
                                                                long long rows, long long cols_per_peer, long long out_pitch, long long out_off)
{
    __shared__ C tile[32][33];
    const long long tiles_c = (cols_per_peer + 31) / 32, tiles_r = (rows + 31) / 32, per = tiles_c * tiles_r;
    const int tx = threadIdx.x % 32, ty = threadIdx.x / 32;   // 32 x 8
    for (long long tI = blockIdx.x; tI < per * npeers; tI += gridDim.x) {
        const int s = (int)(tI / per);
        const long long q = tI % per, r0 = (q / tiles_c) * 32, c0 = (q % tiles_c) * 32;
        C* const out = (C*)peers.p[s];
        for (int j = ty; j < 32; j += 8)
            if (r0 + j < rows && c0 + tx < cols_per_peer) tile[j][tx] = in[(r0 + j) * in_pitch + s * cols_per_peer + c0 + tx];
        __syncthreads();
        for (int j = ty; j < 32; j += 8)
            if (c0 + j < cols_per_peer && r0 + tx < rows) out[(c0 + j) * out_pitch + out_off + r0 + tx] = tile[tx][j];
        __syncthreads();
    }
}

}   // namespace kf
