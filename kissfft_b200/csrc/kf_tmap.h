// kf_tmap.h -- host side of the tensor-map (TMA) input ring of the column modes: eligibility + CUtensorMap encoding.
//
// The array a column pass reads is [planes][nfft rows][columns] with the columns contiguous; a tile is `tpc` adjacent
// columns x nfft rows of one plane.  The map describes it in 4-byte words along the contiguous dimension so that one
// encoding serves every datatype (Q15 complex = 1 word, float / Q31 complex = 2, double complex = 4); a box is
// [tpc columns][<= 256 rows][1 plane] and lands densely in shared memory (no swizzle: the reader walks whole rows).
// cuTensorMapEncodeTiled is a driver-API entry point; it is fetched through the runtime (cudaGetDriverEntryPoint) so the
// library does not link libcuda.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "kf_body.h"

namespace kf {

typedef CUresult (*kf_tmap_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                      const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                      CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline kf_tmap_encode_fn tmap_encoder()
{
    static kf_tmap_encode_fn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (kf_tmap_encode_fn)p;
    }
    return fn;
}

// may the column-ring variant of plan PT take this call?  (whole tiles per plane, 16-byte aligned base and strides)
template <class A, class PT>
bool col_ring_ok(const KParams<A>& P)
{
    constexpr PlanDesc D = PT::D;
    const size_t es = sizeof(typename A::C);
    const long long nc = P.ncols > 0 ? P.ncols : P.howmany;
    if (P.in_dist != 1 || nc <= 0 || nc % D.tpc != 0 || P.howmany % nc != 0) return false;
    if (P.npeers > 0 && (P.cols_per_peer % D.tpc != 0 || P.peer_col_dist % D.tpc != 0)) return false;   // tiles stay inside a peer's block
    // Rows further apart than this land on different 2 MiB pages almost every time; the pass is then bound by address
    // translation per row segment and wants the widest segments (the direct-load plans' 16 columns), measured:
    // profiles/r02/tune_r2c_f32_col*_{a,b}.txt (8 MiB stride: ring of 8 columns 7.4 ms vs 4.35 ms; 8 KiB: 3.39 vs 3.95)
    if ((size_t)P.in_stride * es > ((size_t)128 << 10) && !getenv("KISSFFT_RING_ANY_STRIDE")) return false;   // (override: tuning aid)
    if (((uintptr_t)P.in % 16) != 0 || ((size_t)P.in_stride * es) % 16 != 0) return false;
    if (P.howmany / nc > 1 && ((size_t)P.in_pdist * es) % 16 != 0) return false;
    if ((size_t)nc * (es / 4) >= ((size_t)1 << 32) || (size_t)P.in_stride * es >= ((size_t)1 << 40) ||
        (size_t)P.in_pdist * es >= ((size_t)1 << 40))
        return false;
    return tmap_encoder() != nullptr;
}

// fills P.tmap for a call that passed col_ring_ok; 0 on success
template <class A, class PT, int MODE>
int col_ring_encode(KParams<A>& P)
{
    constexpr PlanDesc D = PT::D;
    typedef FusedLayout<A, PT, MODE> LY;
    const size_t es = sizeof(typename A::C);
    const cuuint32_t wpe = (cuuint32_t)(es / 4);                       // 4-byte words per complex element
    const long long nc = P.ncols > 0 ? P.ncols : P.howmany;
    const long long nplanes = P.howmany / nc;
    // columns the map spans: with peer blocks spread out (peer_col_dist > cols_per_peer) the last block ends further right
    const long long span = (P.npeers > 0) ? (long long)(P.npeers - 1) * P.peer_col_dist + P.cols_per_peer : nc;
    cuuint64_t gdim[3] = {(cuuint64_t)span * wpe, (cuuint64_t)D.N, (cuuint64_t)nplanes};
    cuuint64_t gstr[2] = {(cuuint64_t)P.in_stride * es, nplanes > 1 ? (cuuint64_t)P.in_pdist * es : (cuuint64_t)P.in_stride * es * D.N};
    cuuint32_t box[3] = {(cuuint32_t)D.tpc * wpe, (cuuint32_t)LY::kBoxRows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUtensorMap m;
    const CUresult r = tmap_encoder()(&m, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, (void*)P.in, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                      CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return (int)cudaErrorInvalidValue;
    static_assert(sizeof(CUtensorMap) == sizeof(P.tmap), "opaque tensor-map blob must match CUtensorMap");
    memcpy(P.tmap, &m, sizeof(m));
    return 0;
}

}   // namespace kf
