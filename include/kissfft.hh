/*
 * kissfft.hh -- C++ facade with the interface of the reference's header-only class (reference kissfft.hh:16-189:
 * `kissfft<scalar_t>` with ctor(nfft, inverse), assign(), transform(), transform_real()), implemented over the CUDA
 * libraries' C-ABI (include/kiss_fft.h, kiss_fftr.h, kiss_fft_cuda.h) instead of the reference's recursive CPU code.
 *
 * The C libraries carry one datatype each under identical symbol names (libkissfft-float.so, libkissfft-double.so), so
 * a program that instantiates kissfft<float> AND kissfft<double> cannot link both.  The facade therefore binds the
 * library of its scalar type at run time (dlopen, RTLD_LOCAL):
 *
 *     $KISSFFT_B200_LIB_DIR/libkissfft-<type>.so   when the variable is set, else
 *     libkissfft-<type>.so                          through the normal loader search path / rpath.
 *
 * There is no CPU implementation behind it: a missing library, a missing GPU or a scalar type the GPU build does not have
 * (long double) throws std::runtime_error -- nothing is computed on the host.
 *
 * Beyond the reference's interface: transform_batch() (howmany contiguous rows, host pointers, pipelined H2D/kernel/D2H)
 * and transform_batch_device() (device pointers, one launch on a caller stream).
 *
 * Link with -ldl.  Header-only; C++11.
 */
#ifndef KISSFFT_B200_CLASS_HH
#define KISSFFT_B200_CLASS_HH

#include <dlfcn.h>

#include <complex>
#include <cstddef>
#include <cstdlib>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace kissfft_b200_detail {

template <typename T>
struct lib_of {
    static const char *name() { return 0; }      // no GPU build for this scalar type
};
template <>
struct lib_of<float> {
    static const char *name() { return "libkissfft-float.so"; }
};
template <>
struct lib_of<double> {
    static const char *name() { return "libkissfft-double.so"; }
};

/* the entry points the facade binds; cfg handles are opaque malloc'ed blocks (kiss_fft.h: free() releases them) */
struct api {
    void *handle;
    void *(*fft_alloc)(int, int, void *, std::size_t *);
    void *(*fftr_alloc)(int, int, void *, std::size_t *);
    int (*fft_batch)(void *, const void *, void *, std::size_t);
    int (*fftr_batch)(void *, const void *, void *, std::size_t);
    int (*fft_batch_dev)(void *, const void *, void *, std::size_t, std::size_t, std::size_t, int, void *);
    const char *(*last_error)(void);
    int (*scalar_bytes)(void);
    int (*is_fixed)(void);
};

inline void *sym(void *h, const char *lib, const char *name)
{
    void *p = dlsym(h, name);
    if (!p) throw std::runtime_error(std::string("kissfft: ") + lib + " does not export " + name);
    return p;
}

template <typename T>
inline const api &bind()
{
    static api a = []() -> api {
        const char *lib = lib_of<T>::name();
        if (!lib) throw std::runtime_error("kissfft: no GPU library for this scalar type (float and double only)");
        std::string path = lib;
        if (const char *dir = std::getenv("KISSFFT_B200_LIB_DIR")) path = std::string(dir) + "/" + lib;
        void *h = dlopen(path.c_str(), RTLD_NOW | RTLD_LOCAL);
        if (!h) throw std::runtime_error(std::string("kissfft: cannot load ") + path + ": " + dlerror());
        api r;
        r.handle = h;
        r.fft_alloc = reinterpret_cast<void *(*)(int, int, void *, std::size_t *)>(sym(h, lib, "kiss_fft_alloc"));
        r.fftr_alloc = reinterpret_cast<void *(*)(int, int, void *, std::size_t *)>(sym(h, lib, "kiss_fftr_alloc"));
        r.fft_batch = reinterpret_cast<int (*)(void *, const void *, void *, std::size_t)>(sym(h, lib, "kiss_fft_batch"));
        r.fftr_batch = reinterpret_cast<int (*)(void *, const void *, void *, std::size_t)>(sym(h, lib, "kiss_fftr_batch"));
        r.fft_batch_dev = reinterpret_cast<int (*)(void *, const void *, void *, std::size_t, std::size_t, std::size_t, int, void *)>(
            sym(h, lib, "kiss_fft_batch_dev"));
        r.last_error = reinterpret_cast<const char *(*)(void)>(sym(h, lib, "kiss_fft_cuda_last_error"));
        r.scalar_bytes = reinterpret_cast<int (*)(void)>(sym(h, lib, "kiss_fft_cuda_scalar_bytes"));
        r.is_fixed = reinterpret_cast<int (*)(void)>(sym(h, lib, "kiss_fft_cuda_is_fixed_point"));
        if (r.scalar_bytes() != (int)sizeof(T) || r.is_fixed())
            throw std::runtime_error(std::string("kissfft: ") + path + " was built for another datatype");
        return r;
    }();
    return a;
}

}   // namespace kissfft_b200_detail

template <typename scalar_t>
class kissfft {
public:
    typedef std::complex<scalar_t> cpx_t;

    /* reference kissfft.hh:23-53 */
    kissfft(const std::size_t nfft, const bool inverse) : _nfft(0), _inverse(false), _cfg(0), _rcfg(0) { assign(nfft, inverse); }
    ~kissfft() { release(); }
    kissfft(const kissfft &o) : _nfft(0), _inverse(false), _cfg(0), _rcfg(0) { assign(o._nfft, o._inverse); }
    kissfft &operator=(const kissfft &o)
    {
        if (this != &o) assign(o._nfft, o._inverse);
        return *this;
    }

    /* reference kissfft.hh:61-77: same state as a newly constructed object */
    void assign(const std::size_t nfft, const bool inverse)
    {
        if (_cfg && nfft == _nfft && inverse == _inverse) return;
        const kissfft_b200_detail::api &a = kissfft_b200_detail::bind<scalar_t>();
        void *cfg = nfft ? a.fft_alloc((int)nfft, inverse ? 1 : 0, 0, 0) : 0;
        if (nfft && !cfg) throw std::runtime_error("kissfft: kiss_fft_alloc failed");
        release();
        _cfg = cfg;
        _nfft = nfft;
        _inverse = inverse;
    }

    /* reference kissfft.hh:90-129.  `stage` and `fstride` are the reference's recursion arguments; a caller only ever
     * passes their defaults.  in_stride is the element stride of the input (kiss_fft_stride). */
    void transform(const cpx_t *fft_in, cpx_t *fft_out, const std::size_t stage = 0, const std::size_t fstride = 1,
                   const std::size_t in_stride = 1) const
    {
        if (stage != 0 || fstride != 1) throw std::invalid_argument("kissfft::transform: stage/fstride are internal to the reference's recursion");
        if (_nfft == 0) return;
        if (in_stride == 1) {
            run(kissfft_b200_detail::bind<scalar_t>().fft_batch(_cfg, fft_in, fft_out, 1));
        } else {
            std::vector<cpx_t> packed(_nfft);
            for (std::size_t i = 0; i < _nfft; ++i) packed[i] = fft_in[i * in_stride];
            run(kissfft_b200_detail::bind<scalar_t>().fft_batch(_cfg, packed.data(), fft_out, 1));
        }
    }

    /* howmany contiguous rows of nfft points, host memory; one pipelined pass through the GPU */
    void transform_batch(const cpx_t *in, cpx_t *out, const std::size_t howmany) const
    {
        if (_nfft == 0 || howmany == 0) return;
        run(kissfft_b200_detail::bind<scalar_t>().fft_batch(_cfg, in, out, howmany));
    }

    /* device pointers (rows `in_dist` / `out_dist` elements apart), asynchronous on `cuda_stream` */
    void transform_batch_device(const cpx_t *d_in, cpx_t *d_out, const std::size_t howmany, const std::size_t in_dist,
                                const std::size_t out_dist, void *cuda_stream = 0) const
    {
        if (_nfft == 0 || howmany == 0) return;
        run(kissfft_b200_detail::bind<scalar_t>().fft_batch_dev(_cfg, d_in, d_out, howmany, in_dist, out_dist, 1, cuda_stream));
    }

    /* reference kissfft.hh:154-189: DFT of 2*N real samples, N = nfft of this object.  dst has N entries: dst[0] holds the
     * (real) bins 0 and N in its real and imaginary part, dst[1..N-1] the bins 1..N-1.  With the inverse flag the
     * reference evaluates the same sum with the conjugate kernel, i.e. the conjugate spectrum. */
    void transform_real(const scalar_t *const src, cpx_t *const dst) const
    {
        const std::size_t N = _nfft;
        if (N == 0) return;
        const kissfft_b200_detail::api &a = kissfft_b200_detail::bind<scalar_t>();
        if (!_rcfg) {
            _rcfg = a.fftr_alloc((int)(2 * N), 0, 0, 0);
            if (!_rcfg) throw std::runtime_error("kissfft: kiss_fftr_alloc failed");
        }
        std::vector<cpx_t> half(N + 1);
        run(a.fftr_batch(_rcfg, src, half.data(), 1));
        dst[0] = cpx_t(half[0].real(), half[N].real());
        for (std::size_t k = 1; k < N; ++k) dst[k] = _inverse ? std::conj(half[k]) : half[k];
    }

    std::size_t nfft() const { return _nfft; }
    bool inverse() const { return _inverse; }

private:
    void run(int rc) const
    {
        if (rc != 0) {
            const char *msg = kissfft_b200_detail::bind<scalar_t>().last_error();
            throw std::runtime_error(std::string("kissfft: GPU transform failed: ") + (msg ? msg : "unknown error"));
        }
    }
    void release()
    {
        std::free(_cfg);      /* kiss_fft_free == free (kiss_fft.h) */
        std::free(_rcfg);
        _cfg = 0;
        _rcfg = 0;
    }

    std::size_t _nfft;
    bool _inverse;
    void *_cfg;
    mutable void *_rcfg;      /* real-input plan for 2*nfft samples, made on first use */
};

#endif
