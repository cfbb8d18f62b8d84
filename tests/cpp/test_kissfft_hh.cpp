// Exercises include/kissfft.hh (the C++ facade over the CUDA libraries) the way the reference's test/testcpp.cc
// exercises its header-only class: random input, compare with a direct long-double DFT, print the RMS error.
//   test_kissfft_hh            run every check on the GPU; exit status 0 = all within tolerance
//   test_kissfft_hh --no-gpu   only check that, without a CUDA device, transforms THROW (there is no CPU path)
#include "kissfft.hh"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <typeinfo>

typedef std::complex<long double> lcpx;

template <class T>
static long double rmse_vs_dft(const std::vector<std::complex<T> > &in, const std::vector<std::complex<T> > &out, bool inverse)
{
    const std::size_t n = in.size();
    const long double pi = std::acos(static_cast<long double>(-1));
    long double tot = 0, dif = 0;
    for (std::size_t k0 = 0; k0 < n; ++k0) {
        lcpx acc = 0;
        const long double ph = (inverse ? 2 : -2) * pi * k0 / n;
        for (std::size_t k1 = 0; k1 < n; ++k1) acc += lcpx(in[k1].real(), in[k1].imag()) * std::exp(lcpx(0, k1 * ph));
        tot += std::norm(acc);
        dif += std::norm(acc - lcpx(out[k0].real(), out[k0].imag()));
    }
    return std::sqrt(dif / tot);
}

template <class T>
static int dotest(std::size_t nfft, long double tol)
{
    typedef std::complex<T> cpx;
    int bad = 0;
    std::vector<cpx> in(nfft), out(nfft);
    for (std::size_t k = 0; k < nfft; ++k) in[k] = cpx((T)(rand() / (double)RAND_MAX - .5), (T)(rand() / (double)RAND_MAX - .5));

    kissfft<T> fft(nfft, false);
    fft.transform(&in[0], &out[0]);
    long double e = rmse_vs_dft(in, out, false);
    std::printf("type:%s nfft:%zu forward RMSE:%Lg\n", typeid(T).name(), nfft, e);
    bad += !(e < tol);

    fft.assign(nfft, true);                       // same object, other direction
    fft.transform(&in[0], &out[0]);
    e = rmse_vs_dft(in, out, true);
    std::printf("type:%s nfft:%zu inverse RMSE:%Lg\n", typeid(T).name(), nfft, e);
    bad += !(e < tol);

    // in_stride: every second element of a longer buffer
    std::vector<cpx> wide(2 * nfft);
    for (std::size_t k = 0; k < nfft; ++k) wide[2 * k] = in[k], wide[2 * k + 1] = cpx(99, 99);
    fft.assign(nfft, false);
    fft.transform(&wide[0], &out[0], 0, 1, 2);
    e = rmse_vs_dft(in, out, false);
    std::printf("type:%s nfft:%zu strided RMSE:%Lg\n", typeid(T).name(), nfft, e);
    bad += !(e < tol);

    // batch of rows == row by row
    const std::size_t rows = 5;
    std::vector<cpx> bin(rows * nfft), bout(rows * nfft), one(nfft);
    for (std::size_t k = 0; k < bin.size(); ++k) bin[k] = cpx((T)(rand() / (double)RAND_MAX - .5), (T)(rand() / (double)RAND_MAX - .5));
    fft.transform_batch(&bin[0], &bout[0], rows);
    for (std::size_t r = 0; r < rows; ++r) {
        fft.transform(&bin[r * nfft], &one[0]);
        if (std::memcmp(&one[0], &bout[r * nfft], nfft * sizeof(cpx)) != 0) { std::printf("batch row %zu differs\n", r); ++bad; }
    }

    // transform_real: 2*nfft real samples -> packed half spectrum (reference kissfft.hh:154-189)
    std::vector<T> re(2 * nfft);
    std::vector<cpx> zin(2 * nfft), zout(2 * nfft), packed(nfft);
    for (std::size_t k = 0; k < 2 * nfft; ++k) { re[k] = (T)(rand() / (double)RAND_MAX - .5); zin[k] = cpx(re[k], 0); }
    fft.transform_real(&re[0], &packed[0]);
    kissfft<T> full(2 * nfft, false);
    full.transform(&zin[0], &zout[0]);
    long double tot = 0, dif = 0;
    for (std::size_t k = 0; k < nfft; ++k) {
        const cpx want = k ? zout[k] : cpx(zout[0].real(), zout[nfft].real());
        tot += std::norm(lcpx(want.real(), want.imag()));
        dif += std::norm(lcpx(want.real(), want.imag()) - lcpx(packed[k].real(), packed[k].imag()));
    }
    e = std::sqrt(dif / tot);
    std::printf("type:%s nfft:%zu real RMSE vs complex transform:%Lg\n", typeid(T).name(), nfft, e);
    bad += !(e < tol);
    return bad;
}

template <class T>
static int expect_throw_without_gpu()
{
    kissfft<T> fft(64, false);                    // plans are plain host memory: construction works anywhere
    std::vector<std::complex<T> > in(64), out(64);
    try {
        fft.transform(&in[0], &out[0]);
    } catch (const std::runtime_error &e) {
        std::printf("threw as expected: %s\n", e.what());
        return 0;
    }
    std::printf("transform() returned without a GPU: a CPU path must not exist\n");
    return 1;
}

int main(int argc, char **argv)
{
    if (argc > 1 && !std::strcmp(argv[1], "--no-gpu")) return expect_throw_without_gpu<float>() + expect_throw_without_gpu<double>();
    int bad = 0;
    const std::size_t sizes[] = {1024, 1000, 74, 30};
    for (std::size_t i = 0; i < sizeof(sizes) / sizeof(sizes[0]); ++i) {
        bad += dotest<float>(sizes[i], 2e-6L);
        bad += dotest<double>(sizes[i], 1e-14L);
    }
    try {
        kissfft<long double> nope(16, false);
        std::printf("kissfft<long double> constructed: expected an exception\n");
        ++bad;
    } catch (const std::runtime_error &) {
    }
    std::printf(bad ? "FAILED (%d)\n" : "ALL OK\n", bad);
    return bad ? 1 : 0;
}
