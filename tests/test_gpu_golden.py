"""CUDA path vs the committed golden vectors (outputs of the compiled, unmodified reference)."""
import numpy as np
import pytest

from oracle.loader import TYPES
from tests import golden_util
from tests.test_gpu_parity import check, dev, host

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


@pytest.fixture(scope="module", params=TYPES)
def ctx(request):
    import kissfft_b200
    return request.param, kissfft_b200.get(request.param)


def test_golden_c2c(ctx):
    tname, lib = ctx
    for key, x, y in golden_util.cases(tname, "c2c"):
        nfft, inv = map(int, key.split("_"))
        cfg = lib.alloc(nfft, inv)
        d_in, d_out = dev(x), dev(np.zeros_like(x))
        lib.fft_batch_dev(cfg, d_in, d_out, x.shape[0], nfft, nfft)
        torch.cuda.synchronize()
        check(tname, host(d_out), y, nfft, "golden c2c " + key)
        lib.free(cfg)


def test_golden_real(ctx):
    tname, lib = ctx
    for key, x, y in golden_util.cases(tname, "r2c"):
        nfft = int(key)
        cfg = lib.allocr(nfft, False)
        d_in, d_out = dev(x), dev(np.zeros_like(y))
        lib.fftr_batch_dev(cfg, d_in, d_out, x.shape[0], nfft, nfft // 2 + 1)
        torch.cuda.synchronize()
        check(tname, host(d_out), y, nfft, "golden r2c " + key)
        lib.free(cfg)
    for key, x, y in golden_util.cases(tname, "c2r"):
        nfft = int(key)
        cfg = lib.allocr(nfft, True)
        d_in, d_out = dev(x), dev(np.zeros_like(y))
        lib.fftri_batch_dev(cfg, d_in, d_out, x.shape[0], nfft // 2 + 1, nfft)
        torch.cuda.synchronize()
        check(tname, host(d_out), y, nfft, "golden c2r " + key)
        lib.free(cfg)


def test_golden_nd(ctx):
    tname, lib = ctx
    for key, x, y in golden_util.cases(tname, "nd"):
        inv = int(key.split("_")[-1])
        dims = x.shape[:-1]
        cfg = lib.allocnd(dims, inv)
        d_in, d_out = dev(x), dev(np.zeros_like(x))
        lib.fftnd_dev(cfg, d_in, d_out)
        torch.cuda.synchronize()
        check(tname, host(d_out), y, int(np.prod(dims)), "golden nd " + key)
        lib.free(cfg)
    for key, x, y in golden_util.cases(tname, "ndr"):
        cfg = lib.allocndr(x.shape, False)
        d_in, d_out = dev(x), dev(np.zeros_like(y))
        lib.fftndr_dev(cfg, d_in, d_out)
        torch.cuda.synchronize()
        check(tname, host(d_out), y, x.size, "golden ndr " + key)
        lib.free(cfg)
    for key, x, y in golden_util.cases(tname, "ndri"):
        cfg = lib.allocndr(y.shape, True)
        d_in, d_out = dev(x), dev(np.zeros_like(y))
        lib.fftndri_dev(cfg, d_in, d_out)
        torch.cuda.synchronize()
        check(tname, host(d_out), y, y.size, "golden ndri " + key)
        lib.free(cfg)
