"""Multi-process (gloo, world_size 2 and 4, CPU) test of the slab decomposition's host logic: ownership geometry,
destination-sorted staging layout, all-to-all, transposed-out result.  The three compute steps are supplied by a
numpy backend defined HERE (test infrastructure); the product's SlabFFT3D uses the CUDA backend and has no CPU path."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from kissfft_b200.slab import SlabFFT3D, SlabGeometry, batch_shard, reference_slab_numpy


class NumpyBackend:
    """rows / plane columns / axis 0 with numpy.fft on CPU tensors -- mirrors CudaBackend's layouts exactly"""
    torch_dtype = torch.float64
    device = torch.device("cpu")

    def __init__(self, geo):
        self.geo = geo

    def empty(self, shape):
        return torch.zeros(shape, dtype=torch.float64)

    @staticmethod
    def _c(t):
        a = t.numpy()
        return a[..., 0] + 1j * a[..., 1]

    @staticmethod
    def _put(t, c):
        t[..., 0] = torch.from_numpy(np.ascontiguousarray(c.real))
        t[..., 1] = torch.from_numpy(np.ascontiguousarray(c.imag))

    def set_grid_limit(self, n):
        pass

    def rows_inplace(self, x, stream, p0=0, p1=None):
        p1 = self.geo.planes if p1 is None else p1
        self._put(x[p0:p1], np.fft.fft(self._c(x[p0:p1]), axis=2))

    def planes_cols(self, x, dst, dst_rank, stream, dst_block=0, p0=0, p1=None):
        g = self.geo
        p1 = g.planes if p1 is None else p1
        y = np.fft.fft(self._c(x[p0:p1]), axis=1).transpose(0, 2, 1)   # [p1-p0][d2][d1]
        if dst_rank is None:
            self._put(dst[0], y)
            return
        c0, c1 = g.col_range(dst_rank)
        self._put(dst[dst_block][p0:p1], y[:, c0:c1, :])

    def planes_cols_peers(self, x, ptrs, stream, p0, p1):
        raise NotImplementedError("peer-memory path is CUDA-only")

    def axis0(self, recv, out, stream):
        g = self.geo
        a = self._c(recv).reshape(g.d0, g.cols, g.d1)
        self._put(out, np.fft.fft(a, axis=0).transpose(1, 2, 0))


    # reference-order steps (same layouts as CudaBackend.ref_*)
    def ref_axis0(self, x, work, stream):
        self._put(work, np.fft.fft(self._c(x), axis=0).transpose(1, 2, 0))

    def ref_axis1(self, work, dst, dst_rank, stream):
        g = self.geo
        k0, k1 = g.plane_range(dst_rank)
        y = np.fft.fft(self._c(work)[:, :, k0:k1], axis=0)             # [k1][i2l][k0l]
        self._put(dst[dst_rank], y.transpose(1, 2, 0))

    def ref_axis2(self, recv, out, stream):
        g = self.geo
        a = self._c(recv).reshape(g.d2, g.planes, g.d1)
        self._put(out, np.fft.fft(a, axis=0).transpose(1, 2, 0))


class OracleBackend(NumpyBackend):
    """the reference-order steps in Q15 with the CPU oracle: the distributed result must be bit-identical to kiss_fftnd"""
    torch_dtype = torch.int16

    def __init__(self, geo):
        from oracle.loader import Oracle
        self.geo = geo
        self.o = Oracle("int16_t")

    def empty(self, shape):
        return torch.zeros(shape, dtype=torch.int16)

    def _cols(self, a):
        """1-D transforms along axis 0 of a [n][...][2] int16 array"""
        n = a.shape[0]
        rows = np.ascontiguousarray(np.moveaxis(a, 0, -2).reshape(-1, n, 2))
        y = self.o.fft(rows, 0).reshape(a.shape[1:-1] + (n, 2))
        return y                                                       # [...][n][2]

    def ref_axis0(self, x, work, stream):
        work.copy_(torch.from_numpy(self._cols(x.numpy())))

    def ref_axis1(self, work, dst, dst_rank, stream):
        k0, k1 = self.geo.plane_range(dst_rank)
        dst[dst_rank].copy_(torch.from_numpy(self._cols(np.ascontiguousarray(work.numpy()[:, :, k0:k1]))))

    def ref_axis2(self, recv, out, stream):
        g = self.geo
        out.copy_(torch.from_numpy(self._cols(recv.numpy().reshape(g.d2, g.planes, g.d1, 2))))


def _worker_reference(rank, world, port, dims, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle.loader import Oracle, random_input
        geo = SlabGeometry(*dims, world, rank)
        c0, c1 = geo.col_range()
        # float64 geometry check against numpy
        rng = np.random.default_rng(11)
        full = rng.standard_normal(dims) + 1j * rng.standard_normal(dims)
        plan = SlabFFT3D(dims, tname="double", backend=NumpyBackend(geo))
        x, work, send, recv, out = plan.alloc_reference()
        NumpyBackend._put(x, full[:, :, c0:c1])
        plan.forward_reference(x, work, send, recv, out)
        nat = NumpyBackend._c(plan.gather_reference(out))
        want = np.fft.fftn(full)
        err = float(np.abs(nat - want).max() / np.abs(want).max())
        # Q15: bit-exact against the oracle's kiss_fftnd
        fullq = random_input("int16_t", dims, 5)
        plan = SlabFFT3D(dims, tname="int16_t", backend=OracleBackend(geo))
        x, work, send, recv, out = plan.alloc_reference()
        x.copy_(torch.from_numpy(np.ascontiguousarray(fullq[:, :, c0:c1])))
        plan.forward_reference(x, work, send, recv, out)
        natq = plan.gather_reference(out).numpy()
        exact = bool(np.array_equal(natq, Oracle("int16_t").fftnd(fullq)))
        q.put((rank, err, exact))
    finally:
        dist.destroy_process_group()


def _worker(rank, world, port, dims, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(7)
        full = rng.standard_normal(dims) + 1j * rng.standard_normal(dims)
        geo = SlabGeometry(*dims, world, rank)
        plan = SlabFFT3D(dims, tname="double", backend=NumpyBackend(geo))
        x, send, recv, out = plan.alloc()
        p0, p1 = geo.plane_range()
        NumpyBackend._put(x, full[p0:p1])
        plan.forward(x, send, recv, out)
        want = reference_slab_numpy(full, world)[rank]
        got = NumpyBackend._c(out)
        err = np.abs(got - want).max() / np.abs(want).max()
        nat = NumpyBackend._c(plan.gather_natural(out))
        err2 = np.abs(nat - np.fft.fftn(full)).max() / np.abs(want).max()
        q.put((rank, float(err), float(err2)))
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world,dims", [(2, (4, 6, 8)), (2, (8, 5, 6)), (4, (8, 3, 12))])
def test_slab_exchange_gloo(world, dims):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, dims, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    res = sorted(q.get(timeout=5) for _ in range(world))
    assert [r[0] for r in res] == list(range(world))
    for _, e1, e2 in res:
        assert e1 < 1e-12 and e2 < 1e-12


@pytest.mark.parametrize("world,dims", [(2, (4, 6, 8)), (2, (8, 5, 6)), (4, (8, 3, 12))])
def test_slab_reference_order_gloo(world, dims):
    """axis order 0,1,2 from last-axis slabs: natural-order axis-0 slabs, bit-identical to kiss_fftnd in Q15"""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_reference, args=(r, world, port, dims, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    res = sorted(q.get(timeout=5) for _ in range(world))
    assert [r[0] for r in res] == list(range(world))
    for _, err, exact in res:
        assert err < 1e-12 and exact


def test_geometry_and_batch_shard():
    g = SlabGeometry(1024, 1024, 1024, 8, 3)
    assert g.planes == 128 and g.cols == 128
    assert g.plane_range() == (384, 512) and g.col_range(7) == (896, 1024)
    assert g.block_elems == 128 * 128 * 1024
    # bytes sent per GPU: 8 GiB * (G-1)/G^2
    assert g.a2a_bytes_per_rank(8) == 8 * 2 ** 30 * 7 // 64
    with pytest.raises(ValueError):
        SlabGeometry(10, 4, 8, 4, 0)
    for howmany, world in [(65536, 8), (100000, 8), (7, 4), (3, 8)]:
        shards = [batch_shard(howmany, r, world) for r in range(world)]
        assert sum(c for _, c in shards) == howmany
        assert shards[0][0] == 0
        for (f0, c0), (f1, _) in zip(shards, shards[1:]):
            assert f0 + c0 == f1
        assert max(c for _, c in shards) - min(c for _, c in shards) <= 1
