"""Slab-decomposed 3-D complex transform across G GPUs (one process per GPU).

The reference has no distributed code; kiss_fftnd (kiss_fftnd.c:156-188) sweeps the axes of ONE array.  This module
splits a d0 x d1 x d2 array into slabs of d0/G planes per rank and needs a single exchange:

  A. rows (axis 2) of the local slab          -> kiss_fft_batch_dev, in place            [d0/G][d1][d2]
  B. columns (axis 1) of every local plane    -> kiss_fft_planes_pass_dev, which also transposes each plane and
                                                 writes it pre-sorted by destination rank  [G][d0/G][d2/G][d1]
  X. all-to-all: block s goes to rank s        -> NCCL (torch.distributed.all_to_all_single), or, when the peers'
                                                 receive buffers are mapped (symmetric memory), step B stores straight
                                                 into them over NVLink and X degenerates to a barrier
  C. axis 0, now complete on every rank        -> kiss_fft_axis_pass_dev                   [d2/G][d1][d0]

The result is the 3-D DFT X[k0][k1][k2] stored as out[k2 - r*d2/G][k1][k0] on rank r ("transposed-out", distributed
along k2) -- the usual contract of slab FFTs; `gather_natural` re-assembles the natural order for checks.  Float and
double only agree with kiss_fftnd up to rounding because the axes run in the order 2,1,0 instead of 0,1,2.

Reference-order mode (`forward_reference`): kiss_fftnd sweeps the axes in the order 0,1,2
(kiss_fftnd.c:172-178), and the Q15/Q31 results depend on that order.  Starting from slabs along the LAST axis, axes 0
and 1 are local, one exchange completes axis 2, and the result comes out in natural order as axis-0 slabs:

  A'. axis 0 of the local [d0][d1][d2/G] array   -> kiss_fft_axis_pass_dev                   [d1][d2/G][d0]
  B'. axis 1, k0 columns of destination rank s    -> kiss_fft_planes_pass_dev per s           [G][d2/G][d0/G][d1]
  X'. all-to-all                                                                              [d2][d0/G][d1]
  C'. axis 2                                      -> kiss_fft_axis_pass_dev                   [d0/G][d1][d2]

Same butterflies on the same operands as kiss_fftnd, so every datatype -- fixed point included -- is bit-identical to
the single-GPU kiss_fftnd.

The geometry (who owns what, exchange counts and offsets) is plain integer logic in `SlabGeometry` so that it is
testable on CPU with the gloo backend; the compute steps call the CUDA library and fail without it.
"""
from dataclasses import dataclass

import numpy as np


@dataclass(frozen=True)
class SlabGeometry:
    d0: int
    d1: int
    d2: int
    world: int
    rank: int

    def __post_init__(self):
        if self.d0 % self.world or self.d2 % self.world:
            raise ValueError("d0 and d2 must be divisible by the number of ranks")

    @property
    def planes(self):          # local planes of the input slab
        return self.d0 // self.world

    @property
    def cols(self):            # k2 values owned after the exchange
        return self.d2 // self.world

    @property
    def local_elems(self):
        return self.planes * self.d1 * self.d2

    @property
    def block_elems(self):     # one (source rank, destination rank) block of the exchange
        return self.planes * self.cols * self.d1

    def plane_range(self, rank=None):
        r = self.rank if rank is None else rank
        return r * self.planes, (r + 1) * self.planes

    def col_range(self, rank=None):
        r = self.rank if rank is None else rank
        return r * self.cols, (r + 1) * self.cols

    def send_offset(self, dst):
        """element offset of the block for rank dst inside the step-B output [G][planes][cols][d1]"""
        return dst * self.block_elems

    def recv_offset(self, src):
        """element offset of the block from rank src inside the receive buffer [d0][cols][d1] (= [G][planes][cols][d1])"""
        return src * self.block_elems

    def a2a_bytes_per_rank(self, itemsize):
        """bytes this rank sends to OTHER ranks (NVLink roofline numerator)"""
        return (self.world - 1) * self.block_elems * itemsize


class _null_ctx:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def batch_shard(howmany, rank, world):
    """contiguous share of a batch of independent transforms: (first, count).  No communication is ever needed."""
    base, rem = divmod(howmany, world)
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


class SlabFFT3D:
    """forward/inverse 3-D transform of a slab-distributed array; all buffers are torch CUDA tensors of shape (..., 2)"""

    def __init__(self, dims, tname="float", inverse=False, group=None, backend=None, p2p=False):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.group = group
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.geo = SlabGeometry(int(dims[0]), int(dims[1]), int(dims[2]), world, rank)
        self.inverse = bool(inverse)
        self.tname = tname
        if backend is None:
            import kissfft_b200
            backend = CudaBackend(kissfft_b200.get(tname), self.geo, self.inverse)
        self.backend = backend
        self.p2p = bool(p2p) and world > 1
        self._symm = None
        self._side = None
        import os
        self.chunks = int(os.environ.get("KISSFFT_SLAB_CHUNKS", "4"))
        self.remote_ctas = int(os.environ.get("KISSFFT_SLAB_REMOTE_CTAS", "0"))

    # ---- buffers ----
    def alloc(self):
        """(local input slab [P][d1][d2][2], staging [G][P][C][d1][2], output [C][d1][d0][2])"""
        g, be = self.geo, self.backend
        x = be.empty((g.planes, g.d1, g.d2, 2))
        recv = self._alloc_recv()
        send = None if self.p2p else be.empty((g.world, g.planes, g.cols, g.d1, 2))
        out = be.empty((g.cols, g.d1, g.d0, 2))
        return x, send, recv, out

    def _alloc_recv(self):
        g, be = self.geo, self.backend
        shape = (g.world, g.planes, g.cols, g.d1, 2)
        if not self.p2p:
            return be.empty(shape)
        import torch.distributed._symmetric_memory as symm
        t = symm.empty(shape, dtype=be.torch_dtype, device=be.device)
        self._symm = symm.rendezvous(t, self.group if self.group is not None else self.dist.group.WORLD)
        return t

    # ---- the transform ----
    def forward(self, x, send, recv, out, stream=0):
        """x is overwritten by step A (in place, like kiss_fft with fin == fout)."""
        g, be = self.geo, self.backend
        if g.world == 1:
            be.rows_inplace(x, stream)                               # A
            be.planes_cols(x, recv, dst_rank=None, stream=stream)    # B: whole plane, no exchange
        elif self.p2p:
            self._forward_p2p(x, recv, stream)
        else:
            be.rows_inplace(x, stream)                               # A
            for s in range(g.world):
                be.planes_cols(x, send, dst_rank=s, stream=stream, dst_block=s)
            self._exchange(recv, send, stream)                       # X
        be.axis0(recv, out, stream)                                  # C
        return out

    def _exchange(self, recv, send, stream=0):
        """block s of `send` goes to rank s; moved as bytes so that every datatype (Q15 included) is a type NCCL knows.
        The collective is enqueued on `stream` (torch orders NCCL work against its CURRENT stream, so the caller's raw
        stream is made current for the call -- otherwise the exchange would race with the kernels around it)."""
        torch = self.torch
        u8 = torch.uint8
        ctx = torch.cuda.stream(torch.cuda.ExternalStream(stream)) if (stream and torch.cuda.is_available()) else _null_ctx()
        with ctx:
            self.dist.all_to_all_single(recv.view(-1).view(u8), send.view(-1).view(u8), group=self.group)

    def _forward_p2p(self, x, recv, stream):
        """steps A, B and the exchange fused and overlapped.

        The slab is cut into chunks of planes.  For every chunk the rows (A) run on the main stream; the column pass
        (B) is ONE launch on a second stream (kiss_fft_planes_pass_peers_dev) whose output rows are stored STRAIGHT INTO
        THE RECEIVE BUFFERS of their destination ranks over NVLink (peer stores; its own block locally), so the
        link-bound stores of chunk c overlap the HBM-bound rows of chunk c+1.  Two device-side barriers bracket the
        peer stores (peers finished reading their buffer / all blocks have landed); no staging buffer, no collective."""
        g, be, torch = self.geo, self.backend, self.torch
        main = torch.cuda.ExternalStream(stream) if stream else torch.cuda.current_stream()
        if self._side is None:
            self._side = torch.cuda.Stream()
        side = self._side
        nchunks = max(1, min(self.chunks, g.planes))
        bounds = [g.planes * c // nchunks for c in range(nchunks + 1)]
        with torch.cuda.stream(main):
            self._symm.barrier()          # peers are done reading their receive buffer (step C of the previous call)
        peers = [self._symm.get_buffer(s, recv.shape, recv.dtype) for s in range(g.world)]
        esz = x.element_size() * 2
        for c in range(nchunks):
            p0, p1 = bounds[c], bounds[c + 1]
            if p1 == p0:
                continue
            be.rows_inplace(x, main.cuda_stream, p0, p1)                                      # A on this chunk
            ev = torch.cuda.Event()
            ev.record(main)
            side.wait_event(ev)
            # B + exchange in ONE launch: column block s of every plane goes through the mapped pointer into rank s's
            # receive buffer (its own block for s == rank)
            ptrs = [peers[s].data_ptr() + (g.rank * g.block_elems + p0 * g.cols * g.d1) * esz for s in range(g.world)]
            be.set_grid_limit(self.remote_ctas)
            be.planes_cols_peers(x, ptrs, side.cuda_stream, p0, p1)
            be.set_grid_limit(0)
        done = torch.cuda.Event()
        done.record(side)
        main.wait_event(done)
        with torch.cuda.stream(main):
            self._symm.barrier()          # every rank's blocks have landed in every receive buffer

    # ---- reference axis order 0,1,2 (bit-exact with kiss_fftnd for every datatype) ----
    def alloc_reference(self):
        """(local input [d0][d1][d2/G][2], work [d1][d2/G][d0][2], send and recv [G][d2/G][d0/G][d1][2], out [d0/G][d1][d2][2])"""
        g, be = self.geo, self.backend
        x = be.empty((g.d0, g.d1, g.cols, 2))
        work = be.empty((g.d1, g.cols, g.d0, 2))
        send = be.empty((g.world, g.cols, g.planes, g.d1, 2)) if g.world > 1 else None
        recv = be.empty((g.world, g.cols, g.planes, g.d1, 2))
        out = be.empty((g.planes, g.d1, g.d2, 2))
        return x, work, send, recv, out

    def forward_reference(self, x, work, send, recv, out, stream=0):
        """x: this rank's slab x[:, :, r*d2/G:(r+1)*d2/G] (not modified); out: X[r*d0/G:(r+1)*d0/G, :, :] in natural order"""
        g, be = self.geo, self.backend
        be.ref_axis0(x, work, stream)                                # A'
        if g.world == 1:
            be.ref_axis1(work, recv, 0, stream)                      # B' (whole array)
        else:
            for s in range(g.world):
                be.ref_axis1(work, send, s, stream)                  # B'
            self._exchange(recv, send, stream)                       # X'
        be.ref_axis2(recv, out, stream)                              # C'
        return out

    def gather_reference(self, out):
        """natural-order [d0][d1][d2][2] array on every rank from the axis-0 slabs of forward_reference (testing aid)"""
        g, torch, dist = self.geo, self.torch, self.dist
        if g.world == 1:
            return out
        mine = out.contiguous().view(-1).view(torch.uint8)
        parts = [torch.empty_like(mine) for _ in range(g.world)]
        dist.all_gather(parts, mine, group=self.group)
        return torch.cat([p.view(out.dtype).view(out.shape) for p in parts], dim=0)

    def gather_natural(self, out):
        """collects the distributed transposed result into the natural-order [d0][d1][d2][2] array on every rank (testing aid)"""
        g, torch, dist = self.geo, self.torch, self.dist
        if g.world == 1:
            full = out
        else:
            parts = [torch.empty_like(out) for _ in range(g.world)]
            dist.all_gather(parts, out.contiguous(), group=self.group)
            full = torch.cat(parts, dim=0)                # [d2][d1][d0][2]
        return full.permute(2, 1, 0, 3).contiguous()


class MgpuFFT3D:
    """kiss_fftnd_mgpu_* of the C library (include/kiss_fft_cuda.h, csrc/kf_mgpu.c) bound for torch tensors: the slab
    transform with its orchestration, NCCL communicator, peer mapping and pipelining all inside the library.  This class
    only carries the rendezvous id from rank 0 to the other ranks (through torch.distributed, as an MPI program would use
    MPI_Bcast) and hands device pointers over.  Fast axis order (2, 1, exchange, 0), transposed-out result like SlabFFT3D;
    reference_order=True: the exact mode (axes 0, 1, exchange, 2 -- kiss_fftnd's own order, bit-identical in every datatype),
    input slabs along the last axis [d0][d1][d2/G], natural-order output rows [d0/G][d1][d2]."""

    def __init__(self, dims, tname="float", inverse=False, p2p=True, group=None, reference_order=False):
        import torch
        import torch.distributed as dist
        import kissfft_b200
        self.torch = torch
        self.lib = kissfft_b200.get(tname)
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.geo = SlabGeometry(int(dims[0]), int(dims[1]), int(dims[2]), world, rank)
        self.torch_dtype = {"float": torch.float32, "double": torch.float64, "int16_t": torch.int16, "int32_t": torch.int32}[tname]
        ident = None
        if world > 1:
            box = [self.lib.mgpu_get_id() if rank == 0 else None]
            dist.broadcast_object_list(box, src=0, group=group)
            ident = box[0]
        self.reference_order = bool(reference_order)
        flags = (self.lib.MGPU_P2P if p2p else 0) | (self.lib.MGPU_REFERENCE_ORDER if reference_order else 0)
        self.cfg = self.lib.mgpu_alloc(dims, rank, world, ident, inverse, flags)
        self.info = self.lib.mgpu_info(self.cfg)

    def alloc(self):
        """(local input slab [P][d1][d2][2], output [C][d1][d0][2]); reference order: ([d0][d1][C][2], [P][d1][d2][2])"""
        g = self.geo
        dev = self.torch.device("cuda", self.torch.cuda.current_device())
        if self.reference_order:
            return (self.torch.empty((g.d0, g.d1, g.cols, 2), dtype=self.torch_dtype, device=dev),
                    self.torch.empty((g.planes, g.d1, g.d2, 2), dtype=self.torch_dtype, device=dev))
        return (self.torch.empty((g.planes, g.d1, g.d2, 2), dtype=self.torch_dtype, device=dev),
                self.torch.empty((g.cols, g.d1, g.d0, 2), dtype=self.torch_dtype, device=dev))

    def forward(self, x, out, stream=0):
        """x is overwritten (rows transformed in place); stream-ordered on `stream`"""
        self.lib.mgpu_exec(self.cfg, x, out, stream)
        return out

    def close(self):
        if self.cfg:
            self.lib.mgpu_free(self.cfg)
            self.cfg = None


class CudaBackend:
    """the three compute steps on the CUDA library (kissfft_b200.KissFFT); no CPU path"""

    def __init__(self, lib, geo, inverse):
        import torch
        self.torch = torch
        self.lib, self.geo = lib, geo
        self.device = torch.device("cuda", torch.cuda.current_device())
        self.torch_dtype = {"float": torch.float32, "double": torch.float64, "int16_t": torch.int16, "int32_t": torch.int32}[lib.tname]
        self.cfg2 = lib.alloc(geo.d2, inverse)
        self.cfg1 = lib.alloc(geo.d1, inverse)
        self.cfg0 = lib.alloc(geo.d0, inverse)

    def empty(self, shape):
        return self.torch.empty(shape, dtype=self.torch_dtype, device=self.device)

    def set_grid_limit(self, n):
        self.lib.set_grid_limit(n)

    def rows_inplace(self, x, stream, p0=0, p1=None):
        g = self.geo
        p1 = g.planes if p1 is None else p1
        esz = x.element_size() * 2
        ptr = x.data_ptr() + p0 * g.d1 * g.d2 * esz
        self.lib.fft_batch_dev(self.cfg2, ptr, ptr, (p1 - p0) * g.d1, g.d2, g.d2, 1, stream)

    def planes_cols(self, x, dst, dst_rank, stream, dst_block=0, p0=0, p1=None):
        """axis 1 of local planes [p0, p1); k2 columns of destination rank `dst_rank` (all if None) -> dst block"""
        g = self.geo
        p1 = g.planes if p1 is None else p1
        esz = x.element_size() * 2
        if dst_rank is None:
            self.lib.planes_pass_dev(self.cfg1, x, dst, g.planes, g.d2, g.d2, g.d1 * g.d2, g.d2 * g.d1, stream)
            return
        c0, _ = g.col_range(dst_rank)
        src_ptr = x.data_ptr() + (p0 * g.d1 * g.d2 + c0) * esz
        dst_ptr = dst.data_ptr() + (dst_block * g.block_elems + p0 * g.cols * g.d1) * esz
        self.lib.planes_pass_dev(self.cfg1, src_ptr, dst_ptr, p1 - p0, g.cols, g.d2, g.d1 * g.d2, g.cols * g.d1, stream)

    def planes_cols_peers(self, x, ptrs, stream, p0, p1):
        g = self.geo
        esz = x.element_size() * 2
        src_ptr = x.data_ptr() + p0 * g.d1 * g.d2 * esz
        self.lib.planes_pass_peers_dev(self.cfg1, src_ptr, ptrs, p1 - p0, g.cols, g.d2, g.d1 * g.d2, g.cols * g.d1, stream)

    def axis0(self, recv, out, stream):
        g = self.geo
        ncols = g.cols * g.d1
        self.lib.axis_pass_dev(self.cfg0, recv, out, ncols, ncols, stream)

    # reference-order steps
    def ref_axis0(self, x, work, stream):
        g = self.geo
        ncols = g.d1 * g.cols                       # [d0][d1*d2l] -> [d1*d2l][d0]
        self.lib.axis_pass_dev(self.cfg0, x, work, ncols, ncols, stream)

    def ref_axis1(self, work, dst, dst_rank, stream):
        """work [d1][d2l][d0]: plane = i2l (distance d0), row = i1 (stride d2l*d0), columns = the k0 slab of dst_rank;
        output block [d2l][d0/G][d1]"""
        g = self.geo
        esz = work.element_size() * 2
        src_ptr = work.data_ptr() + dst_rank * g.planes * esz
        dst_ptr = dst.data_ptr() + dst_rank * g.block_elems * esz
        self.lib.planes_pass_dev(self.cfg1, src_ptr, dst_ptr, g.cols, g.planes, g.cols * g.d0, g.d0, g.planes * g.d1, stream)

    def ref_axis2(self, recv, out, stream):
        g = self.geo
        ncols = g.planes * g.d1                     # [d2][d0l*d1] -> [d0l*d1][d2]
        self.lib.axis_pass_dev(self.cfg2, recv, out, ncols, ncols, stream)


def reference_slab_numpy(x_full, world):
    """what every rank must end up with, from numpy: list over ranks of out[k2_local][k1][k0] (complex128).  Testing aid."""
    X = np.fft.fftn(x_full)
    d2 = X.shape[2]
    c = d2 // world
    return [np.ascontiguousarray(X[:, :, r * c:(r + 1) * c].transpose(2, 1, 0)) for r in range(world)]
