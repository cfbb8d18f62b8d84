"""Multi-GPU checks, run under torchrun on a box with >= 2 GPUs (not collected by pytest):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/multi_gpu_check.py [--dims 256 256 256] [--p2p]

Every rank transforms its slab with kissfft_b200.slab.SlabFFT3D (NCCL all-to-all, or the fused peer-memory stores
with --p2p) and the distributed result is compared with numpy.fft.fftn of the full array; it also times the steps."""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kissfft_b200.slab import SlabFFT3D  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dims", type=int, nargs=3, default=[256, 256, 256])
    ap.add_argument("--p2p", action="store_true")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--no-check", action="store_true")
    ap.add_argument("--reference-order", action="store_true",
                    help="axis order 0,1,2 from last-axis slabs; checked BIT-EXACT against the single-GPU kiss_fftnd_dev")
    ap.add_argument("--tname", default="float", choices=["float", "double", "int16_t", "int32_t"])
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dims = tuple(args.dims)
    if args.reference_order:
        return reference_order(args, dims, rank, world)
    plan = SlabFFT3D(dims, tname="float", p2p=args.p2p)
    g = plan.geo
    x, send, recv, out = plan.alloc()
    gen = torch.Generator(device="cuda")
    gen.manual_seed(100 + rank)
    x0 = torch.rand(x.shape, generator=gen, device="cuda") * 2 - 1
    stream = torch.cuda.current_stream().cuda_stream
    err = None
    if not args.no_check:
        x.copy_(x0)
        plan.forward(x, send, recv, out, stream)
        torch.cuda.synchronize()
        parts = [torch.empty_like(x0) for _ in range(world)]
        dist.all_gather(parts, x0)
        full = torch.cat(parts, 0).cpu().numpy()
        want = np.fft.fftn(full[..., 0].astype(np.float64) + 1j * full[..., 1])
        c0, c1 = g.col_range()
        mine = np.ascontiguousarray(want[:, :, c0:c1].transpose(2, 1, 0))
        got = out.cpu().numpy()
        got = got[..., 0].astype(np.float64) + 1j * got[..., 1]
        err = float(np.sqrt(np.sum(np.abs(got - mine) ** 2) / np.sum(np.abs(mine) ** 2)))
    # timing: max over ranks of the CUDA-event time of `iters` forward calls
    for _ in range(3):
        x.copy_(x0)
        plan.forward(x, send, recv, out, stream)
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.iters):
        plan.forward(x, send, recv, out, stream)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / args.iters], device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    errs = [None] * world
    dist.all_gather_object(errs, err)
    if rank == 0:
        n = float(np.prod(dims))
        print(json.dumps({"dims": dims, "world": world, "p2p": args.p2p, "ms": float(ms[0]), "rel_rms_per_rank": errs,
                          "gflops": 5 * n * np.log2(n) / (float(ms[0]) * 1e-3) / 1e9,
                          "a2a_bytes_per_rank": g.a2a_bytes_per_rank(8)}))
        if err is not None:
            assert max(errs) <= 1e-6 * np.log2(n), errs
    dist.barrier()
    dist.destroy_process_group()


def reference_order(args, dims, rank, world):
    """SlabFFT3D.forward_reference on `world` GPUs must equal kiss_fftnd_dev on one GPU bit for bit (all datatypes)"""
    import kissfft_b200
    tname = args.tname
    plan = SlabFFT3D(dims, tname=tname)
    g = plan.geo
    x, work, send, recv, out = plan.alloc_reference()
    gen = torch.Generator(device="cuda")
    gen.manual_seed(5)                                     # every rank generates the same full array
    if tname in ("float", "double"):
        full = (torch.rand(dims + (2,), generator=gen, device="cuda", dtype=torch.float64) * 2 - 1).to(x.dtype)
    else:
        half = (32767 if tname == "int16_t" else 2147483647) // 2
        full = torch.randint(-half, half + 1, dims + (2,), generator=gen, device="cuda", dtype=torch.int64).to(x.dtype)
    c0, c1 = g.col_range()
    x.copy_(full[:, :, c0:c1])
    stream = torch.cuda.current_stream().cuda_stream
    plan.forward_reference(x, work, send, recv, out, stream)
    torch.cuda.synchronize()
    lib = kissfft_b200.get(tname)
    cfg = lib.allocnd(list(dims), False)
    want = torch.empty_like(full)
    lib.fftnd_dev(cfg, full, want, None, stream)
    torch.cuda.synchronize()
    p0, p1 = g.plane_range()
    exact = bool(torch.equal(out, want[p0:p1]))
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.iters):
        plan.forward_reference(x, work, send, recv, out, stream)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / args.iters], device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    flags = [None] * world
    dist.all_gather_object(flags, exact)
    if rank == 0:
        print(json.dumps({"mode": "reference-order", "tname": tname, "dims": dims, "world": world, "ms": float(ms[0]),
                          "bit_exact_vs_single_gpu_fftnd_per_rank": flags}))
        assert all(flags), flags
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
