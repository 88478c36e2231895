"""Multi-GPU plumbing: one process per GPU, ``torch.distributed`` (NCCL over NVLink on the GPU
box, gloo in the CPU tests).  The path shards two ways:

* Step 1: every rank holds the whole normalised matrix and takes a contiguous slice of the S
  subsample index sets; the per-gene accumulators double[2][G] are combined with ONE all-reduce
  (16*G bytes).
* Step 3 has no subsample axis: contiguous gene slices per rank; the o vector is completed with
  a second all-reduce (ranks write 0 outside their slice, so SUM is exact).

Nothing here touches the data path of a single GPU.
"""
from __future__ import annotations

import numpy as np


def shard_range(n: int, rank: int, world: int):
    """Contiguous, balanced partition of range(n): first (n % world) ranks get one extra."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    hi = lo + base + (1 if rank < extra else 0)
    return lo, hi


def _backend_device(dist):
    import torch
    if dist.get_backend() == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def broadcast_numpy(arr: np.ndarray, dist, src: int = 0) -> np.ndarray:
    """Broadcast a numpy array (same shape/dtype on all ranks) from ``src``; returns the array."""
    import torch
    dev = _backend_device(dist)
    t = torch.from_numpy(np.ascontiguousarray(arr)).to(dev)
    dist.broadcast(t, src=src)
    return t.cpu().numpy()


class _DeviceView:
    """Minimal __cuda_array_interface__ carrier so torch can wrap a raw device pointer."""

    def __init__(self, ptr: int, nbytes: int, typestr: str, itemsize: int):
        self.__cuda_array_interface__ = {"shape": (nbytes // itemsize,), "typestr": typestr,
                                         "data": (ptr, False), "version": 2}


def device_tensor(ctx, which, typestr="<f8", itemsize=8):
    """A torch tensor aliasing a context buffer (no copy, no ownership change); cached per (pointer, size)."""
    import torch
    ptr, nbytes = ctx.device_ptr(which)
    cache = ctx.__dict__.setdefault("_tensor_cache", {})
    key = (which, ptr, nbytes, typestr)
    if key not in cache:
        cache[key] = torch.as_tensor(_DeviceView(ptr, nbytes, typestr, itemsize), device="cuda:%d" % ctx.device)
    return cache[key]


def allreduce_device_buffer(ctx, which, dist):
    """In-place SUM all-reduce of a double context buffer.  NCCL: directly on the device memory
    (the context runs on torch's current stream, so ordering is stream order).  gloo: staged
    through the host."""
    if dist.get_backend() == "nccl":
        t = device_tensor(ctx, which)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return
    import torch
    _, nbytes = ctx.device_ptr(which)
    host = ctx.read(which, np.float64, (nbytes // 8,))
    t = torch.from_numpy(host)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    ctx.write(which, t.numpy())


def allreduce_numpy(arr: np.ndarray, dist) -> np.ndarray:
    """SUM all-reduce of a host array (used by the CPU tests of the sharding logic)."""
    import torch
    t = torch.from_numpy(np.ascontiguousarray(arr)).to(_backend_device(dist))
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()
