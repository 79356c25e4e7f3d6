"""Multi-GPU plumbing for the branch-and-cut frontier: one process per GPU, `torch.distributed`
(NCCL on GPUs, gloo for CPU tests) carrying only the per-round all-gather of node summaries
(128 bytes per node).  The LP itself never shards; the frontier manager in
libjslp_b200 calls back into `all_gather` once per speculative round and every rank commits the
same results in the same order, so no frontier state is ever exchanged."""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import ALL_GATHER_FN


def is_active() -> bool:
    try:
        import torch.distributed as dist
        return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
    except Exception:
        return False


def rank_and_world():
    import torch.distributed as dist
    return dist.get_rank(), dist.get_world_size()


_COMM: dict = {}


def nccl_communicator(context):
    """The in-library NCCL communicator of this process (include/jslp_b200.h: jslp_comm_*), created once per
    device context: rank 0's 128-byte unique id travels over torch.distributed (the bootstrap this host layer
    happens to have), then every rank calls jslp_comm_create.  Returns None when the process group is not NCCL
    (gloo in the CPU tests and when several ranks share one GPU): the caller falls back to the host hook."""
    import torch
    import torch.distributed as dist
    if dist.get_backend() != "nccl":
        return None
    key = id(context)
    if key in _COMM:
        return _COMM[key][0]
    from . import _lib
    L = context.lib
    rank, world = dist.get_rank(), dist.get_world_size()
    buf = (C.c_uint8 * 128)()
    if rank == 0:
        _lib.check(L.jslp_comm_unique_id(buf))
    t = torch.tensor(list(buf), dtype=torch.uint8, device=torch.device("cuda", torch.cuda.current_device()))
    dist.broadcast(t, 0)
    ident = (C.c_uint8 * 128)(*t.cpu().tolist())
    h = C.c_void_p()
    _lib.check(L.jslp_comm_create(context.handle, ident, rank, world, C.byref(h)))
    _COMM[key] = (h, context)
    return h


def destroy_communicators():
    for h, ctx in _COMM.values():
        ctx.lib.jslp_comm_destroy(h)
    _COMM.clear()


def make_all_gather_hook(device=None):
    """Returns (ctypes callback, keepalive).  The callback all-gathers `bytes_per_rank` bytes per rank
    in place in the rank-major host buffer `buf` (jslp_bnb_opts.all_gather)."""
    import torch
    import torch.distributed as dist

    rank, world = dist.get_rank(), dist.get_world_size()
    backend = dist.get_backend()
    use_cuda = backend == "nccl"
    dev = torch.device("cuda", torch.cuda.current_device()) if use_cuda and device is None else device

    def hook(user, buf, bytes_per_rank):
        try:
            n = int(bytes_per_rank)
            host = np.ctypeslib.as_array(C.cast(buf, C.POINTER(C.c_uint8)), shape=(world * n,))
            mine = torch.from_numpy(host[rank * n:(rank + 1) * n].copy())
            if use_cuda:
                mine = mine.to(dev)
                out = torch.empty(world * n, dtype=torch.uint8, device=dev)
            else:
                out = torch.empty(world * n, dtype=torch.uint8)
            dist.all_gather_into_tensor(out, mine)
            host[:] = out.cpu().numpy()
            return 0
        except Exception:  # never let an exception cross the C boundary
            import traceback
            traceback.print_exc()
            return 1

    cb = ALL_GATHER_FN(hook)
    return cb, (cb, hook)
