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
