"""
Multi-GPU plumbing of the bulk path: one process per GPU (torch.distributed), reads sharded by
rank, the annotation index replicated, per-GPU counters merged with ONE all-reduce.

The bulk units are independent and the tally is a commutative integer sum, so there
is no data-path collective besides the merge of `n_ensg + 8` int64 counters (about 0.3 MB, latency
bound).  On GPUs the all-reduce runs over NCCL directly on the library's device counter block
(`tec_bulk_counts_dev`) and on the library's stream: no host round trip, no extra copy.
"""
import numpy as np


def shard_units(n_units, rank, world, align=1):
    """[lo, hi) of this rank's contiguous slice of n_units units (any contiguous split is valid:
    the reference's result does not depend on read order in bulk mode)."""
    lo = (n_units * rank // world) // align * align
    hi = n_units if rank == world - 1 else (n_units * (rank + 1) // world) // align * align
    return lo, hi


class _DevArray:
    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


def counts_tensor(engine, n_ensg, device):
    """torch view (no copy) of the engine's device counter block: int64[n_ensg + BULK_NSTATS]."""
    import torch
    from . import _lib
    return torch.as_tensor(_DevArray(engine.bulk_counts_dev(), n_ensg + _lib.BULK_NSTATS, "<i8"), device=device)


def allreduce_counts_device(engine, n_ensg, device, group=None):
    """Sum the counter blocks of all ranks in place (NCCL), ordered on the engine's stream."""
    import torch
    import torch.distributed as dist
    t = counts_tensor(engine, n_ensg, device)
    with torch.cuda.stream(torch.cuda.ExternalStream(engine.stream, device=device)):
        dist.all_reduce(t, group=group)
    return t


def allreduce_counts_host(counts, stats, group=None):
    """Host-side merge of (counts, stats) as returned by Engine.bulk_finish: used where the ranks
    finished independently (e.g. one te_count process per BAM shard), any backend."""
    import torch
    import torch.distributed as dist
    t = torch.from_numpy(np.concatenate([np.asarray(counts, np.int64), np.asarray(stats, np.int64)]))
    dist.all_reduce(t, group=group)
    out = t.numpy()
    return out[:len(counts)].copy(), out[len(counts):].copy()
