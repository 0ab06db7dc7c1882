"""Multi-GPU sharding of the matrix build: one process per GPU, plumbing by
``torch.distributed`` (NCCL over NVLink on the GPU box, gloo in the CPU tests).

How the path shards:

* gcs / jc / pocp / af -- independent output tiles.  Every rank holds the
  whole (small) packed table; rank r computes a contiguous block of genome
  ROWS of the upper triangle chosen so that every rank owns the same number of
  pairs.  No collective during compute; ONE all-gather of the row blocks at the
  end assembles the packed triangle.
* aai / peq -- independent alignments.  Phams are dealt round-robin (in
  descending cost order) over the ranks; each rank aligns its phams' member
  pairs and scatter-adds best matches into dense int64 pair accumulators
  ``(matched << 32) | length``.  ONE integer all-reduce (exact, order
  independent) merges them, then the row-block epilogue + all-gather as above.

The compute callables are injectable so the host logic (partitioning,
all-reduce, assembly) is covered on CPU with world_size-2 gloo.
"""
import os

import numpy as np


def _dist():
    import torch.distributed as dist
    return dist


def is_active():
    """True inside an initialised process group of more than one rank."""
    try:
        import sys
        if "torch" not in sys.modules:
            return False
        dist = _dist()
        return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
    except Exception:
        return False


def tri_offset(n, i):
    """Packed-triangle index of (i, i)."""
    return i * n - (i * (i - 1)) // 2


def row_partition(n, world):
    """Row boundaries b[0..world] such that rows [b[r], b[r+1]) hold (nearly)
    the same number of upper-triangle entries for every rank."""
    total = n * (n + 1) // 2
    bounds = [0]
    for r in range(1, world):
        want = total * r // world
        # smallest i with tri_offset(n, i) >= want
        lo, hi = bounds[-1], n
        while lo < hi:
            mid = (lo + hi) // 2
            if tri_offset(n, mid) >= want:
                hi = mid
            else:
                lo = mid + 1
        bounds.append(lo)
    bounds.append(n)
    return bounds


def assemble(n, bounds, shards):
    """Concatenate per-rank row-block slices (already trimmed) into the packed
    triangle."""
    out = np.empty(n * (n + 1) // 2, dtype=np.int32)
    for r, shard in enumerate(shards):
        lo, hi = tri_offset(n, bounds[r]), tri_offset(n, bounds[r + 1])
        out[lo:hi] = shard[:hi - lo]
    return out


def sharded_matrix(table, metric, as_distance=True, ppos=False, backend_ops=None):
    """Packed int32 micro-unit triangle computed by all ranks of the default
    process group; every rank returns the full triangle (host numpy).

    ``backend_ops`` (tests): object with ``device``, ``accumulate(part, n_parts)
    -> int64 tensor`` and ``rows(r0, r1, acc) -> int32 tensor``; the default
    drives libphamclust_b200 on this rank's GPU.
    """
    import torch
    dist = _dist()
    rank, world = dist.get_rank(), dist.get_world_size()
    n = len(table.genome_ptr) - 1
    ops = backend_ops or _GpuOps(table, metric, as_distance, ppos)
    bounds = row_partition(n, world)
    acc = None
    if metric in ("aai", "peq"):
        acc = ops.accumulate(rank, world)
        dist.all_reduce(acc, op=dist.ReduceOp.SUM)       # exact integer sum
    local = ops.rows(bounds[rank], bounds[rank + 1], acc)
    # equal-sized shards for all_gather: pad to the largest block
    sizes = [tri_offset(n, bounds[r + 1]) - tri_offset(n, bounds[r]) for r in range(world)]
    width = max(sizes)
    padded = torch.zeros(width, dtype=torch.int32, device=local.device)
    padded[:local.numel()] = local
    gathered = torch.empty(world * width, dtype=torch.int32, device=local.device)
    dist.all_gather_into_tensor(gathered, padded)
    host = gathered.cpu().numpy().reshape(world, width)
    return assemble(n, bounds, [host[r] for r in range(world)])


class _GpuOps:
    """The real thing: this rank's GPU through the C ABI, on torch's stream so
    that NCCL collectives order after the kernels."""

    def __init__(self, table, metric, as_distance, ppos):
        import torch
        from phamclust_b200.matrix import get_context
        self.torch = torch
        local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local_rank)
        self.device = torch.device("cuda", local_rank)
        self.ctx = get_context(local_rank)
        self.ctx.set_stream(torch.cuda.current_stream().cuda_stream)
        self.ctx.load_table(table)
        if metric in ("aai", "peq"):
            self.ctx.load_sequences(table)
        self.metric, self.as_distance, self.ppos = metric, as_distance, ppos
        self.n = len(table.genome_ptr) - 1

    def accumulate(self, part, n_parts):
        acc = self.torch.zeros(self.n * (self.n + 1) // 2, dtype=self.torch.int64,
                               device=self.device)
        self.ctx.align_accumulate(acc.data_ptr(), part=part, n_parts=n_parts, ppos=self.ppos)
        return acc

    def rows(self, r0, r1, acc):
        count = tri_offset(self.n, r1) - tri_offset(self.n, r0)
        out = self.torch.empty(max(count, 1), dtype=self.torch.int32, device=self.device)
        if count:
            self.ctx.matrix_rows(self.metric, self.as_distance, r0, r1,
                                 acc_ptr=acc.data_ptr() if acc is not None else None,
                                 out_device_ptr=out.data_ptr(), ppos=self.ppos)
        return out[:count]
