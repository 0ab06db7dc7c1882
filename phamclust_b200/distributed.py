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

When the consumer of the matrix is the HOST (the reference hands it to host
clustering), the all-gather is the wrong exchange: ``sharded_matrix_to_host``
lets every rank copy its finished row block straight from its GPU into ONE
triangle in POSIX shared memory (``SharedTriangle``, page-locked by every rank),
over that GPU's own PCIe link and in parallel with the others -- no device-side
assembly, no single-link D2H of the whole triangle.

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


class SharedTriangle:
    """The packed int32 triangle of an n-genome matrix in POSIX shared memory:
    created by rank 0, attached by every rank of the node; with a GPU present each
    rank page-locks the rows it will write (``pin_rows=(r0, r1)``, rounded out to
    whole pages; ``None`` = its ``row_partition`` block) so that its D2H is a
    full-speed DMA.  ``array`` is the numpy view; ``close()`` unmaps, rank 0 also
    unlinks."""

    def __init__(self, n, pin_rows=None):
        from multiprocessing import shared_memory
        dist = _dist()
        self.rank = dist.get_rank()
        self.n = n
        count = n * (n + 1) // 2
        nbytes = max(count, 1) * 4
        name = [None]
        if self.rank == 0:
            self._shm = shared_memory.SharedMemory(create=True, size=nbytes)
            name[0] = self._shm.name
        dist.broadcast_object_list(name, src=0)
        if self.rank != 0:
            self._shm = shared_memory.SharedMemory(name=name[0])
            try:        # the creator unlinks; keep this process's resource tracker out of it
                import multiprocessing
                from multiprocessing import resource_tracker
                # (a multiprocessing child shares its parent's tracker with the creator:
                # attaching registered nothing new there, so there is nothing to take back)
                if multiprocessing.parent_process() is None:
                    resource_tracker.unregister(self._shm._name, "shared_memory")
            except Exception:
                pass
        self.array = np.ndarray((count,), dtype=np.int32, buffer=self._shm.buf)
        self._registered = None
        if pin_rows is None:
            bounds = row_partition(n, dist.get_world_size())
            pin_rows = (bounds[self.rank], bounds[self.rank + 1])
        self._pin(tri_offset(n, pin_rows[0]) * 4, tri_offset(n, pin_rows[1]) * 4, nbytes)

    def _pin(self, lo, hi, nbytes):
        import sys
        if hi <= lo or "torch" not in sys.modules:
            return
        import torch
        if not torch.cuda.is_available():
            return
        page = 4096                          # the mapping itself starts on a page boundary
        lo = lo // page * page
        hi = min((hi + page - 1) // page * page, (nbytes + page - 1) // page * page)
        addr = self.array.ctypes.data + lo
        rc = torch.cuda.cudart().cudaHostRegister(addr, hi - lo, 0)
        if int(rc) != 0:
            raise RuntimeError(f"cudaHostRegister of the shared triangle failed ({rc})")
        self._registered = addr

    def rows(self, r0, r1):
        """View of the rows [r0, r1) of the triangle."""
        return self.array[tri_offset(self.n, r0):tri_offset(self.n, r1)]

    def close(self):
        if self._registered is not None:
            import torch
            torch.cuda.cudart().cudaHostUnregister(self._registered)
            self._registered = None
        self.array = None
        self._shm.close()
        if self.rank == 0:
            self._shm.unlink()


def sharded_matrix_to_host(table, metric, shared, as_distance=True, ppos=False, backend_ops=None):
    """Every rank computes its row block and copies it from its own GPU straight
    into ``shared`` (a SharedTriangle); returns after a barrier, when the whole
    triangle is in host memory (``shared.array``, valid on every rank of the node).

    ``backend_ops`` (tests): object with ``accumulate(part, n_parts) -> int64
    tensor`` and ``rows_to_host(r0, r1, acc, out)`` filling the int32 numpy view
    ``out``; the default drives libphamclust_b200 on this rank's GPU.
    """
    dist = _dist()
    rank, world = dist.get_rank(), dist.get_world_size()
    n = len(table.genome_ptr) - 1
    ops = backend_ops or _GpuOps(table, metric, as_distance, ppos)
    bounds = row_partition(n, world)
    acc = None
    if metric in ("aai", "peq"):
        acc = ops.accumulate(rank, world)
        dist.all_reduce(acc, op=dist.ReduceOp.SUM)       # exact integer sum
    ops.rows_to_host(bounds[rank], bounds[rank + 1], acc, shared.rows(bounds[rank], bounds[rank + 1]))
    dist.barrier()
    return shared.array


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

    def rows_to_host(self, r0, r1, acc, out):
        """Row block computed on this GPU, D2H by the library into ``out`` (page-locked
        shared memory: a full-speed DMA on this GPU's own link)."""
        if r1 > r0:
            self.ctx.matrix_rows(self.metric, self.as_distance, r0, r1,
                                 acc_ptr=acc.data_ptr() if acc is not None else None, out=out,
                                 ppos=self.ppos)
