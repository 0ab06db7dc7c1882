"""Multi-GPU sharding of the matrix build: one process per GPU, plumbing by
``torch.distributed`` (NCCL over NVLink on the GPU box, gloo in the CPU tests).

How the path shards; the reference's only parallelism is a process pool over
genome pairs (/root/reference/src/phamclust/matrix.py:471-491):

* gcs / jc / pocp / af -- independent output tiles.  Every rank holds the (small) packed table;
  rank r computes a contiguous block of genome ROWS of the upper triangle.  The blocks are cut
  by WORK, not by pair count: the int8 engine only streams the K blocks a tile's two genome
  groups have in common, so late rows (whose phams sit in more K blocks) cost more than early
  ones (``contraction_row_weights``).  No collective during compute.
* aai / peq -- independent alignments.  Phams are dealt over the ranks in descending cost order;
  each rank ships and canonicalises only the genes of ITS phams, aligns their member pairs and
  turns every (pham, genome pair) best match into one 16-byte record addressed to the rank that
  owns the pair's row.  ONE all-to-all moves the records (only pairs that share a pham exist:
  a few per cent of the triangle), the owner adds them into accumulators of its rows only --
  per-rank accumulator memory is 1/N of the triangle, nothing dense is reduced -- and runs the
  row-block epilogue.

Assembly of the row blocks, two consumers:

* device (``sharded_matrix``): every rank ends up with the triangle.  On GPUs the triangle lives in
  symmetric memory (``FusedTriangle``) and no collective is called: either copy engines push each
  finished row sub-block into the peers' triangles while the next sub-block is computed (default),
  or the producing kernel's epilogue itself stores every row segment into all triangles over NVLink
  (``pc_matrix_rows_fused``); one device-side barrier ends the step.  Where symmetric memory is not
  available (and in the CPU tests) the row blocks are exchanged with one NCCL / gloo all-gather;
* host (``sharded_matrix_to_host``; what the reference does next is host clustering): every rank
  copies its finished block from its own GPU over its own PCIe link into ONE triangle in POSIX
  shared memory (``SharedTriangle``) -- no device-side assembly, no single-link D2H.

``ShardedBuilder`` is the one implementation of a rank's share; ``bench.py`` times exactly its
methods.  The per-rank compute is injectable (``backend_ops``) so that partitioning, exchange and
assembly are covered on CPU with world_size-2 gloo.
"""
import os

import numpy as np

ALIGN_METRICS = ("aai", "peq")


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


def row_partition(n, world, row_weight=None):
    """Row boundaries b[0..world]: rows [b[r], b[r+1]) carry (nearly) the same share of the
    total weight for every rank.  ``row_weight[i]`` defaults to the number of upper-triangle
    entries of row i (n - i): equal pair counts."""
    if row_weight is None:
        total = n * (n + 1) // 2
        bounds = [0]
        for r in range(1, world):
            want = total * r // world
            lo, hi = bounds[-1], n              # smallest i with tri_offset(n, i) >= want
            while lo < hi:
                mid = (lo + hi) // 2
                if tri_offset(n, mid) >= want:
                    hi = mid
                else:
                    lo = mid + 1
            bounds.append(lo)
        bounds.append(n)
        return bounds
    w = np.asarray(row_weight, dtype=np.float64)
    assert w.shape == (n,) and (w >= 0).all()
    cum = np.concatenate(([0.0], np.cumsum(w)))
    bounds = [0]
    for r in range(1, world):
        cut = int(np.searchsorted(cum, cum[-1] * r / world, side="left"))
        bounds.append(min(max(cut, bounds[-1]), n))
    bounds.append(n)
    return bounds


def contraction_row_weights(table, group=128, block=128, tile_cols=256, out_cost=48.0):
    """Estimated cost of every genome row of the shared-count contraction under the block-sparse
    K schedule of the int8 engine (csrc/pc_contract_i8.cu): a (128-row, 256-column) tile streams
    the K blocks (128 operand columns) that its row group and one of its column groups both
    occupy, then pays a K-independent epilogue (``out_cost`` K-block equivalents: TMEM drain,
    fp64 normalise, 4 B/pair store).  Operand columns follow the library's order: phams held by
    >= 2 genomes, by first genome.  Returns float64 [n_genomes]."""
    n = len(table.genome_ptr) - 1
    pham_id = np.asarray(table.pham_id, dtype=np.int64)
    n_phams = int(table.n_phams)
    counts = np.bincount(pham_id, minlength=n_phams)
    first = np.full(n_phams, len(pham_id), dtype=np.int64)
    seen, first_at = np.unique(pham_id, return_index=True)
    first[seen] = first_at
    eligible = np.nonzero(counts >= 2)[0]
    col = np.full(n_phams, -1, dtype=np.int64)
    col[eligible[np.argsort(first[eligible], kind="stable")]] = np.arange(len(eligible))
    entry_col = col[pham_id]
    entry_genome = np.repeat(np.arange(n), np.diff(table.genome_ptr))
    keep = entry_col >= 0
    n_groups = (n + group - 1) // group
    n_blocks = max(1, (len(eligible) + block - 1) // block)
    occ = np.zeros((n_groups, n_blocks), dtype=np.float32)
    occ[entry_genome[keep] // group, entry_col[keep] // block] = 1.0
    common = occ @ occ.T                                   # K blocks two groups have in common
    per_col_tile = tile_cols // group
    work = np.zeros(n_groups)
    for g in range(n_groups):
        tiles = common[g, g:]
        # a column tile spans `per_col_tile` groups: it visits the union, bounded by the sum
        work[g] = tiles.sum() / per_col_tile * 1.3 + out_cost * ((n_groups - g) / per_col_tile)
    return np.repeat(work / group, group)[:n].astype(np.float64)


def assemble(n, bounds, shards):
    """Concatenate per-rank row-block slices (already trimmed) into the packed triangle."""
    out = np.empty(n * (n + 1) // 2, dtype=np.int32)
    for r, shard in enumerate(shards):
        lo, hi = tri_offset(n, bounds[r]), tri_offset(n, bounds[r + 1])
        out[lo:hi] = shard[:hi - lo]
    return out


class FusedTriangle:
    """The packed int32 triangle of every rank in symmetric memory (torch.distributed
    ._symmetric_memory: CUDA IPC mappings of every rank's buffer into every process), so that a
    kernel of one rank can store straight into the others' triangles."""

    def __init__(self, n, device):
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        self.n = n
        count = max(n * (n + 1) // 2, 1)
        group = dist.group.WORLD
        self.tensor = symm_mem.empty(count, dtype=torch.int32, device=device)
        self.handle = symm_mem.rendezvous(self.tensor, group)
        self.rank, self.world = self.handle.rank, self.handle.world_size
        self.ptrs = [int(p) for p in self.handle.buffer_ptrs]
        assert len(self.ptrs) == self.world and self.ptrs[self.rank] == self.tensor.data_ptr()

    def slice_ptrs(self, r0):
        """(own, [peers]) addresses of the slice that starts at row r0."""
        off = 4 * tri_offset(self.n, r0)
        return self.ptrs[self.rank] + off, [p + off for k, p in enumerate(self.ptrs) if k != self.rank]

    def barrier(self):
        """Device-side barrier on the current stream: every rank's stores have landed after it."""
        self.handle.barrier()


class ShardedBuilder:
    """This rank's share of one matrix build (the default process group's world).

    ``ops`` drives the rank's device (``_GpuOps``) or, in the CPU tests, the oracle:
        load(table, part, n_parts)                        host buffers -> device
        rows(r0, r1, acc_rows) -> int32 tensor            row block [r0, r1) of the triangle
        rows_to_host(r0, r1, acc_rows, out)               the same, written into a host view
        contributions(part, n_parts, bounds) -> (records int64 [k, 2] grouped by destination,
                                                 counts per destination)            aai / peq
        accumulate(records, r0, r1) -> int64 tensor       accumulators of rows [r0, r1) only
    """

    def __init__(self, table, metric, as_distance=True, ppos=False, backend_ops=None, balance="work"):
        dist = _dist()
        self.dist = dist
        grouped = dist.is_available() and dist.is_initialized()
        self.rank, self.world = (dist.get_rank(), dist.get_world_size()) if grouped else (0, 1)
        self.table, self.metric = table, metric
        self.n = len(table.genome_ptr) - 1
        self.align = metric in ALIGN_METRICS
        self.ops = backend_ops or _GpuOps(metric, as_distance, ppos)
        self._row_weight = None
        if self.align or balance != "work" or self.world == 1:
            # aai/peq: the row-block stage is the dense epilogue, whose cost is the pair count
            self.bounds = row_partition(self.n, self.world)
        else:
            self._row_weight = contraction_row_weights(table)
            bounds = row_partition(self.n, self.world, self._row_weight)
            # whole 256-row tiles per rank (a tile cut by a boundary would be multiplied twice)
            if self.n >= 512 * self.world:
                for r in range(1, self.world):
                    bounds[r] = min(max((bounds[r] + 128) // 256 * 256, bounds[r - 1]), self.n)
            self.bounds = bounds
        self.r0, self.r1 = self.bounds[self.rank], self.bounds[self.rank + 1]
        self.sizes = [tri_offset(self.n, self.bounds[r + 1]) - tri_offset(self.n, self.bounds[r])
                      for r in range(self.world)]
        self.exchange_records = 0           # records this rank received in the last exchange
        self.fused = None                   # FusedTriangle once the fused gather has been set up
        self.fused_error = None
        self.gather_mode = os.environ.get("PC_GATHER", "push")     # "push" (copy engines) | "stores" (kernel)
        self.push_blocks = int(os.environ.get("PC_PUSH_BLOCKS", "4"))
        self._push_streams = None

    def load(self, table=None):
        """Host buffers -> this rank's device: the table, and for aai/peq only the genes of the
        phams this rank aligns."""
        self.ops.load(table if table is not None else self.table, self.rank, self.world)

    def accumulators(self):
        """aai/peq: align this rank's phams, exchange the contribution records, return the
        accumulators of this rank's rows (None for the other metrics)."""
        if not self.align:
            return None
        import torch
        dist = self.dist
        records, counts = self.ops.contributions(self.rank, self.world, self.bounds)
        if self.world > 1:
            send_counts = torch.as_tensor(np.asarray(counts, dtype=np.int64), device=records.device)
            recv_counts = torch.empty_like(send_counts)
            dist.all_to_all_single(recv_counts, send_counts)
            recv_counts = recv_counts.tolist()
            received = torch.empty((int(sum(recv_counts)), 2), dtype=torch.int64, device=records.device)
            dist.all_to_all_single(received, records[:int(sum(counts))], recv_counts, [int(c) for c in counts])
        else:
            received = records[:int(sum(counts))]
        self.exchange_records = int(received.shape[0])
        return self.ops.accumulate(received, self.r0, self.r1)

    def row_block(self):
        """This rank's rows of the packed triangle on its device (int32 tensor)."""
        return self.ops.rows(self.r0, self.r1, self.accumulators())

    def row_block_to_host(self, out):
        """... written straight into ``out`` (a host view of exactly this block)."""
        acc = self.accumulators()
        if self.r1 > self.r0:
            self.ops.rows_to_host(self.r0, self.r1, acc, out)

    def enable_fused_gather(self):
        """Set up the symmetric-memory triangle for ``gathered_triangle``; returns False (and keeps
        the all-gather path) where that is not possible.  Collective: every rank calls it."""
        if self.fused is not None:
            return True
        if self.world == 1 or not hasattr(self.ops, "rows_fused"):
            return False
        import torch
        ok = torch.ones(1, dtype=torch.int32, device=self.ops.device)
        try:
            fused = FusedTriangle(self.n, self.ops.device)
        except Exception as exc:              # no symmetric memory here: say why, fall back
            self.fused_error = f"{type(exc).__name__}: {exc}"
            fused = None
            ok.zero_()
        self.dist.all_reduce(ok, op=self.dist.ReduceOp.MIN)
        if int(ok.item()) == 0:
            return False
        self.fused = fused
        return True

    def gathered_triangle(self, mode=None):
        """Device consumer: the whole packed triangle on this rank's device after every rank has
        computed its row block.  With the symmetric-memory triangle set up, two ways to get a row
        block into the peers' triangles (``mode``, default ``self.gather_mode``):

        * ``"push"``: the kernels write this rank's triangle only; the block is cut into a few
          row sub-blocks and, as soon as a sub-block is done, copy engines push it to every peer
          (one stream per peer, ``cudaMemcpyAsync`` between peer-mapped buffers) while the next
          sub-block is being computed -- no SM takes part in the transfer;
        * ``"stores"``: the producing kernel itself stores every row segment into all triangles
          (``pc_matrix_rows_fused``).  Measured: free at 2 GPUs, but an SM sustains well under
          1 GB/s of peer stores, so from 4 GPUs on the epilogue becomes the bottleneck.

        Without symmetric memory: row blocks + one all-gather (returned padded per rank: use
        ``assemble_host``)."""
        if self.fused is None:
            return self.all_gather(self.row_block())
        mode = mode or self.gather_mode
        acc = self.accumulators()
        if mode == "stores":
            own, peers = self.fused.slice_ptrs(self.r0)
            if self.r1 > self.r0:
                self.ops.rows_fused(self.r0, self.r1, acc, own, peers)
            self.fused.barrier()
            return self.fused.tensor
        import torch
        fused = self.fused
        peers = [p for p in range(self.world) if p != self.rank]
        if self._push_streams is None:
            self._push_streams = [torch.cuda.Stream(device=self.ops.device) for _ in peers]
        main = torch.cuda.current_stream(self.ops.device)
        cuts = self._push_cuts()
        for a, b in zip(cuts, cuts[1:]):
            if b <= a:
                continue
            lo, hi = tri_offset(self.n, a), tri_offset(self.n, b)
            block = fused.tensor[lo:hi]
            if self.align:
                off = tri_offset(self.n, a) - tri_offset(self.n, self.r0)
                self.ops.rows(a, b, acc[off:], out=block)
            else:
                self.ops.rows(a, b, None, out=block)
            done = torch.cuda.Event()
            done.record(main)
            for stream, peer in zip(self._push_streams, peers):
                with torch.cuda.stream(stream):
                    stream.wait_event(done)
                    fused.handle.get_buffer(peer, (hi - lo,), torch.int32, lo).copy_(block, non_blocking=True)
        for stream in self._push_streams:
            main.wait_stream(stream)
        fused.barrier()
        return fused.tensor

    def _push_cuts(self):
        """Row boundaries of the sub-blocks of this rank's block (about equal work each)."""
        k = max(1, min(self.push_blocks, self.r1 - self.r0))
        if 4 * (tri_offset(self.n, self.r1) - tri_offset(self.n, self.r0)) < (64 << 20):
            k = 1                                          # a small block: one launch, one push
        if k == 1 or self.align:
            sub = row_partition(self.r1 - self.r0, k) if k > 1 else [0, self.r1 - self.r0]
            # (pair-balanced cut of a block that starts at row r0 of an n-row triangle: good enough)
            return [self.r0 + x for x in sub]
        w = self._row_weight[self.r0:self.r1] if self._row_weight is not None else None
        if w is None:
            return [self.r0 + (self.r1 - self.r0) * i // k for i in range(k + 1)]
        cum = np.concatenate(([0.0], np.cumsum(w)))
        cuts = [self.r0]
        for i in range(1, k):
            c = int(np.searchsorted(cum, cum[-1] * i / k, side="left"))
            cut = (self.r0 + c) // 256 * 256               # whole 256-row tiles of the contraction
            cuts.append(min(max(cut, cuts[-1]), self.r1))
        cuts.append(self.r1)
        return cuts

    def all_gather(self, local, gathered=None):
        """Device consumer: every rank receives every row block (padded to the widest)."""
        import torch
        width = max(max(self.sizes), 1)
        padded = local
        if local.numel() != width:
            padded = torch.zeros(width, dtype=torch.int32, device=local.device)
            padded[:local.numel()] = local
        if self.world == 1:
            return padded
        if gathered is None:
            gathered = torch.empty(self.world * width, dtype=torch.int32, device=local.device)
        self.dist.all_gather_into_tensor(gathered, padded)
        return gathered

    def assemble_host(self, gathered):
        width = max(max(self.sizes), 1)
        host = gathered.cpu().numpy().reshape(self.world, width)
        return assemble(self.n, self.bounds, [host[r] for r in range(self.world)])


def sharded_matrix(table, metric, as_distance=True, ppos=False, backend_ops=None, fused=True):
    """Packed int32 micro-unit triangle computed by all ranks of the default process group;
    every rank returns the full triangle (host numpy)."""
    builder = ShardedBuilder(table, metric, as_distance, ppos, backend_ops)
    builder.load()
    if fused and builder.enable_fused_gather():
        return builder.gathered_triangle().cpu().numpy()
    return builder.assemble_host(builder.all_gather(builder.row_block()))


class SharedTriangle:
    """The packed int32 triangle of an n-genome matrix in POSIX shared memory:
    created by rank 0, attached by every rank of the node; with a GPU present each
    rank page-locks the rows it will write (``pin_rows=(r0, r1)``, rounded out to
    whole pages; ``None`` = its pair-balanced ``row_partition`` block) so that its D2H is a
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


def sharded_matrix_to_host(table, metric, shared=None, as_distance=True, ppos=False, backend_ops=None,
                           builder=None):
    """Every rank computes its row block and copies it from its own GPU straight into
    ``shared`` (a SharedTriangle pinned for the builder's rows); returns after a barrier, when
    the whole triangle is in host memory (``shared.array``, valid on every rank of the node)."""
    if builder is None:
        builder = ShardedBuilder(table, metric, as_distance, ppos, backend_ops)
        builder.load()
    if shared is None:
        shared = SharedTriangle(builder.n, pin_rows=(builder.r0, builder.r1))
    builder.row_block_to_host(shared.rows(builder.r0, builder.r1))
    if builder.world > 1:
        builder.dist.barrier()
    return shared.array


class _GpuOps:
    """The real thing: this rank's GPU through the C ABI, on torch's stream so that NCCL
    collectives order after the kernels."""

    def __init__(self, metric, as_distance, ppos):
        import torch
        from phamclust_b200.matrix import get_context
        self.torch = torch
        local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local_rank)
        self.device = torch.device("cuda", local_rank)
        self.ctx = get_context(local_rank)
        self.ctx.set_stream(torch.cuda.current_stream().cuda_stream)
        self.metric, self.as_distance, self.ppos = metric, as_distance, ppos
        self.n = 0
        self._records = None
        self._acc = None

    def load(self, table, part, n_parts):
        self.ctx.load_table(table)
        self.n = len(table.genome_ptr) - 1
        if self.metric in ALIGN_METRICS:
            if table.seq_off is None:
                raise ValueError(f"metric {self.metric} needs translations")
            self.ctx.load_sequences(table, part, n_parts)

    def contributions(self, part, n_parts, bounds):
        torch = self.torch
        n_units = self.ctx.align_units(part, n_parts)
        if self._records is None or self._records.shape[0] < n_units:
            self._records = torch.empty((max(n_units, 1), 2), dtype=torch.int64, device=self.device)
        counts = self.ctx.align_contributions(self._records.data_ptr(), bounds, part, n_parts, ppos=self.ppos)
        return self._records, counts

    def accumulate(self, records, r0, r1):
        torch = self.torch
        count = tri_offset(self.n, r1) - tri_offset(self.n, r0)
        if self._acc is None or self._acc.numel() != max(count, 1):
            self._acc = torch.empty(max(count, 1), dtype=torch.int64, device=self.device)
        self._acc.zero_()
        self.ctx.accumulate_records(int(records.shape[0]), records.data_ptr(), count, self._acc.data_ptr())
        return self._acc

    def rows(self, r0, r1, acc_rows, out=None):
        count = tri_offset(self.n, r1) - tri_offset(self.n, r0)
        if out is None:
            out = self.torch.empty(max(count, 1), dtype=self.torch.int32, device=self.device)
        if count:
            self.ctx.matrix_rows_owned(self.metric, self.as_distance, r0, r1,
                                       acc_rows.data_ptr() if acc_rows is not None else None,
                                       out_device_ptr=out.data_ptr(), ppos=self.ppos)
        return out[:count]

    def rows_fused(self, r0, r1, acc_rows, out_ptr, peer_ptrs):
        """Row block stored into this rank's triangle and, by the same kernel, into the peers'."""
        self.ctx.matrix_rows_fused(self.metric, self.as_distance, r0, r1,
                                   acc_rows.data_ptr() if acc_rows is not None else None, out_ptr, peer_ptrs,
                                   ppos=self.ppos)

    def rows_to_host(self, r0, r1, acc_rows, out):
        """Row block computed on this GPU, D2H by the library into ``out`` (page-locked
        shared memory: a full-speed DMA on this GPU's own link)."""
        if r1 > r0:
            self.ctx.matrix_rows_owned(self.metric, self.as_distance, r0, r1,
                                       acc_rows.data_ptr() if acc_rows is not None else None, out=out,
                                       ppos=self.ppos)
