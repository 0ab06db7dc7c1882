"""world_size-2 gloo run of the multi-GPU host logic (row partition, all-to-all exchange of
the sparse contribution records, owner-side accumulation, all-gather / shared-memory assembly)
with the CPU oracle standing in for the per-rank GPU ops -- the N>1 path without a GPU."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import pham_oracle as po
from oracle import testdata
from phamclust_b200 import distributed
from phamclust_b200.genome import Genome
from phamclust_b200.packing import pack_genomes


class OracleOps:
    """Per-rank compute restated with the oracle (CPU tensors): the interface ShardedBuilder
    drives (load / contributions / accumulate / rows / rows_to_host)."""

    def __init__(self, genomes, metric, as_distance):
        self.genomes, self.metric, self.as_distance = genomes, metric, as_distance
        self.n = len(genomes)
        vocab = sorted({p for g in genomes for p in g.phams})
        self.pham_rank = {p: k for k, p in enumerate(vocab)}
        self.loaded = None

    def load(self, table, part, n_parts):
        self.loaded = (part, n_parts)

    def contributions(self, part, n_parts, bounds):
        """One record per (pham, genome pair) of this partition's phams, grouped by the row
        block that owns the pair: [index inside the owner's block, (hit << 32) | length]."""
        buckets = [[] for _ in range(len(bounds) - 1)]
        for i in range(self.n):
            dest = next(d for d in range(len(bounds) - 1) if bounds[d] <= i < bounds[d + 1])
            for j in range(i + 1, self.n):
                for pham, hit, length in po.aai_best_matches(self.genomes[i], self.genomes[j]):
                    if self.pham_rank[pham] % n_parts == part:
                        local = distributed.tri_offset(self.n, i) + (j - i) - distributed.tri_offset(self.n, bounds[dest])
                        buckets[dest].append([local, (hit << 32) | length])
        flat = [rec for b in buckets for rec in b]
        records = torch.tensor(flat, dtype=torch.int64).reshape(-1, 2)
        return records, [len(b) for b in buckets]

    def accumulate(self, records, r0, r1):
        count = distributed.tri_offset(self.n, r1) - distributed.tri_offset(self.n, r0)
        acc = torch.zeros(max(count, 1), dtype=torch.int64)
        for local, packed in records.tolist():
            acc[local] += packed
        return acc

    def rows(self, r0, r1, acc):
        out = []
        base = distributed.tri_offset(self.n, r0)
        for i in range(r0, r1):
            for j in range(i, self.n):
                if i == j:
                    out.append(0 if self.as_distance else 1000000)
                    continue
                s, t = self.genomes[i], self.genomes[j]
                if self.metric in ("aai", "peq"):
                    v = int(acc[distributed.tri_offset(self.n, i) + (j - i) - base])
                    hit, length = v >> 32, v & 0xFFFFFFFF
                    sim = round(hit / length, 6) if length else 0.0
                    if self.metric == "peq":
                        sim = po.alignment_fraction(s, t) * sim
                    w = round(1.0 - sim, 6) if self.as_distance else round(sim, 6)
                else:
                    w = po.METRICS[self.metric](s, t, as_distance=self.as_distance)
                out.append(po.round6_micro(w))
        return torch.tensor(out, dtype=torch.int32)

    def rows_to_host(self, r0, r1, acc, out):
        out[:] = self.rows(r0, r1, acc).numpy()

    def single_rank(self):
        """The whole triangle through the same ops with one partition and one row block."""
        records, counts = self.contributions(0, 1, [0, self.n])
        acc = self.accumulate(records, 0, self.n) if self.metric in ("aai", "peq") else None
        return self.rows(0, self.n, acc).tolist()


def _worker(rank, world, port, rows, metric, queue):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        genomes = testdata.rows_to_genomes(rows, cls=Genome)
        table = pack_genomes(genomes)
        assert distributed.is_active()
        got = distributed.sharded_matrix(table, metric, as_distance=True,
                                         backend_ops=OracleOps(genomes, metric, True))
        # the host-consumer exchange: every rank writes its row block into one shared triangle
        shared = distributed.SharedTriangle(len(genomes))
        try:
            ops = OracleOps(genomes, metric, True)
            host = distributed.sharded_matrix_to_host(table, metric, shared, as_distance=True, backend_ops=ops)
            assert ops.loaded == (rank, world)
            assert host.tolist() == got.tolist()
            dist.barrier()
        finally:
            shared.close()
        queue.put((rank, got.tolist()))
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("metric", ["gcs", "peq"])
def test_two_rank_sharding_equals_single_rank(metric):
    rows = testdata.genome_rows(seed=21, n_genomes=7, n_phams=10, lo=10, hi=40, paralog_p=0.3)
    genomes = testdata.rows_to_genomes(rows, cls=Genome)
    ops = OracleOps(genomes, metric, True)
    single = ops.single_rank()
    # and the single-rank result is the oracle's matrix
    w = po.matrix_de_novo(genomes, po.METRICS[metric], 1)
    assert single == [po.round6_micro(w[(i, j)]) for i in range(len(genomes))
                      for j in range(i, len(genomes))]
    ctx = mp.get_context("spawn")
    queue = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, rows, metric, queue)) for r in range(2)]
    for p in procs:
        p.start()
    results = dict(queue.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert results[0] == single and results[1] == single


def test_more_ranks_than_rows_leaves_empty_blocks():
    """3 ranks, 2 genomes: one rank owns no row at all -- both exchanges must still agree."""
    rows = testdata.genome_rows(seed=4, n_genomes=2, n_phams=6, lo=10, hi=30, paralog_p=0.3)
    genomes = testdata.rows_to_genomes(rows, cls=Genome)
    assert distributed.row_partition(2, 3) == [0, 1, 1, 2]
    ops = OracleOps(genomes, "peq", True)
    single = ops.single_rank()
    ctx = mp.get_context("spawn")
    queue = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 3, port, rows, "peq", queue)) for r in range(3)]
    for p in procs:
        p.start()
    results = dict(queue.get(timeout=120) for _ in range(3))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert results[0] == results[1] == results[2] == single


def test_assemble_trims_padding():
    n, bounds = 5, [0, 2, 5]
    full = np.arange(15, dtype=np.int32)
    shards = [np.concatenate((full[:9], [99, 99])), np.concatenate((full[9:], [77] * 5))]
    assert np.array_equal(distributed.assemble(n, bounds, shards), full)


def test_work_balanced_row_partition():
    """Row blocks cut by estimated contraction work: late rows of an unclustered table occupy
    more K blocks, so their blocks hold fewer pairs; every row lands in exactly one block and the
    estimated work per block is even."""
    from phamclust_b200 import synthetic
    table = synthetic.sparse_table(3000, 20000, seed=9)
    w = distributed.contraction_row_weights(table)
    assert w.shape == (3000,) and (w > 0).all()
    for world in (2, 3, 8):
        b = distributed.row_partition(3000, world, w)
        assert b[0] == 0 and b[-1] == 3000 and all(x <= y for x, y in zip(b, b[1:]))
        share = [w[b[r]:b[r + 1]].sum() / w.sum() for r in range(world)]
        assert max(share) - min(share) < 0.02, share
    assert distributed.row_partition(5, 2) == [0, 2, 5]          # pair-balanced default unchanged
