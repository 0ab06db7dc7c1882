"""The several-GPU data plane on real hardware.

* One GPU is enough to check the kernels behind the exchange: partition by partition, ship only
  that partition's genes (pc_load_sequences_part), emit its contribution records
  (pc_align_contributions), route the buckets by hand to their owners, accumulate
  (pc_accumulate_records) and run the owned-rows epilogue (pc_matrix_rows_owned) -- the assembled
  triangle must equal the single-partition dense path entry for entry, and the oracle.
* With two GPUs visible, two NCCL ranks run ``ShardedBuilder`` for real (all-to-all of the
  records, all-gather of the row blocks, the shared host triangle) and compare the WHOLE triangle
  with the 1-GPU result -- array_equal, not a checksum.
"""
import os
import socket

import numpy as np
import pytest

from oracle import pham_oracle as po
from phamclust_b200 import distributed, synthetic
from phamclust_b200._native import tri_len
from phamclust_b200.packing import unpack_genomes

pytestmark = pytest.mark.gpu


def _torch():
    import torch
    return torch


@pytest.mark.parametrize("n_parts,metric,ship_all", [(2, "peq", "0"), (3, "aai", "1"), (5, "peq", "0"), (2, "peq", "1")])
def test_partitioned_records_equal_the_dense_path(gpu_ctx, monkeypatch, n_parts, metric, ship_all):
    # PC_SHIP_ALL: a rank either gathers its genes on the host or ships the whole arena and lets
    # the device skip the others (what pc_load_sequences_part picks for page-locked buffers below 8 ranks)
    monkeypatch.setenv("PC_SHIP_ALL", ship_all)
    torch = _torch()
    dev = torch.device("cuda", gpu_ctx.device)
    table = synthetic.phage_table(260, seed=11, genomes_per_cluster=40, pool=90)
    n = table.n_genomes
    gpu_ctx.set_stream(torch.cuda.current_stream(dev).cuda_stream)
    try:
        gpu_ctx.load_table(table)
        gpu_ctx.load_sequences(table)
        dense = gpu_ctx.matrix(metric, as_distance=True)
        hit_d, len_d = gpu_ctx.numerators("aai")
        bounds = distributed.row_partition(n, n_parts)
        accs = [torch.zeros(max(tri_len(n, bounds[d], bounds[d + 1]), 1), dtype=torch.int64, device=dev)
                for d in range(n_parts)]
        total_records = 0
        for part in range(n_parts):
            gpu_ctx.load_sequences(table, part, n_parts)           # only this partition's genes
            n_units = gpu_ctx.align_units(part, n_parts)
            records = torch.empty((max(n_units, 1), 2), dtype=torch.int64, device=dev)
            counts = gpu_ctx.align_contributions(records.data_ptr(), bounds, part, n_parts)
            assert int(counts.sum()) == n_units
            total_records += n_units
            start = 0
            for d in range(n_parts):                               # the all-to-all, by hand
                bucket = records[start:start + int(counts[d])].contiguous()
                start += int(counts[d])
                if bucket.shape[0]:
                    assert int(bucket[:, 0].min()) >= 0 and int(bucket[:, 0].max()) < accs[d].numel()
                gpu_ctx.accumulate_records(int(bucket.shape[0]), bucket.data_ptr(), accs[d].numel(), accs[d].data_ptr())
        assert total_records > 0
        gpu_ctx.load_sequences(table, 0, n_parts)
        got = np.concatenate([
            gpu_ctx.matrix_rows_owned(metric, True, bounds[d], bounds[d + 1], accs[d].data_ptr())
            for d in range(n_parts) if bounds[d + 1] > bounds[d]])
        assert np.array_equal(got, dense)
        # the integer sums themselves
        acc_all = torch.cat([a[:tri_len(n, bounds[d], bounds[d + 1])] for d, a in enumerate(accs)]).cpu().numpy()
        off_diag = np.ones(tri_len(n), dtype=bool)
        off_diag[[distributed.tri_offset(n, i) for i in range(n)]] = False
        assert np.array_equal((acc_all >> 32)[off_diag], hit_d.astype(np.int64)[off_diag])
        assert np.array_equal((acc_all & 0xFFFFFFFF)[off_diag], len_d.astype(np.int64)[off_diag])
        # and a few rows against the oracle
        genomes = unpack_genomes(table)
        for i, j in ((0, 1), (3, 39), (41, 77), (100, 119), (200, 259)):
            want = po.round6_micro(po.METRICS[metric](genomes[i], genomes[j], as_distance=True))
            assert abs(int(dense[distributed.tri_offset(n, i) + j - i]) - want) <= 1
        # a plan for another partition than the one whose genes are on the device is refused
        with pytest.raises(ValueError):
            gpu_ctx.align_units(1, n_parts)
    finally:
        gpu_ctx.set_stream(0, external=False)
        gpu_ctx.load_table(table)
        gpu_ctx.load_sequences(table)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _nccl_worker(rank, world, port, queue):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), LOCAL_RANK=str(rank),
                      RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        out = {}
        for metric, table in (("peq", synthetic.phage_table(300, seed=5)),
                              ("jc", synthetic.sparse_table(1500, 9000, seed=6))):
            n = table.n_genomes
            got = distributed.sharded_matrix(table, metric, fused=False)        # row blocks + NCCL all-gather
            builder = distributed.ShardedBuilder(table, metric)
            builder.load()
            # the gather fused into the producing kernels (peer stores over NVLink, symmetric memory)
            fused_ok = builder.enable_fused_gather()
            if fused_ok:
                for mode in ("push", "stores", "push"):                         # reuse of the triangle too
                    tri = builder.gathered_triangle(mode)
                    torch.cuda.synchronize()
                    assert np.array_equal(tri.cpu().numpy(), got), (metric, mode)
                    tri.zero_()
                    torch.cuda.synchronize()
                    dist.barrier()
            shared = distributed.SharedTriangle(n, pin_rows=(builder.r0, builder.r1))
            try:
                host = distributed.sharded_matrix_to_host(table, metric, shared, builder=builder).copy()
                dist.barrier()
            finally:
                shared.close()
            assert np.array_equal(host, got)
            out[metric] = (got, builder.bounds, builder.exchange_records, fused_ok, builder.fused_error)
        queue.put((rank, out))
    finally:
        dist.destroy_process_group()


def test_two_nccl_ranks_equal_one_gpu(gpu_ctx):
    torch = _torch()
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 visible GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    from phamclust_b200 import packed_matrix
    ctx = mp.get_context("spawn")
    queue = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_nccl_worker, args=(r, 2, port, queue)) for r in range(2)]
    for p in procs:
        p.start()
    results = dict(queue.get(timeout=600) for _ in range(2))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for metric, table in (("peq", synthetic.phage_table(300, seed=5)),
                          ("jc", synthetic.sparse_table(1500, 9000, seed=6))):
        single = packed_matrix(table, metric)
        for rank in (0, 1):
            got, bounds, records, fused_ok, fused_error = results[rank][metric]
            assert fused_ok, fused_error                 # a B200 box has NVLink peer access: the fused path must run
            assert np.array_equal(got, single), (metric, rank)
            assert bounds[0] == 0 and bounds[-1] == table.n_genomes
        if metric == "peq":
            assert results[0][metric][2] + results[1][metric][2] > 0


def test_resident_matrix_clusters_without_shipping_the_triangle(gpu_ctx, monkeypatch):
    """The nr / clustering hand-off on the triangle kept in HBM: threshold edges and component
    blocks come from the device (pc_resident_threshold_edges / pc_resident_gather); partitions,
    sub-matrices and medoids equal those of the host SymMatrix path, for a fraction of the D2H."""
    from phamclust_b200 import METRICS, clustering, matrix_de_novo, matrix_de_novo_resident
    monkeypatch.setattr(clustering, "DECOMPOSE_MIN_NODES", 64)
    table = synthetic.phage_table(900, seed=13, genomes_per_cluster=30, pool=100, with_sequences=False)
    host = matrix_de_novo(table, METRICS["gcs"], 1)
    res = matrix_de_novo_resident(table, METRICS["gcs"], 1)
    assert len(res) == 900 and res.is_distance and res.nodes == host.nodes
    for eps in (0.25, 0.6):
        want_labels = host.threshold_components(int(round(eps * 1e6)) + 1)
        got_labels = res.threshold_components(int(round(eps * 1e6)) + 1)
        # same partition (label values may differ)
        assert len(set(zip(want_labels.tolist(), got_labels.tolist()))) == len(set(want_labels.tolist()))
        a = clustering.hierarchical_clustering(host, "average", eps=eps)
        b = clustering.hierarchical_clustering(res, "average", eps=eps)
        assert [sorted(m.nodes) for m in a] == [sorted(m.nodes) for m in b]
        for x, y in zip(a, b):
            assert np.array_equal(x.extract_submatrix(sorted(x.nodes)).micro, y.extract_submatrix(sorted(y.nodes)).micro)
            assert x.extract_submatrix(sorted(x.nodes)).medoid == y.extract_submatrix(sorted(y.nodes)).medoid
    assert 0 < res.d2h_bytes < host.micro.nbytes // 2
    assert np.array_equal(res.to_symmatrix().micro, host.micro)
    gpu_ctx_other = matrix_de_novo(table, METRICS["jc"], 1)          # reloads the table on the same ctx
    with pytest.raises(RuntimeError):
        res.threshold_components(250001)
    assert len(gpu_ctx_other) == 900
