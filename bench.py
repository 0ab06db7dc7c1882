#!/usr/bin/env python
"""bench.py -- genome pairs/second of the all-vs-all matrix build (BASELINE.json).

    python bench.py --gpus 1 --steps 5 --warmup 3                 # this repo, 1 B200
    torchrun --nproc-per-node N ... bench.py --gpus N ...         # N ranks, NCCL
    python bench.py --impl reference ...                          # host-CPU arm

Headline workload (``config.workload``): BASELINE.json configs[1] -- metric=peq on the bundled
``benchmark_data.tsv``.  That file is absent from the reference mount (.MISSING_LARGE_BLOBS), so
the deterministic stand-in of SURVEY.md 8(d) is used: ``synthetic.phage_table(2000, seed=42)``.
A *step* is one pass of the hot path (matrix_de_novo for that metric over the whole table).

* ``value``  : pairs/s with the packed table + residues already resident in HBM (CUDA events on
               the stream the kernels run on, L2 flushed between steps, max over ranks).  N = 1:
               ``pc_matrix`` into device memory; N > 1: ``distributed.ShardedBuilder.row_block`` +
               ``all_gather`` (pham-dealt alignments, all-to-all of the sparse contribution
               records, owner-side accumulation, row-block epilogue, NCCL all-gather).
* ``e2e``    : the same metric through the C ABI with (pinned) HOST buffers every step:
               pc_load_table + pc_load_sequences[_part] + the matrix, H2D and D2H inside the timed
               region.  N > 1: every rank ships only its phams' genes and copies its row block from
               its own GPU into one page-locked shared-memory triangle (the consumer is the host).
* ``e2e_api``: N = 1: ``matrix_de_novo(list[Genome], METRICS[m], cpus)`` -- the reference's own
               boundary (matrix.py:432) -- timed from the Python objects: the C packer + e2e.
* ``workloads``: the other BASELINE.json configs the driver's default command should see, each a
               full record (value / e2e / roofline / clocks): N = 1 -> configs[2] C3 gcs and jc
               (tensor-bound contraction); N > 1 -> configs[4] C5 gcs; N = 8 -> configs[3] C4 peq.
* ``roofline``: the dominant kernel (the alignment DP for peq/aai, the shared-count contraction
               for gcs/jc/pocp/af).  Instruction counts per DP cell come from the SASS of the
               shipped library (tools/sass_roofline.py -> profiles/sass_dp_instr.json), not from
               literals; ``traffic`` from the committed ncu summary (profiles/ncu_traffic.json).
* ``cpu_baseline``: the oracle port of the reference's CPU path on a bounded RANDOM sample
               (``random.Random(42).sample``, as scripts/benchmark.py:96-100 samples), timed on
               this box's host cores (rank 0, N = 1 only), with alignments per pair and CPU GCUPS
               so that the two arms can be compared on equal work.
"""
import argparse
import copy
import json
import os
import random
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ALIGN_METRICS = ("aai", "peq")
SM_COUNT = 148
DEFAULTS = {"c2": ("peq", 2000, None), "c3": ("jc", 10000, 60000), "c4": ("peq", 50000, None),
            "c5": ("gcs", 100000, 200000)}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return json.load(fh), "measured"
    # fallback stated in /opt/skills/guides/B200_PROFILING.md
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0,
            "sm_max_mhz": 1965.0}, "fallback"


def dp_instruction_counts():
    """Per-cell SASS instruction counts of the DP kernel families, derived from the shipped
    library by tools/sass_roofline.py.  Regenerated on the spot when the library changed since
    (cuobjdump is part of the image); marked stale if that is impossible."""
    from phamclust_b200 import build_native
    path = os.path.join(ROOT, "profiles", "sass_dp_instr.json")
    sha = build_native.library_sha256()
    data = None
    if os.path.exists(path):
        with open(path) as fh:
            data = json.load(fh)
    state = "profiles/sass_dp_instr.json (same library sha256)"
    if data is None or data.get("library_sha256") != sha:
        try:
            subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sass_roofline.py")], check=True,
                           capture_output=True, timeout=300)
            with open(path) as fh:
                data = json.load(fh)
            state = "re-derived from this library's SASS at run time (tools/sass_roofline.py)"
        except Exception as exc:               # no cuobjdump: keep the committed counts, say so
            if data is None:
                raise SystemExit(f"no SASS instruction counts and cuobjdump failed: {exc}")
            state = "STALE: committed counts belong to another build of the library"
    fam = data["families"]
    return ({k: fam[k]["alu_per_cell"] for k in fam}, {k: fam[k]["instr_per_cell"] for k in fam}, state)


def ncu_traffic(key):
    """dram__bytes_read + dram__bytes_write per launch of the dominant kernel from the committed
    ncu --set full summary (profiles/ncu_traffic.json), or None."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(path):
        return None, None
    with open(path) as fh:
        data = json.load(fh)
    rec = data.get(key)
    return (rec["dram_bytes_per_launch"], rec.get("source")) if rec else (None, None)


class ClockSampler:
    """nvidia-smi clocks + throttle reasons DURING the timed region."""
    QUERY = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.rows, self.proc = device, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.device}", f"--query-gpu={self.QUERY}",
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def __exit__(self, *exc):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for row in self.rows:
            try:
                sm.append(float(row[0])); mx.append(float(row[1]))
            except (ValueError, IndexError):
                continue
            for name, val in zip(names, row[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)),
                "reasons": sorted(reasons), "samples": len(sm)}


def build_workload(workload, metric, genomes, phams):
    from phamclust_b200 import synthetic
    if workload == "c2":
        table = synthetic.phage_table(genomes, seed=42)
        name = (f"configs[1]: metric={metric} on stand-in for the missing benchmark_data.tsv = "
                f"synthetic.phage_table(n_genomes={genomes}, seed=42) [SURVEY 8(d) C1/C2]")
    elif workload == "c3":
        table = synthetic.sparse_table(genomes, phams, seed=1)
        if metric in ALIGN_METRICS:
            synthetic.attach_m_translations(table)
        name = (f"configs[2]: metric={metric} on synthetic.sparse_table("
                f"n_genomes={genomes}, n_phams={phams}, seed=1)")
    elif workload == "c4":
        table = synthetic.phage_table(genomes, seed=2)
        name = (f"configs[3]: metric={metric}, pham-sharded alignment, on synthetic.phage_table("
                f"n_genomes={genomes}, seed=2) [SURVEY 8(d) C4: clusters of 50 genomes x 150-pham pools]")
    elif workload == "c5":
        table = synthetic.sparse_table(genomes, phams, seed=3)
        name = (f"configs[4]: metric={metric}, full upper triangle in row blocks + NCCL gather, on "
                f"synthetic.sparse_table(n_genomes={genomes}, n_phams={phams}, seed=3) [SURVEY 8(d) C5]")
    else:
        raise SystemExit(f"unknown workload {workload}")
    return table, name


def alignment_work(table):
    """(alignments, DP cells) the REFERENCE runs for this table: for every pham and every pair of
    member genomes all copies x copies gene pairs are aligned (metrics.py:203-221; no dedup there),
    cells = len_a * len_b summed = product of the two members' summed lengths."""
    pid = np.asarray(table.pham_id, dtype=np.int64)
    c = np.asarray(table.copies, dtype=np.float64)
    ln = np.asarray(table.len_sum, dtype=np.float64)
    n_phams = int(table.n_phams)
    sc, sc2 = np.bincount(pid, weights=c, minlength=n_phams), np.bincount(pid, weights=c * c, minlength=n_phams)
    sl, sl2 = np.bincount(pid, weights=ln, minlength=n_phams), np.bincount(pid, weights=ln * ln, minlength=n_phams)
    return float(((sc * sc - sc2) / 2).sum()), float(((sl * sl - sl2) / 2).sum())


def pinned_copy(torch, arr):
    t = torch.from_numpy(np.ascontiguousarray(arr))
    try:
        return t.pin_memory()
    except RuntimeError:
        return t


# --------------------------------------------------------------------------------------
# the reference arm / cpu_baseline: oracle port of the reference's CPU path
# --------------------------------------------------------------------------------------
def cpu_reference_rate(table, metric, sample_genomes, cpus_list, seed=42):
    """pairs/s of the oracle port of matrix_de_novo (one joblib task per pair, exactly like
    matrix.py:471-491) on ``sample_genomes`` genomes drawn with ``random.Random(seed).sample`` --
    the reference benchmark's own protocol (scripts/benchmark.py:23, 96-100), so that the share
    of genome pairs that have phams in common (and therefore alignments to run) is the workload's.
    The worker count is picked on a small calibration sample (the reference's per-pair joblib
    dispatch makes more workers SLOWER for the cheap metrics)."""
    from oracle import nw_c, pham_oracle as po
    from phamclust_b200.packing import take_genomes, unpack_genomes
    nw_c.build()
    n_all = table.n_genomes
    k = min(sample_genomes, n_all)
    idx = random.Random(seed).sample(range(n_all), k)
    sub = take_genomes(table, idx)
    genomes = unpack_genomes(sub)
    n = len(genomes)
    pairs = n * (n - 1) // 2
    runs = {}
    if len(cpus_list) > 1:
        calib = genomes[:min(n, 60 if metric in ALIGN_METRICS else 300)]
        cp = len(calib) * (len(calib) - 1) // 2
        for cpus in cpus_list:
            t0 = time.perf_counter()
            po.matrix_de_novo(calib, po.METRICS[metric], cpus)
            runs[cpus] = cp / (time.perf_counter() - t0)
        cpus = max(runs, key=runs.get)
    else:
        cpus = cpus_list[0]
    t0 = time.perf_counter()
    po.matrix_de_novo(genomes, po.METRICS[metric], cpus)
    dt = time.perf_counter() - t0
    out = {"value": pairs / dt, "cores": cpus, "seconds": dt, "pairs": pairs,
           "runs": {str(k_): v for k_, v in runs.items()}, "n": n}
    if metric in ALIGN_METRICS and sub.seq_off is not None:
        aln, cells = alignment_work(sub)
        out.update(alignments=aln, alignments_per_pair=aln / max(pairs, 1), alignments_per_s=aln / dt,
                   gcups=cells / dt / 1e9, gcups_per_core=cells / dt / 1e9 / cpus)
    return out


def cpu_baseline_record(res, metric):
    txt = (f"random.Random(42).sample of {res['n']} genomes of the workload ({res['pairs']} pairs, "
           f"{res['seconds']:.1f} s), as scripts/benchmark.py:96-100 samples; oracle port of matrix_de_novo "
           f"(one joblib task per pair) + C restatement of parasail NW (scalar, not parasail's SIMD kernel: "
           f"parasail publishes ~1-3 GCUPS per core for nw_trace_diag_16, rescale with gcups_per_core); "
           f"cpus chosen on a calibration sample, pairs/s by cpus there: "
           f"{ {k: round(v, 1) for k, v in res['runs'].items()} }")
    rec = {"value": res["value"], "unit": "pairs/s", "cores": res["cores"], "kind": "port", "sample": txt}
    for key in ("alignments_per_pair", "alignments_per_s", "gcups", "gcups_per_core"):
        if key in res:
            rec[key] = res[key]
    return rec


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    table, workload = build_workload(args.workload, args.metric, args.genomes, args.phams)
    n_cpus = os.cpu_count() or 1
    sample = args.cpu_sample or (400 if args.metric in ALIGN_METRICS else 1000)
    # warm-up: page the oracle in and pick the worker count once
    cal = cpu_reference_rate(table, args.metric, min(sample, 120), sorted({1, n_cpus}))
    cpus_list = [cal["cores"]]
    for _ in range(max(args.warmup - 1, 0)):
        cpu_reference_rate(table, args.metric, min(sample, 120), cpus_list)
    values, secs = [], []
    res = None
    for _ in range(args.steps):
        res = cpu_reference_rate(table, args.metric, sample, cpus_list)
        values.append(res["value"]); secs.append(res["seconds"])
    res["runs"] = cal["runs"]
    value = float(np.mean(values))
    res["value"] = value
    line = {"impl": "reference", "metric": "genome_pairs_per_sec", "value": value,
            "unit": "pairs/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * float(np.mean(secs)), "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": workload, "metric": args.metric},
            "cpu_baseline": cpu_baseline_record(res, args.metric),
            "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0}}
    if "alignments_per_pair" in res:
        line["config"]["alignments_per_pair"] = res["alignments_per_pair"]
    emit(line)
    return 0


# --------------------------------------------------------------------------------------
# the native arm
# --------------------------------------------------------------------------------------
class Env:
    pass


def host_buffers(torch, table, need_seqs):
    """The table with every array in pinned host memory (what a caller keeps between steps)."""
    arrays = [table.genome_ptr, table.pham_id, table.copies, table.len_sum]
    if need_seqs:
        arrays += [table.seq_off, table.residues]
    pinned = [pinned_copy(torch, a) for a in arrays]
    host_table = copy.copy(table)
    host_table.genome_ptr, host_table.pham_id, host_table.copies, host_table.len_sum = \
        [p.numpy() for p in pinned[:4]]
    if need_seqs:
        host_table.seq_off, host_table.residues = pinned[4].numpy(), pinned[5].numpy()
    return host_table, pinned, int(sum(a.nbytes for a in arrays))


def measure(env, workload, metric, genomes, phams, steps, warmup, engine=None, api=False):
    """One workload on this process group: resident ``value``, ``e2e``, roofline, clocks."""
    torch, dist = env.torch, env.dist
    from phamclust_b200 import _native, distributed
    world, rank, dev, ctx = env.world, env.rank, env.dev, env.ctx
    table, name = build_workload(workload, metric, genomes, phams)
    n = table.n_genomes
    pairs = n * (n - 1) // 2
    tri = _native.tri_len(n)
    need_seqs = metric in ALIGN_METRICS
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    ctx.set_engine(engine or "auto")
    flush = env.flush

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    builder = None
    if world == 1:
        ctx.load_table(table)
        if need_seqs:
            ctx.load_sequences(table)
        out_dev = torch.empty(tri, dtype=torch.int32, device=dev)
        r0, r1 = 0, n

        def step():
            ctx.matrix(metric, True, out_device_ptr=out_dev.data_ptr())
        local_view = lambda: out_dev
    else:
        builder = distributed.ShardedBuilder(table, metric)
        builder.load()
        r0, r1 = builder.r0, builder.r1
        fused = builder.enable_fused_gather()        # symmetric-memory triangle: gather fused into the kernels
        state = {}
        if fused:
            def step():
                state["tri"] = builder.gathered_triangle()
            local_view = lambda: state["tri"][_native.tri_len(n, 0, r0):]
        else:
            width = max(max(builder.sizes), 1)
            gathered = torch.empty(world * width, dtype=torch.int32, device=dev)

            def step():
                state["local"] = builder.row_block()
                builder.all_gather(state["local"], gathered)
            local_view = lambda: state["local"]

    for _ in range(max(warmup, 0)):
        flush.zero_()
        step()
    barrier()
    ctx.reset_counters()
    events = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
              for _ in range(steps)]
    with ClockSampler(env.local_rank) as clocks:
        barrier()
        for start, stop in events:
            flush.zero_()                     # L2 flush between timed steps (not timed)
            start.record()
            step()
            stop.record()
        barrier()
        step_ms = [s.elapsed_time(e) for s, e in events]
        total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
        total_ms = float(total_ms.item())
        counters = ctx.counters()
        stats = ctx.stats()
        timings = ctx.timings()
        # nvidia-smi polls every 100 ms (and takes a while to start): a timed region shorter than
        # ~0.7 s is followed by UNTIMED steps of the same load until the sampler has had time to
        # see it.  The count derives from the max-over-ranks time, so every rank runs the same
        # number of (collective) steps; counters and timings were read above.
        clocks_extended = total_ms < 700.0
        if clocks_extended:
            # rank 0 decides (it owns the sampler): repeat until >= 4 samples were seen or 3 s passed
            t_ext = time.perf_counter()
            batch = max(1, min(400, int(np.ceil(100.0 / max(total_ms / steps, 1e-3)))))   # ~0.1 s of steps
            while True:
                for _ in range(batch):
                    step()
                barrier()
                go = torch.tensor([1 if (len(clocks.rows) < 4 and time.perf_counter() - t_ext < 3.0) else 0],
                                  dtype=torch.int32, device=dev)
                if world > 1:
                    dist.broadcast(go, src=0)
                if int(go.item()) == 0:
                    break
        # per-rank time of the dominant stage (load balance of the sharding)
        key = "align" if need_seqs else "contract"
        rank_ms = torch.tensor([timings[key] / max(steps, 1)], dtype=torch.float64, device=dev)
        if world > 1:
            all_ms = [torch.empty_like(rank_ms) for _ in range(world)]
            dist.all_gather(all_ms, rank_ms)
            rank_ms = torch.cat(all_ms)
        rank_kernel_ms = [round(float(v), 3) for v in rank_ms.tolist()]
    local_sum = local_view()[:_native.tri_len(n, r0, r1)].to(torch.int64).sum().reshape(1)
    if world > 1:
        dist.all_reduce(local_sum, op=dist.ReduceOp.SUM)
    checksum = int(local_sum.item())          # sum of every entry of the packed triangle
    if world > 1 and builder.fused is not None:
        # every rank must hold the WHOLE triangle after the fused gather: its sum is the checksum
        whole = int(state["tri"].to(torch.int64).sum().item())
        assert whole == checksum, f"rank {rank}: gathered triangle sums to {whole}, row blocks to {checksum}"

    # ---- e2e: host buffers in, host buffer out, through the C ABI, every step ----
    host_table, pinned, h2d = host_buffers(torch, table, need_seqs)
    if world == 1:
        host_out = torch.empty(tri, dtype=torch.int32).pin_memory().numpy()     # pinned result buffer
        ctx.set_stream(0, external=False)
        e2e_secs = []
        for it in range(warmup + steps):
            if it == warmup:
                ctx.reset_counters()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            ctx.load_table(host_table)
            if need_seqs:
                ctx.load_sequences(host_table)
            ctx.matrix(metric, as_distance=True, out=host_out)
            ctx.sync()
            dt = time.perf_counter() - t0
            if it >= warmup:
                e2e_secs.append(dt)
        e2e = {"value": pairs / float(np.mean(e2e_secs)), "unit": "pairs/s",
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": int(host_out.nbytes),
               "ms_per_step": 1e3 * float(np.mean(e2e_secs)),
               "stage_ms_per_step": {k: v / max(steps, 1) for k, v in ctx.timings().items()},
               "api": "pc_load_table + pc_load_sequences + pc_matrix (host buffers)"}
        assert int(host_out.astype(np.int64).sum()) == checksum, "e2e and resident results differ"
        ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    else:
        # N > 1: every rank loads the table and ITS phams' genes from pinned host buffers, aligns,
        # joins the all-to-all of contribution records, and copies its row block from its own GPU
        # straight into ONE triangle in page-locked POSIX shared memory: the consumer is the host
        shared = distributed.SharedTriangle(n, pin_rows=(r0, r1))
        my_rows = shared.rows(r0, r1)
        secs = []
        h2d_rank = None
        for it in range(min(warmup, 2) + steps):
            if it == min(warmup, 2):
                ctx.reset_counters()
            barrier()
            t0 = time.perf_counter()
            builder.load(host_table)
            builder.row_block_to_host(my_rows)
            barrier()
            if it >= min(warmup, 2):
                secs.append(time.perf_counter() - t0)
        worst = torch.tensor([float(np.mean(secs))], dtype=torch.float64, device=dev)
        dist.all_reduce(worst, op=dist.ReduceOp.MAX)
        if rank == 0:
            assert int(shared.array.sum(dtype=np.int64)) == checksum, "e2e and resident results differ"
        # what the ranks actually handed to their H2D copies (counted inside the library)
        shipped = torch.tensor([ctx.stats()["h2d_bytes"]], dtype=torch.int64, device=dev)
        dist.all_reduce(shipped, op=dist.ReduceOp.SUM)
        e2e = {"value": pairs / float(worst.item()), "unit": "pairs/s",
               "h2d_bytes_per_step": int(shipped.item()) // max(steps, 1),
               "d2h_bytes_per_step": int(tri * 4), "ms_per_step": 1e3 * float(worst.item()),
               "stage_ms_per_step_rank0": {k: v / max(steps, 1) for k, v in ctx.timings().items()},
               "api": "per rank: pc_load_table + pc_load_sequences_part (its phams' genes: gathered on the host at 8+ ranks, "
                      "the whole page-locked arena shipped and filtered on the device below that) + "
                      "pc_align_contributions + NCCL all-to-all of the records + pc_accumulate_records + "
                      "pc_matrix_rows_owned into the node's shared page-locked host triangle "
                      "(distributed.SharedTriangle), every GPU over its own PCIe link"}
        barrier()
        shared.close()
        ctx.load_table(table)            # back to resident state for whoever measures next
    del pinned

    # ---- the reference's boundary: matrix_de_novo(list[Genome], func, cpus) -------
    e2e_api = None
    if api and world == 1:
        from phamclust_b200 import METRICS, matrix_de_novo
        from phamclust_b200.packing import unpack_genomes
        genomes_py = unpack_genomes(table)
        func = METRICS[metric]
        ctx.set_stream(0, external=False)
        api_secs = []
        for it in range(2 + steps):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            result = matrix_de_novo(genomes_py, func, 1)
            dt = time.perf_counter() - t0
            if it >= 2:
                api_secs.append(dt)
        assert int(result._micro.astype(np.int64).sum()) == checksum, "matrix_de_novo and resident results differ"
        e2e_api = {"value": pairs / float(np.mean(api_secs)), "unit": "pairs/s",
                   "ms_per_step": 1e3 * float(np.mean(api_secs)), "h2d_bytes_per_step": h2d,
                   "d2h_bytes_per_step": int(tri * 4),
                   "api": "matrix_de_novo(list[Genome], METRICS[metric], cpus=1) -> SymMatrix "
                          "(matrix.py:432): C packer over the Genome dicts + pc_load_* + pc_matrix"}
        ctx.set_stream(torch.cuda.current_stream().cuda_stream)
        del genomes_py, result

    # ---- roofline of the dominant kernel --------------------------------------------
    peaks, peak_kind = measured_peaks()
    clk = clocks.summary()
    if clocks_extended:
        clk["window"] = ("timed region + untimed repeats of the same step until the sampler had seen the load "
                         "(nvidia-smi polls every 100 ms and takes a while to start)")
    record_extra = {}
    if need_seqs:
        alu, instr, counts_state = dp_instruction_counts()
        cells = counters["dp_cells"]
        lex_cells, lexr_cells, lexc_cells = stats["lex_cells"], stats["lexr_cells"], stats.get("lexc_cells", 0)
        cls_cells = cells - lex_cells - lexr_cells - lexc_cells
        align_ms = timings["align"]
        launches = max(counters["align_launches"], 1)
        gcups = cells / (align_ms * 1e-3) / 1e9 if align_ms > 0 else 0.0
        sm_mhz = clk["sm_mhz"] or peaks.get("sm_max_mhz", 1965.0)
        # pipe-bound time of this mix of cells: each kernel's cells x its ALU-pipe (resp. all)
        # instructions per cell / the pipe's (resp. the issue ports') lanes per second
        alu_lanes = SM_COUNT * 64 * sm_mhz * 1e6
        issue_lanes = SM_COUNT * 128 * sm_mhz * 1e6
        alu_work = (lex_cells * alu["lex"] + lexr_cells * alu["lexr"] + lexc_cells * alu["lexc"]
                    + cls_cells * alu["classic"])
        issue_work = (lex_cells * instr["lex"] + lexr_cells * instr["lexr"] + lexc_cells * instr["lexc"]
                      + cls_cells * instr["classic"])
        peak = cells / (alu_work / alu_lanes) / 1e9 if alu_work else 0.0
        issue_peak = cells / (issue_work / issue_lanes) / 1e9 if issue_work else 0.0
        all_lex_peak = alu_lanes / alu["lex"] / 1e9
        traffic, traffic_src = ncu_traffic("dp")
        roofline = {"bound": "int-alu-pipe",
                    "kernel": "nw_lex_kernel<G,K,B,REBASE> + nw_lexc_kernel<K> + nw_kernel<G,K> (global-alignment DP, "
                              "one launch per config)",
                    "achieved": gcups, "peak": peak, "unit": "GCUPS", "frac": gcups / peak if peak else None,
                    "frac_of_all_lex_bound": gcups / all_lex_peak if all_lex_peak else None,
                    "all_lex_peak": all_lex_peak,
                    "traffic": traffic, "traffic_source": traffic_src,
                    "cells_per_step": cells / max(steps, 1),
                    "cell_fractions": {"lex": lex_cells / cells if cells else None,
                                       "lexr": lexr_cells / cells if cells else None,
                                       "lexc": lexc_cells / cells if cells else None,
                                       "classic": cls_cells / cells if cells else None},
                    "lexc_fallbacks_per_step": stats.get("lexc_fallbacks", 0) / max(steps, 1),
                    "alu_instr_per_cell": alu, "instr_per_cell": instr, "instr_counts": counts_state,
                    "kernel_ms_per_step": align_ms / max(steps, 1),
                    "launches_per_step": launches / max(steps, 1),
                    "share_of_step": align_ms / total_ms if total_ms else None,
                    "issue_slot_peak": issue_peak,
                    "peak_basis": (f"ALU pipe: {SM_COUNT} SMs x 64 lanes/clk x {sm_mhz:.0f} MHz (median SM clock sampled "
                                   "under load) / ALU-pipe SASS instructions per cell of each kernel family (counted "
                                   "from the shipped library's SASS), weighted by the cells each family ran; neither "
                                   "HBM nor tensor bound (SURVEY 8(d)); frac_of_all_lex_bound: the same cells if all "
                                   "ran at the plain LEX kernel's count ('of derived pipe peak'); achieved counts only "
                                   "real cells (padding, pipeline fill excluded)")}
        aln_ref, cells_ref = alignment_work(table)
        record_extra = {"alignments_per_step": counters["alignments"] // max(steps, 1),
                        "alignments_per_pair": aln_ref / max(pairs, 1),
                        "distinct_alignments_per_pair": counters["alignments"] / max(steps, 1) / max(pairs, 1),
                        "reference_alignments_per_step": aln_ref, "reference_cells_per_step": cells_ref}
    else:
        contract_ms = timings["contract"]
        launches = max(counters["contract_launches"], 1)
        # one step may take several launches (row sub-blocks of the push gather): rates are per STEP
        per_launch_ms = contract_ms / max(steps, 1)
        af_limbs = 1 if int(table.len_sum.max(initial=0)) <= 255 else 2
        limbs = {"gcs": 1, "jc": 1, "pocp": 2, "af": 2 * af_limbs}[metric]     # K multiplier of the operand
        if (engine or "auto") in ("i8mma", "i8mma_v1", "auto"):
            # algorithmic int8 ops: 2 * K * N(N+1)/2 with K = P (gcs/jc), 2P (pocp), 2P per af limb
            # K counts only phams held by >= 2 genomes: single-genome phams can never be
            # shared and get no operand column (exact)
            k_eff = int((np.bincount(table.pham_id, minlength=table.n_phams) >= 2).sum())
            ops = 2.0 * k_eff * limbs * _native.tri_len(n, r0, r1)   # this rank's row block
            # block-sparse K schedule: only the (tile, K block) pairs both operands occupy are
            # multiplied; the algorithmic ops of the launch shrink by the same fraction
            visited = (stats["i8_tile_steps"] / stats["i8_tile_steps_dense"]) if stats["i8_tile_steps_dense"] else 1.0
            dense_tops = ops / (per_launch_ms * 1e-3) / 1e12
            tops = dense_tops * visited
            # int8 peak: MEASURED_PEAKS.json has no int8 figure.  2 x the measured bf16 number (3349)
            # is provably too low -- this engine sustains 4.1 POP/s executed on C5 -- so the NOMINAL
            # dense int8 rate is the denominator (4500 TOP/s = 2 x nominal bf16) and the measured-bf16
            # derivative is reported next to it
            peak = 4500.0
            derived = 2.0 * peaks["bf16_tflops"]
            traffic, traffic_src = ncu_traffic(f"i8_{workload}_{metric}") if world == 1 else (None, None)
            algo_bytes = n * k_eff * limbs + 4.0 * _native.tri_len(n, r0, r1)
            roofline = {"bound": "tensor", "kernel": "i8 contraction (tcgen05 cta_group::2 kind::i8, TMA, TMEM) + fused epilogue",
                        "achieved": tops, "peak": peak, "unit": "TOP/s", "frac": tops / peak,
                        "frac_of_2x_measured_bf16": tops / derived, "two_x_measured_bf16": derived,
                        "k_blocks_visited_frac": visited, "dense_equivalent_tops": dense_tops,
                        "traffic": traffic, "traffic_source": traffic_src, "algorithmic_bytes_per_step": algo_bytes,
                        "kernel_ms_per_step": contract_ms / max(steps, 1),
                        "share_of_step": contract_ms / total_ms if total_ms else None,
                        "algorithmic_ops_per_step": ops, "launches_per_step": launches / max(steps, 1),
                        "k_columns": k_eff * limbs,
                        "peak_basis": ("NOMINAL dense int8 tensor rate of the B200 (4500 TOP/s); the driver's "
                                       f"MEASURED_PEAKS.json ({peak_kind}) holds no int8 figure and 2 x its bf16 number "
                                       f"({derived:.0f}) is below what this kernel sustains on C5 (4.1 POP/s executed); "
                                       "ncu's sm__pipe_tensor_cycles_active is the independent check (profiles/)")}
        else:
            words = (table.n_phams + 1023) // 1024 * 32
            algo_bytes = n * words * 4 + tri * 4            # operand read once + result written
            gbs = algo_bytes / (per_launch_ms * 1e-3) / 1e9
            roofline = {"bound": "hbm", "kernel": "popc_contract_kernel (+fused epilogue)",
                        "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                        "frac": gbs / peaks["hbm_gbs"], "traffic": None,
                        "kernel_ms_per_step": contract_ms / max(steps, 1),
                        "share_of_step": contract_ms / total_ms if total_ms else None,
                        "note": "POPC-issue bound in practice (16 lanes/clk/SM), far from the HBM roof",
                        "peak_basis": f"MEASURED_PEAKS.json hbm_gbs ({peak_kind})"}

    gather_txt = ""
    if world > 1:
        gather_txt = ((f"no collective: symmetric-memory triangles, {builder.push_blocks} row sub-blocks pushed to the peers "
                       "by copy engines while the next one is computed, one device-side barrier"
                       if builder.gather_mode == "push" else
                       "gather FUSED into the producing kernels (every row segment stored into all ranks' symmetric-"
                       "memory triangles over NVLink, one device-side barrier)") if builder.fused is not None else
                      f"one NCCL all-gather of the row blocks (no symmetric memory: {builder.fused_error})")
    if world == 1:
        parallelism = "1 GPU"
    elif need_seqs:
        parallelism = (f"{world} ranks: phams dealt by cost, each rank ships only its genes; sparse contribution "
                       f"records exchanged by one NCCL all-to-all ({builder.exchange_records} records received by "
                       "rank 0 per step); pair-balanced row-block epilogue; " + gather_txt)
    else:
        parallelism = (f"{world} ranks: row blocks cut by estimated contraction work "
                       f"(rank 0 rows [{r0},{r1}) of {n}), no collective during compute; " + gather_txt)
    value = pairs * steps / (total_ms * 1e-3)
    record = {"metric": "genome_pairs_per_sec", "value": value, "unit": "pairs/s",
              "n_gpus": world, "steps": steps, "warmup": warmup,
              "ms_per_step": total_ms / steps, "higher_is_better": True, "scaling": "strong",
              "vs_baseline": None, "dtype": "int32", "data": "synthetic",
              "config": dict({"workload": name, "metric": metric, "n_genomes": n,
                              "n_phams": table.n_phams, "pairs": pairs,
                              "l2": "256 MiB buffer written between timed steps (L2 flush)",
                              "engine": engine or "auto", "parallelism": parallelism}, **record_extra),
              "clocks": clk, "e2e": e2e, "gpu_launches": counters["launches"],
              "roofline": roofline, "stage_ms_per_step": {k: v / steps for k, v in timings.items()},
              "rank_kernel_ms_per_step": rank_kernel_ms, "checksum": checksum}
    if e2e_api:
        record["e2e_api"] = e2e_api
    return record, table


def run_native(args):
    import torch
    import torch.distributed as dist
    from phamclust_b200 import build_native
    from phamclust_b200.matrix import get_context

    env = Env()
    env.torch, env.dist = torch, dist
    env.world = int(os.environ.get("WORLD_SIZE", "1"))
    env.rank = int(os.environ.get("RANK", "0"))
    env.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if env.world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={env.world}: launch with torchrun")
    torch.cuda.set_device(env.local_rank)
    env.dev = torch.device("cuda", env.local_rank)
    if env.world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=env.dev)
    env.ctx = get_context(env.local_rank)
    env.flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=env.dev)   # > 126 MB L2

    line, table = measure(env, args.workload, args.metric, args.genomes, args.phams, args.steps, args.warmup,
                          engine=args.engine, api=True)
    line["library_sha256"] = build_native.library_sha256()
    extras = []
    if args.extras == "auto" and args.workload == "c2" and args.genomes == 2000:
        if env.world == 1:
            plan = [("c3", "gcs"), ("c3", "jc")]
        else:
            plan = [("c5", "gcs")] + ([("c4", "peq")] if env.world >= 8 else [])
        for wl, metric in plan:
            torch.cuda.empty_cache()
            _, genomes, phams = DEFAULTS[wl]
            rec, _ = measure(env, wl, metric, genomes, phams, args.steps, args.warmup)
            extras.append(rec)
    if extras:
        line["workloads"] = extras
    if env.world == 1 and env.rank == 0 and not args.no_cpu_baseline:
        n_cpus = os.cpu_count() or 1
        sample = args.cpu_sample or (400 if args.metric in ALIGN_METRICS else 1000)
        res = cpu_reference_rate(table, args.metric, sample, sorted({1, n_cpus}))
        line["cpu_baseline"] = cpu_baseline_record(res, args.metric)
    if env.rank == 0:
        emit(line)
    if env.world > 1:
        dist.destroy_process_group()
    return 0


def emit(line):
    """The ONE JSON line goes to the real stdout; everything else any library prints
    (e.g. NCCL's version banner) was re-routed to stderr by main()."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = 1


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default="c2", choices=["c2", "c3", "c4", "c5"])
    ap.add_argument("--metric", default=None)
    ap.add_argument("--genomes", type=int, default=None)
    ap.add_argument("--phams", type=int, default=None)
    ap.add_argument("--engine", default=None, choices=[None, "auto", "popc", "i8mma", "i8mma_v1"])
    ap.add_argument("--cpu-sample", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--extras", default="auto", choices=["auto", "none"],
                    help="auto: also measure the other BASELINE.json configs that fit this N (see docstring)")
    args = ap.parse_args()
    d_metric, d_genomes, d_phams = DEFAULTS[args.workload]
    if args.metric is None:
        args.metric = d_metric
    if args.genomes is None:
        args.genomes = d_genomes
    if args.phams is None:
        args.phams = d_phams or 60000
    if args.impl == "reference":
        return run_reference(args)
    return run_native(args)


if __name__ == "__main__":
    sys.exit(main())
