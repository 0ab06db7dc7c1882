#!/usr/bin/env python
"""bench.py -- genome pairs/second of the all-vs-all matrix build (BASELINE.json).

    python bench.py --gpus 1 --steps 5 --warmup 3                 # this repo, 1 B200
    torchrun --nproc-per-node N ... bench.py --gpus N ...         # N ranks, NCCL
    python bench.py --impl reference ...                          # host-CPU arm

Workload (``config.workload``): BASELINE.json configs[1] -- metric=peq on the
bundled ``benchmark_data.tsv``.  That file is absent from the reference mount
(.MISSING_LARGE_BLOBS), so the deterministic stand-in of SURVEY.md 8(d) is used:
``synthetic.phage_table(2000, seed=42)``.  A *step* is one pass of the hot path
(matrix_de_novo for that metric over the whole table).

* ``value``  : pairs/s with the packed table + residues already resident in HBM
               (CUDA events on the stream the kernels run on, L2 flushed between
               steps, max over ranks).
* ``e2e``    : the same metric through the C ABI with (pinned) HOST buffers every step:
               pc_load_table + pc_load_sequences + pc_matrix, H2D and D2H inside
               the timed region.  N > 1: every rank loads from host buffers, aligns
               its phams, joins the NCCL all-reduce and copies its row block from
               its own GPU into one page-locked shared-memory triangle
               (distributed.SharedTriangle) -- the consumer is the host.
* other workloads: ``--workload c3`` (10k x 60k, jc/pocp/af/gcs), ``c4`` (50k genomes,
               peq/aai, the 8-GPU target of BASELINE.json), ``c5`` (100k x 200k, gcs).
* ``roofline``: the dominant kernel (the alignment DP for peq/aai, the
               shared-count contraction for gcs/jc/pocp/af).
* ``cpu_baseline``: the oracle port of the reference's CPU path on a bounded
               sample, timed on this box's host cores (rank 0, N=1 only).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ALIGN_METRICS = ("aai", "peq")
# SASS instructions per DP cell in the row loop (cuobjdump -sass), counted
# over the unrolled K-cell row step of representative configs:
#   classic nw_kernel<32,20> / <16,16>: 18.5 per cell, 10.7 of them on the ALU pipe
#   LEX     nw_lex_kernel<16,16> / <16,24>: 11.0 per cell, 6.6 on the ALU pipe
# ncu (profiles/r1b_ncu_dp_launches.txt) shows the ALU pipe (16 lanes/clk/SMSP) 88-95 % busy in both
# kernels and the FMA pipe 14-27 %, so the ALU pipe is the binding unit of each
DP_INSTR_PER_CELL = {"classic": 18.5, "lex": 11.0}
DP_ALU_INSTR_PER_CELL = {"classic": 10.7, "lex": 6.6}
SM_COUNT = 148


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return json.load(fh), "measured"
    # fallback stated in /opt/skills/guides/B200_PROFILING.md
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0,
            "sm_max_mhz": 1965.0}, "fallback"


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


def build_workload(args):
    from phamclust_b200 import synthetic
    if args.workload == "c2":
        table = synthetic.phage_table(args.genomes, seed=42)
        name = (f"configs[1]: metric={args.metric} on stand-in for the missing benchmark_data.tsv = "
                f"synthetic.phage_table(n_genomes={args.genomes}, seed=42) [SURVEY 8(d) C1/C2]")
    elif args.workload == "c3":
        table = synthetic.sparse_table(args.genomes, args.phams, seed=1)
        if args.metric in ALIGN_METRICS:
            synthetic.attach_m_translations(table)
        name = (f"configs[2]: metric={args.metric} on synthetic.sparse_table("
                f"n_genomes={args.genomes}, n_phams={args.phams}, seed=1)")
    elif args.workload == "c4":
        table = synthetic.phage_table(args.genomes, seed=2)
        name = (f"configs[3]: metric={args.metric}, pham-sharded alignment, on synthetic.phage_table("
                f"n_genomes={args.genomes}, seed=2) [SURVEY 8(d) C4: clusters of 50 genomes x 150-pham pools]")
    elif args.workload == "c5":
        table = synthetic.sparse_table(args.genomes, args.phams, seed=3)
        name = (f"configs[4]: metric={args.metric}, full upper triangle in row blocks + NCCL gather, on "
                f"synthetic.sparse_table(n_genomes={args.genomes}, n_phams={args.phams}, seed=3) [SURVEY 8(d) C5]")
    else:
        raise SystemExit(f"unknown workload {args.workload}")
    return table, name


def pinned_copy(torch, arr):
    t = torch.from_numpy(np.ascontiguousarray(arr))
    try:
        return t.pin_memory()
    except RuntimeError:
        return t


# --------------------------------------------------------------------------------------
# the reference arm / cpu_baseline: oracle port of the reference's CPU path
# --------------------------------------------------------------------------------------
def cpu_reference_rate(table, metric, sample_genomes, cpus_list):
    """pairs/s of the oracle port of matrix_de_novo (one joblib task per pair,
    exactly like matrix.py:471-491) on the first ``sample_genomes`` genomes.
    The worker count is picked on a small calibration sample (the reference's
    per-pair joblib dispatch makes more workers SLOWER for the cheap metrics,
    BASELINE.md section 2), then the sample is timed once with it."""
    from oracle import nw_c, pham_oracle as po
    from phamclust_b200.packing import unpack_genomes
    nw_c.build()
    genomes = unpack_genomes(table.subset(sample_genomes))
    n = len(genomes)
    pairs = n * (n - 1) // 2
    runs = {}
    if len(cpus_list) > 1:
        calib = genomes[:min(n, 24 if metric in ALIGN_METRICS else 300)]
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
    return {"value": pairs / dt, "cores": cpus, "seconds": dt, "pairs": pairs,
            "runs": {str(k): v for k, v in runs.items()}, "n": n}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    table, workload = build_workload(args)
    n_cpus = os.cpu_count() or 1
    sample = args.cpu_sample or (80 if args.metric in ALIGN_METRICS else 1000)
    # warm-up: page the oracle in and pick the worker count once
    cal = cpu_reference_rate(table, args.metric, min(sample, 40), sorted({1, n_cpus}))
    cpus_list = [cal["cores"]]
    for _ in range(max(args.warmup - 1, 0)):
        cpu_reference_rate(table, args.metric, min(sample, 40), cpus_list)
    values, secs = [], []
    res = None
    for _ in range(args.steps):
        res = cpu_reference_rate(table, args.metric, sample, cpus_list)
        values.append(res["value"]); secs.append(res["seconds"])
    res["runs"] = cal["runs"]
    value = float(np.mean(values))
    sample_txt = (f"first {res['n']} genomes of the workload ({res['pairs']} pairs) per step; oracle port "
                  f"of matrix_de_novo (one joblib task per pair) + C restatement of parasail NW "
                  f"(not parasail's SIMD kernel); cpus={res['cores']} chosen on a calibration sample, "
                  f"pairs/s by cpus there: { {k: round(v, 1) for k, v in res['runs'].items()} }")
    line = {"impl": "reference", "metric": "genome_pairs_per_sec", "value": value,
            "unit": "pairs/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * float(np.mean(secs)), "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": workload, "metric": args.metric},
            "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": res["cores"],
                             "kind": "port", "sample": sample_txt},
            "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0}}
    emit(line)
    return 0


# --------------------------------------------------------------------------------------
# the native arm
# --------------------------------------------------------------------------------------
def run_native(args):
    import torch
    import torch.distributed as dist
    from phamclust_b200 import _native, distributed
    from phamclust_b200.matrix import get_context

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    table, workload = build_workload(args)
    metric = args.metric
    n = table.n_genomes
    pairs = n * (n - 1) // 2
    tri = _native.tri_len(n)
    need_seqs = metric in ALIGN_METRICS

    ctx = get_context(local_rank)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    if args.engine:
        ctx.set_engine(args.engine)
    ctx.load_table(table)
    if need_seqs:
        ctx.load_sequences(table)

    bounds = distributed.row_partition(n, world)
    r0, r1 = bounds[rank], bounds[rank + 1]
    width = max(_native.tri_len(n, bounds[r], bounds[r + 1]) for r in range(world))
    out_local = torch.empty(max(width, 1), dtype=torch.int32, device=dev)
    gathered = torch.empty(world * width, dtype=torch.int32, device=dev) if world > 1 else None
    acc = torch.zeros(tri, dtype=torch.int64, device=dev) if need_seqs else None
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def step():
        """One pass of the hot path with inputs resident in HBM."""
        if need_seqs:
            acc.zero_()
            ctx.align_accumulate(acc.data_ptr(), part=rank, n_parts=world)
            if world > 1:
                dist.all_reduce(acc, op=dist.ReduceOp.SUM)
        ctx.matrix_rows(metric, True, r0, r1, acc_ptr=acc.data_ptr() if need_seqs else None,
                        out_device_ptr=out_local.data_ptr())
        if world > 1:
            dist.all_gather_into_tensor(gathered, out_local[:width])

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(max(args.warmup, 0)):
        flush.zero_()
        step()
    barrier()
    ctx.reset_counters()
    events = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
              for _ in range(args.steps)]
    with ClockSampler(local_rank) as clocks:
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
        lex_cells = stats["lex_cells"] if need_seqs else 0
        timings = ctx.timings()
        # nvidia-smi polls every 100 ms (and takes a while to start): a timed region shorter than
        # ~0.7 s is followed by UNTIMED steps of the same load until the sampler has had time to
        # see it.  The count derives from the max-over-ranks time, so every rank runs the same
        # number of (collective) steps; counters and timings were read above.
        clocks_extended = total_ms < 700.0
        if clocks_extended:
            for _ in range(min(200, int(np.ceil((700.0 - total_ms) / max(total_ms / args.steps, 1e-3))))):
                step()
            barrier()
    local_sum = out_local[:_native.tri_len(n, r0, r1)].to(torch.int64).sum().reshape(1)
    if world > 1:
        dist.all_reduce(local_sum, op=dist.ReduceOp.SUM)
    checksum = int(local_sum.item())          # sum of every entry of the packed triangle

    # ---- e2e: host buffers in, host buffer out, through the C ABI, every step ----
    e2e = None
    if world == 1:
        host_out = torch.empty(tri, dtype=torch.int32).pin_memory().numpy()     # pinned result buffer
        arrays = [table.genome_ptr, table.pham_id, table.copies, table.len_sum]
        if need_seqs:
            arrays += [table.seq_off, table.residues]
        pinned = [pinned_copy(torch, a) for a in arrays]
        import copy
        host_table = copy.copy(table)
        host_table.genome_ptr, host_table.pham_id, host_table.copies, host_table.len_sum = \
            [p.numpy() for p in pinned[:4]]
        if need_seqs:
            host_table.seq_off, host_table.residues = pinned[4].numpy(), pinned[5].numpy()
        h2d = int(sum(a.nbytes for a in arrays))
        ctx.set_stream(0, external=False)
        e2e_secs = []
        for it in range(args.warmup + args.steps):
            if it == args.warmup:
                ctx.reset_counters()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            ctx.load_table(host_table)
            if need_seqs:
                ctx.load_sequences(host_table)
            ctx.matrix(metric, as_distance=True, out=host_out)
            ctx.sync()
            dt = time.perf_counter() - t0
            if it >= args.warmup:
                e2e_secs.append(dt)
        e2e = {"value": pairs / float(np.mean(e2e_secs)), "unit": "pairs/s",
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": int(host_out.nbytes),
               "ms_per_step": 1e3 * float(np.mean(e2e_secs)),
               "stage_ms_per_step": {k: v / max(args.steps, 1) for k, v in ctx.timings().items()},
               "api": "pc_load_table + pc_load_sequences + pc_matrix (host buffers)"}
        assert int(host_out.astype(np.int64).sum()) == checksum, "e2e and resident results differ"
    else:
        # N > 1: the same C-ABI sequence on every rank from pinned HOST buffers (load table /
        # sequences, this rank's share of the work, NCCL all-reduce of the pair accumulators),
        # then every rank copies its row block from its own GPU straight into ONE triangle in
        # page-locked POSIX shared memory (distributed.SharedTriangle): the consumer is the host,
        # so no device-side all-gather and no single-link D2H of the whole triangle
        import copy
        arrays = [table.genome_ptr, table.pham_id, table.copies, table.len_sum] + \
            ([table.seq_off, table.residues] if need_seqs else [])
        pinned = [pinned_copy(torch, a) for a in arrays]
        host_table = copy.copy(table)
        host_table.genome_ptr, host_table.pham_id, host_table.copies, host_table.len_sum = \
            [p.numpy() for p in pinned[:4]]
        if need_seqs:
            host_table.seq_off, host_table.residues = pinned[4].numpy(), pinned[5].numpy()
        shared = distributed.SharedTriangle(n)
        my_rows = shared.rows(r0, r1)
        secs = []
        for it in range(min(args.warmup, 2) + args.steps):
            if it == min(args.warmup, 2):
                ctx.reset_counters()
            barrier()
            t0 = time.perf_counter()
            ctx.load_table(host_table)
            if need_seqs:
                ctx.load_sequences(host_table)
                acc.zero_()
                ctx.align_accumulate(acc.data_ptr(), part=rank, n_parts=world)
                dist.all_reduce(acc, op=dist.ReduceOp.SUM)
            if r1 > r0:
                ctx.matrix_rows(metric, True, r0, r1, acc_ptr=acc.data_ptr() if need_seqs else None, out=my_rows)
            barrier()
            if it >= min(args.warmup, 2):
                secs.append(time.perf_counter() - t0)
        worst = torch.tensor([float(np.mean(secs))], dtype=torch.float64, device=dev)
        dist.all_reduce(worst, op=dist.ReduceOp.MAX)
        if rank == 0:
            assert int(shared.array.sum(dtype=np.int64)) == checksum, "e2e and resident results differ"
        e2e = {"value": pairs / float(worst.item()), "unit": "pairs/s",
               "h2d_bytes_per_step": int(sum(a.nbytes for a in arrays)) * world,
               "d2h_bytes_per_step": int(tri * 4), "ms_per_step": 1e3 * float(worst.item()),
               "stage_ms_per_step_rank0": {k: v / max(args.steps, 1) for k, v in ctx.timings().items()},
               "api": "per rank: pc_load_table + pc_load_sequences + pc_align_accumulate (pinned host buffers) + "
                      "NCCL all-reduce + pc_matrix_rows into the node's shared page-locked host triangle "
                      "(distributed.SharedTriangle), every GPU over its own PCIe link"}
        barrier()
        shared.close()

    # ---- roofline of the dominant kernel --------------------------------------------
    peaks, peak_kind = measured_peaks()
    clk = clocks.summary()
    if clocks_extended:
        clk["window"] = "timed region + untimed repeats of the same step until 0.7 s had passed (nvidia-smi polls every 100 ms)"
    if need_seqs:
        cells = counters["dp_cells"]
        cls_cells = cells - lex_cells
        align_ms = timings["align"]
        launches = max(counters["align_launches"], 1)
        gcups = cells / (align_ms * 1e-3) / 1e9 if align_ms > 0 else 0.0
        sm_mhz = clk["sm_mhz"] or peaks.get("sm_max_mhz", 1965.0)
        # pipe-bound time of this mix of cells: each kernel's cells x its ALU-pipe (resp. all)
        # instructions per cell / the pipe's (resp. the issue ports') lanes per second
        alu_lanes = SM_COUNT * 64 * sm_mhz * 1e6
        issue_lanes = SM_COUNT * 128 * sm_mhz * 1e6
        alu_work = lex_cells * DP_ALU_INSTR_PER_CELL["lex"] + cls_cells * DP_ALU_INSTR_PER_CELL["classic"]
        issue_work = lex_cells * DP_INSTR_PER_CELL["lex"] + cls_cells * DP_INSTR_PER_CELL["classic"]
        peak = cells / (alu_work / alu_lanes) / 1e9 if alu_work else 0.0
        issue_peak = cells / (issue_work / issue_lanes) / 1e9 if issue_work else 0.0
        roofline = {"bound": "int-alu-pipe",
                    "kernel": "nw_lex_kernel<G,K> + nw_kernel<G,K> (global-alignment DP, one launch per config)",
                    "achieved": gcups, "peak": peak, "unit": "GCUPS", "frac": gcups / peak if peak else None,
                    "traffic": 4.0e6, "cells_per_step": cells / max(args.steps, 1),
                    "lex_cell_fraction": lex_cells / cells if cells else None,
                    "kernel_ms_per_step": align_ms / max(args.steps, 1),
                    "launches_per_step": launches / max(args.steps, 1),
                    "share_of_step": align_ms / total_ms if total_ms else None,
                    "issue_slot_peak": issue_peak,
                    "peak_basis": (f"ALU pipe: {SM_COUNT} SMs x 64 lanes/clk x {sm_mhz:.0f} MHz (median SM clock sampled "
                                   f"under load) / ALU-pipe SASS instr per cell, weighted by the cells each kernel ran: "
                                   f"LEX {DP_ALU_INSTR_PER_CELL['lex']} (of {DP_INSTR_PER_CELL['lex']} total), classic "
                                   f"{DP_ALU_INSTR_PER_CELL['classic']} (of {DP_INSTR_PER_CELL['classic']}); neither HBM nor "
                                   "tensor bound (SURVEY 8(d)); ncu: sm__pipe_alu_cycles_active 88-95 %, DRAM traffic "
                                   "~4 MB per launch ('of derived pipe peak'); achieved counts only real cells "
                                   "(padding, pipeline fill excluded)")}
    else:
        contract_ms = timings["contract"]
        launches = max(counters["contract_launches"], 1)
        per_launch_ms = contract_ms / launches
        engine = args.engine or "auto"
        af_limbs = 1 if int(table.len_sum.max(initial=0)) <= 255 else 2
        limbs = {"gcs": 1, "jc": 1, "pocp": 2, "af": 2 * af_limbs}[metric]     # K multiplier of the operand
        if engine in ("i8mma", "auto"):
            # algorithmic int8 ops: 2 * K * N(N+1)/2 with K = P (gcs/jc), 2P (pocp), 2P per af limb
            # K counts only phams held by >= 2 genomes: single-genome phams can never be
            # shared and get no operand column (exact)
            k_eff = int((np.bincount(table.pham_id, minlength=table.n_phams) >= 2).sum())
            ops = 2.0 * k_eff * limbs * _native.tri_len(n, r0, r1)   # this rank's row block
            # block-sparse K schedule: only the (tile, K block) pairs both operands occupy are
            # multiplied; the algorithmic ops of the launch shrink by the same fraction (1.0 on
            # dense tables such as C3)
            visited = (stats["i8_tile_steps"] / stats["i8_tile_steps_dense"]) if stats["i8_tile_steps_dense"] else 1.0
            dense_tops = ops / (per_launch_ms * 1e-3) / 1e12
            tops = dense_tops * visited
            peak = 2.0 * peaks["bf16_tflops"]
            roofline = {"bound": "tensor", "kernel": "i8_contract_kernel (tcgen05 kind::i8 + fused epilogue)",
                        "achieved": tops, "peak": peak, "unit": "TOP/s", "frac": tops / peak,
                        "k_blocks_visited_frac": visited, "dense_equivalent_tops": dense_tops,
                        # dram__bytes_read+write of one launch, ncu --set full (profiles/): only
                        # captured for the C3 jc launch on one GPU
                        "traffic": 3.58e9 if (args.workload == "c3" and metric == "jc" and world == 1
                                              and n == 10000) else None, "kernel_ms_per_step": contract_ms / max(args.steps, 1),
                        "share_of_step": contract_ms / total_ms if total_ms else None,
                        "algorithmic_ops_per_launch": ops, "k_columns": k_eff * limbs,
                        "peak_basis": (f"2 x MEASURED_PEAKS.json bf16_tflops ({peak_kind}; int8 dense is not "
                                       "measured by the driver: derived, nominal 4500 TOP/s)")}
        else:
            words = (table.n_phams + 1023) // 1024 * 32
            algo_bytes = n * words * 4 + tri * 4            # operand read once + result written
            gbs = algo_bytes / (per_launch_ms * 1e-3) / 1e9
            roofline = {"bound": "hbm", "kernel": "popc_contract_kernel (+fused epilogue)",
                        "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                        "frac": gbs / peaks["hbm_gbs"], "traffic": None,
                        "kernel_ms_per_step": contract_ms / max(args.steps, 1),
                        "share_of_step": contract_ms / total_ms if total_ms else None,
                        "note": "POPC-issue bound in practice (16 lanes/clk/SM), far from the HBM roof",
                        "peak_basis": f"MEASURED_PEAKS.json hbm_gbs ({peak_kind})"}

    value = pairs * args.steps / (total_ms * 1e-3)
    line = {"metric": "genome_pairs_per_sec", "value": value, "unit": "pairs/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": workload, "metric": metric, "n_genomes": n,
                       "n_phams": table.n_phams, "pairs": pairs,
                       "alignments_per_step": counters["alignments"] // max(args.steps, 1),
                       "l2": "256 MiB buffer written between timed steps (L2 flush)",
                       "engine": args.engine or "auto",
                       "parallelism": (f"{world} rank(s): pham-dealt alignments + int64 all-reduce, "
                                       "row-block epilogue + all-gather") if world > 1 else "1 GPU"},
            "clocks": clk, "e2e": e2e, "gpu_launches": counters["launches"],
            "roofline": roofline, "stage_ms_per_step": {k: v / args.steps for k, v in timings.items()},
            "checksum": checksum}
    if world == 1 and rank == 0 and not args.no_cpu_baseline:
        n_cpus = os.cpu_count() or 1
        sample = args.cpu_sample or (80 if need_seqs else 1000)
        res = cpu_reference_rate(table, metric, sample, sorted({1, n_cpus}))
        line["cpu_baseline"] = {
            "value": res["value"], "unit": "pairs/s", "cores": res["cores"], "kind": "port",
            "sample": (f"first {res['n']} genomes ({res['pairs']} pairs, {res['seconds']:.1f} s); oracle port of "
                       f"matrix_de_novo (joblib task per pair) + C restatement of parasail NW; cpus chosen on a "
                       f"calibration sample, pairs/s by cpus there: { {k: round(v, 1) for k, v in res['runs'].items()} }")}
    if rank == 0:
        emit(line)
    if world > 1:
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
    ap.add_argument("--phams", type=int, default=60000)
    ap.add_argument("--engine", default=None, choices=[None, "auto", "popc", "i8mma"])
    ap.add_argument("--cpu-sample", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.metric is None:
        args.metric = {"c2": "peq", "c3": "jc", "c4": "peq", "c5": "gcs"}[args.workload]
    if args.genomes is None:
        args.genomes = {"c2": 2000, "c3": 10000, "c4": 50000, "c5": 100000}[args.workload]
    if args.workload == "c5" and args.phams == 60000:
        args.phams = 200000
    if args.impl == "reference":
        return run_reference(args)
    return run_native(args)


if __name__ == "__main__":
    sys.exit(main())
