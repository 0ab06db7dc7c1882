#!/usr/bin/env python
"""tools/probe_lock_threshold.py -- when does the pair kernel's lock-step (producers of all CTA
pairs wait for each other at the start of every round, so that the tiles of a round stream the
same K window through L2) pay?  Times pc_matrix('gcs') resident for several operand shapes with
PC_I8_LOCK = 1 (always) and 2 (never); one line per case.

    python tools/probe_lock_threshold.py
"""
import os
import subprocess
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)

CASES = [(10000, 60000), (10000, 100000), (10000, 140000), (16000, 100000), (20000, 200000)]


def child(n, p):
    import torch
    from phamclust_b200 import synthetic
    from phamclust_b200._native import Context, tri_len
    table = synthetic.sparse_table(n, p, seed=1)
    ctx = Context(0)
    dev = torch.device("cuda", 0)
    ctx.set_stream(torch.cuda.current_stream(dev).cuda_stream)
    ctx.load_table(table)
    out = torch.empty(tri_len(n), dtype=torch.int32, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for _ in range(3):
        ctx.matrix("gcs", True, out_device_ptr=out.data_ptr())
    times = []
    for _ in range(6):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        ctx.matrix("gcs", True, out_device_ptr=out.data_ptr())
        b.record()
        torch.cuda.synchronize()
        times.append(a.elapsed_time(b))
    times.sort()
    k = ctx.stats()["operand_columns"]
    print(f"n={n} phams={p} K={k} strip={256 * k / 1e6:.1f}MB lock={os.environ.get('PC_I8_LOCK')} "
          f"median {times[len(times) // 2]:.3f} ms min {times[0]:.3f} checksum {int(out.sum())}", flush=True)


if __name__ == "__main__":
    if len(sys.argv) == 3:
        child(int(sys.argv[1]), int(sys.argv[2]))
    else:
        for n, p in CASES:
            for lock in ("1", "2"):
                env = dict(os.environ, PC_I8_LOCK=lock)
                subprocess.run([sys.executable, os.path.abspath(__file__), str(n), str(p)], env=env, check=False)
