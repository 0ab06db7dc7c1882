"""Break the end-to-end C-ABI call (host buffers in, host buffer out) into its
host and device stages -- used to decide where e2e time goes.
usage: e2e_probe.py [c2|c3] [metric] [genomes]"""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from phamclust_b200 import _native, synthetic  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "c2"
metric = sys.argv[2] if len(sys.argv) > 2 else ("peq" if workload == "c2" else "jc")
n = int(sys.argv[3]) if len(sys.argv) > 3 else (2000 if workload == "c2" else 10000)
table = synthetic.phage_table(n, seed=42) if workload == "c2" else synthetic.sparse_table(n, 60000, seed=1)
ctx = _native.Context(0)
out = torch.empty(_native.tri_len(n), dtype=torch.int32).pin_memory().numpy()
for it in range(4):
    ctx.reset_counters()
    t0 = time.perf_counter()
    ctx.load_table(table)
    t1 = time.perf_counter()
    if metric in ("aai", "peq"):
        ctx.load_sequences(table)
    t2 = time.perf_counter()
    ctx.matrix(metric, out=out)
    t3 = time.perf_counter()
    tm = ctx.timings()
    print(f"iter {it}: load_table {1e3*(t1-t0):.1f} ms, load_sequences {1e3*(t2-t1):.1f} ms, "
          f"matrix {1e3*(t3-t2):.1f} ms | " + ", ".join(f"{k}={v:.1f}" for k, v in tm.items()))
