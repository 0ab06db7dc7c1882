#!/usr/bin/env python
"""tools/probe_partial_load.py -- where does pc_load_sequences_part spend its host time?
One GPU is enough: the probe loads partition 0 of 8 (what rank 0 of an 8-GPU job does) with
PC_DEBUG=1, which makes the library print the time between its internal marks.

    PC_DEBUG=1 LOCAL_WORLD_SIZE=8 python tools/probe_partial_load.py [n_genomes ...]
"""
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from phamclust_b200 import synthetic          # noqa: E402
from phamclust_b200._native import Context    # noqa: E402


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [2000]
    ctx = Context(0)
    for n in sizes:
        t0 = time.perf_counter()
        table = synthetic.phage_table(n, seed=42)
        print(f"# {n} genomes: {table.n_genes} genes, {len(table.residues) / 1e6:.1f} MB of residues, "
              f"generated in {time.perf_counter() - t0:.1f} s", flush=True)
        for rep in range(3):
            for part, parts in ((0, 1), (0, 8), (3, 8)):
                ctx.reset_counters()
                t0 = time.perf_counter()
                ctx.load_table(table)
                t1 = time.perf_counter()
                ctx.load_sequences(table, part, parts)
                t2 = time.perf_counter()
                n_units = ctx.align_units(part, parts) if parts > 1 else 0
                t3 = time.perf_counter()
                tm = ctx.timings()
                print(f"## rep {rep} part {part}/{parts}: load_table {1e3 * (t1 - t0):.2f} ms, load_sequences "
                      f"{1e3 * (t2 - t1):.2f} ms, plan {1e3 * (t3 - t2):.2f} ms ({n_units} units); stages "
                      + ", ".join(f"{k} {v:.2f}" for k, v in tm.items() if v), flush=True)


if __name__ == "__main__":
    main()
