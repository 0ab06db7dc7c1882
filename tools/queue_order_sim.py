#!/usr/bin/env python
"""tools/queue_order_sim.py -- how the tile queue of the persistent pair kernel should be ordered.

One round of the kernel's static schedule is a window of #clusters consecutive queue entries that
run in lock step (csrc/pc_contract_i8p.cu): a round lasts as long as its LONGEST tile, and its
tiles share operand strips in L2 only if they are close in the tile grid.  This script rebuilds
the K-block occupancy of the C3 and C5-shape tables on the host (the library's column order:
phams held by >= 2 genomes, by first genome), and for several queue keys reports
  * round-max efficiency = (sum of all tiles' K steps / #clusters) / (sum over rounds of the
    longest tile), i.e. what lock step costs, and
  * distinct operand strips per round, i.e. the DRAM traffic a round needs at least.
DESIGN.md section 3 quotes its output; the key the kernel uses is "work>>s then block" with s
chosen so that the launch has at most 32 work classes.

    PYTHONPATH=. python tools/queue_order_sim.py
"""
import numpy as np
from phamclust_b200 import synthetic
def tiles_of(t):
    n = t.n_genomes
    pid = t.pham_id.astype(np.int64)
    counts = np.bincount(pid, minlength=t.n_phams)
    seen, first_at = np.unique(pid, return_index=True)
    first = np.full(t.n_phams, len(pid)); first[seen] = first_at
    elig = np.nonzero(counts >= 2)[0]
    col = np.full(t.n_phams, -1); col[elig[np.argsort(first[elig], kind="stable")]] = np.arange(len(elig))
    ec = col[pid]; eg = np.repeat(np.arange(n), np.diff(t.genome_ptr)); keep = ec >= 0
    ng = (n + 255)//256; nb = (len(elig)+127)//128
    occ = np.zeros((ng, nb), dtype=np.float32); occ[eg[keep]//256, ec[keep]//128] = 1
    common = occ @ occ.T
    bi, bj = np.triu_indices(ng)
    return bi, bj, common[bi, bj].astype(np.int64), ng
def evaluate(bi, bj, steps, order, C=74):
    s = steps[order]; tot=0; strips=0
    for i in range(0, len(s), C):
        sel = order[i:i+C]
        tot += s[i:i+C].max()
        strips += len(set(bi[sel].tolist()) | set((bj[sel]).tolist()))
    return steps.sum()/C/tot, strips/ (len(s)/C)
for name, t in (("c3", synthetic.sparse_table(10000, 60000, seed=1)), ("c5-30k", synthetic.sparse_table(30000, 200000, seed=3))):
    bi, bj, steps, ng = tiles_of(t)
    print(name, "tiles", len(steps), "steps mean/min/max", steps.mean(), steps.min(), steps.max())
    blk = (bi//8)*4096 + (4095 - np.minimum(bj//9, 4095))
    schemes = {
      "block(current)": (blk.astype(np.int64) << 10) | np.minimum(steps >> 3, 1023),
      "work": steps.astype(np.int64) << 20,
    }
    for sh in (3, 4, 5, 6):
        schemes[f"work>>{sh} then block"] = ((steps >> sh).astype(np.int64) << 32) | (blk.astype(np.int64) << 10) | np.minimum(steps>>3,1023)
    for k, key in schemes.items():
        order = np.argsort(-key, kind="stable")
        eff, strips = evaluate(bi, bj, steps, order)
        print(f"   {k:24s} round-max efficiency {eff:.3f}   distinct strips per round {strips:.1f}")
