"""How far can the UNPINNED tie rules of the parasail restatement (SURVEY.md Appendix A,
U2-U4) move the results?  TEST INFRASTRUCTURE (imports oracle/).

For every alignment the reference would run on a sample of the C2 stand-in, oracle
nw_tie_range gives the min/max of (identical columns, length) over ALL co-optimal
alignments of the score model.  If min == max the pair's contribution is the same under
any tie rule, i.e. pinned by the score model alone.  For the rest, aai / peq are
re-evaluated with every alignment pushed to its lowest (ident_min / length_max) and
highest (ident_max / length_min) ratio: the spread is what a different tie rule could
change at most.

    python tools/tie_exposure.py [--genomes 100] > profiles/r1_tie_exposure.txt
"""
import argparse
import sys

import numpy as np

sys.path.insert(0, ".")
from oracle import nw_c  # noqa: E402
from oracle import pham_oracle as po  # noqa: E402
from phamclust_b200 import synthetic  # noqa: E402
from phamclust_b200.packing import unpack_genomes  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--genomes", type=int, default=100)
ap.add_argument("--seed", type=int, default=42)
args = ap.parse_args()

table = synthetic.phage_table(args.genomes, seed=args.seed)
genomes = unpack_genomes(table)
cache = {}


def ranges(a, b):
    key = (a, b)
    if key not in cache:
        cache[key] = nw_c.tie_range(a, b)
    return cache[key]


def align_rule(a, b):
    return nw_c.forward_stats(a, b)[:3]


def align_lo(a, b):
    r = ranges(a, b)
    return r[0], r[2], r[5]


def align_hi(a, b):
    r = ranges(a, b)
    return r[1], r[3], r[4]


def aai(s, t, align):
    best = po.aai_best_matches(s, t, align=align)
    hit, length = sum(b[1] for b in best), sum(b[2] for b in best)
    return (hit / length) if length else 0.0, hit, length


n = len(genomes)
spread_aai, spread_peq, pairs_with_alignments = [], [], 0
outside = 0
for i in range(n):
    for j in range(i + 1, n):
        s, t = genomes[i], genomes[j]
        if not po.shared_phams(s, t):
            continue
        pairs_with_alignments += 1
        lo, hi, mid = aai(s, t, align_lo)[0], aai(s, t, align_hi)[0], aai(s, t, align_rule)[0]
        outside += not (lo - 1e-12 <= mid <= hi + 1e-12)
        af = po.alignment_fraction(s, t)
        spread_aai.append(hi - lo)
        spread_peq.append(af * (hi - lo))

r = np.array(list(cache.values()), dtype=np.int64)
dep = (r[:, 0] != r[:, 1]) | (r[:, 4] != r[:, 5])
d_id, d_ln = r[:, 1] - r[:, 0], r[:, 5] - r[:, 4]
spread_aai, spread_peq = np.array(spread_aai), np.array(spread_peq)
print(f"# tools/tie_exposure.py --genomes {args.genomes} --seed {args.seed} (synthetic.phage_table: the C2 stand-in's generator)")
print(f"distinct ordered alignments: {len(r)}")
print(f"  tie-INDEPENDENT (ident and length identical over every co-optimal alignment): "
      f"{int((~dep).sum())} = {100.0 * (~dep).mean():.2f} %")
print(f"  tie-dependent: {int(dep.sum())} = {100.0 * dep.mean():.2f} %; "
      f"ident spread max {int(d_id.max())} mean-over-dependent {d_id[dep].mean():.2f}; "
      f"length spread max {int(d_ln.max())} mean-over-dependent {d_ln[dep].mean():.2f}")
print(f"genome pairs with at least one alignment: {pairs_with_alignments}")
print(f"  aai  spread (highest-ratio rule minus lowest-ratio rule): max {spread_aai.max():.3e}, "
      f"median {np.median(spread_aai):.3e}, pairs above 1e-6: {int((spread_aai > 1e-6).sum())} "
      f"({100.0 * (spread_aai > 1e-6).mean():.1f} %)")
print(f"  peq  spread: max {spread_peq.max():.3e}, median {np.median(spread_peq):.3e}, "
      f"pairs above 1e-6: {int((spread_peq > 1e-6).sum())} ({100.0 * (spread_peq > 1e-6).mean():.1f} %)")
print(f"  restated rules outside [lowest, highest]: {outside} pairs")
