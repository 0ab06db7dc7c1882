"""Parity tests proper: the CUDA path, called through the C ABI (ctypes), against
the CPU oracle and the committed golden vectors.  Integer work is bit-exact;
the only floating-point tolerance is aai/peq's <= 1e-6 (1 micro-unit), which
north_star states and which exists because upstream sums floats over a
hash-ordered set (SURVEY.md 7 hard part 4) -- the integer sums under it are exact.
"""
import random

import numpy as np
import pytest

from conftest import load_golden
from oracle import nw_c
from oracle import pham_oracle as po
from oracle import testdata
from phamclust_b200 import (METRICS, Genome, SymMatrix, matrix_de_novo, pack_genomes,
                            packed_matrix)
from phamclust_b200 import metrics as pm
from phamclust_b200 import synthetic
from phamclust_b200._native import tri_len
from phamclust_b200.packing import unpack_genomes

pytestmark = pytest.mark.gpu

CHEAP = ("gcs", "jc", "pocp", "af")
ALL = CHEAP + ("aai", "peq")
FLOAT_TOL_MICRO = 1          # aai / peq: 1e-6 absolute (north_star)


def _oracle_triangle(genomes, key, as_distance):
    w = po.matrix_de_novo(genomes, po.METRICS[key], 1, as_distance=as_distance)
    n = len(genomes)
    return np.array([po.round6_micro(w[(i, j)]) for i in range(n) for j in range(i, n)], np.int64)


def _load(ctx, table, sequences=True):
    ctx.load_table(table)
    if sequences and table.seq_off is not None:
        ctx.load_sequences(table)


# ----------------------------------------------------------------------------- rounding
def test_device_round6_equals_cpython_round(gpu_ctx):
    cases = load_golden("round6.json")
    xs = np.array([float.fromhex(c["x"]) for c in cases])
    assert gpu_ctx.debug_round6(xs).tolist() == [c["micro"] for c in cases]
    rng = np.random.default_rng(3)
    den = rng.integers(1, 200000, size=200000)
    num = (rng.random(200000) * (den + 1)).astype(np.int64)
    xs = np.concatenate((num / den, 1.0 - num / den, rng.random(100000)))
    expect = [int(round(round(float(x), 6) * 1e6)) for x in xs]
    assert gpu_ctx.debug_round6(xs).tolist() == expect


def test_integer_epilogue_equals_cpython_arithmetic(gpu_ctx):
    """The production epilogue of gcs/jc/pocp/af works in integers (exact rational, round half
    even) and only asks fp64 on exact .5 ties; both it and the fp64 path must equal what the
    reference's float arithmetic gives (metrics.py:45-53 etc. + matrix.py:323) -- on random
    operands, on every small denominator exhaustively, and on the tie-prone ones (denominators
    that are multiples of 128, where 10^6 p / q can end in .5)."""
    rng = np.random.default_rng(99)
    cases = []
    for q in list(range(1, 260)) + [384, 512, 640, 768, 896, 1024, 1280, 2560, 12800, 65536, 10 ** 6, 2 ** 20, 2 ** 30 - 128]:
        ps = range(0, q + 1) if q <= 260 else sorted({int(x) for x in rng.integers(0, q + 1, size=300)} | {0, 1, q // 2, q - 1, q})
        cases += [(p, q) for p in ps]
    big = rng.integers(1, 2 ** 31 - 2, size=20000)
    cases += [(int(rng.integers(0, q + 1)), int(q)) for q in big]
    p = np.array([c[0] for c in cases], dtype=np.int64)
    q = np.array([c[1] for c in cases], dtype=np.int64)

    def python_micro(sim, as_distance):
        return int(round(round(1.0 - sim if as_distance else sim, 6) * 1e6))

    for metric in ("gcs", "jc", "pocp", "af"):
        # operands (numer, den_i, den_j) whose ratio is p / q for that metric
        if metric == "gcs":                      # sim = 2 numer / den: use even p
            keep = (p % 2 == 0)
            numer, den = p[keep] // 2, q[keep]
        elif metric == "jc":                     # sim = numer / (den - numer)
            numer, den = p, q + p
            keep = den < 2 ** 31
            numer, den = numer[keep], den[keep]
        else:
            numer, den = p, q
        den_i = den // 2
        den_j = den - den_i
        for as_distance in (True, False):
            got, got64 = gpu_ctx.debug_metric_micro(metric, as_distance, numer, den_i, den_j)
            want = []
            for a, d in zip(numer.tolist(), den.tolist()):
                if a == 0:
                    sim = 0.0
                elif metric == "gcs":
                    sim = (2.0 * a) / d
                elif metric == "jc":
                    sim = a / (d - a)
                else:
                    sim = a / d
                want.append(python_micro(sim, as_distance))
            want = np.array(want, dtype=np.int32)
            assert np.array_equal(got64, want), (metric, as_distance, "fp64 path")
            bad = np.nonzero(got != want)[0]
            assert bad.size == 0, (metric, as_distance, [(int(numer[k]), int(den[k]), int(got[k]), int(want[k])) for k in bad[:5]])


# ----------------------------------------------------------------------------- golden matrices
@pytest.mark.parametrize("dataset", sorted(testdata.DATASETS))
def test_golden_matrices_from_the_verbatim_reference(gpu_ctx, dataset):
    golden = load_golden("matrices.json")[dataset]
    genomes = testdata.rows_to_genomes([tuple(r) for r in golden["rows"]], cls=Genome)
    table = pack_genomes(genomes)
    _load(gpu_ctx, table)
    for key in ALL:
        for mode, as_distance in (("distance", True), ("similarity", False)):
            got = gpu_ctx.matrix(key, as_distance=as_distance).astype(np.int64)
            want = np.array(golden[mode][key], dtype=np.int64)
            if key in CHEAP:
                assert np.array_equal(got, want), (key, mode)
            else:
                assert np.abs(got - want).max() <= FLOAT_TOL_MICRO, (key, mode)


def test_known_answer_vector(gpu_ctx):
    ka = load_golden("known_answer.json")
    a, b = testdata.rows_to_genomes([tuple(r) for r in ka["rows"]], cls=Genome)
    for key in ALL:
        assert METRICS[key](a, b) == ka["similarity"][key]
        assert METRICS[key](a, b, as_distance=True) == ka["distance"][key]
    _load(gpu_ctx, pack_genomes([a, b]))
    assert gpu_ctx.numerators("gcs").tolist() == [0, 2, 0]
    assert gpu_ctx.numerators("pocp").tolist() == [0, 4, 0]
    assert gpu_ctx.numerators("af").tolist() == [0, 14, 0]


# ----------------------------------------------------------------------------- integers, bit-exact
@pytest.mark.parametrize("seed", [31, 32, 33])
def test_integer_numerators_bit_exact(gpu_ctx, seed):
    rows = testdata.genome_rows(seed=seed, n_genomes=10, n_phams=14, phams_per_genome=(2, 12),
                                paralog_p=0.45, low_complexity=(seed % 2 == 0))
    genomes = testdata.rows_to_genomes(rows, cls=Genome)
    n = len(genomes)
    _load(gpu_ctx, pack_genomes(genomes))
    pairs = [(i, j) for i in range(n) for j in range(i, n)]
    shared = gpu_ctx.numerators("gcs")
    pocp = gpu_ctx.numerators("pocp")
    af = gpu_ctx.numerators("af")
    hit, length = gpu_ctx.numerators("aai")
    hit_p, length_p = gpu_ctx.numerators("aai", ppos=True)
    for k, (i, j) in enumerate(pairs):
        if i == j:
            continue
        s, t = genomes[i], genomes[j]
        assert shared[k] == po.gcs_parts(s, t)[0]
        assert pocp[k] == po.pocp_parts(s, t)[0]
        assert af[k] == po.af_parts(s, t)[0]
        assert (hit[k], length[k]) == po.aai_parts(s, t)
        assert (hit_p[k], length_p[k]) == po.aai_parts(s, t, ppos=True)
    # and every metric through the full path
    for key in ALL:
        got = gpu_ctx.matrix(key, as_distance=True).astype(np.int64)
        want = _oracle_triangle(genomes, key, True)
        tol = 0 if key in CHEAP else FLOAT_TOL_MICRO
        assert np.abs(got - want).max() <= tol, key
    got = gpu_ctx.matrix("aai", as_distance=False, ppos=True).astype(np.int64)
    want = np.array([1000000 if i == j else po.round6_micro(
        po.average_aminoacid_identity(genomes[i], genomes[j], ppos=True)) for i, j in pairs])
    assert np.abs(got - want).max() <= FLOAT_TOL_MICRO


# ----------------------------------------------------------------------------- alignments
def _arena(seqs):
    lens = np.array([len(s) for s in seqs], dtype=np.int64)
    off = np.concatenate(([0], np.cumsum(lens)))
    res = np.frombuffer("".join(seqs).encode("latin-1"), dtype=np.uint8).copy()
    return off, res


def _genes_table(seqs):
    """One genome per sequence, each in its own pham: an arena for the hook."""
    from phamclust_b200.packing import PackedTable
    n = len(seqs)
    off, res = _arena(seqs)
    return PackedTable([f"s{k}" for k in range(n)], [f"p{k}" for k in range(n)],
                       np.arange(n + 1, dtype=np.int64), np.arange(n, dtype=np.int32),
                       np.ones(n, dtype=np.uint32), np.diff(off).astype(np.uint32), off, res)


@pytest.fixture(params=["auto", "classic"])
def dp_mode(request, gpu_ctx):
    """Both DP kernels: "auto" routes every alignment whose word fits to the LEX
    kernel (counters inside the score word), "classic" never does."""
    gpu_ctx.set_dp_mode(request.param)
    yield request.param
    gpu_ctx.set_dp_mode("auto")


def test_golden_alignment_vectors(gpu_ctx, dp_mode):
    cases = load_golden("alignments.json")
    seqs, q, r = [], [], []
    for c in cases:
        q.append(len(seqs)); seqs.append(c["query"])
        r.append(len(seqs)); seqs.append(c["ref"])
    _load(gpu_ctx, _genes_table(seqs))
    got = gpu_ctx.alignment_identities(q, r)
    want = np.array([[c["ident"], c["similar"], c["length"], c["score"]] for c in cases])
    bad = np.nonzero((got != want).any(axis=1))[0]
    assert bad.size == 0, [(cases[k]["query"], cases[k]["ref"], got[k].tolist(), want[k].tolist())
                           for k in bad[:3]]


def test_random_alignments_bit_exact_all_kernel_configs(gpu_ctx, dp_mode):
    """Lengths chosen to land in every DP config (32..1024 columns per pass and
    the multi-chunk configs with 2-4 column chunks), tie-rich alphabets included."""
    rng = random.Random(77)
    seqs = []
    for cap in (8, 30, 33, 64, 65, 96, 100, 128, 150, 192, 200, 256, 300, 384, 400, 512, 513,
                700, 1100, 1600, 2100, 2650, 3300):
        for alpha in ("AG", testdata.AA):
            a = testdata.random_protein(rng, max(1, cap - 6), cap, alpha)
            seqs += [a, testdata.mutate(rng, a, 0.25, alpha), testdata.mutate(rng, a, 0.6, alpha)]
    seqs += ["", "A", "W" * 40, "AG" * 300]
    n = len(seqs)
    q = [rng.randrange(n) for _ in range(1500)] + list(range(n))
    r = [rng.randrange(n) for _ in range(1500)] + list(range(n))
    _load(gpu_ctx, _genes_table(seqs))
    got = gpu_ctx.alignment_identities(q, r)
    off, res = _arena(seqs)
    want = nw_c.batch(res, off, q, r, mode=0)          # explicit trace table + traceback
    bad = np.nonzero((got != want).any(axis=1))[0]
    assert bad.size == 0, [(len(seqs[q[k]]), len(seqs[r[k]]), got[k].tolist(), want[k].tolist())
                           for k in bad[:5]]


def test_every_reachable_dp_config_is_launched_and_bit_exact(gpu_ctx, dp_mode):
    """One ref length just under every G x K capacity (K even: 8..32 classic, 8..48
    LEX), against short queries (LEX-admissible: min length < 512 and the score word
    fits), a same-length mutant and a 4090-residue query (never LEX-admissible: the
    classic kernel at every K); then ref lengths beyond the single-pass grid for the
    multi-chunk configs.  The library must report that every config its chooser
    can return was launched, and every alignment must equal the traceback oracle."""
    rng = random.Random(4848)
    short = [testdata.random_protein(rng, m, m, testdata.AA) for m in (1, 37, 300, 4090)]
    seqs = list(short)
    q, r = [], []
    lengths = sorted({g * k - d for g in (4, 8, 16, 32) for k in range(8, 50, 2) for d in (0, 1)}
                     | {1537, 1700, 2049, 2300, 2561, 2700, 3000}
                     # twins of these lengths land in every LEXC config (K = 16 .. 24, 2 column chunks)
                     | {1030, 1100, 1270, 1400, 1530})
    for n in lengths:
        ref = testdata.random_protein(rng, n, n, testdata.AA if n % 3 else "AG")
        twin = testdata.mutate(rng, ref, 0.3, testdata.AA)
        base = len(seqs)
        seqs += [ref, twin]
        for k in range(len(short)):
            q.append(k); r.append(base)
        q.append(base + 1); r.append(base)
        q.append(base); r.append(base + 1)
    _load(gpu_ctx, _genes_table(seqs))
    gpu_ctx.reset_counters()
    got = gpu_ctx.alignment_identities(q, r)
    stats = gpu_ctx.stats()
    assert stats["dp_configs_used"] == stats["dp_configs"], stats
    assert stats["dp_configs"] >= (65 if dp_mode == "auto" else 40), stats
    off, res = _arena(seqs)
    want = nw_c.batch(res, off, q, r, mode=0)
    bad = np.nonzero((got != want).any(axis=1))[0]
    assert bad.size == 0, [(len(seqs[q[k]]), len(seqs[r[k]]), got[k].tolist(), want[k].tolist())
                           for k in bad[:5]]


def test_lex_kernel_at_the_edges_of_its_word(gpu_ctx):
    """The LEX DP kernel packs score | priority | diagonal steps | matches into 32
    bits (9-bit counters, 12-bit biased score).  Lengths around 511, score bounds
    around the admission limit (W scores 11 against itself), one very long partner
    (the m + n floor), all-gap and tie-heavy inputs: every alignment must equal the
    explicit-traceback oracle whichever kernel pc_nw_lex_ok routes it to."""
    rng = random.Random(511)
    seqs = ["W" * k for k in (200, 250, 262, 263, 264, 270, 300, 372, 511, 512)]
    seqs += ["A" * k for k in (100, 509, 510, 511, 512, 513)]
    seqs += ["AG" * 255, "GA" * 255 + "A", "AG" * 256, "P" * 511, "C" * 400]
    for cap in (500, 505, 509, 510, 511, 512, 513, 520):
        for alpha in ("AG", "WC", testdata.AA):
            a = testdata.random_protein(rng, cap, cap, alpha)
            seqs += [a, testdata.mutate(rng, a, 0.1, alpha), testdata.mutate(rng, a, 0.7, alpha)]
    for cap in (60, 100, 300, 511):
        seqs.append(testdata.random_protein(rng, cap, cap, testdata.AA))
    seqs += [testdata.random_protein(rng, k, k, testdata.AA) for k in (1400, 2300, 3100)]
    n = len(seqs)
    q = [rng.randrange(n) for _ in range(2500)] + list(range(n))
    r = [rng.randrange(n) for _ in range(2500)] + list(range(n))
    _load(gpu_ctx, _genes_table(seqs))
    off, res = _arena(seqs)
    want = nw_c.batch(res, off, q, r, mode=0)
    for mode in ("auto", "classic"):
        gpu_ctx.set_dp_mode(mode)
        gpu_ctx.reset_counters()
        try:
            got = gpu_ctx.alignment_identities(q, r)
        finally:
            gpu_ctx.set_dp_mode("auto")
        lex = gpu_ctx.stats()["lex_alignments"]
        # the hook runs every task twice (identity and ppos tables)
        assert (lex > len(q)) if mode == "auto" else (lex == 0)
        bad = np.nonzero((got != want).any(axis=1))[0]
        assert bad.size == 0, (mode, [(len(seqs[q[k]]), len(seqs[r[k]]), got[k].tolist(), want[k].tolist())
                                      for k in bad[:5]])


def test_rebased_lex_kernel_long_alignments(gpu_ctx):
    """LEXR: alignments with 512 <= min length < 1024 (or a score range the plain LEX word cannot
    hold) and a ref of at most 1024 columns run on the rebased LEX kernel -- 10-bit counters, 10-bit
    score field re-centred per lane.  Steepest possible score slopes (all-W), flat ones, tie-rich
    alphabets, lengths at every edge (511/512, 1023/1024, refs at 32 x K boundaries): every
    alignment must equal the explicit-traceback oracle, and the LEXR kernel must have run."""
    rng = random.Random(1023)
    seqs = ["W" * k for k in (512, 600, 767, 768, 1000, 1023, 1024)]
    seqs += ["A" * k for k in (513, 1023, 1024, 1025)] + ["AG" * 400, "GA" * 511 + "A", "WC" * 300, "P" * 900]
    for cap in (512, 513, 540, 576, 577, 640, 700, 704, 768, 769, 832, 900, 960, 1000, 1022, 1023, 1024, 1030):
        for alpha in ("AG", "WC", testdata.AA):
            a = testdata.random_protein(rng, cap, cap, alpha)
            seqs += [a, testdata.mutate(rng, a, 0.15, alpha), testdata.mutate(rng, a, 0.6, alpha)]
    seqs += [testdata.random_protein(rng, k, k, testdata.AA) for k in (30, 400, 1400, 2500)]
    n = len(seqs)
    q = [rng.randrange(n) for _ in range(1200)] + list(range(n))
    r = [rng.randrange(n) for _ in range(1200)] + list(range(n))
    _load(gpu_ctx, _genes_table(seqs))
    off, res = _arena(seqs)
    want = nw_c.batch(res, off, q, r, mode=0)
    gpu_ctx.set_dp_mode("auto")
    gpu_ctx.reset_counters()
    got = gpu_ctx.alignment_identities(q, r)
    stats = gpu_ctx.stats()
    assert stats["lexr_cells"] > 0, stats
    bad = np.nonzero((got != want).any(axis=1))[0]
    assert bad.size == 0, [(len(seqs[q[k]]), len(seqs[r[k]]), got[k].tolist(), want[k].tolist()) for k in bad[:5]]


def test_lexc_kernel_any_length(gpu_ctx):
    """LEXC: what neither LEX nor LEXR can take (min length >= 1024, or a ref of more than 1024
    columns) and whose two lengths differ by at most 128 runs on the rebased LEX word with REBASED
    COUNTERS, in column chunks.  The kernel checks its counter band and hands an alignment to the
    classic kernel when a check fails, so every result must equal the explicit-traceback oracle
    whichever way it went: homologs with small indels (stay on LEXC), lengths at every chunk edge
    (32 x K columns, K = 16 .. 24), all-W / single-letter / tie-rich inputs, and pairs built to
    break the band (a + b against b + a, a 300-residue block moved, unrelated sequences)."""
    rng = random.Random(1536)
    seqs, q, r = [], [], []

    def pair(a, b):
        seqs.extend([a, b])
        q.extend([len(seqs) - 2, len(seqs) - 1])
        r.extend([len(seqs) - 1, len(seqs) - 2])

    for n in (1024, 1025, 1030, 1087, 1088, 1151, 1152, 1153, 1279, 1280, 1281, 1407, 1408, 1409, 1500, 1535, 1536,
              1537, 1600, 1727, 1728, 1900, 2047, 2048, 2049, 2111, 2112, 2303, 2304, 2305, 2600, 3071, 3072, 3300):
        for alpha, rate in ((testdata.AA, 0.05), (testdata.AA, 0.3), ("AG", 0.1)):
            a = testdata.random_protein(rng, n, n, alpha)
            pair(a, testdata.mutate(rng, a, rate, alpha))
    # short query / long ref inside the length window: m < 1024 <= n
    for m, n in ((960, 1060), (1000, 1128), (1023, 1025), (950, 1078)):
        a = testdata.random_protein(rng, n, n, testdata.AA)
        pair(testdata.mutate(rng, a, 0.2, testdata.AA)[:m], a)
    for k in (1024, 1100, 1536, 2000):
        pair("W" * k, "W" * (k + 7))
        pair("A" * k, "A" * (k - 3))
        pair("AG" * (k // 2), "GA" * (k // 2) + "A")
    n_band_ok = len(q)
    # built to leave the band: the checks must catch all of them
    for k in (600, 800, 1100):
        a, b = testdata.random_protein(rng, k, k, testdata.AA), testdata.random_protein(rng, k, k, testdata.AA)
        pair(a + b, b + a)
    for k in (1300, 1800):
        a = testdata.random_protein(rng, k, k, testdata.AA)
        pair(a, a[:200] + a[500:] + a[200:500])
        pair(a, testdata.random_protein(rng, k + 40, k + 40, testdata.AA))
        pair(a, testdata.mutate(rng, a, 0.7, testdata.AA))
    _load(gpu_ctx, _genes_table(seqs))
    off, res = _arena(seqs)
    want = nw_c.batch(res, off, q, r, mode=0)
    gpu_ctx.set_dp_mode("auto")
    gpu_ctx.reset_counters()
    got = gpu_ctx.alignment_identities(q, r)
    stats = gpu_ctx.stats()
    bad = np.nonzero((got != want).any(axis=1))[0]
    assert bad.size == 0, [(len(seqs[q[k]]), len(seqs[r[k]]), got[k].tolist(), want[k].tolist()) for k in bad[:5]]
    assert stats["lexc_cells"] > 0, stats
    # the hook runs every task twice (identity and ppos tables); the band-breaking pairs fall back,
    # the homologs do not (a handful may, where a mutated copy happens to open a long gap)
    assert 2 * 6 <= stats["lexc_fallbacks"] <= 2 * (len(q) - n_band_ok) + n_band_ok // 4, stats
    # and the classic kernel alone gives the same
    gpu_ctx.set_dp_mode("classic")
    try:
        assert np.array_equal(gpu_ctx.alignment_identities(q, r), want)
    finally:
        gpu_ctx.set_dp_mode("auto")


def test_lexc_gives_up_on_data_with_long_gaps(gpu_ctx):
    """One pham of twelve 1300-residue genes that are rearrangements of each other (a + b, b + a,
    a block moved): every alignment breaks LEXC's counter band and is re-run by the classic kernel
    -- the matrix must equal the oracle's -- and since more than a third fell back, the next call
    on the same sequences does not try LEXC again."""
    rng = random.Random(77)
    a, b, c = (testdata.random_protein(rng, k, k, testdata.AA) for k in (650, 650, 300))
    variants = [a + b, b + a, a[:300] + b + a[300:], b[:200] + a + b[200:], c + a + b[:350], a + b[:350] + c,
                b + c + a[:350], a[350:] + b + c + a[:50], b[:325] + a + b[325:], a[:500] + c + b + a[500:550],
                c + b + a[:350], b[300:] + a + b[:300]]
    variants = [v[:1300] for v in variants]
    genomes = [Genome(f"g{k}") for k in range(len(variants))]
    for k, (g, v) in enumerate(zip(genomes, variants)):
        g.add("pham1", testdata.mutate(rng, v, 0.05))
        g.add(f"own{k}", testdata.random_protein(rng, 50, 50, testdata.AA))
    table = pack_genomes(genomes)
    _load(gpu_ctx, table)
    gpu_ctx.reset_counters()
    first = gpu_ctx.matrix("peq", as_distance=True).astype(np.int64)
    s1 = gpu_ctx.stats()
    second = gpu_ctx.matrix("aai", as_distance=True).astype(np.int64)
    s2 = gpu_ctx.stats()
    third = gpu_ctx.matrix("peq", as_distance=True).astype(np.int64)
    s3 = gpu_ctx.stats()
    assert s1["lexc_fallbacks"] > 22, s1                      # 66 alignments, more than a third gave up
    assert s2["lexc_fallbacks"] >= s1["lexc_fallbacks"]       # the second call learns it at its sync ...
    assert s3["lexc_fallbacks"] == s2["lexc_fallbacks"] and s3["lexc_cells"] == s2["lexc_cells"], (s2, s3)   # ... the third skips LEXC
    ref_genomes = unpack_genomes(table)
    assert np.abs(first - _oracle_triangle(ref_genomes, "peq", True)).max() <= FLOAT_TOL_MICRO
    assert np.abs(second - _oracle_triangle(ref_genomes, "aai", True)).max() <= FLOAT_TOL_MICRO
    assert np.array_equal(first, third)


def test_non_standard_residues_and_case(gpu_ctx, dp_mode):
    seqs = ["mktwACDE", "MKTWacde", "MKUXO*BZ", "MKOXU*ZB", "JJJJ", "jJjJ", "M-K.T", "M-K.T"]
    _load(gpu_ctx, _genes_table(seqs))
    q, r = [0, 2, 4, 6, 1, 3], [1, 3, 5, 7, 2, 0]
    got = gpu_ctx.alignment_identities(q, r)
    off, res = _arena(seqs)
    assert np.array_equal(got, nw_c.batch(res, off, q, r, mode=0))
    assert got[0].tolist()[:3] == [8, 0, 8]


# ----------------------------------------------------------------------------- the Python boundary
def test_matrix_de_novo_drop_in(gpu_ctx):
    rows = testdata.genome_rows(seed=41, n_genomes=9, n_phams=12, paralog_p=0.3)
    genomes = testdata.rows_to_genomes(rows, cls=Genome)
    for key in ALL:
        for as_distance in (True, False):
            m = matrix_de_novo(genomes, METRICS[key], cpus=4, as_distance=as_distance)
            assert isinstance(m, SymMatrix) and m.is_distance == as_distance
            assert m.nodes == [g.name for g in genomes]
            want = po.matrix_de_novo(genomes, po.METRICS[key], 1, as_distance=as_distance)
            tol = 0 if key in CHEAP else 1.0000001e-6
            for i, a in enumerate(genomes):
                assert m.get_weight(a.name, a.name) == 1.0 - as_distance
                for j in range(i + 1, len(genomes)):
                    got = m.get_weight(a.name, genomes[j].name)
                    assert 0.0 <= got <= 1.0 and abs(got - want[(i, j)]) <= tol
    # per-pair call keeps the reference signature and orientation
    s, t = genomes[0], genomes[1]
    assert pm.gene_content_similarity(s, t) == po.gene_content_similarity(s, t)
    assert pm.proteomic_equivalence_quotient(s, t, as_distance=True) == pytest.approx(
        po.proteomic_equivalence_quotient(s, t, as_distance=True), abs=1.0000001e-6)
    assert pm.average_aminoacid_identity(s, t, ppos=True) == pytest.approx(
        po.average_aminoacid_identity(s, t, ppos=True), abs=1.0000001e-6)


def test_reference_function_objects_are_recognised(gpu_ctx):
    # a caller holding the REFERENCE's function objects (same __name__) still dispatches
    def gene_content_similarity(source, target, as_distance=False):
        raise AssertionError("must not be called per pair")
    genomes = testdata.rows_to_genomes(testdata.genome_rows(seed=43, n_genomes=4, n_phams=6), cls=Genome)
    m = matrix_de_novo(genomes, gene_content_similarity, 1)
    assert m.get_weight(genomes[0].name, genomes[1].name) == \
        po.gene_content_similarity(genomes[0], genomes[1], as_distance=True)


def test_cluster_assignments_identical_through_unchanged_clustering(gpu_ctx):
    from phamclust_b200.clustering import hierarchical_clustering
    for dataset, golden in load_golden("matrices.json").items():
        genomes = testdata.rows_to_genomes([tuple(r) for r in golden["rows"]], cls=Genome)
        for key in ALL:
            dist = matrix_de_novo(genomes, METRICS[key], 1)
            for spec, expect in golden["clusters"][key].items():
                linkage, eps = spec.split(":")
                got = hierarchical_clustering(dist, linkage, eps=float(eps))
                assert sorted(sorted(c.nodes) for c in got) == expect, (dataset, key, spec)


# ----------------------------------------------------------------------------- edge cases
def test_edge_cases(gpu_ctx):
    solo = Genome("solo")
    solo.add("p", "MKT")
    m = matrix_de_novo([solo], METRICS["peq"], 1)
    assert list(m) == [("solo", "solo", 0.0)]
    a, b = Genome("a"), Genome("b")
    a.add("x", "MKT")
    b.add("y", "MKTW")
    for key in ALL:                       # nothing shared -> similarity exactly 0.0
        assert METRICS[key](a, b) == 0.0 and METRICS[key](a, b, as_distance=True) == 1.0
    twin = Genome("twin")
    twin.add("x", "MKT")
    for key in ALL:                       # identical genomes -> distance 0.0
        assert METRICS[key](a, twin, as_distance=True) == 0.0
    with pytest.raises(ValueError):
        matrix_de_novo([], METRICS["gcs"], 1)
    table = pack_genomes([a, b], with_sequences=False)
    gpu_ctx.load_table(table)
    with pytest.raises(ValueError):       # aai without translations: loud, not silent
        gpu_ctx.matrix("aai")
    bad = pack_genomes([a, b])
    bad.pham_id = np.array([5, 1], dtype=np.int32)
    with pytest.raises(ValueError):
        gpu_ctx.load_table(bad)
    long_gene = Genome("long")
    long_gene.add("x", "A" * 40000)
    with pytest.raises(ValueError):
        METRICS["aai"](a, long_gene)


def test_empty_genomes_heavy_paralogy_and_big_phams(gpu_ctx):
    """An empty genome, 6-copy paralog families with exact duplicates (the
    'last maximum wins' tie of metrics.py:223), and one pham shared by 70 genomes."""
    rng = random.Random(5)
    genomes = [Genome("empty")]
    anc = testdata.random_protein(rng, 40, 60)
    for g in range(70):
        x = Genome(f"g{g:02d}")
        x.add("everyone", testdata.mutate(rng, anc, 0.3))
        if g % 7 == 0:
            fam = testdata.random_protein(rng, 25, 45)
            copies = [testdata.mutate(rng, fam, 0.2) for _ in range(3)]
            for s in copies + copies[:2] + [copies[0]]:      # 6 copies, duplicates included
                x.add("family", s)
        if g % 5 == 0:
            x.add("pairwise", anc[:30])
            x.add("pairwise", anc[:30])
        genomes.append(x)
    genomes.append(Genome("empty2"))
    _load(gpu_ctx, pack_genomes(genomes))
    n = len(genomes)
    for key in ALL:
        got = gpu_ctx.matrix(key, as_distance=True).astype(np.int64)
        want = _oracle_triangle(genomes, key, True)
        tol = 0 if key in CHEAP else FLOAT_TOL_MICRO
        assert np.abs(got - want).max() <= tol, key
    hit, length = gpu_ctx.numerators("aai")
    k = 0
    for i in range(n):
        for j in range(i, n):
            if i != j and (i % 9 == 1 or j == n - 1 or i == 0):
                assert (hit[k], length[k]) == po.aai_parts(genomes[i], genomes[j]), (i, j)
            k += 1


def test_ragged_sizes_cross_tile_boundaries(gpu_ctx):
    """N and P that are not multiples of the 128-pair / 1024-pham tiles."""
    for n, p, seed in ((130, 1500, 5), (257, 40000, 6), (1, 10, 7), (129, 33, 8)):
        table = synthetic.sparse_table(n, max(p, 400), seed=seed, clusters=4, mean_phams=30)
        gpu_ctx.load_table(table)
        for key in CHEAP:
            want_num = po.table_numerators(table, key)
            got_num = gpu_ctx.numerators(key)
            iu = np.triu_indices(n)
            want_tri = want_num[iu].copy()
            want_tri[iu[0] == iu[1]] = 0
            assert np.array_equal(got_num, want_tri), (n, p, key)
        for key in CHEAP:
            for as_distance in (True, False):
                assert np.array_equal(gpu_ctx.matrix(key, as_distance=as_distance),
                                      po.table_micro(table, key, as_distance)), (n, key)


# ----------------------------------------------------------------------------- properties at size
def test_checksums_at_benchmark_size_c3(gpu_ctx):
    """10k genomes x 60k phams (config 3): sum over pairs of each numerator equals
    the closed form over phams (a checksum of checksums), and a sample of rows
    equals the array oracle."""
    table = synthetic.sparse_table(10000, 60000, seed=1)
    gpu_ctx.load_table(table)
    n = table.n_genomes
    members = np.bincount(table.pham_id, minlength=table.n_phams).astype(np.int64)
    csum = np.bincount(table.pham_id, weights=table.copies, minlength=table.n_phams).astype(np.int64)
    shared = gpu_ctx.numerators("gcs").astype(np.int64)
    assert shared.sum() == (members * (members - 1) // 2).sum()
    pocp = gpu_ctx.numerators("pocp").astype(np.int64)
    assert pocp.sum() == ((members - 1) * csum).sum()
    # 50 random FULL rows (plus the first and last) of the shared counts and of the pocp numerators
    # against the sparse-product oracle, entry for entry
    from scipy.sparse import csr_matrix
    ones = np.ones(len(table.pham_id), dtype=np.int64)
    b = csr_matrix((ones, table.pham_id, table.genome_ptr), shape=(n, table.n_phams))
    c = csr_matrix((table.copies.astype(np.int64), table.pham_id, table.genome_ptr), shape=(n, table.n_phams))
    rows = sorted(set(np.random.default_rng(50).choice(n, size=50, replace=False).tolist()) | {0, n - 1})
    want_shared = np.asarray((b[rows] @ b.T).todense())
    want_pocp = np.asarray((c[rows] @ b.T).todense()) + np.asarray((b[rows] @ c.T).todense())
    for k, i in enumerate(rows):
        lo = i * n - (i * (i - 1)) // 2
        assert np.array_equal(shared[lo + 1: lo + (n - i)], want_shared[k, i + 1:]), i
        assert np.array_equal(pocp[lo + 1: lo + (n - i)], want_pocp[k, i + 1:]), i
    jc = gpu_ctx.matrix("jc").astype(np.int64)
    gcs = gpu_ctx.matrix("gcs").astype(np.int64)
    assert jc.min() >= 0 and jc.max() <= 1000000 and (jc >= gcs).all()   # jc sim <= gcs sim
    # the pipelined host path (row blocks shipped while the next is computed) equals
    # the single-launch device path
    import torch
    out = torch.empty(tri_len(n), dtype=torch.int32, device=torch.device("cuda", gpu_ctx.device))
    gpu_ctx.matrix_rows("jc", True, 0, n, out_device_ptr=out.data_ptr())
    gpu_ctx.sync()
    assert np.array_equal(out.cpu().numpy().astype(np.int64), jc)


def test_peq_properties_on_phage_table(gpu_ctx):
    """Config-2-like table: duplicates align perfectly, results stay in range,
    peq == round(af)*round(aai), and a subsample equals the oracle."""
    table = synthetic.phage_table(150, seed=42)
    genomes = unpack_genomes(table)
    genomes.append(genomes[0].__class__("copy_of_0"))
    genomes[-1].phams = {k: list(v) for k, v in genomes[0].phams.items()}
    full = pack_genomes(genomes)
    _load(gpu_ctx, full)
    n = len(genomes)
    peq = gpu_ctx.matrix("peq", as_distance=False).astype(np.int64)
    af = gpu_ctx.matrix("af", as_distance=False).astype(np.int64)
    aai = gpu_ctx.matrix("aai", as_distance=False).astype(np.int64)
    assert peq.min() >= 0 and peq.max() <= 1000000
    expect = np.array([po.round6_micro(round((a / 1e6) * (b / 1e6), 6)) for a, b in zip(af, aai)])
    assert np.array_equal(peq, expect)
    assert peq[(n - 1)] == 1000000                     # genome 0 vs its copy
    dist = gpu_ctx.matrix("peq", as_distance=True).astype(np.int64)
    assert np.array_equal(dist, np.array([po.round6_micro(round(1.0 - (a / 1e6) * (b / 1e6), 6))
                                          for a, b in zip(af, aai)]))
    # oracle on a sub-sample of pairs (C NW on the host)
    hit, length = gpu_ctx.numerators("aai")
    rng = random.Random(1)
    for _ in range(25):
        i = rng.randrange(n - 1)
        j = rng.randrange(i + 1, n)
        k = i * n - (i * (i - 1)) // 2 + (j - i)
        assert (hit[k], length[k]) == po.aai_parts(genomes[i], genomes[j])


def test_aai_sums_at_full_c2_size_against_the_oracle(gpu_ctx):
    """The benchmark's own workload (BASELINE.json configs[1] stand-in: 2000 genomes, 3.1 M distinct
    alignments through LEX, rebased LEX and classic kernels): the integer sums behind aai/peq --
    (sum of identical columns, sum of alignment lengths) per genome pair -- for 200 pairs, half of
    them inside a phage cluster (dozens of shared phams each), against the oracle's C traceback."""
    table = synthetic.phage_table(2000, seed=42)
    _load(gpu_ctx, table)
    n = table.n_genomes
    hit, length = gpu_ctx.numerators("aai")
    stats = gpu_ctx.stats()
    assert stats["lexr_cells"] > 0 and stats["lex_cells"] > 0
    genomes = unpack_genomes(table)
    rng = random.Random(2000)
    pairs = []
    for _ in range(100):
        i = rng.randrange(n - 1)
        pairs.append((i, rng.randrange(i + 1, n)))
    for _ in range(100):                                  # same cluster of 50 consecutive genomes
        c = rng.randrange(n // 50)
        i, j = sorted(rng.sample(range(50 * c, 50 * c + 50), 2))
        pairs.append((i, j))
    shared_pairs = 0
    for i, j in pairs:
        k = i * n - (i * (i - 1)) // 2 + (j - i)
        want = po.aai_parts(genomes[i], genomes[j])
        assert (int(hit[k]), int(length[k])) == want, (i, j)
        shared_pairs += want[1] > 0
    assert shared_pairs >= 100


# ----------------------------------------------------------------------------- sharding on one GPU
def test_row_blocks_and_pham_partitions_recompose(gpu_ctx):
    import torch
    from phamclust_b200.distributed import row_partition
    rows = testdata.genome_rows(seed=51, n_genomes=11, n_phams=13, paralog_p=0.3)
    genomes = testdata.rows_to_genomes(rows, cls=Genome)
    table = pack_genomes(genomes)
    _load(gpu_ctx, table)
    n = len(genomes)
    for key in ("gcs", "af"):
        full = gpu_ctx.matrix(key)
        bounds = row_partition(n, 3)
        parts = [gpu_ctx.matrix_rows(key, True, bounds[r], bounds[r + 1]) for r in range(3)]
        assert np.array_equal(np.concatenate(parts), full)
    dev = torch.device("cuda", gpu_ctx.device)
    total = torch.zeros(tri_len(n), dtype=torch.int64, device=dev)
    for part in range(3):
        acc = torch.zeros(tri_len(n), dtype=torch.int64, device=dev)
        gpu_ctx.align_accumulate(acc.data_ptr(), part=part, n_parts=3)
        gpu_ctx.sync()
        total += acc
    torch.cuda.synchronize()
    for key in ("aai", "peq"):
        full = gpu_ctx.matrix(key)
        bounds = row_partition(n, 2)
        parts = [gpu_ctx.matrix_rows(key, True, bounds[r], bounds[r + 1], acc_ptr=total.data_ptr())
                 for r in range(2)]
        assert np.array_equal(np.concatenate(parts), full)
        out = torch.empty(tri_len(n), dtype=torch.int32, device=dev)
        gpu_ctx.matrix_rows(key, True, 0, n, acc_ptr=total.data_ptr(), out_device_ptr=out.data_ptr())
        gpu_ctx.sync()
        assert np.array_equal(out.cpu().numpy(), full)


def test_identical_sequences_are_aligned_once(gpu_ctx):
    """Duplicate-rich table: 40 genomes drawing their genes from 3 variants per
    pham -> few distinct ordered sequence pairs, same integers as the oracle."""
    rng = random.Random(9)
    variants = {f"ph{p}": [testdata.random_protein(rng, 60, 90)] for p in range(6)}
    for p, v in variants.items():
        v += [testdata.mutate(rng, v[0], 0.15), testdata.mutate(rng, v[0], 0.3)]
    genomes = []
    for g in range(40):
        x = Genome(f"g{g:02d}")
        for p, v in variants.items():
            if rng.random() < 0.8:
                x.add(p, rng.choice(v))
                if rng.random() < 0.2:
                    x.add(p, rng.choice(v))
        genomes.append(x)
    _load(gpu_ctx, pack_genomes(genomes))
    gpu_ctx.reset_counters()
    hit, length = gpu_ctx.numerators("aai")
    c, st = gpu_ctx.counters(), gpu_ctx.stats()
    assert st["unique_genes"] <= 18 and st["genes"] > 150
    assert c["alignments"] <= 6 * 9 and st["logical_alignments"] > 2000
    n, k = len(genomes), 0
    for i in range(n):
        for j in range(i, n):
            if i != j and (i + j) % 5 == 0:
                assert (hit[k], length[k]) == po.aai_parts(genomes[i], genomes[j]), (i, j)
            k += 1
    got = gpu_ctx.matrix("peq").astype(np.int64)
    assert np.abs(got - _oracle_triangle(genomes, "peq", True)).max() <= FLOAT_TOL_MICRO


def test_counters_report_kernel_launches(gpu_ctx):
    table = synthetic.phage_table(40, seed=3)
    _load(gpu_ctx, table)
    gpu_ctx.reset_counters()
    gpu_ctx.matrix("peq")
    c = gpu_ctx.counters()
    t = gpu_ctx.timings()
    assert c["launches"] >= 4 and c["align_launches"] >= 1 and c["contract_launches"] >= 1
    assert c["alignments"] > 0 and c["dp_cells"] > 0 and t["align"] > 0.0


def test_host_gather_through_the_shared_triangle(gpu_ctx):
    """distributed.sharded_matrix_to_host on a 1-rank group: the row block is computed on
    the GPU and DMA'd by the library into page-locked POSIX shared memory."""
    import socket
    import torch.distributed as dist
    from phamclust_b200 import distributed
    table = synthetic.phage_table(120, seed=9)
    _load(gpu_ctx, table)
    want = {m: gpu_ctx.matrix(m, as_distance=True) for m in ("jc", "peq")}
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=0, world_size=1)
    try:
        shared = distributed.SharedTriangle(table.n_genomes)
        try:
            assert shared._registered is not None
            for m in ("jc", "peq"):
                got = distributed.sharded_matrix_to_host(table, m, shared, as_distance=True)
                assert np.array_equal(got, want[m]), m
        finally:
            shared.close()
    finally:
        dist.destroy_process_group()


def test_threaded_table_load(gpu_ctx, monkeypatch):
    """pc_load_table validates and derives the CSR/CSC bookkeeping on several host threads:
    same operand columns and same matrix as the single-thread pass, and of two bad entries
    the one with the lower index is reported whichever thread meets it."""
    import copy
    table = synthetic.sparse_table(3000, 20000, seed=5)
    results = []
    for threads in ("1", "7"):
        monkeypatch.setenv("PC_HOST_THREADS", threads)
        gpu_ctx.load_table(table)
        results.append((gpu_ctx.stats()["operand_columns"], gpu_ctx.matrix("pocp"), gpu_ctx.numerators("af")))
    assert results[0][0] == results[1][0]
    assert np.array_equal(results[0][1], results[1][1]) and np.array_equal(results[0][2], results[1][2])
    bad = copy.copy(table)
    bad.pham_id = table.pham_id.copy()
    nnz = len(bad.pham_id)
    lo, hi = nnz // 5, nnz - 7
    bad.pham_id[lo] = table.meta["n_phams"] + 3                 # out of range, early
    g_hi = int(np.searchsorted(table.genome_ptr, hi, side="right") - 1)
    if hi == table.genome_ptr[g_hi]:
        hi += 1
    bad.pham_id[hi] = bad.pham_id[hi - 1]                       # listed twice, late
    for threads in ("1", "7"):
        monkeypatch.setenv("PC_HOST_THREADS", threads)
        with pytest.raises(ValueError, match=f"out of range at entry {lo}"):
            gpu_ctx.load_table(bad)
        with pytest.raises(ValueError):                          # no table left half-loaded
            gpu_ctx.matrix("gcs")
    monkeypatch.delenv("PC_HOST_THREADS")
    gpu_ctx.load_table(table)
    assert np.array_equal(gpu_ctx.matrix("pocp"), results[0][1])
