"""The LEX DP word of nw_lex_kernel, restated on the CPU (oracle/lex_model.py), against the
explicit-traceback oracle: the one-word max-plus recurrence with priorities, drift and packed
counters gives parasail's (identical columns, length, score) for every alignment its admission
test lets in -- including the ones at the edge of the 12-bit score field."""
import random

from oracle import lex_model, nw_c, testdata


def _check_pair(q, r):
    ident, similar, length, score = nw_c.forward_stats(q, r)
    assert lex_model.lex_stats(q, r) == (ident, length, score), (q, r)
    assert lex_model.lex_stats(q, r, ppos=True) == (ident + similar, length, score), (q, r)


def test_lex_word_equals_the_traceback_oracle():
    rng = random.Random(20)
    for t in range(160):
        alpha = ["AG", "AGS", "WC", "NDB", testdata.AA][t % 5]
        a = testdata.random_protein(rng, 0 if t % 17 == 0 else 1, 70, alpha)
        if t % 2 and a:
            b = testdata.mutate(rng, a, rng.random() * 0.6, alpha)
        else:
            b = testdata.random_protein(rng, 0 if t % 19 == 0 else 1, 70, alpha)
        assert lex_model.lex_ok(a, b)
        _check_pair(a, b)
    _check_pair("mktwACDE", "MKTWacde")
    _check_pair("MKUXO*BZ", "MKOXU*ZB")


def test_admission_bound_keeps_every_real_cell_inside_the_score_field():
    """All-W sequences score 11 per residue, the steepest climb there is: the longest pair the
    admission test lets in must keep every H word of a real cell inside the signed 12-bit score
    field (no wrap of anything that can flow to the answer), the next longer one is refused."""
    k = 1
    while lex_model.lex_ok("W" * (k + 1), "W" * (k + 1)):
        k += 1
    assert not lex_model.lex_ok("W" * (k + 1), "W" * (k + 1))
    lo, hi = -(1 << (lex_model.SCORE_BITS - 1)), (1 << (lex_model.SCORE_BITS - 1)) - 1
    seen = []

    def check(word):
        seen.append(word >> lex_model.SHIFT)

    for q, r in (("W" * k, "W" * k), ("W" * k, "A" * k), ("W" * k, "W" * (k // 3)), ("C" * k, "W" * 5)):
        assert lex_model.lex_ok(q, r)
        seen.clear()
        got = lex_model.lex_stats(q, r, check=check)
        assert lo < min(seen) and max(seen) < hi, (min(seen), max(seen))
        ident, similar, length, score = nw_c.forward_stats(q, r)
        assert got == (ident, length, score)
    # counters: min length 511 is the last one admitted (9-bit fields)
    assert lex_model.lex_ok("A" * 511, "G" * 600) and not lex_model.lex_ok("A" * 512, "G" * 600)


def test_rebased_lex_word_equals_the_traceback_oracle():
    """LEXR (nw_lex_kernel<..., B = 10, REBASE>): 10-bit counters next to a 10-bit score field that
    every lane re-centres every REBASE row-steps.  The model runs the kernel's lane schedule; it
    must give the oracle's (identical, length, score) and never leave the field's window -- on
    tie-rich, ordinary and adversarial inputs (all-W sequences climb 11 per residue, the steepest
    slope the Lipschitz bound PC_NW_LEXR_LIP allows for)."""
    rng = random.Random(1010)
    half = 1 << (32 - (2 * 10 + 2) - 1)
    for G, K, R in ((32, 16, 8), (32, 24, 8), (32, 26, 4), (32, 32, 4)):
        assert lex_model.lexr_window_ok(10, K, R)
    assert not lex_model.lexr_window_ok(10, 26, 8) and not lex_model.lexr_window_ok(10, 34, 4)
    cases = []
    for t in range(30):
        alpha = ["AG", "AGS", "WC", "NDB", testdata.AA][t % 5]
        a = testdata.random_protein(rng, 1, 70, alpha)
        b = testdata.mutate(rng, a, rng.random() * 0.6, alpha) if t % 2 else testdata.random_protein(rng, 1, 70, alpha)
        cases.append((a, b, 4, 24, 8))
        cases.append((a, b, 8, 16, 8))
    long_a = testdata.random_protein(rng, 640, 640)
    cases.append((long_a, testdata.mutate(rng, long_a, 0.35), 32, 22, 8))
    cases += [("W" * 300, "W" * 300, 32, 24, 8), ("W" * 300, "A" * 310, 32, 24, 8), ("W" * 330, "W" * 90, 32, 32, 4),
              ("WA" * 150, "AW" * 150, 32, 26, 4), ("mktwACDE", "MKTWacde", 4, 16, 8)]
    for q, r, G, K, R in cases:
        if len(r) > G * K:
            continue
        ident, similar, length, score = nw_c.forward_stats(q, r)
        got, worst = lex_model.lexr_stats(q, r, G, K, R, bits=10)
        assert got == (ident, length, score), (q, r, G, K, R)
        assert worst < half - 8, (worst, G, K, R)
    got, _ = lex_model.lexr_stats(long_a[:200], long_a[5:190], 32, 16, 8, bits=10, ppos=True)
    ident, similar, length, score = nw_c.forward_stats(long_a[:200], long_a[5:190])
    assert got == (ident + similar, length, score)


def test_lexc_model_equals_the_oracle_or_raises_its_flag():
    """LEXC (nw_lexc_kernel): score AND counters re-centred per lane, column chunks, band checks.
    oracle/lexc_model.c executes the kernel's arithmetic word for word.  Whatever the input, the
    model either reports the oracle's (identical, length, score) or raises the fallback flag --
    and it never sees a counter field borrow or carry while the flag is still clear (``broke``),
    which is the induction the kernel's exactness rests on."""
    from oracle import lexc_c
    rng = random.Random(1536)
    cases = []
    for n in (1024, 1025, 1100, 1279, 1280, 1281, 1536, 1537, 2049, 2305):
        a = testdata.random_protein(rng, n, n, testdata.AA)
        cases.append((a, testdata.mutate(rng, a, 0.05), False))
        cases.append((testdata.mutate(rng, a, 0.3), a, False))
    ag = testdata.random_protein(rng, 1200, 1200, "AG")
    cases += [(ag, testdata.mutate(rng, ag, 0.1, "AG"), False), ("W" * 1100, "W" * 1107, False),
              ("A" * 1030, "A" * 1027, False), ("AG" * 550, "GA" * 550 + "A", False)]
    a, b = testdata.random_protein(rng, 700, 700, testdata.AA), testdata.random_protein(rng, 700, 700, testdata.AA)
    c = testdata.random_protein(rng, 1400, 1400, testdata.AA)
    cases += [(a + b, b + a, True), (c, c[:200] + c[500:] + c[200:500], True),
              (c, testdata.random_protein(rng, 1440, 1440, testdata.AA), None),
              (c[:300], c, None), (c, c[:300], None)]            # far outside MAXDIFF: flagged or exact
    flagged = 0
    for q, r, expect_flag in cases:
        ident, similar, length, score = nw_c.forward_stats(q, r)
        for ppos in (False, True):
            for K in (lexc_c.choose_k(len(q), len(r)), 16, 24):
                got = lexc_c.align(q, r, K=K, ppos=ppos)
                assert not (got["broke"] and not got["flagged"]), (len(q), len(r), K, got)
                if expect_flag is not None:
                    assert bool(got["flagged"]) == expect_flag, (len(q), len(r), K, got)
                if got["flagged"]:
                    flagged += 1
                    continue
                assert (got["matched"], got["columns"], got["score"]) == (
                    ident + (similar if ppos else 0), length, score), (len(q), len(r), K, got)
                assert got["field_dev"] <= lexc_c.BAND + lexc_c.REBASE and got["frame_dev"] <= 256
                assert got["score_field"] < 512
    assert flagged >= 12


def test_lexc_model_small_and_degenerate_inputs():
    """Short sequences run through the same code (one chunk, most lanes padding only)."""
    from oracle import lexc_c
    rng = random.Random(5)
    for t in range(60):
        alpha = ["AG", "WC", testdata.AA][t % 3]
        a = testdata.random_protein(rng, 1, 90, alpha)
        b = testdata.mutate(rng, a, 0.3, alpha) if t % 2 else testdata.random_protein(rng, 1, 90, alpha)
        ident, similar, length, score = nw_c.forward_stats(a, b)
        got = lexc_c.align(a, b, K=16)
        if not got["flagged"]:
            assert (got["matched"], got["columns"], got["score"]) == (ident, length, score), (a, b, got)
        assert not (got["broke"] and not got["flagged"])
    got = lexc_c.align("MKT", "", K=16)
    assert (got["matched"], got["columns"], got["score"], got["flagged"]) == (0, 3, -13, 0)
