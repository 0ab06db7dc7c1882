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
