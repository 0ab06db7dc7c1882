"""The alignment oracle (restated parasail, PARITY UNPINNED at that boundary):
what CAN be pinned is pinned here -- SURVEY.md 8(c) items (ii)-(iv)."""
import random

import pytest

from conftest import load_golden
from oracle import nw_c, nw_py, testdata


def _pairs(n, seed, max_len=70):
    rng = random.Random(seed)
    for t in range(n):
        alpha = ["AG", "AGS", testdata.AA][t % 3]
        a = testdata.random_protein(rng, 0 if t % 7 == 0 else 1, max_len, alpha)
        if t % 2:
            b = testdata.mutate(rng, a, rng.random() * 0.6, alpha) if a else ""
        else:
            b = testdata.random_protein(rng, 0 if t % 11 == 0 else 1, max_len, alpha)
        yield a, b


def test_c_and_python_twins_agree():
    for a, b in _pairs(400, 1):
        assert nw_c.traceback(a, b)[3] == nw_py.nw_stats(a, b)


def test_score_is_optimal_against_independent_gotoh():
    for a, b in _pairs(1500, 2):
        assert nw_c.traceback(a, b)[3][3] == nw_c.gotoh_score(a, b)


def test_traceback_is_a_valid_alignment_that_rescores_to_the_dp_score():
    for a, b in _pairs(1500, 3):
        q, c, r, stats = nw_c.traceback(a, b)
        assert q.replace("-", "") == a and r.replace("-", "") == b
        assert len(q) == len(c) == len(r) == stats[2]
        assert nw_py.rescore(q, r) == stats[3]
        assert c.count("|") == stats[0] and c.count("+") == stats[1]
        assert all(x == " " for x, y, z in zip(c, q, r) if y == "-" or z == "-")


def test_forward_counters_equal_explicit_traceback():
    # the formulation the CUDA kernel implements
    for a, b in _pairs(3000, 4, max_len=90):
        assert nw_c.forward_stats(a, b) == nw_c.traceback(a, b)[3]


def test_long_sequences_forward_equals_traceback():
    rng = random.Random(5)
    for _ in range(5):
        a = testdata.random_protein(rng, 600, 1400)
        b = testdata.mutate(rng, a, 0.35)
        assert nw_c.forward_stats(a, b) == nw_c.traceback(a, b)[3]
        assert nw_c.traceback(a, b)[3][3] == nw_c.gotoh_score(a, b)


def test_case_insensitive_match_and_unknown_symbols():
    # U5/U6 of the survey's uncertainty register, as restated
    assert nw_c.traceback("mktw", "MKTW")[3][:3] == (4, 0, 4)
    # U and O are outside the alphabet: both map to '*' ('*' vs '*' scores +1),
    # but they are different characters, so the column is '+', not '|'
    q, c, r, stats = nw_c.traceback("MKUW", "MKOW")
    assert c == "||+|" and stats[:3] == (3, 1, 4)
    assert nw_c.lib().nw_map_char(ord("U")) == 23 and nw_c.lib().nw_map_char(ord("a")) == 0


def test_golden_alignment_vectors():
    for case in load_golden("alignments.json"):
        got = nw_c.forward_stats(case["query"], case["ref"])
        assert got == (case["ident"], case["similar"], case["length"], case["score"])
        if len(case["query"]) < 80:
            assert nw_py.nw_stats(case["query"], case["ref"]) == got


def test_parasail_stub_traceback_shape():
    import sys
    import pathlib
    stub = pathlib.Path(__file__).resolve().parent.parent / "oracle" / "parasail_stub"
    sys.path.insert(0, str(stub))
    try:
        import parasail
        tb = parasail.nw_trace_diag_16("MKTAA", "MKT", 11, 1, parasail.blosum62).get_traceback(
            mch="|", sim="+", neg=" ")
        assert tb.query == "MKTAA" and tb.ref == "MKT--" and tb.comp == "|||  "
    finally:
        sys.path.remove(str(stub))


# ---- how much of the result the unpinned tie rules can move ------------------------
def _all_alignments(a, b):
    """Every global alignment of a and b as (query, ref) strings with '-' (no
    column of two gaps): brute force, tiny inputs only."""
    if not a and not b:
        yield "", ""
        return
    if a and b:
        for q, r in _all_alignments(a[1:], b[1:]):
            yield a[0] + q, b[0] + r
    if a:
        for q, r in _all_alignments(a[1:], b):
            yield a[0] + q, "-" + r
    if b:
        for q, r in _all_alignments(a, b[1:]):
            yield "-" + q, b[0] + r


def test_tie_range_equals_brute_force_over_all_cooptimal_alignments():
    """nw_tie_range (min/max of ident, similar, length over EVERY co-optimal alignment of
    the score model) against explicit enumeration; and the restated tie rules always pick
    one of those alignments."""
    rng = random.Random(6)
    seen_ties = 0
    for t in range(300):
        alpha = ["AG", "AGS", "AS", "AGWC", "NDB", testdata.AA][t % 6]
        a = testdata.random_protein(rng, 0 if t % 9 == 0 else 1, 7, alpha)
        b = testdata.random_protein(rng, 0 if t % 13 == 0 else 1, 5, alpha)
        best, stats = None, set()
        for q, r in _all_alignments(a, b):
            score = nw_py.rescore(q, r)
            if best is None or score > best:
                best, stats = score, set()
            if score == best:
                ident = sum(1 for x, y in zip(q, r) if x == y and x != "-")
                similar = sum(1 for x, y in zip(q, r) if x != y and "-" not in (x, y)
                              and nw_c.lib().nw_blosum62(nw_c.lib().nw_map_char(ord(x)),
                                                         nw_c.lib().nw_map_char(ord(y))) > 0)
                stats.add((ident, similar, len(q)))
        got = nw_c.tie_range(a, b)
        assert got[6] == best
        assert (got[0], got[1]) == (min(s[0] for s in stats), max(s[0] for s in stats))
        assert (got[2], got[3]) == (min(s[1] for s in stats), max(s[1] for s in stats))
        assert (got[4], got[5]) == (min(s[2] for s in stats), max(s[2] for s in stats))
        assert nw_c.forward_stats(a, b)[:3] in stats          # the restated rules are co-optimal
        seen_ties += len(stats) > 1
    assert seen_ties >= 3           # the sample does exercise tie-dependent cases


def test_restated_rules_land_inside_the_tie_range_on_protein_sized_inputs():
    for a, b in _pairs(600, 8, max_len=120):
        lo_hi = nw_c.tie_range(a, b)
        ident, similar, length, score = nw_c.forward_stats(a, b)
        assert score == lo_hi[6]
        assert lo_hi[0] <= ident <= lo_hi[1] and lo_hi[2] <= similar <= lo_hi[3]
        assert lo_hi[4] <= length <= lo_hi[5]
