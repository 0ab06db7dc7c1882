"""oracle/lex_model.py -- TEST INFRASTRUCTURE: an executable CPU model of the LEX DP word.

``nw_lex_kernel`` (phamclust_b200/csrc/pc_align.cu) keeps score, tie priority, diagonal
steps and matches of a DP state in ONE signed 32-bit word and runs the affine-gap
recurrence of parasail's nw_trace (reference call site
/root/reference/src/phamclust/metrics.py:160-175) as bare max-plus on it.  This file
restates that arithmetic in plain Python, 32-bit wrap-around included, with the constants of
phamclust_b200/csrc/pc_common.cuh (PC_NW_LEX_*), so the formulation and its admission bound
(``lex_ok`` == pc_nw_lex_ok) can be checked against the traceback oracle without a GPU
(tests/test_lex_model.py); the GPU parity tests then check the kernel against the same oracle.
"""
from oracle import nw_py

OPEN, EXT = 11, 1
BITS = 9                                   # PC_NW_LEX_BITS
SHIFT = 2 * BITS + 2                       # PC_NW_LEX_SHIFT
SCORE_BITS = 32 - SHIFT
FLOOR = 40                                 # PC_NW_LEX_FLOOR
U = 1 << SHIFT
CMASK = (1 << BITS) - 1
PRIO_E, PRIO_F, PRIO_D = 1 << (2 * BITS), 2 << (2 * BITS), 3 << (2 * BITS)     # diag > del(F) > ins(E)
PMASK = 3 << (2 * BITS)
BIAS = -(1 << (SCORE_BITS - 1)) + FLOOR
CU = (EXT - OPEN) * U
ORIGIN = BIAS * U
EDGE = (BIAS + EXT - OPEN) * U
NEG = (-(1 << (SCORE_BITS - 1)) + 2) * U


def wrap32(x):
    """Two's-complement wrap of a Python int to a signed 32-bit value."""
    return ((x + (1 << 31)) & 0xFFFFFFFF) - (1 << 31)


def row_max(ch):
    """Best substitution score of a residue against the 24 matrix letters, >= 0."""
    return max(0, max(nw_py.BLOSUM62[nw_py.map_char(ch)]))


def lex_ok(q, r):
    """pc_nw_lex_ok: both counters fit BITS bits and the drifted score range fits SCORE_BITS."""
    m, n = len(q), len(r)
    hi = min(sum(map(row_max, q)), sum(map(row_max, r)))
    return min(m, n) < (1 << BITS) and hi + m + n + FLOOR + 24 < (1 << SCORE_BITS)


def lex_stats(q, r, ppos=False, check=None):
    """(matched columns, alignment length, score) of query q against ref r through the LEX
    word.  ``check(word)`` (tests) sees every H word of a real cell."""
    m, n = len(q), len(r)
    if n == 0:
        return 0, m, (0 if m == 0 else -OPEN - (m - 1) * EXT)
    qi = [nw_py.map_char(c) for c in q]
    ri = [nw_py.map_char(c) for c in r]
    Hp = [EDGE] * n                        # H^[i-1][j], drifted by (i + j) ext
    Fp = [NEG] * n
    for i in range(1, m + 1):
        hd = ORIGIN if i == 1 else EDGE    # H^[i-1][0]
        hl, el = EDGE, NEG                 # H^[i][0], E^[i][0]
        for x in range(n):
            s = nw_py.BLOSUM62[qi[i - 1]][ri[x]]
            same = q[i - 1].upper() == r[x].upper()
            hit = 1 if (same or (ppos and s > 0)) else 0
            tw = (s + 2 * EXT) * U + PRIO_D + (1 << BITS) + hit
            f = max(wrap32(Hp[x] + CU), Fp[x]) | PRIO_F
            e = max(wrap32(hl + CU), el) | PRIO_E
            hdg = wrap32(hd + tw)
            h = max(hdg, f, e) & ~PMASK
            if check:
                check(h)
            hd = Hp[x]
            hl, el = h, e
            Hp[x], Fp[x] = h, f
    word = Hp[n - 1]
    score = (word >> SHIFT) - BIAS - (m + n) * EXT
    matched = word & CMASK
    columns = m + n - ((word >> BITS) & CMASK)
    return matched, columns, score
