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


# ---------------------------------------------------------------------------------------------
# LEXR: the same word with a NARROW score field that every lane keeps centred by rebasing.
#
# DP values are Lipschitz: neighbouring cells differ by at most LIP = max substitution score +
# open + ext (drifted units), and E/F sit within open + LIP below their H.  A lane that owns K
# adjacent columns therefore only ever combines values within LIP * (K/2 + rows since its last
# rebase) + slack of the H of its centre column.  ``nw_lex_kernel<.., REBASE>`` subtracts that centre
# score from all of the lane's state every REBASE row-steps and keeps the running sum (``base``)
# in a separate register, so a 10-bit score field (window +-512) is enough for ANY score range
# and the counters can take 10 bits each: alignments with min(m, n) < 1024.  This model follows
# the kernel's schedule exactly (G lanes x K columns, lane l one row behind lane l-1, rebase at
# the top of every REBASE-th step, values converted between lane frames when they are shuffled)
# and reports the largest |score field| it met, so the window claim is checked, not assumed.
# ---------------------------------------------------------------------------------------------
LIP = 11 + OPEN + EXT                       # max BLOSUM62 entry (W/W) + open + ext


def lexr_params(bits):
    shift = 2 * bits + 2
    return {"bits": bits, "shift": shift, "score_bits": 32 - shift, "U": 1 << shift,
            "cmask": (1 << bits) - 1, "prio_e": 1 << (2 * bits), "prio_f": 2 << (2 * bits),
            "prio_d": 3 << (2 * bits), "pmask": 3 << (2 * bits)}


def lexr_window_ok(bits, K, rebase):
    """The static guarantee the launcher relies on: LIP * (K/2 + rebase) + slack inside the field."""
    half = 1 << (32 - (2 * bits + 2) - 1)
    return LIP * (K // 2 + rebase) + OPEN + LIP + 8 < half


def lexr_stats(q, r, G, K, rebase, bits=10, ppos=False):
    """((matched, columns, score), max |score field| seen) through the rebased LEX word, with the
    kernel's lane schedule.  Raises AssertionError if any intermediate sum wraps 32 bits."""
    P = lexr_params(bits)
    U, SH = P["U"], P["shift"]
    m, n = len(q), len(r)
    if n == 0:
        return (0, m, (0 if m == 0 else -OPEN - (m - 1) * EXT)), 0
    assert n <= G * K
    qi = [nw_py.map_char(c) for c in q]
    ri = [nw_py.map_char(c) for c in r] + [0] * (G * K - n)          # padding columns: code 0
    rch = list(r.upper()) + ["A"] * (G * K - n)
    CUr = (EXT - OPEN) * U
    EDGE0 = (EXT - OPEN) * U
    NEG0 = EDGE0 + CUr - 4 * U                   # below every opening from a boundary cell
    lanes = []
    for l in range(G):
        lanes.append({"Hp": [EDGE0] * K, "Fp": [NEG0] * K, "diag": 0 if l == 0 else EDGE0,
                      "outH": 0, "outE": NEG0, "base": 0})
    worst = 0

    def add(a, b):
        s = a + b
        assert wrap32(s) == s, "32-bit wrap in the rebased LEX word"
        return s

    for s in range(1, m + G):
        if rebase and s % rebase == 0:
            for L in lanes:
                delta = L["Hp"][K // 2] >> SH
                if delta:
                    d = delta * U
                    L["Hp"] = [add(v, -d) for v in L["Hp"]]
                    L["Fp"] = [add(v, -d) for v in L["Fp"]]
                    L["diag"] = add(L["diag"], -d)
                    L["outH"] = add(L["outH"], -d)
                    L["outE"] = add(L["outE"], -d)
                    L["base"] += delta
        shuffled = [(L["outH"], L["outE"], L["base"]) for L in lanes]
        for l, L in enumerate(lanes):
            i = s - l
            inH, inE, inB = (EDGE0, NEG0, 0) if l == 0 else shuffled[l - 1]
            conv = (inB - L["base"]) * U
            inH, inE = add(inH, conv), add(inE, conv)
            if not (1 <= i <= m):
                continue
            hl, el, hd = inH, inE, L["diag"]
            L["diag"] = inH
            for x in range(K):
                j = l * K + x
                sc = nw_py.BLOSUM62[qi[i - 1]][ri[j]]
                same = q[i - 1].upper() == rch[j]
                hit = 1 if (same or (ppos and sc > 0)) else 0
                tw = (sc + 2 * EXT) * U + P["prio_d"] + (1 << bits) + hit
                f = max(add(L["Hp"][x], CUr), L["Fp"][x]) | P["prio_f"]
                e = max(add(hl, CUr), el) | P["prio_e"]
                hdg = add(hd, tw)
                h = max(hdg, f, e) & ~P["pmask"]
                if j < n:
                    worst = max(worst, abs(h >> SH), abs(f >> SH), abs(e >> SH))
                hd = L["Hp"][x]
                hl, el = h, e
                L["Hp"][x], L["Fp"][x] = h, f
            L["outH"], L["outE"] = hl, el
    owner = lanes[(n - 1) // K]
    word = owner["Hp"][(n - 1) % K]
    score = (word >> SH) + owner["base"] - (m + n) * EXT
    matched = word & P["cmask"]
    columns = m + n - ((word >> bits) & P["cmask"])
    return (matched, columns, score), worst
