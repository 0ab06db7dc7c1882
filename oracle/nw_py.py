"""oracle/nw_py.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Pure-Python twin of oracle/nw_oracle.c: parasail's
``nw_trace_diag_16(s1, s2, open, extend, blosum62).get_traceback(mch, sim, neg)``
as stated in SURVEY.md Appendix A (reference call sites
/root/reference/src/phamclust/metrics.py:160-175).

*** PARITY UNPINNED AT THE PARASAIL BOUNDARY *** -- parasail (PyPI
``parasail~=1.3.0``, requirements.txt:5) is a third-party wheel that is neither
in /root/reference nor installable offline, and the reference holds no golden
vectors; this is a restatement of its published algorithm.  Small cases only
(pure-Python loops); the C twin is used for anything larger.
"""

ALPHABET = "ARNDCQEGHILKMFPSTWYVBZX*"

_ROWS = """
 4 -1 -2 -2  0 -1 -1  0 -2 -1 -1 -1 -1 -2 -1  1  0 -3 -2  0 -2 -1  0 -4
-1  5  0 -2 -3  1  0 -2  0 -3 -2  2 -1 -3 -2 -1 -1 -3 -2 -3 -1  0 -1 -4
-2  0  6  1 -3  0  0  0  1 -3 -3  0 -2 -3 -2  1  0 -4 -2 -3  3  0 -1 -4
-2 -2  1  6 -3  0  2 -1 -1 -3 -4 -1 -3 -3 -1  0 -1 -4 -3 -3  4  1 -1 -4
 0 -3 -3 -3  9 -3 -4 -3 -3 -1 -1 -3 -1 -2 -3 -1 -1 -2 -2 -1 -3 -3 -2 -4
-1  1  0  0 -3  5  2 -2  0 -3 -2  1  0 -3 -1  0 -1 -2 -1 -2  0  3 -1 -4
-1  0  0  2 -4  2  5 -2  0 -3 -3  1 -2 -3 -1  0 -1 -3 -2 -2  1  4 -1 -4
 0 -2  0 -1 -3 -2 -2  6 -2 -4 -4 -2 -3 -3 -2  0 -2 -2 -3 -3 -1 -2 -1 -4
-2  0  1 -1 -3  0  0 -2  8 -3 -3 -1 -2 -1 -2 -1 -2 -2  2 -3  0  0 -1 -4
-1 -3 -3 -3 -1 -3 -3 -4 -3  4  2 -3  1  0 -3 -2 -1 -3 -1  3 -3 -3 -1 -4
-1 -2 -3 -4 -1 -2 -3 -4 -3  2  4 -2  2  0 -3 -2 -1 -2 -1  1 -4 -3 -1 -4
-1  2  0 -1 -3  1  1 -2 -1 -3 -2  5 -1 -3 -1  0 -1 -3 -2 -2  0  1 -1 -4
-1 -1 -2 -3 -1  0 -2 -3 -2  1  2 -1  5  0 -2 -1 -1 -1 -1  1 -3 -1 -1 -4
-2 -3 -3 -3 -2 -3 -3 -3 -1  0  0 -3  0  6 -4 -2 -2  1  3 -1 -3 -3 -1 -4
-1 -2 -2 -1 -3 -1 -1 -2 -2 -3 -3 -1 -2 -4  7 -1 -1 -4 -3 -2 -2 -1 -2 -4
 1 -1  1  0 -1  0  0  0 -1 -2 -2  0 -1 -2 -1  4  1 -3 -2 -2  0  0  0 -4
 0 -1  0 -1 -1 -1 -1 -2 -2 -1 -1 -1 -1 -2 -1  1  5 -2 -2  0 -1 -1  0 -4
-3 -3 -4 -4 -2 -2 -3 -2 -2 -3 -2 -3 -1  1 -4 -3 -2 11  2 -3 -4 -3 -2 -4
-2 -2 -2 -3 -2 -1 -2 -3  2 -1 -1 -2 -1  3 -3 -2 -2  2  7 -1 -3 -2 -1 -4
 0 -3 -3 -3 -1 -2 -2 -3 -3  3  1 -2  1 -1 -2 -2  0 -3 -1  4 -3 -2 -1 -4
-2 -1  3  4 -3  0  1 -1  0 -3 -4  0 -3 -3 -2  0 -1 -4 -3 -3  4  1 -1 -4
-1  0  0  1 -3  3  4 -2  0 -3 -3  1 -1 -3 -1  0 -1 -3 -2 -2  1  4 -1 -4
 0 -1 -1 -1 -2 -1 -1 -1 -1 -1 -1 -1 -1 -1 -2  0  0 -2 -1 -1 -1 -1 -1 -4
-4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4  1
"""
BLOSUM62 = [[int(x) for x in line.split()] for line in _ROWS.strip().splitlines()]

NEG_INF = -(1 << 28)
# rule switches, same meaning as in nw_oracle.c
GAP_TIE_EXTENDS = True          # U3
H_PRIORITY_DIAG_DEL_INS = True  # U2
CASE_SENSITIVE = False          # U5


def map_char(ch):
    """U6: characters outside the matrix alphabet map to the '*' column."""
    k = ALPHABET.find(ch.upper())
    return k if k >= 0 and len(ch) == 1 else 23


def _match(a, b):
    return a == b if CASE_SENSITIVE else a.upper() == b.upper()


def nw_traceback(s1, s2, gap_open=11, gap_extend=1, mch="|", sim="+", neg=" "):
    """Return (query, comp, ref, score) by explicit trace table + traceback."""
    m, n = len(s1), len(s2)
    i1 = [map_char(c) for c in s1]
    i2 = [map_char(c) for c in s2]
    DIAG, DEL, INS, DIAG_F, DIAG_E = 1, 2, 4, 8, 16
    H = [0] + [-gap_open - (j - 1) * gap_extend for j in range(1, n + 1)]
    F = [NEG_INF] * (n + 1)
    trace = [[0] * n for _ in range(m)]
    for i in range(1, m + 1):
        row = BLOSUM62[i1[i - 1]]
        north = H[0]
        west = -gap_open - (i - 1) * gap_extend
        e = NEG_INF
        H[0] = west
        for j in range(1, n + 1):
            northwest = north
            north = H[j]
            f_opn, f_ext = north - gap_open, F[j] - gap_extend
            e_opn, e_ext = west - gap_open, e - gap_extend
            F[j] = max(f_opn, f_ext)
            e = max(e_opn, e_ext)
            if GAP_TIE_EXTENDS:
                t = (DIAG_F if f_opn > f_ext else 0) | (DIAG_E if e_opn > e_ext else 0)
            else:
                t = (DIAG_F if f_opn >= f_ext else 0) | (DIAG_E if e_opn >= e_ext else 0)
            h_dag = northwest + row[i2[j - 1]]
            west = max(h_dag, F[j], e)
            if west == h_dag:
                t |= DIAG
            elif H_PRIORITY_DIAG_DEL_INS:
                t |= DEL if west == F[j] else INS
            else:
                t |= INS if west == e else DEL
            H[j] = west
            trace[i - 1][j - 1] = t
    if m == 0 and n == 0:
        score = 0
    elif m == 0 or n == 0:
        score = -gap_open - (max(m, n) - 1) * gap_extend
    else:
        score = H[n]
    q, c, r = [], [], []
    i, j, where = m - 1, n - 1, DIAG
    while i >= 0 or j >= 0:
        if i < 0:
            q.append("-"); r.append(s2[j]); c.append(" "); j -= 1
            continue
        if j < 0:
            q.append(s1[i]); r.append("-"); c.append(" "); i -= 1
            continue
        t = trace[i][j]
        if where == DIAG:
            if t & DIAG:
                q.append(s1[i]); r.append(s2[j])
                if _match(s1[i], s2[j]):
                    c.append(mch)
                elif BLOSUM62[i1[i]][i2[j]] > 0:
                    c.append(sim)
                else:
                    c.append(neg)
                i -= 1; j -= 1
            elif t & INS:
                where = INS
            else:
                where = DEL
        elif where == INS:
            q.append("-"); r.append(s2[j]); c.append(" "); j -= 1
            where = DIAG if t & DIAG_E else INS
        else:
            q.append(s1[i]); r.append("-"); c.append(" "); i -= 1
            where = DIAG if t & DIAG_F else DEL
    return "".join(reversed(q)), "".join(reversed(c)), "".join(reversed(r)), score


def nw_stats(s1, s2, gap_open=11, gap_extend=1):
    """(identical, similar, length, score) via the explicit traceback."""
    q, c, _, score = nw_traceback(s1, s2, gap_open, gap_extend)
    return c.count("|"), c.count("+"), len(q), score


def rescore(query, ref, gap_open=11, gap_extend=1):
    """Score of an emitted alignment (pins traceback validity in tests)."""
    total, in_q_gap, in_r_gap = 0, False, False
    for a, b in zip(query, ref):
        if a == "-":
            total -= gap_extend if in_q_gap else gap_open
            in_q_gap, in_r_gap = True, False
        elif b == "-":
            total -= gap_extend if in_r_gap else gap_open
            in_q_gap, in_r_gap = False, True
        else:
            total += BLOSUM62[map_char(a)][map_char(b)]
            in_q_gap = in_r_gap = False
    return total
