"""oracle/parasail_stub/parasail.py -- TEST INFRASTRUCTURE.

A module NAMED ``parasail`` exposing exactly the two names PhamClust imports
(/root/reference/src/phamclust/metrics.py:20): ``blosum62`` and
``nw_trace_diag_16``.  Put this directory first on sys.path and the VERBATIM
reference ``metrics.py`` imports and runs; the alignment arithmetic is the
CPU restatement in oracle/nw_oracle.c (or oracle/nw_py.py when
``PHAMCLUST_ORACLE_PURE_PYTHON=1``).  PARITY UNPINNED at this boundary.
"""
import os


class _Matrix:
    name = "blosum62"


blosum62 = _Matrix()


class _Traceback:
    __slots__ = ("query", "comp", "ref")

    def __init__(self, query, comp, ref):
        self.query, self.comp, self.ref = query, comp, ref


class _Result:
    def __init__(self, s1, s2, gap_open, gap_extend):
        self._args = (s1, s2, gap_open, gap_extend)
        self.saturated = False

    def get_traceback(self, mch="|", sim=":", neg="."):
        if os.environ.get("PHAMCLUST_ORACLE_PURE_PYTHON") == "1":
            from oracle import nw_py
            q, c, r, _ = nw_py.nw_traceback(*self._args, mch="|", sim="+", neg=" ")
        else:
            from oracle import nw_c
            q, c, r, _ = nw_c.traceback(*self._args)
        if (mch, sim, neg) != ("|", "+", " "):
            c = c.translate(str.maketrans({"|": mch, "+": sim, " ": neg}))
            # gap columns are always ' ' in parasail; restore them
            c = "".join(" " if a == "-" or b == "-" else x for a, x, b in zip(q, c, r))
        return _Traceback(q, c, r)


def nw_trace_diag_16(s1, s2, gap_open, gap_extend, matrix):
    if matrix is not blosum62:
        raise NotImplementedError("the stub only carries blosum62")
    return _Result(s1, s2, gap_open, gap_extend)
