"""oracle/lexc_c.py -- TEST INFRASTRUCTURE: ctypes view of oracle/lexc_model.c, the executable CPU
model of the LEXC DP kernel (rebased LEX word with rebased counters, column chunks, band checks)."""
import ctypes
import pathlib
import subprocess

_HERE = pathlib.Path(__file__).resolve().parent
_SO = _HERE / "_build" / "liblexc_model.so"
_lib = None

# phamclust_b200/csrc/pc_common.cuh: PC_NW_LEXC_BITS, _REBASE, _BAND, _MAXDIFF
BITS, REBASE, BAND, MAXDIFF = 10, 8, 128, 128
K_VALUES = (16, 18, 20, 22, 24)


def build(force=False):
    srcs = [_HERE / "lexc_model.c", _HERE / "nw_oracle.c"]
    if force or not _SO.exists() or _SO.stat().st_mtime < max(s.stat().st_mtime for s in srcs):
        _SO.parent.mkdir(exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-std=c11", "-shared", "-o", str(_SO)] + [str(s) for s in srcs])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(str(_SO))
        L.lexc_align.argtypes = ([ctypes.c_char_p, ctypes.c_int, ctypes.c_char_p, ctypes.c_int] + [ctypes.c_int] * 5
                                 + [ctypes.POINTER(ctypes.c_int)])
        _lib = L
    return _lib


def choose_k(m, n):
    """pc_nw_choose_cfg for PC_NW_MODE_LEXC: fewest chunk x row-step x (K cells + overhead) slots."""
    best = None
    for K in K_VALUES:
        chunks = (n + 32 * K - 1) // (32 * K)
        cost = chunks * (m + 31) * (K * 11 + 22)
        if best is None or cost < best[0]:
            best = (cost, K)
    return best[1]


def align(q, r, K=None, rebase=REBASE, bits=BITS, band=BAND, ppos=False):
    """dict(matched, columns, score, flagged, field_dev, frame_dev, broke, score_field, lane_spread) of the model run;
    matched / columns / score are only meaningful when ``flagged`` is False."""
    a, b = q.encode("latin-1"), r.encode("latin-1")
    out = (ctypes.c_int * 9)()
    rc = lib().lexc_align(a, len(a), b, len(b), K or choose_k(len(a), len(b)), rebase, bits, band, int(ppos), out)
    if rc:
        raise ValueError("lexc_align: parameters outside the soundness conditions")
    keys = ("matched", "columns", "score", "flagged", "field_dev", "frame_dev", "broke", "score_field", "lane_spread")
    return dict(zip(keys, out))
