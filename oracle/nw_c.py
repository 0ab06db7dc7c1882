"""oracle/nw_c.py -- TEST INFRASTRUCTURE: ctypes view of oracle/nw_oracle.c."""
import ctypes
import pathlib
import subprocess

import numpy as np

_HERE = pathlib.Path(__file__).resolve().parent
_SO = _HERE / "_build" / "libnw_oracle.so"
_lib = None


def build(force=False):
    """Compile nw_oracle.c with gcc (seconds); idempotent."""
    src = _HERE / "nw_oracle.c"
    if force or not _SO.exists() or _SO.stat().st_mtime < src.stat().st_mtime:
        _SO.parent.mkdir(exist_ok=True)
        subprocess.check_call(["gcc", "-O3", "-fPIC", "-std=c11", "-shared",
                               "-o", str(_SO), str(src)])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(str(_SO))
        ip = ctypes.POINTER(ctypes.c_int)
        L.nw_trace_stats.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_char_p,
                                     ctypes.c_int, ctypes.c_int, ctypes.c_int, ip,
                                     ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p]
        L.nw_forward_stats.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_char_p,
                                       ctypes.c_int, ctypes.c_int, ctypes.c_int, ip]
        L.nw_gotoh_score.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_char_p,
                                     ctypes.c_int, ctypes.c_int, ctypes.c_int]
        L.nw_batch.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                               ctypes.c_void_p, ctypes.c_int64, ctypes.c_int,
                               ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
        L.nw_tie_range.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_char_p,
                                   ctypes.c_int, ctypes.c_int, ctypes.c_int, ip]
        L.nw_tie_batch.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                   ctypes.c_void_p, ctypes.c_int64, ctypes.c_int,
                                   ctypes.c_int, ctypes.c_void_p]
        _lib = L
    return _lib


def traceback(s1, s2, gap_open=11, gap_extend=1):
    """(query, comp, ref, (ident, similar, length, score)) -- explicit traceback."""
    a, b = s1.encode("latin-1"), s2.encode("latin-1")
    n = len(a) + len(b) + 1
    q, c, r = (ctypes.create_string_buffer(n) for _ in range(3))
    out = (ctypes.c_int * 4)()
    rc = lib().nw_trace_stats(a, len(a), b, len(b), gap_open, gap_extend, out, q, c, r)
    if rc:
        raise MemoryError("nw_trace_stats failed")
    return (q.value.decode("latin-1"), c.value.decode("latin-1"),
            r.value.decode("latin-1"), tuple(out))


def forward_stats(s1, s2, gap_open=11, gap_extend=1):
    a, b = s1.encode("latin-1"), s2.encode("latin-1")
    out = (ctypes.c_int * 4)()
    lib().nw_forward_stats(a, len(a), b, len(b), gap_open, gap_extend, out)
    return tuple(out)


def gotoh_score(s1, s2, gap_open=11, gap_extend=1):
    a, b = s1.encode("latin-1"), s2.encode("latin-1")
    return lib().nw_gotoh_score(a, len(a), b, len(b), gap_open, gap_extend)


def batch(residues, seq_off, q_idx, r_idx, gap_open=11, gap_extend=1, mode=1):
    """Run many alignments over a residue arena; returns int32 [n_tasks, 4]."""
    residues = np.ascontiguousarray(residues, dtype=np.uint8)
    seq_off = np.ascontiguousarray(seq_off, dtype=np.int64)
    q_idx = np.ascontiguousarray(q_idx, dtype=np.int32)
    r_idx = np.ascontiguousarray(r_idx, dtype=np.int32)
    out = np.empty((len(q_idx), 4), dtype=np.int32)
    rc = lib().nw_batch(residues.ctypes.data, seq_off.ctypes.data, q_idx.ctypes.data,
                        r_idx.ctypes.data, len(q_idx), gap_open, gap_extend, mode,
                        out.ctypes.data)
    if rc:
        raise MemoryError("nw_batch failed")
    return out


def tie_range(s1, s2, gap_open=11, gap_extend=1):
    """(id_min, id_max, sm_min, sm_max, ln_min, ln_max, score) over ALL co-optimal
    alignments of the score model -- what any tie-break rule could produce."""
    a, b = s1.encode("latin-1"), s2.encode("latin-1")
    out = (ctypes.c_int * 7)()
    if lib().nw_tie_range(a, len(a), b, len(b), gap_open, gap_extend, out):
        raise MemoryError("nw_tie_range failed")
    return tuple(out)


def tie_batch(residues, seq_off, q_idx, r_idx, gap_open=11, gap_extend=1):
    """tie_range over a residue arena; returns int32 [n_tasks, 7]."""
    residues = np.ascontiguousarray(residues, dtype=np.uint8)
    seq_off = np.ascontiguousarray(seq_off, dtype=np.int64)
    q_idx = np.ascontiguousarray(q_idx, dtype=np.int32)
    r_idx = np.ascontiguousarray(r_idx, dtype=np.int32)
    out = np.empty((len(q_idx), 7), dtype=np.int32)
    if lib().nw_tie_batch(residues.ctypes.data, seq_off.ctypes.data, q_idx.ctypes.data,
                          r_idx.ctypes.data, len(q_idx), gap_open, gap_extend, out.ctypes.data):
        raise MemoryError("nw_tie_batch failed")
    return out
