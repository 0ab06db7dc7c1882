"""ctypes binding of libphamclust_b200.so (include/phamclust_b200.h).

There is no CPU fallback anywhere in this package: if the shared library is
missing, or no sm_100 GPU is visible, the calls below raise.
"""
import ctypes
import os
import pathlib

import numpy as np

_HERE = pathlib.Path(__file__).resolve().parent
# PHAMCLUST_B200_LIB: an alternative build of the same library (kernel experiments)
LIB_PATH = pathlib.Path(os.environ.get("PHAMCLUST_B200_LIB") or _HERE / "libphamclust_b200.so")

METRIC_IDS = {"gcs": 0, "jc": 1, "pocp": 2, "af": 3, "aai": 4, "peq": 5}
ENGINE_IDS = {"auto": 0, "popc": 1, "i8mma": 2, "i8mma_v1": 3}
DP_MODE_IDS = {"auto": 0, "classic": 1}
STAGES = ("h2d", "pack", "contract", "align", "gather", "epilogue", "d2h",
          "host_table", "host_seqs", "host_plan")

EXPORTS = ("pc_init", "pc_destroy", "pc_last_error", "pc_version", "pc_set_stream",
           "pc_set_engine", "pc_set_dp_mode", "pc_sync", "pc_load_table", "pc_load_sequences", "pc_matrix",
           "pc_matrix_rows", "pc_align_accumulate", "pc_numerators", "pc_load_sequences_part", "pc_align_units",
           "pc_align_contributions", "pc_accumulate_records", "pc_matrix_rows_owned", "pc_matrix_rows_fused",
           "pc_matrix_resident", "pc_resident_threshold_edges", "pc_resident_gather", "pc_resident_fetch",
           "pc_alignment_identities", "pc_debug_round6", "pc_debug_metric_micro", "pc_timings", "pc_counters", "pc_reset_counters", "pc_stat",
           "pc_write_matrix", "pc_parse_squareform", "pc_scan_tsv", "pc_tsv_intern", "pc_gather_fields",
           "pc_dense_from_micro", "pc_node_sums", "pc_scan_fasta",
           "pc_io_last_error")

_lib = None


def load_library():
    """dlopen the in-tree .so and declare every prototype; raises if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m phamclust_b200.build_native` "
            "(nvcc, sm_100a). phamclust_b200 has no CPU fallback.")
    lib = ctypes.CDLL(str(LIB_PATH))
    vp, i32, i64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64
    lib.pc_init.argtypes = [i32, ctypes.POINTER(vp)]
    lib.pc_destroy.argtypes = [vp]
    lib.pc_destroy.restype = None
    lib.pc_last_error.argtypes = [vp]
    lib.pc_last_error.restype = ctypes.c_char_p
    lib.pc_version.argtypes = []
    lib.pc_set_stream.argtypes = [vp, vp, i32]
    lib.pc_set_engine.argtypes = [vp, i32]
    lib.pc_set_dp_mode.argtypes = [vp, i32]
    lib.pc_sync.argtypes = [vp]
    lib.pc_load_table.argtypes = [vp, i64, i64, vp, vp, vp, vp]
    lib.pc_load_sequences.argtypes = [vp, i64, vp, vp]
    lib.pc_matrix.argtypes = [vp, i32, i32, i32, vp, i32]
    lib.pc_matrix_rows.argtypes = [vp, i32, i32, i32, i64, i64, vp, vp, i32]
    lib.pc_align_accumulate.argtypes = [vp, i32, i32, i32, vp]
    lib.pc_numerators.argtypes = [vp, i32, i32, vp, vp]
    lib.pc_load_sequences_part.argtypes = [vp, i64, vp, vp, i32, i32]
    lib.pc_align_units.argtypes = [vp, i32, i32, vp]
    lib.pc_align_contributions.argtypes = [vp, i32, i32, i32, i32, vp, vp, vp]
    lib.pc_accumulate_records.argtypes = [vp, i64, vp, i64, vp]
    lib.pc_matrix_rows_owned.argtypes = [vp, i32, i32, i32, i64, i64, vp, vp, i32]
    lib.pc_matrix_rows_fused.argtypes = [vp, i32, i32, i32, i64, i64, vp, vp, i32, vp]
    lib.pc_matrix_resident.argtypes = [vp, i32, i32, i32]
    lib.pc_resident_threshold_edges.argtypes = [vp, i32, i64, vp, vp]
    lib.pc_resident_gather.argtypes = [vp, i64, vp, vp]
    lib.pc_resident_fetch.argtypes = [vp, vp]
    lib.pc_alignment_identities.argtypes = [vp, i64, vp, vp, vp]
    lib.pc_debug_round6.argtypes = [vp, i64, vp, vp]
    lib.pc_debug_metric_micro.argtypes = [vp, i32, i32, i64, vp, vp, vp, vp, vp]
    lib.pc_timings.argtypes = [vp, vp, i32]
    lib.pc_counters.argtypes = [vp, vp, vp, vp, vp, vp]
    lib.pc_reset_counters.argtypes = [vp]
    lib.pc_stat.argtypes = [vp, i32, vp]
    lib.pc_write_matrix.argtypes = [ctypes.c_char_p, i64, vp, vp, ctypes.c_char_p, vp, i32]
    lib.pc_parse_squareform.argtypes = [ctypes.c_char_p, i64, i64, vp, vp, vp]
    lib.pc_scan_tsv.argtypes = [ctypes.c_char_p, i64, i64, vp]
    lib.pc_scan_tsv.restype = ctypes.c_int64
    lib.pc_dense_from_micro.argtypes = [i64, vp, vp, i32, vp]
    lib.pc_tsv_intern.argtypes = [ctypes.c_char_p, i64, vp, i32, vp, vp]
    lib.pc_tsv_intern.restype = ctypes.c_int64
    lib.pc_gather_fields.argtypes = [ctypes.c_char_p, i64, vp, vp, ctypes.c_uint8, vp]
    lib.pc_node_sums.argtypes = [i64, vp, vp]
    lib.pc_scan_fasta.argtypes = [ctypes.c_char_p, i64, i64, vp, vp, i64, vp, i64, vp]
    lib.pc_scan_fasta.restype = ctypes.c_int64
    lib.pc_io_last_error.argtypes = []
    lib.pc_io_last_error.restype = ctypes.c_char_p
    _lib = lib
    return lib


def _metric_id(metric):
    if isinstance(metric, str):
        return METRIC_IDS[metric]
    return int(metric)


def tri_len(n, r0=0, r1=None):
    """Entries of the packed upper triangle covering rows [r0, r1)."""
    r1 = n if r1 is None else r1
    idx = lambda i: i * n - (i * (i - 1)) // 2
    return idx(r1) - idx(r0)


class Context:
    """One pc_ctx (one GPU).  Methods mirror the C entry points 1:1."""

    def __init__(self, device=0):
        self._lib = load_library()
        handle = ctypes.c_void_p()
        rc = self._lib.pc_init(int(device), ctypes.byref(handle))
        if rc:
            msg = self._lib.pc_last_error(None).decode()
            raise RuntimeError(f"pc_init failed ({rc}): {msg}")
        self._h = handle
        self.device = int(device)
        self.n = 0
        self._keep = []

    def close(self):
        if getattr(self, "_h", None):
            self._lib.pc_destroy(self._h)
            self._h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, rc):
        if rc == 0:
            return
        msg = self._lib.pc_last_error(self._h).decode()
        if rc in (1, 4):            # PC_EINVAL / PC_ESTATE
            raise ValueError(msg)
        if rc == 5:
            raise MemoryError(msg)
        raise RuntimeError(f"libphamclust_b200 error {rc}: {msg}")

    # ---- configuration -----------------------------------------------------
    def set_stream(self, cuda_stream, external=True):
        """Run on the caller's CUDA stream handle (0 = legacy default stream);
        ``external=False`` goes back to the ctx's own stream."""
        self._check(self._lib.pc_set_stream(self._h, ctypes.c_void_p(int(cuda_stream or 0)),
                                            int(bool(external))))

    def set_engine(self, engine):
        self._check(self._lib.pc_set_engine(self._h, ENGINE_IDS.get(engine, engine)))

    def set_dp_mode(self, mode):
        """"auto": LEX DP kernel where its 32-bit word provably fits; "classic": never."""
        self._check(self._lib.pc_set_dp_mode(self._h, DP_MODE_IDS.get(mode, mode)))

    def sync(self):
        self._check(self._lib.pc_sync(self._h))

    # ---- inputs --------------------------------------------------------------
    def load_table(self, table):
        gp = np.ascontiguousarray(table.genome_ptr, dtype=np.int64)
        pid = np.ascontiguousarray(table.pham_id, dtype=np.int32)
        cp = np.ascontiguousarray(table.copies, dtype=np.uint32)
        ls = np.ascontiguousarray(table.len_sum, dtype=np.uint32)
        self._check(self._lib.pc_load_table(self._h, len(gp) - 1, table.n_phams, gp.ctypes.data,
                                            pid.ctypes.data, cp.ctypes.data, ls.ctypes.data))
        self.n = len(gp) - 1
        self.resident_generation = getattr(self, "resident_generation", 0) + 1   # the resident matrix is gone

    def load_sequences(self, table, part=0, n_parts=1):
        """All translations, or (``n_parts`` > 1) only the genes of the phams partition ``part``
        aligns."""
        so = np.ascontiguousarray(table.seq_off, dtype=np.int64)
        rs = np.ascontiguousarray(table.residues, dtype=np.uint8)
        if n_parts > 1:
            self._check(self._lib.pc_load_sequences_part(self._h, len(so) - 1, so.ctypes.data,
                                                         rs.ctypes.data if rs.size else None, part, n_parts))
        else:
            self._check(self._lib.pc_load_sequences(self._h, len(so) - 1, so.ctypes.data,
                                                    rs.ctypes.data if rs.size else None))

    # ---- the path --------------------------------------------------------------
    def matrix(self, metric, as_distance=True, ppos=False, out=None, out_device_ptr=None):
        """Packed upper triangle (int32 micro-units) of the whole table: into a host array
        (returned), or into caller-owned device memory (``out_device_ptr``)."""
        count = tri_len(self.n)
        if out_device_ptr is not None:
            self._check(self._lib.pc_matrix(self._h, _metric_id(metric), int(bool(as_distance)),
                                            int(bool(ppos)), ctypes.c_void_p(out_device_ptr), 1))
            return None
        if out is None:
            out = np.empty(count, dtype=np.int32)
        assert out.dtype == np.int32 and out.size == count and out.flags.c_contiguous
        self._check(self._lib.pc_matrix(self._h, _metric_id(metric), int(bool(as_distance)),
                                        int(bool(ppos)), out.ctypes.data, 0))
        return out

    def matrix_rows(self, metric, as_distance, row_begin, row_end, acc_ptr=None, out=None,
                    out_device_ptr=None, ppos=False):
        """Rows [row_begin, row_end); writes to a host array or a device pointer."""
        count = tri_len(self.n, row_begin, row_end)
        if out_device_ptr is not None:
            self._check(self._lib.pc_matrix_rows(
                self._h, _metric_id(metric), int(bool(as_distance)), int(bool(ppos)), row_begin,
                row_end, ctypes.c_void_p(acc_ptr or 0), ctypes.c_void_p(out_device_ptr), 1))
            return None
        if out is None:
            out = np.empty(count, dtype=np.int32)
        assert out.dtype == np.int32 and out.size == count and out.flags.c_contiguous
        self._check(self._lib.pc_matrix_rows(
            self._h, _metric_id(metric), int(bool(as_distance)), int(bool(ppos)), row_begin,
            row_end, ctypes.c_void_p(acc_ptr or 0), out.ctypes.data, 0))
        return out

    def align_accumulate(self, acc_device_ptr, part=0, n_parts=1, ppos=False):
        self._check(self._lib.pc_align_accumulate(self._h, int(bool(ppos)), part, n_parts,
                                                  ctypes.c_void_p(acc_device_ptr)))

    def align_units(self, part=0, n_parts=1):
        """Upper bound of the contribution records partition ``part`` emits."""
        v = ctypes.c_int64()
        self._check(self._lib.pc_align_units(self._h, part, n_parts, ctypes.byref(v)))
        return v.value

    def align_contributions(self, records_device_ptr, row_bounds, part=0, n_parts=1, ppos=False):
        """Align partition ``part`` and write one record per (pham, genome pair) unit into the
        bucket of the row block (``row_bounds``) that owns it; returns the bucket sizes."""
        bounds = np.ascontiguousarray(row_bounds, dtype=np.int64)
        counts = np.zeros(len(bounds) - 1, dtype=np.int64)
        self._check(self._lib.pc_align_contributions(self._h, int(bool(ppos)), part, n_parts, len(bounds) - 1,
                                                     bounds.ctypes.data, ctypes.c_void_p(records_device_ptr or 0),
                                                     counts.ctypes.data))
        return counts

    def accumulate_records(self, n_records, records_device_ptr, n_acc, acc_device_ptr):
        self._check(self._lib.pc_accumulate_records(self._h, n_records, ctypes.c_void_p(records_device_ptr or 0),
                                                    n_acc, ctypes.c_void_p(acc_device_ptr or 0)))

    def matrix_rows_owned(self, metric, as_distance, row_begin, row_end, acc_rows_ptr, out=None,
                          out_device_ptr=None, ppos=False):
        """``matrix_rows`` with accumulators that hold only the rows [row_begin, row_end)."""
        count = tri_len(self.n, row_begin, row_end)
        if out_device_ptr is not None:
            self._check(self._lib.pc_matrix_rows_owned(
                self._h, _metric_id(metric), int(bool(as_distance)), int(bool(ppos)), row_begin,
                row_end, ctypes.c_void_p(acc_rows_ptr or 0), ctypes.c_void_p(out_device_ptr), 1))
            return None
        if out is None:
            out = np.empty(count, dtype=np.int32)
        assert out.dtype == np.int32 and out.size == count and out.flags.c_contiguous
        self._check(self._lib.pc_matrix_rows_owned(
            self._h, _metric_id(metric), int(bool(as_distance)), int(bool(ppos)), row_begin,
            row_end, ctypes.c_void_p(acc_rows_ptr or 0), out.ctypes.data, 0))
        return out

    def matrix_rows_fused(self, metric, as_distance, row_begin, row_end, acc_rows_ptr, out_device_ptr,
                          peer_ptrs, ppos=False):
        """``matrix_rows_owned`` whose result is also stored into the same slice of the other ranks'
        triangles (``peer_ptrs``: peer-mapped device addresses of those slices)."""
        arr = (ctypes.c_void_p * max(len(peer_ptrs), 1))(*[ctypes.c_void_p(int(p)) for p in peer_ptrs])
        self._check(self._lib.pc_matrix_rows_fused(
            self._h, _metric_id(metric), int(bool(as_distance)), int(bool(ppos)), row_begin, row_end,
            ctypes.c_void_p(acc_rows_ptr or 0), ctypes.c_void_p(out_device_ptr), len(peer_ptrs), arr))

    # ---- the triangle kept in HBM ---------------------------------------------------
    def matrix_resident(self, metric, as_distance=True, ppos=False):
        """Compute the whole triangle into the ctx's own device buffer (no D2H); returns a
        generation number that the queries below check."""
        self._check(self._lib.pc_matrix_resident(self._h, _metric_id(metric), int(bool(as_distance)), int(bool(ppos))))
        self.resident_generation = getattr(self, "resident_generation", 0) + 1
        return self.resident_generation

    def resident_threshold_edges(self, max_micro):
        """(i, j) storage indices, i < j, of the pairs with weight <= max_micro; int32 [k, 2]."""
        cap = max(1024, 4 * self.n)
        while True:
            edges = np.empty((cap, 2), dtype=np.int32)
            found = ctypes.c_int64()
            self._check(self._lib.pc_resident_threshold_edges(self._h, int(max_micro), cap, edges.ctypes.data,
                                                              ctypes.byref(found)))
            if found.value <= cap:
                return edges[:found.value]
            cap = found.value

    def resident_gather(self, tri_index):
        idx = np.ascontiguousarray(tri_index, dtype=np.int64)
        out = np.empty(idx.size, dtype=np.int32)
        self._check(self._lib.pc_resident_gather(self._h, idx.size, idx.ctypes.data, out.ctypes.data))
        return out

    def resident_fetch(self):
        out = np.empty(tri_len(self.n), dtype=np.int32)
        self._check(self._lib.pc_resident_fetch(self._h, out.ctypes.data))
        return out

    # ---- integer hooks -----------------------------------------------------------
    def numerators(self, metric, ppos=False):
        count = tri_len(self.n)
        num = np.empty(count, dtype=np.int32)
        aux = np.empty(count, dtype=np.int32) if _metric_id(metric) == 4 else None
        self._check(self._lib.pc_numerators(self._h, _metric_id(metric), int(bool(ppos)),
                                            num.ctypes.data, aux.ctypes.data if aux is not None else None))
        return (num, aux) if aux is not None else num

    def alignment_identities(self, query_gene, ref_gene):
        q = np.ascontiguousarray(query_gene, dtype=np.int32)
        r = np.ascontiguousarray(ref_gene, dtype=np.int32)
        out = np.empty((len(q), 4), dtype=np.int32)
        self._check(self._lib.pc_alignment_identities(self._h, len(q), q.ctypes.data,
                                                      r.ctypes.data, out.ctypes.data))
        return out

    def debug_round6(self, values):
        x = np.ascontiguousarray(values, dtype=np.float64)
        out = np.empty(x.size, dtype=np.int32)
        self._check(self._lib.pc_debug_round6(self._h, x.size, x.ctypes.data, out.ctypes.data))
        return out

    def debug_metric_micro(self, metric, as_distance, numer, den_i, den_j):
        """(production path, fp64 path) of the cheap-metric epilogue on the given operands."""
        a = np.ascontiguousarray(numer, dtype=np.int32)
        b = np.ascontiguousarray(den_i, dtype=np.int32)
        c = np.ascontiguousarray(den_j, dtype=np.int32)
        out, out64 = np.empty(a.size, dtype=np.int32), np.empty(a.size, dtype=np.int32)
        self._check(self._lib.pc_debug_metric_micro(self._h, _metric_id(metric), int(bool(as_distance)), a.size,
                                                    a.ctypes.data, b.ctypes.data, c.ctypes.data,
                                                    out.ctypes.data, out64.ctypes.data))
        return out, out64

    # ---- measurement ---------------------------------------------------------------
    def timings(self):
        ms = (ctypes.c_double * len(STAGES))()
        self._check(self._lib.pc_timings(self._h, ms, len(STAGES)))
        return dict(zip(STAGES, ms))

    def counters(self):
        vals = [ctypes.c_int64() for _ in range(5)]
        self._check(self._lib.pc_counters(self._h, *[ctypes.byref(v) for v in vals]))
        keys = ("launches", "dp_cells", "alignments", "contract_launches", "align_launches")
        return {k: v.value for k, v in zip(keys, vals)}

    def stats(self):
        out = {}
        for which, key in enumerate(("logical_alignments", "unique_genes", "operand_columns", "genes",
                                     "lex_alignments", "lex_cells", "dp_configs_used", "dp_configs", "i8_tile_steps",
                                     "i8_tile_steps_dense", "lexr_cells", "lexc_cells", "lexc_fallbacks", "h2d_bytes")):
            v = ctypes.c_int64()
            self._check(self._lib.pc_stat(self._h, which, ctypes.byref(v)))
            out[key] = v.value
        return out

    def reset_counters(self):
        self._check(self._lib.pc_reset_counters(self._h))
