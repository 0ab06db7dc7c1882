"""Flatten ``list[Genome]`` into the arrays the C ABI takes (SURVEY.md 7
step 1): a CSR genome x pham table (copies and summed translation length per
entry) and one residue arena with genes in CSR entry order.

Only this module walks the Python dicts, once per matrix; the per-pair set
construction of the reference (genome.py:135) has no equivalent here.
"""
import itertools
from dataclasses import dataclass, field

import numpy as np


@dataclass
class PackedTable:
    names: list                      # genome names, caller order
    phams: list                      # pham vocabulary (id -> name)
    genome_ptr: np.ndarray           # int64 [N+1]
    pham_id: np.ndarray              # int32 [nnz]
    copies: np.ndarray               # uint32 [nnz]
    len_sum: np.ndarray              # uint32 [nnz]
    seq_off: np.ndarray = None       # int64 [n_genes+1]
    residues: np.ndarray = None      # uint8 [total residues]
    meta: dict = field(default_factory=dict)

    def __len__(self):
        return self.n_genomes

    @property
    def n_genomes(self):
        return len(self.genome_ptr) - 1

    @property
    def n_phams(self):
        return len(self.phams) if self.phams is not None else int(self.meta["n_phams"])

    @property
    def n_genes(self):
        return int(self.copies.sum())

    @property
    def n_pairs(self):
        n = self.n_genomes
        return n * (n - 1) // 2

    def subset(self, n):
        """The first ``n`` genomes as their own table (CPU-baseline samples)."""
        n = min(n, self.n_genomes)
        nnz = int(self.genome_ptr[n])
        out = PackedTable(self.names[:n], self.phams, self.genome_ptr[:n + 1].copy(),
                          self.pham_id[:nnz].copy(), self.copies[:nnz].copy(),
                          self.len_sum[:nnz].copy(), meta=dict(self.meta))
        if self.seq_off is not None:
            n_genes = int(self.copies[:nnz].sum())
            out.seq_off = self.seq_off[:n_genes + 1].copy()
            out.residues = self.residues[:int(out.seq_off[-1])].copy()
        return out


_pypack = None


def _pypack_lib():
    """libphamclust_pypack.so (csrc/pc_pypack.c) through ctypes.PyDLL: the calls keep the GIL
    and a Python exception set in C is raised here."""
    global _pypack
    if _pypack is None:
        import ctypes
        import pathlib
        path = pathlib.Path(__file__).resolve().parent / "libphamclust_pypack.so"
        if not path.exists():
            raise ImportError(f"{path} is missing: build it with `python -m phamclust_b200.build_native`")
        lib = ctypes.PyDLL(str(path))
        vp = ctypes.c_void_p
        lib.pcpy_pack.argtypes = [ctypes.py_object, ctypes.py_object, ctypes.c_int, vp]
        lib.pcpy_pack.restype = vp
        lib.pcpy_buffers.argtypes = [vp, vp]
        lib.pcpy_buffers.restype = ctypes.c_int
        lib.pcpy_free.argtypes = [vp]
        lib.pcpy_free.restype = None
        _pypack = lib
    return _pypack


def take_genomes(table, indices):
    """The genomes ``indices`` (any order, no repeats needed) of ``table`` as their own table, in
    that order -- random samples for the CPU arm (scripts/benchmark.py:96-100 samples genomes
    with ``random.sample``, it does not take a prefix)."""
    idx = np.asarray(indices, dtype=np.int64)
    lo, hi = table.genome_ptr[idx], table.genome_ptr[idx + 1]
    counts = hi - lo
    ptr = np.zeros(len(idx) + 1, dtype=np.int64)
    np.cumsum(counts, out=ptr[1:])
    ent = np.repeat(lo - ptr[:-1], counts) + np.arange(int(ptr[-1]))          # source entry ids
    out = PackedTable([table.names[i] for i in idx.tolist()], table.phams, ptr,
                      table.pham_id[ent].copy(), table.copies[ent].copy(), table.len_sum[ent].copy(),
                      meta=dict(table.meta))
    if table.seq_off is not None:
        first = np.zeros(len(table.copies) + 1, dtype=np.int64)
        np.cumsum(table.copies, out=first[1:])
        c = table.copies[ent].astype(np.int64)
        g_ptr = np.zeros(len(ent) + 1, dtype=np.int64)
        np.cumsum(c, out=g_ptr[1:])
        genes = np.repeat(first[ent] - g_ptr[:-1], c) + np.arange(int(g_ptr[-1]))   # source gene ids
        glen = table.seq_off[genes + 1] - table.seq_off[genes]
        out.seq_off = np.zeros(len(genes) + 1, dtype=np.int64)
        np.cumsum(glen, out=out.seq_off[1:])
        src = np.repeat(table.seq_off[genes] - out.seq_off[:-1], glen) + np.arange(int(out.seq_off[-1]))
        out.residues = table.residues[src].copy()
    return out


def pack_genomes(genomes, with_sequences=True):
    """list[Genome] (anything with ``.name`` and a ``.phams`` dict of lists) -> PackedTable.

    The dicts are walked ONCE, in C (csrc/pc_pypack.c): pham ids in first-seen order, copies,
    summed translation lengths, and the translations' bytes copied straight from the str objects
    into the residue arena -- this pass sits on the critical path of ``matrix_de_novo`` and, written
    in Python, cost more than the GPU work.  Wide (non latin-1) characters raise ValueError."""
    genomes = genomes if isinstance(genomes, list) else list(genomes)
    lib = _pypack_lib()
    names = [genome.name for genome in genomes]
    sizes = np.zeros(4, dtype=np.int64)
    vocab = {}
    owner = _PackOwner(lib, lib.pcpy_pack(genomes, vocab, int(with_sequences), sizes.ctypes.data))
    _, nnz, n_genes, n_res = (int(v) for v in sizes)
    ptrs = np.zeros(6, dtype=np.uint64)
    lib.pcpy_buffers(owner.handle, ptrs.ctypes.data)
    view = lambda k, count, typestr: np.asarray(_Borrowed(int(ptrs[k]), count, typestr, owner))
    table = PackedTable(names, list(vocab), view(0, len(genomes) + 1, "<i8"), view(1, nnz, "<i4"),
                        view(2, nnz, "<u4"), view(3, nnz, "<u4"))
    if with_sequences:
        table.seq_off, table.residues = view(4, n_genes + 1, "<i8"), view(5, n_res, "|u1")
    return table


class _PackOwner:
    """Owns the C buffers of one pcpy_pack; freed when the last array built on them dies."""

    def __init__(self, lib, handle):
        self.lib, self.handle = lib, handle

    def __del__(self):
        if self.handle:
            self.lib.pcpy_free(self.handle)
            self.handle = None


class _Borrowed:
    """Array-interface shim: numpy wraps the C buffer in place and keeps this object (and through
    it the owner of the buffers) alive as the array's base -- the single buffer hand-off."""

    def __init__(self, address, count, typestr, owner):
        self.owner = owner
        if count == 0 or not address:
            self.__array_interface__ = np.empty(0, dtype=np.dtype(typestr)).__array_interface__
            self._empty = np.empty(0, dtype=np.dtype(typestr))
            self.__array_interface__ = self._empty.__array_interface__
        else:
            self.__array_interface__ = {"data": (address, False), "shape": (count,), "typestr": typestr,
                                        "version": 3}


def pack_genomes_py(genomes, with_sequences=True):
    """The same packing in plain Python/numpy (the pre-C implementation): kept as the executable
    specification the C packer is tested against."""
    vocab = {}
    intern_pham = vocab.setdefault
    names, ptr = [], [0]
    pham_ids, gene_lists = [], []
    for genome in genomes:
        names.append(genome.name)
        phams = genome.phams
        for pham in phams:
            if not isinstance(pham, str):
                raise TypeError(f"type(pham) should be 'str', not '{type(pham)}'")
        pham_ids += [intern_pham(pham, len(vocab)) for pham in phams]
        gene_lists += phams.values()
        ptr.append(len(pham_ids))
    nnz = len(pham_ids)
    copies = np.fromiter(map(len, gene_lists), dtype=np.int64, count=nnz)
    seqs = list(itertools.chain.from_iterable(gene_lists))
    lens = np.fromiter(map(len, seqs), dtype=np.int64, count=len(seqs))
    seq_off = np.zeros(len(seqs) + 1, dtype=np.int64)
    np.cumsum(lens, out=seq_off[1:])
    first_gene = np.zeros(nnz + 1, dtype=np.int64)
    np.cumsum(copies, out=first_gene[1:])
    len_sums = seq_off[first_gene[1:]] - seq_off[first_gene[:-1]]      # exact, also for 0 copies
    table = PackedTable(names, list(vocab), np.asarray(ptr, dtype=np.int64),
                        np.asarray(pham_ids, dtype=np.int32),
                        copies.astype(np.uint32), len_sums.astype(np.uint32))
    if with_sequences:
        table.seq_off = seq_off
        try:
            blob = "".join(seqs).encode("latin-1")
        except UnicodeEncodeError:
            raise ValueError("translations must be made of single-byte (latin-1) characters") from None
        table.residues = np.frombuffer(blob, dtype=np.uint8).copy()
    return table


def unpack_genomes(table):
    """PackedTable -> list[Genome] (used by tests and the CPU baseline)."""
    from phamclust_b200.genome import Genome
    genomes = []
    gene = 0
    raw = table.residues.tobytes().decode("latin-1") if table.residues is not None else None
    for g, name in enumerate(table.names):
        genome = Genome(name)
        for e in range(int(table.genome_ptr[g]), int(table.genome_ptr[g + 1])):
            pham = table.phams[table.pham_id[e]] if table.phams is not None else f"p{table.pham_id[e]}"
            for _ in range(int(table.copies[e])):
                if raw is None:
                    genome.add(pham)
                else:
                    genome.add(pham, raw[int(table.seq_off[gene]):int(table.seq_off[gene + 1])])
                gene += 1
        genomes.append(genome)
    return genomes


def _table_from_rows(names, row_genome, row_name, n_names, name_of, lens, residues_of, with_sequences):
    """Rows (genome id, pham-name id, translation length) in input order -> PackedTable with the
    semantics of the reference loaders followed by ``pack_genomes``: entries = distinct
    (genome, pham) in order of first appearance inside the genome, paralogs in input order, pham
    ids in the order ``pack_genomes`` would meet them.  ``name_of(ids)`` -> list of pham names;
    ``residues_of(row_order, lens)`` -> uint8 arena of the rows' translations in that order."""
    n_genomes = len(names)
    key = row_genome.astype(np.int64) * max(n_names, 1) + row_name
    uniq, first_row, row_entry = np.unique(key, return_index=True, return_inverse=True)
    entry_genome = uniq // max(n_names, 1)
    entry_name = uniq % max(n_names, 1)
    order = np.lexsort((first_row, entry_genome))
    rank = np.empty(len(order), dtype=np.int64)
    rank[order] = np.arange(len(order))
    row_slot = rank[row_entry]
    row_order = np.argsort(row_slot, kind="stable")          # genes in CSR-entry order, input order inside
    copies = np.bincount(row_slot, minlength=len(order)).astype(np.uint32)
    len_sum = np.bincount(row_slot, weights=lens, minlength=len(order)).astype(np.uint32)
    ptr = np.zeros(n_genomes + 1, dtype=np.int64)
    np.cumsum(np.bincount(entry_genome, minlength=n_genomes), out=ptr[1:])
    # pham ids in the order pack_genomes would meet them (genome by genome, CSR order)
    csr_name = entry_name[order]
    met, met_at = np.unique(csr_name, return_index=True)
    vocab_rank = np.empty(max(n_names, 1), dtype=np.int64)
    vocab_order = met[np.argsort(met_at, kind="stable")]      # interned name ids in vocabulary order
    vocab_rank[vocab_order] = np.arange(len(vocab_order))
    pham_id = vocab_rank[csr_name].astype(np.int32)
    table = PackedTable(names, name_of(vocab_order), ptr, pham_id, copies, len_sum)
    if with_sequences:
        glen = np.ascontiguousarray(lens[row_order], dtype=np.int64)
        table.seq_off = np.concatenate(([0], np.cumsum(glen))).astype(np.int64)
        table.residues = residues_of(row_order, glen)
    return table


def table_from_fasta_dir(dirpath, with_sequences=True):
    """A directory of single-genome FASTA files straight into a PackedTable, without Genome
    dicts.  Same semantics as ``load_genomes_from_fasta_dir``
    (scripts/phamclust.py:50-83: suffix .fasta/.faa/.fa, name = file stem, duplicate stem ->
    ValueError, directory order) + ``Genome.load`` (genome.py:66-83: first ``pham=`` field wins,
    none -> GenomeLoadError) + ``read_fasta`` (fasta.py:4-40: lines rstrip()ped and joined)
    followed by ``pack_genomes``.  Every file is scanned once in the library (pc_scan_fasta);
    Python only loops over FILES."""
    import logging
    import pathlib
    from phamclust_b200 import _native
    from phamclust_b200.genome import GenomeLoadError
    lib = _native.load_library()
    names, seen = [], set()
    row_genome, name_fields, lens_all, name_blobs, seq_blobs = [], [], [], [], []
    name_base = 0
    for f in pathlib.Path(dirpath).iterdir():
        if f.suffix not in (".fasta", ".faa", ".fa"):
            logging.debug(f"{f} does not appear to be FASTA - skipping")
            continue
        if f.is_dir():
            logging.debug(f"{f} appears to be a directory - skipping")
            continue
        if f.stem in seen:
            raise ValueError(f"duplicate genome name detected {f.stem}")
        seen.add(f.stem)
        data = f.read_bytes()
        if not data:                     # upstream: next() on an empty file inside a generator (fasta.py:17)
            raise RuntimeError("generator raised StopIteration")
        max_rec = data.count(b">") + 1
        fields = np.empty((max_rec, 4), dtype=np.int64)
        name_out = np.empty(len(data) + 1, dtype=np.uint8)
        seq_out = np.empty(len(data) + 1, dtype=np.uint8)
        used = np.zeros(2, dtype=np.int64)
        rc = int(lib.pc_scan_fasta(data, len(data), max_rec, fields.ctypes.data, name_out.ctypes.data,
                                   len(name_out), seq_out.ctypes.data, len(seq_out), used.ctypes.data))
        if rc < 0:
            msg = lib.pc_io_last_error().decode()
            raise (GenomeLoadError if rc == -3 else ValueError)(msg)
        fields = fields[:rc]
        g = len(names)
        names.append(f.stem)
        row_genome.append(np.full(rc, g, dtype=np.int32))
        name_fields.append(fields[:, 0:2] + name_base)
        lens_all.append(fields[:, 3] - fields[:, 2])
        name_blobs.append(name_out[:int(used[0])].tobytes())
        seq_blobs.append(seq_out[:int(used[1])].copy() if with_sequences else None)
        name_base += int(used[0])
    rows = int(sum(len(x) for x in row_genome))
    row_genome = np.concatenate(row_genome) if row_genome else np.zeros(0, dtype=np.int32)
    lens = np.concatenate(lens_all) if lens_all else np.zeros(0, dtype=np.int64)
    blob = b"".join(name_blobs)
    # pham names interned in first-seen order by the library (the TSV helper: column 1 of a
    # 6-wide field table)
    six = np.zeros((max(rows, 1), 6), dtype=np.int64)
    if rows:
        six[:rows, 2:4] = np.concatenate(name_fields)
    row_name = np.empty(rows, dtype=np.int32)
    first_name_row = np.empty(max(rows, 1), dtype=np.int64)
    n_names = int(lib.pc_tsv_intern(blob, rows, six.ctypes.data, 1, row_name.ctypes.data,
                                    first_name_row.ctypes.data)) if rows else 0
    if n_names < 0:
        raise ValueError(lib.pc_io_last_error().decode())

    def name_of(ids):
        return [blob[b:e].decode() for b, e in six[first_name_row[ids], 2:4].tolist()]

    def residues_of(row_order, glen):
        arena = np.concatenate([b for b in seq_blobs if b is not None]) if seq_blobs else np.zeros(0, dtype=np.uint8)
        start = np.concatenate(([0], np.cumsum(lens)))[:-1]          # rows are laid out in input order
        out_off = np.concatenate(([0], np.cumsum(glen)))
        src = np.repeat(start[row_order] - out_off[:-1], glen) + np.arange(int(out_off[-1]))
        return arena[src]

    return _table_from_rows(names, row_genome, row_name, n_names, name_of, lens, residues_of, with_sequences)


def table_from_tsv(filepath, with_sequences=True):
    """2/3-column TSV straight into a PackedTable, without materialising Genome
    dicts.  Same semantics as ``load_genomes_from_tsv``
    (scripts/phamclust.py:21-47) followed by ``pack_genomes``: genomes and phams
    in first-seen order, paralogs in file order, 2 columns -> translation "M",
    ``rstrip()`` eats a trailing empty third column."""
    from phamclust_b200 import _native
    lib = _native.load_library()
    with open(filepath, "rb") as handle:
        data = handle.read()
    max_rows = data.count(b"\n") + 1
    fields = np.empty((max_rows, 6), dtype=np.int64)
    rows = int(lib.pc_scan_tsv(data, len(data), max_rows, fields.ctypes.data))
    if rows < 0:
        raise ValueError(lib.pc_io_last_error().decode())
    fields = np.ascontiguousarray(fields[:rows])
    # genome and pham names interned in first-seen order by the library (one hash lookup per
    # row in C++, no per-row Python)
    row_genome = np.empty(rows, dtype=np.int32)
    row_name = np.empty(rows, dtype=np.int32)
    first_genome_row = np.empty(max(rows, 1), dtype=np.int64)
    first_name_row = np.empty(max(rows, 1), dtype=np.int64)
    n_genomes = int(lib.pc_tsv_intern(data, rows, fields.ctypes.data, 0, row_genome.ctypes.data,
                                      first_genome_row.ctypes.data))
    n_names = int(lib.pc_tsv_intern(data, rows, fields.ctypes.data, 1, row_name.ctypes.data,
                                    first_name_row.ctypes.data))
    if n_genomes < 0 or n_names < 0:
        raise ValueError(lib.pc_io_last_error().decode())
    names = [data[b:e].decode() for b, e in fields[first_genome_row[:n_genomes], 0:2].tolist()]
    t0, t1 = fields[:, 4], fields[:, 5]
    lens = np.where(t0 < 0, 1, t1 - t0)

    def name_of(ids):
        return [data[b:e].decode() for b, e in fields[first_name_row[ids], 2:4].tolist()]

    def residues_of(row_order, glen):
        starts = np.ascontiguousarray(t0[row_order], dtype=np.int64)
        res = np.empty(int(glen.sum()), dtype=np.uint8)
        # every translation's bytes ("M" for 2-column rows), one memcpy per gene in the library
        if lib.pc_gather_fields(data, len(starts), starts.ctypes.data, glen.ctypes.data, ord("M"),
                                res.ctypes.data):
            raise ValueError(lib.pc_io_last_error().decode())
        return res

    return _table_from_rows(names, row_genome, row_name, n_names, name_of, lens, residues_of, with_sequences)
