"""Flatten ``list[Genome]`` into the arrays the C ABI takes (SURVEY.md 7
step 1): a CSR genome x pham table (copies and summed translation length per
entry) and one residue arena with genes in CSR entry order.

Only this module walks the Python dicts, once per matrix; the per-pair set
construction of the reference (genome.py:135) has no equivalent here.
"""
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


def pack_genomes(genomes, with_sequences=True):
    """list[Genome] (anything with ``.name`` and ``.phams``) -> PackedTable."""
    vocab = {}
    names, ptr = [], [0]
    pham_ids, copies, len_sums, seqs = [], [], [], []
    for genome in genomes:
        names.append(genome.name)
        for pham, translations in genome.phams.items():
            if not isinstance(pham, str):
                raise TypeError(f"type(pham) should be 'str', not '{type(pham)}'")
            pham_ids.append(vocab.setdefault(pham, len(vocab)))
            copies.append(len(translations))
            len_sums.append(sum(map(len, translations)))
            if with_sequences:
                seqs.extend(translations)
        ptr.append(len(pham_ids))
    table = PackedTable(names, list(vocab), np.asarray(ptr, dtype=np.int64),
                        np.asarray(pham_ids, dtype=np.int32),
                        np.asarray(copies, dtype=np.uint32),
                        np.asarray(len_sums, dtype=np.uint32))
    if with_sequences:
        lens = np.fromiter(map(len, seqs), dtype=np.int64, count=len(seqs))
        table.seq_off = np.concatenate(([0], np.cumsum(lens))).astype(np.int64)
        blob = "".join(seqs).encode("latin-1", errors="replace")
        if len(blob) != int(table.seq_off[-1]):
            raise ValueError("translations must be single-byte characters")
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
