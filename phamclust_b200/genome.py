"""Host-side genome container with PhamClust's ``Genome`` API.

Mirrors /root/reference/src/phamclust/genome.py:12-225 (same attribute and
method names, argument meaning and error behaviour) so code written against the
reference keeps working; the device path never walks these dicts per pair --
``phamclust_b200.packing`` flattens a list of genomes ONCE into CSR + a residue
arena.
"""
from phamclust_b200.fasta import read_fasta


class GenomeLoadError(Exception):
    """Raised when a FASTA record carries no ``pham=`` field (genome.py:8-10)."""


def _need_genome(other):
    if not isinstance(other, Genome):
        raise TypeError(f"cannot compare Genome to '{type(other)}'")


class Genome:
    """``phams`` maps pham name -> list of translations (paralogs in insertion
    order, which decides best-match ties downstream: metrics.py:223)."""

    def __init__(self, name):
        self.name = name
        self.phams = dict()

    # -- construction ------------------------------------------------------
    def add(self, pham, translation="M"):
        """genome.py:32-48."""
        if not isinstance(pham, str):
            raise TypeError(f"type(pham) should be 'str', not '{type(pham)}'")
        if not isinstance(translation, str):
            raise TypeError(f"type(translation) should be 'str', not '{type(translation)}'")
        self.phams.setdefault(pham, []).append(translation)

    def load(self, fasta):
        """genome.py:66-83 -- header fields ``k=v`` joined by ``|``; the first
        ``pham=`` wins."""
        for header, translation in read_fasta(fasta):
            pham = None
            for field in header.split("|"):
                key, value = field.split("=")
                if key == "pham":
                    pham = value
                    break
            if pham is None:
                raise GenomeLoadError("unable to get pham from FASTA header")
            self.add(pham, translation)

    def pop(self, pham):
        if pham not in self:
            return None
        return self.phams.pop(pham)

    def save(self, filepath):
        with open(filepath, "w") as handle:
            handle.write(str(self))
        return filepath

    # -- set algebra on pham names (genome.py:50-64, 108-135, 177-225) ------
    def difference(self, other):
        return self - other

    def intersection(self, other):
        return self & other

    def symmetric_difference(self, other):
        return self ^ other

    def union(self, other):
        return self | other

    def __and__(self, other):
        _need_genome(other)
        return set(self.phams.keys()).intersection(other.phams.keys())

    def __or__(self, other):
        _need_genome(other)
        return set(self.phams.keys()).union(other.phams.keys())

    def __sub__(self, other):
        _need_genome(other)
        return set(self.phams.keys()).difference(other.phams.keys())

    def __xor__(self, other):
        _need_genome(other)
        return set(self.phams.keys()).symmetric_difference(other.phams.keys())

    # -- container protocol --------------------------------------------------
    def __contains__(self, item):
        if not isinstance(item, str):
            raise TypeError(f"type(item) should be 'str', not '{type(item)}'")
        return item in self.phams

    def __getitem__(self, item):
        if item not in self:
            raise KeyError(f"node '{item}' not in matrix")
        return self.phams[item]

    def __iter__(self):
        yield from self.phams.items()

    def __len__(self):
        """Total number of GENES, not phams (genome.py:168-169)."""
        return sum(len(v) for v in self.phams.values())

    def __lt__(self, other):
        _need_genome(other)
        return len(self) < len(other)

    def __str__(self):
        """FASTA text, ``>name=..|pham=..|n=..`` (genome.py:192-199); the
        pipeline's md5 cache key hashes exactly this text."""
        parts = []
        for pham, translations in self:
            for i, translation in enumerate(translations):
                parts.append(f">name={self.name}|pham={pham}|n={i + 1}\n{translation}\n")
        return "".join(parts)

    __repr__ = __str__
