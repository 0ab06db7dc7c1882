"""oracle/testdata.py -- TEST INFRASTRUCTURE: seeded small genome sets used by
the golden-vector generator and the parity tests (same inputs on both sides)."""
import random

AA = "ARNDCQEGHILKMFPSTWYV"


class PlainGenome:
    """Minimal genome (``name`` + ``phams`` dict) accepted by oracle/pham_oracle.py."""

    def __init__(self, name):
        self.name = name
        self.phams = {}

    def add(self, pham, translation="M"):
        self.phams.setdefault(pham, []).append(translation)


def random_protein(rng, lo=20, hi=90, alphabet=AA):
    return "".join(rng.choice(alphabet) for _ in range(rng.randint(lo, hi)))


def mutate(rng, seq, rate, alphabet=AA):
    out = []
    for ch in seq:
        r = rng.random()
        if r < rate * 0.6:
            out.append(rng.choice(alphabet))
        elif r < rate * 0.8:
            continue
        elif r < rate:
            out.append(ch)
            out.append(rng.choice(alphabet))
        else:
            out.append(ch)
    return "".join(out) or rng.choice(alphabet)


def genome_rows(seed, n_genomes=12, n_phams=30, phams_per_genome=(4, 14), lo=20, hi=90,
                paralog_p=0.2, low_complexity=False, dup_p=0.15):
    """Rows (name, pham, translation) of a small random dataset: shared phams
    with diverged homologs, paralogs (up to 3 copies, sometimes identical),
    an isolated genome, optional low-complexity alphabet (tie-rich alignments)."""
    rng = random.Random(seed)
    alphabet = "AGS" if low_complexity else AA
    ancestors = {f"pham{p}": random_protein(rng, lo, hi, alphabet) for p in range(n_phams)}
    rows = []
    for g in range(n_genomes):
        name = f"g{seed}_{g:02d}"
        k = rng.randint(*phams_per_genome)
        if g == n_genomes - 1:
            chosen = [f"solo{seed}_{x}" for x in range(3)]     # shares nothing
            for p in chosen:
                ancestors[p] = random_protein(rng, lo, hi, alphabet)
        else:
            chosen = rng.sample(sorted(ancestors)[:n_phams], min(k, n_phams))
        for p in chosen:
            copies = 1
            while copies < 3 and rng.random() < paralog_p:
                copies += 1
            prev = None
            for _ in range(copies):
                if prev is not None and rng.random() < dup_p:
                    seq = prev
                else:
                    seq = mutate(rng, ancestors[p], rng.random() * 0.5, alphabet)
                rows.append((name, p, seq))
                prev = seq
    return rows


def rows_to_genomes(rows, cls=PlainGenome):
    genomes = {}
    for name, pham, translation in rows:
        if name not in genomes:
            genomes[name] = cls(name)
        genomes[name].add(pham, translation)
    return list(genomes.values())


DATASETS = {
    "mixed12": dict(seed=11, n_genomes=12, n_phams=30),
    "paralogs9": dict(seed=12, n_genomes=9, n_phams=10, phams_per_genome=(5, 9), paralog_p=0.6),
    "lowcomplex8": dict(seed=13, n_genomes=8, n_phams=12, low_complexity=True, lo=10, hi=60),
    "tiny3": dict(seed=14, n_genomes=3, n_phams=4, phams_per_genome=(2, 4), lo=5, hi=30),
}
