"""oracle/make_golden.py -- TEST INFRASTRUCTURE: generate tests/golden/*.json by
running the VERBATIM reference (/root/reference/src/phamclust, imported through
oracle/reference_loader.py) in the build container.

    python -m oracle.make_golden          # rewrites tests/golden/

What is pinned by the verbatim reference: every Python-level behaviour of
genome.py / metrics.py / matrix.py / clustering.py (set algebra, anchor rule,
best-match tie, double rounding, SymMatrix storage).  The alignment call inside
aai/peq is the CPU restatement of parasail (oracle/nw_oracle.c): PARITY UNPINNED
at that boundary -- those numbers are regression vectors, not reference truth.
"""
import json
import pathlib
import random
import sys

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from oracle import nw_c, reference_loader, testdata  # noqa: E402

GOLDEN = ROOT / "tests" / "golden"
METRIC_FUNCS = {"gcs": "gene_content_similarity", "jc": "jaccard_coefficient",
                "pocp": "percentage_of_conserved_proteins", "af": "alignment_fraction",
                "aai": "average_aminoacid_identity", "peq": "proteomic_equivalence_quotient"}


def micro(x):
    return int(round(x * 1e6))


def matrices_fixture(ref):
    out = {}
    for name, kwargs in testdata.DATASETS.items():
        rows = testdata.genome_rows(**kwargs)
        genomes = testdata.rows_to_genomes(rows, cls=ref.genome.Genome)
        entry = {"rows": rows, "names": [g.name for g in genomes], "distance": {},
                 "similarity": {}, "clusters": {}}
        for key, fname in METRIC_FUNCS.items():
            func = getattr(ref.metrics, fname)
            for mode, as_distance in (("distance", True), ("similarity", False)):
                m = ref.matrix.matrix_de_novo(genomes, func, 1, as_distance=as_distance)
                tri = []
                for i, a in enumerate(genomes):
                    for b in genomes[i:]:
                        tri.append(micro(m.get_weight(a.name, b.name)))
                entry[mode][key] = tri
            # partitions from the unchanged host clustering on the distance matrix
            dist = ref.matrix.matrix_de_novo(genomes, func, 1)
            parts = {}
            for linkage, eps in (("complete", 0.25), ("average", 0.75), ("single", 0.4)):
                clusters = ref.clustering.hierarchical_clustering(dist, linkage, eps=eps)
                parts[f"{linkage}:{eps}"] = sorted(sorted(c.nodes) for c in clusters)
            entry["clusters"][key] = parts
        out[name] = entry
    return out


def known_answer_fixture(ref):
    """the hand-checked vector."""
    a = ref.genome.Genome("a")
    b = ref.genome.Genome("b")
    for p, t in (("p1", "MKT"), ("p1", "MKTAA"), ("p2", "MKT"), ("p3", "MKT")):
        a.add(p, t)
    for p in ("p2", "p3", "p4", "p5"):
        b.add(p, "MKTW")
    rows = [("a", "p1", "MKT"), ("a", "p1", "MKTAA"), ("a", "p2", "MKT"), ("a", "p3", "MKT")] + \
           [("b", p, "MKTW") for p in ("p2", "p3", "p4", "p5")]
    sims = {k: getattr(ref.metrics, f)(a, b) for k, f in METRIC_FUNCS.items()}
    dists = {k: getattr(ref.metrics, f)(a, b, as_distance=True) for k, f in METRIC_FUNCS.items()}
    return {"rows": rows, "similarity": sims, "distance": dists}


def alignment_fixture():
    """(query, ref) -> ident, similar, length, score from the restated parasail."""
    rng = random.Random(2024)
    cases = [("", ""), ("A", ""), ("", "WW"), ("A", "A"), ("MKT", "MKTAA"), ("MKTAA", "MKT"),
             ("ACDEFGHIKL", "ACDEFGHIKL"), ("AAAAAAAA", "AAAA"), ("AGAGAGAG", "GAGAGA"),
             ("mktw", "MKTW"), ("MKUX*", "MKOXB"), ("WWWW", "PPPP")]
    for _ in range(120):
        alpha = rng.choice(["AG", "AGS", testdata.AA, testdata.AA])
        a = testdata.random_protein(rng, 1, 70, alpha)
        b = testdata.mutate(rng, a, rng.random() * 0.6, alpha) if rng.random() < 0.7 else \
            testdata.random_protein(rng, 1, 70, alpha)
        cases.append((a, b))
    for _ in range(6):                       # long enough to need several column chunks
        a = testdata.random_protein(rng, 500, 1300)
        cases.append((a, testdata.mutate(rng, a, 0.3)))
    out = []
    for a, b in cases:
        q, c, r, stats = nw_c.traceback(a, b)
        out.append({"query": a, "ref": b, "ident": stats[0], "similar": stats[1],
                    "length": stats[2], "score": stats[3]})
    return out


def round6_fixture():
    """x -> round(x, 6) * 1e6 by CPython, including binary half-way cases."""
    rng = random.Random(7)
    xs = [0.0, 1.0, 0.5, 2 ** -7, 3 * 2 ** -7, 1 / 640, 0.0000005, 0.0000015, 0.0000025,
          0.9999995, 0.9999985, 1 / 3, 2 / 3, 0.1234565, 0.1234575, 0.428571428571, 4 / 7]
    for _ in range(800):
        den = rng.randint(1, 5000)
        num = rng.randint(0, den)
        xs.append(num / den)
        xs.append(1.0 - num / den)
    for k in range(64):                      # exactly representable ties: odd/128 * 1e6 = n + 0.5
        xs.append((2 * k + 1) / 128.0)
    return [{"x": float.hex(x), "micro": micro(round(x, 6))} for x in xs]


def main():
    ref = reference_loader.load()
    GOLDEN.mkdir(parents=True, exist_ok=True)
    (GOLDEN / "matrices.json").write_text(json.dumps(matrices_fixture(ref), indent=0))
    (GOLDEN / "known_answer.json").write_text(json.dumps(known_answer_fixture(ref), indent=1))
    (GOLDEN / "alignments.json").write_text(json.dumps(alignment_fixture(), indent=0))
    (GOLDEN / "round6.json").write_text(json.dumps(round6_fixture(), indent=0))
    for f in sorted(GOLDEN.glob("*.json")):
        print(f.name, f.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
