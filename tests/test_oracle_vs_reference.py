"""Pin the CPU oracle (oracle/pham_oracle.py): against the golden vectors the
VERBATIM reference produced (always), and against the verbatim reference itself
when /root/reference is mounted (build container only)."""
import random

import pytest

from conftest import load_golden
from oracle import pham_oracle as po
from oracle import testdata

METRICS = ("gcs", "jc", "pocp", "af", "aai", "peq")


def _oracle_triangle(genomes, key, as_distance):
    w = po.matrix_de_novo(genomes, po.METRICS[key], 1, as_distance=as_distance)
    n = len(genomes)
    return [po.round6_micro(w[(i, j)]) for i in range(n) for j in range(i, n)]


@pytest.mark.parametrize("dataset", sorted(testdata.DATASETS))
def test_oracle_matches_golden_matrices(dataset):
    golden = load_golden("matrices.json")[dataset]
    genomes = testdata.rows_to_genomes([tuple(r) for r in golden["rows"]])
    assert [g.name for g in genomes] == golden["names"]
    for key in METRICS:
        assert _oracle_triangle(genomes, key, True) == golden["distance"][key], key
        assert _oracle_triangle(genomes, key, False) == golden["similarity"][key], key


def test_oracle_matches_known_answer():
    ka = load_golden("known_answer.json")
    a, b = testdata.rows_to_genomes([tuple(r) for r in ka["rows"]])
    for key in METRICS:
        assert po.METRICS[key](a, b) == ka["similarity"][key]
        assert po.METRICS[key](a, b, as_distance=True) == ka["distance"][key]
    # the hand-derived integers of SURVEY.md 4
    assert po.gcs_parts(a, b) == (2, 3, 4)
    assert po.pocp_parts(a, b) == (4, 8)
    assert po.af_parts(a, b) == (14, 30)


def test_oracle_matches_verbatim_reference(reference):
    rng = random.Random(99)
    for trial in range(6):
        rows = testdata.genome_rows(seed=1000 + trial, n_genomes=rng.randint(2, 9),
                                    n_phams=rng.randint(3, 15), phams_per_genome=(1, 8),
                                    paralog_p=0.4, low_complexity=bool(trial % 2))
        mine = testdata.rows_to_genomes(rows)
        theirs = testdata.rows_to_genomes(rows, cls=reference.genome.Genome)
        names = {"gcs": "gene_content_similarity", "jc": "jaccard_coefficient",
                 "pocp": "percentage_of_conserved_proteins", "af": "alignment_fraction",
                 "aai": "average_aminoacid_identity", "peq": "proteomic_equivalence_quotient"}
        for key in METRICS:
            func = getattr(reference.metrics, names[key])
            for as_distance in (True, False):
                m = reference.matrix.matrix_de_novo(theirs, func, 1, as_distance=as_distance)
                w = po.matrix_de_novo(mine, po.METRICS[key], 1, as_distance=as_distance)
                for i, a in enumerate(theirs):
                    for j in range(i, len(theirs)):
                        assert m.get_weight(a.name, theirs[j].name) == w[(i, j)]
        # ppos flag of aai (metrics.py:219-220)
        for a, b, x, y in zip(theirs, theirs[1:], mine, mine[1:]):
            assert reference.metrics.average_aminoacid_identity(a, b, ppos=True) == \
                po.average_aminoacid_identity(x, y, ppos=True)


def test_outside_in_iterator(reference):
    for n in range(0, 12):
        assert po.outside_in(n) == list(reference.matrix._outside_in_index_iterator(n))


def test_empty_input_raises():
    with pytest.raises(ValueError):
        po.matrix_de_novo([], po.gene_content_similarity)
