"""The ctypes stub printed in INTEGRATION.md is executed verbatim (only the
library path is substituted) and checked against the oracle, so the document
cannot drift from the ABI."""
import re

import numpy as np
import pytest

from conftest import ROOT
from oracle import pham_oracle as po
from oracle import testdata
from phamclust_b200 import Genome, _native

pytestmark = pytest.mark.gpu


def test_integration_md_stub_runs_and_matches_the_oracle():
    text = (ROOT / "INTEGRATION.md").read_text()
    blocks = re.findall(r"```python\n(.*?)```", text, re.S)
    stub = next(b for b in blocks if "def distance_triangle" in b)
    stub = stub.replace('"libphamclust_b200.so"', repr(str(_native.LIB_PATH)))
    ns = {}
    exec(compile(stub, "INTEGRATION.md", "exec"), ns)
    rows = testdata.genome_rows(seed=61, n_genomes=7, n_phams=9, paralog_p=0.3)
    genomes = testdata.rows_to_genomes(rows, cls=Genome)
    n = len(genomes)
    for key, func in (("gcs", po.gene_content_similarity), ("peq", po.proteomic_equivalence_quotient)):
        tri = ns["distance_triangle"](genomes, func, True)
        want = po.matrix_de_novo(genomes, func, 1)
        k = 0
        for i in range(n):
            for j in range(i, n):
                assert abs(tri[k] / 1e6 - want[(i, j)]) <= (0 if key == "gcs" else 1.0000001e-6)
                k += 1
    with pytest.raises(ValueError):
        ns["distance_triangle"]([], po.gene_content_similarity)
