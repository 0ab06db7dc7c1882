"""oracle/reference_loader.py -- TEST INFRASTRUCTURE.

Imports the VERBATIM reference (``/root/reference/src/phamclust``) in the build
container, with oracle/parasail_stub first on sys.path so that
``from parasail import blosum62, nw_trace_diag_16`` (metrics.py:20) resolves to
the CPU restatement.  /root/reference does not exist on the GPU box: callers
must check ``available()`` and nothing on the ``-m gpu`` / smoke / bench paths
may depend on it.
"""
import importlib
import pathlib
import sys

REFERENCE_SRC = pathlib.Path("/root/reference/src")
_STUB = pathlib.Path(__file__).resolve().parent / "parasail_stub"


def available():
    return (REFERENCE_SRC / "phamclust" / "metrics.py").is_file()


def load():
    """Return a namespace with the reference's genome / metrics / matrix /
    clustering modules (heatmap needs plotly, which is absent: not imported)."""
    if not available():
        raise RuntimeError("reference sources are not mounted in this environment")
    for p in (str(REFERENCE_SRC), str(_STUB)):
        if p not in sys.path:
            sys.path.insert(0, p)
    mods = {}
    for name in ("genome", "statistics", "metrics", "matrix", "clustering", "fasta"):
        mods[name] = importlib.import_module(f"phamclust.{name}")
    return type("Reference", (), mods)
