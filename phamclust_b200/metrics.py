"""PhamClust's six genome-pair metrics with the reference's signatures
(/root/reference/src/phamclust/metrics.py:26,56,83,118,178,235):
``f(source, target, as_distance=False) -> float``.

They exist so that ``METRICS[name]`` / ``matrix_de_novo(genomes, func, cpus)``
keep working unchanged; ``matrix_de_novo`` recognises these objects and computes
the whole triangle on the GPU in one call.  Calling one of them directly on two
genomes also runs on the GPU (a 2-genome table) -- there is no CPU arithmetic
path in this package.
"""
from phamclust_b200.genome import Genome

__all__ = ["alignment_fraction", "average_aminoacid_identity",
           "gene_content_similarity", "jaccard_coefficient",
           "percentage_of_conserved_proteins", "proteomic_equivalence_quotient"]


def _pair(metric, source, target, as_distance, ppos=False):
    if not isinstance(source, Genome) or not isinstance(target, Genome):
        raise TypeError(f"cannot compare Genome to '{type(target)}'")
    from phamclust_b200.matrix import get_context
    from phamclust_b200.packing import pack_genomes
    table = pack_genomes([source, target], with_sequences=metric in ("aai", "peq"))
    ctx = get_context()
    ctx.load_table(table)
    if metric in ("aai", "peq"):
        ctx.load_sequences(table)
    micro = ctx.matrix(metric, as_distance=as_distance, ppos=ppos)
    return int(micro[1]) / 1e6          # packed triangle of 2 nodes: (0,0) (0,1) (1,1)


def gene_content_similarity(source, target, as_distance=False):
    """2|S n T| / (|S| + |T|) over distinct phams (metrics.py:26-53)."""
    return _pair("gcs", source, target, as_distance)


def jaccard_coefficient(source, target, as_distance=False):
    """|S n T| / |S u T| (metrics.py:56-80)."""
    return _pair("jc", source, target, as_distance)


def percentage_of_conserved_proteins(source, target, as_distance=False):
    """Genes in shared phams over all genes, both genomes (metrics.py:83-115)."""
    return _pair("pocp", source, target, as_distance)


def alignment_fraction(source, target, as_distance=False):
    """Residues in shared phams over all residues (metrics.py:118-157)."""
    return _pair("af", source, target, as_distance)


def average_aminoacid_identity(source, target, ppos=False, as_distance=False):
    """Length-weighted identity of best matches in shared phams
    (metrics.py:178-232); ``ppos`` counts positive-scoring columns too."""
    return _pair("aai", source, target, as_distance, ppos=ppos)


def proteomic_equivalence_quotient(source, target, as_distance=False):
    """round(af, 6) * round(aai, 6) (metrics.py:235-253)."""
    return _pair("peq", source, target, as_distance)


METRICS = {"gcs": gene_content_similarity,
           "jc": jaccard_coefficient,
           "pocp": percentage_of_conserved_proteins,
           "af": alignment_fraction,
           "aai": average_aminoacid_identity,
           "peq": proteomic_equivalence_quotient}

_BY_NAME = {f.__name__: k for k, f in METRICS.items()}


def metric_key(func):
    """'gcs' .. 'peq' for one of the six functions (ours, or an object with the
    same ``__name__``, e.g. the reference's own function), or a key string."""
    if isinstance(func, str):
        return func if func in METRICS else None
    for key, f in METRICS.items():
        if func is f:
            return key
    return _BY_NAME.get(getattr(func, "__name__", None))
