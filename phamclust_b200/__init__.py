"""phamclust_b200 -- B200-native (sm_100a) all-vs-all genome similarity matrix
for PhamClust, behind the reference's Python API.

    from phamclust_b200 import METRICS, matrix_de_novo, load_genomes_from_tsv
    genomes = load_genomes_from_tsv("phams.tsv")
    dist = matrix_de_novo(genomes, METRICS["peq"], cpus=1)     # a SymMatrix

The arithmetic runs in hand-written CUDA kernels inside libphamclust_b200.so
(C ABI: include/phamclust_b200.h); this package only packs inputs, calls it via
ctypes and wraps the packed result.  There is no CPU fallback.
"""
from phamclust_b200.genome import Genome, GenomeLoadError
from phamclust_b200.matrix import (ResidentMatrix, SymMatrix, calculate_adjacency, matrix_de_novo,
                                   matrix_de_novo_resident, matrix_from_adjacency, matrix_from_squareform,
                                   matrix_to_adjacency, matrix_to_squareform, packed_matrix)
from phamclust_b200.metrics import (METRICS, alignment_fraction, average_aminoacid_identity,
                                    gene_content_similarity, jaccard_coefficient,
                                    percentage_of_conserved_proteins,
                                    proteomic_equivalence_quotient)
from phamclust_b200.packing import PackedTable, pack_genomes, table_from_fasta_dir, table_from_tsv
from phamclust_b200.parsers import load_genomes_from_fasta_dir, load_genomes_from_tsv

__version__ = "0.1.0"
