/*
 * phamclust_pypack.h -- entry points of libphamclust_pypack.so (phamclust_b200/csrc/pc_pypack.c):
 * the host-side half of the drop-in boundary that has to touch Python objects.
 *
 * matrix_de_novo(genomes, func, cpus, as_distance=True)
 *                              /root/reference/src/phamclust/matrix.py:432
 * receives list[Genome]; Genome.phams (genome.py:30) is a dict pham -> list[str].  These
 * functions flatten that list ONCE into the arrays pc_load_table / pc_load_sequences take
 * (include/phamclust_b200.h).  They use the CPython C API and must be called with the GIL held
 * (ctypes.PyDLL); on error they return NULL / -1 with a Python exception set.  No GPU involved.
 */
#ifndef PHAMCLUST_PYPACK_H
#define PHAMCLUST_PYPACK_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* PyObject* arguments are declared void* so that the header needs no Python.h. */

/* One walk over genomes (a list): vocab (an empty dict) receives pham name -> id in first-seen
 * order; sizes[4] = n_genomes, nnz, n_genes, n_residues.  Returns an opaque pack. */
void *pcpy_pack(void *genomes_list, void *vocab_dict, int with_sequences, int64_t *sizes);
/* Addresses of the pack's arrays, to be wrapped in place (no copy):
 * ptrs[0] genome_ptr int64[n+1], [1] pham_id int32[nnz], [2] copies uint32[nnz],
 * [3] len_sum uint32[nnz], [4] seq_off int64[n_genes+1] or NULL, [5] residues uint8[] or NULL. */
int pcpy_buffers(void *pack, void **ptrs);
/* Release the pack (its residue arena is kept as a spare for the next pack). */
void pcpy_free(void *pack);

#ifdef __cplusplus
}
#endif
#endif /* PHAMCLUST_PYPACK_H */
