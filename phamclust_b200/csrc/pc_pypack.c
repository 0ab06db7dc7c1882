/* pc_pypack.c -- list[Genome] -> CSR table + residue arena in ONE walk of the Python
 * objects, in C (CPython C API, loaded with ctypes.PyDLL so the GIL is held).
 *
 * Replaces the per-entry Python steps of packing.pack_genomes on the critical path of
 * matrix_de_novo(genomes, func, cpus)  (/root/reference/src/phamclust/matrix.py:432): the
 * reference hands every worker two whole Genome objects per pair (matrix.py:483-488); here
 * the dicts are read once per matrix.  Genome layout: genome.py:30 (`phams`: dict pham ->
 * list of translations, paralogs in insertion order, genome.py:45-48).
 *
 * The str / list / dict objects of 2000 genomes are ~600 000 separate heap objects: touching
 * them is what this pass costs, so they are touched exactly once:
 *   pcpy_pack(genomes, vocab, with_seqs, sizes[4]) -> opaque handle (NULL + Python exception on
 *       error); walks the objects once, into growable C buffers; `vocab` (an empty dict)
 *       receives pham name -> id in first-seen order, exactly the ids the pure Python packer
 *       hands out; sizes = n_genomes, nnz, n_genes, n_residues
 *   pcpy_buffers(handle, ptrs[6])  addresses of genome_ptr, pham_id, copies, len_sum, seq_off,
 *       residues: the caller wraps them as numpy arrays WITHOUT copying (single buffer hand-off)
 *       and calls pcpy_free(handle) when the last of those arrays is gone.
 *   The first capacities are the sizes of the previous pack, so a repeated call never reallocs.
 * Translations must be str made of single-byte (latin-1) characters: anything wider raises
 * ValueError instead of being replaced by '?' (two different wide characters would compare
 * equal in the alignment kernel).
 */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static PyObject *phams_of(PyObject *genome)
{
    static PyObject *attr = NULL;
    if (!attr) {
        attr = PyUnicode_InternFromString("phams");
        if (!attr) return NULL;
    }
    PyObject *phams = PyObject_GetAttr(genome, attr);     /* new reference */
    if (!phams) return NULL;
    if (!PyDict_Check(phams)) {
        Py_DECREF(phams);
        PyErr_SetString(PyExc_TypeError, "genome.phams should be a dict (pham -> list of translations)");
        return NULL;
    }
    return phams;
}

static int bad_pham(PyObject *pham)
{
    PyErr_Format(PyExc_TypeError, "type(pham) should be 'str', not '%s'", Py_TYPE(pham)->tp_name);
    return -1;
}

typedef struct {
    int64_t n, nnz, n_genes, n_res;
    int64_t cap_e, cap_g, cap_r;
    int64_t *genome_ptr;
    int32_t *pham_id;
    uint32_t *copies, *len_sum;
    int64_t *seq_off;
    uint8_t *residues;
} Pack;

/* Fresh pages are the expensive part of a 40 MB arena (one soft fault per 4 KB page, several
 * microseconds each under a hypervisor): the residue buffer of the last freed pack is kept as
 * a spare and handed to the next pack, so repeated matrix builds touch no new memory. */
static uint8_t *spare_res = NULL;
static int64_t spare_res_cap = 0;

void pcpy_free(void *h)
{
    Pack *p = (Pack *)h;
    if (!p) return;
    free(p->genome_ptr); free(p->pham_id); free(p->copies); free(p->len_sum);
    free(p->seq_off);
    if (p->residues && p->cap_r > spare_res_cap) {
        free(spare_res);
        spare_res = p->residues;
        spare_res_cap = p->cap_r;
    } else {
        free(p->residues);
    }
    free(p);
}

static int grow(void **buf, int64_t *cap, int64_t need, size_t elem)
{
    if (need <= *cap) return 0;
    int64_t c = *cap ? *cap : 1024;
    while (c < need) c += c / 2 + 1024;
    void *q = realloc(*buf, (size_t)c * elem);
    if (!q) { PyErr_NoMemory(); return -1; }
    *buf = q;
    *cap = c;
    return 0;
}

void *pcpy_pack(PyObject *genomes, PyObject *vocab, int with_seqs, int64_t *sizes)
{
    if (!PyList_Check(genomes) || !PyDict_Check(vocab)) {
        PyErr_SetString(PyExc_TypeError, "pcpy_pack(list, dict, ...)");
        return NULL;
    }
    Pack *p = (Pack *)calloc(1, sizeof(Pack));
    if (!p) { PyErr_NoMemory(); return NULL; }
    const Py_ssize_t n = PyList_GET_SIZE(genomes);
    p->n = n;
    p->genome_ptr = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n + 1));
    if (!p->genome_ptr) { PyErr_NoMemory(); pcpy_free(p); return NULL; }
    p->genome_ptr[0] = 0;
    static int64_t last_nnz = 0, last_genes = 0, last_res = 0;   /* sizes of the previous pack (GIL held) */
    int64_t cap_c = 0, cap_l = 0, cap_off = 0;
    if (grow((void **)&p->pham_id, &p->cap_e, last_nnz + 1, sizeof(int32_t)) ||
        grow((void **)&p->copies, &cap_c, last_nnz + 1, sizeof(uint32_t)) ||
        grow((void **)&p->len_sum, &cap_l, last_nnz + 1, sizeof(uint32_t))) { pcpy_free(p); return NULL; }
    if (with_seqs) {
        if (spare_res && spare_res_cap >= last_res + 1) {      /* already-touched pages */
            p->residues = spare_res;
            p->cap_r = spare_res_cap;
            spare_res = NULL;
            spare_res_cap = 0;
        }
        if (grow((void **)&p->seq_off, &cap_off, last_genes + 1, sizeof(int64_t)) ||
            grow((void **)&p->residues, &p->cap_r, last_res + 1, 1)) { pcpy_free(p); return NULL; }
        p->seq_off[0] = 0;
    }
    for (Py_ssize_t g = 0; g < n; ++g) {
        PyObject *phams = phams_of(PyList_GET_ITEM(genomes, g));
        if (!phams) { pcpy_free(p); return NULL; }
        const int64_t need_e = p->nnz + PyDict_GET_SIZE(phams);
        if (grow((void **)&p->pham_id, &p->cap_e, need_e, sizeof(int32_t)) ||
            grow((void **)&p->copies, &cap_c, need_e, sizeof(uint32_t)) ||
            grow((void **)&p->len_sum, &cap_l, need_e, sizeof(uint32_t))) {
            Py_DECREF(phams); pcpy_free(p); return NULL;
        }
        Py_ssize_t pos = 0;
        PyObject *pham, *genes;
        while (PyDict_Next(phams, &pos, &pham, &genes)) {
            if (!PyUnicode_Check(pham)) { bad_pham(pham); goto fail; }
            if (!PyList_Check(genes)) {
                PyErr_SetString(PyExc_TypeError, "genome.phams values should be lists of translations");
                goto fail;
            }
            PyObject *id = PyDict_GetItemWithError(vocab, pham);       /* borrowed */
            long pid;
            if (id) {
                pid = PyLong_AsLong(id);
            } else {
                if (PyErr_Occurred()) goto fail;
                pid = (long)PyDict_GET_SIZE(vocab);
                PyObject *fresh = PyLong_FromLong(pid);
                if (!fresh || PyDict_SetItem(vocab, pham, fresh) < 0) { Py_XDECREF(fresh); goto fail; }
                Py_DECREF(fresh);
            }
            const Py_ssize_t c = PyList_GET_SIZE(genes);
            if (with_seqs && grow((void **)&p->seq_off, &cap_off, p->n_genes + c + 1, sizeof(int64_t))) goto fail;
            int64_t total = 0;
            for (Py_ssize_t k = 0; k < c; ++k) {
                PyObject *t = PyList_GET_ITEM(genes, k);
                if (!PyUnicode_Check(t)) {
                    PyErr_Format(PyExc_TypeError, "type(translation) should be 'str', not '%s'",
                                 Py_TYPE(t)->tp_name);
                    goto fail;
                }
                const Py_ssize_t len = PyUnicode_GET_LENGTH(t);
                if (with_seqs) {
                    if (len && PyUnicode_KIND(t) != PyUnicode_1BYTE_KIND) {
                        PyErr_SetString(PyExc_ValueError,
                                        "translations must be made of single-byte (latin-1) characters");
                        goto fail;
                    }
                    if (grow((void **)&p->residues, &p->cap_r, p->n_res + len, 1)) goto fail;
                    memcpy(p->residues + p->n_res, PyUnicode_1BYTE_DATA(t), (size_t)len);
                    p->seq_off[p->n_genes + k + 1] = p->n_res + len;
                }
                p->n_res += len;
                total += len;
            }
            p->n_genes += c;
            p->pham_id[p->nnz] = (int32_t)pid;
            p->copies[p->nnz] = (uint32_t)c;
            p->len_sum[p->nnz] = (uint32_t)total;
            p->nnz++;
            continue;
        fail:
            Py_DECREF(phams);
            pcpy_free(p);
            return NULL;
        }
        Py_DECREF(phams);
        p->genome_ptr[g + 1] = p->nnz;
    }
    sizes[0] = p->n; sizes[1] = p->nnz; sizes[2] = p->n_genes; sizes[3] = p->n_res;
    last_nnz = p->nnz; last_genes = p->n_genes; last_res = with_seqs ? p->n_res : last_res;
    return p;
}

int pcpy_buffers(void *h, void **ptrs)
{
    Pack *p = (Pack *)h;
    if (!p) { PyErr_SetString(PyExc_ValueError, "pcpy_buffers: no pack"); return -1; }
    ptrs[0] = p->genome_ptr; ptrs[1] = p->pham_id; ptrs[2] = p->copies; ptrs[3] = p->len_sum;
    ptrs[4] = p->seq_off; ptrs[5] = p->residues;
    return 0;
}
