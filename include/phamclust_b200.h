/*
 * phamclust_b200.h -- C ABI of libphamclust_b200.so (sm_100a only).
 *
 * The reference (chg60/phamclust, pure Python) has no FFI; its boundary for
 * this path is one Python call,
 *
 *     matrix_de_novo(genomes, func, cpus, as_distance=True)
 *                                  /root/reference/src/phamclust/matrix.py:432
 *
 * with func one of the six metric functions of metrics.py (:26 gcs, :56 jc,
 * :83 pocp, :118 af, :178 aai, :235 peq).  This header declares exactly what a
 * ctypes/cffi stub for that call binds: load the packed
 * genome x pham table and the protein arena once, ask for a metric, get back
 * the upper triangle in 6-decimal micro-units.
 *
 * Conventions
 *  - every entry point returns 0 on success, a PC_E* code otherwise; the
 *    message is available from pc_last_error().  No exceptions cross the ABI.
 *  - plain pointers and sizes only.  "host" pointers are ordinary (pinned or
 *    pageable) host memory owned by the caller for the duration of the call;
 *    "dev" pointers are device memory on the ctx's GPU owned by the caller.
 *    The library owns whatever it allocates inside pc_ctx and never hands
 *    library-allocated memory out.
 *  - there is NO CPU fallback: pc_init fails without a usable sm_100 GPU.
 *  - one pc_ctx per process per GPU; calls on one ctx must be serialised by
 *    the caller.
 *
 * Packed upper triangle: entry (i, j), i <= j, of an N x N symmetric matrix
 * sits at  idx(i, j) = i*N - i*(i-1)/2 + (j - i)  (row-major, diagonal
 * included); rows [r0, r1) therefore form one contiguous slice.
 */
#ifndef PHAMCLUST_B200_H
#define PHAMCLUST_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pc_ctx pc_ctx;

/* error codes */
enum {
    PC_OK = 0,
    PC_EINVAL = 1,      /* bad argument / malformed table  -> ValueError   */
    PC_ECUDA = 2,       /* CUDA runtime failure            -> RuntimeError */
    PC_ENOGPU = 3,      /* no sm_100 device                -> RuntimeError */
    PC_ESTATE = 4,      /* call order (e.g. aai before pc_load_sequences)  */
    PC_ENOMEM = 5
};

/* metric ids, in the order of cli.METRICS (/root/reference/src/phamclust/cli.py:30-35) */
enum {
    PC_GCS = 0,         /* gene_content_similarity          metrics.py:26  */
    PC_JC = 1,          /* jaccard_coefficient              metrics.py:56  */
    PC_POCP = 2,        /* percentage_of_conserved_proteins metrics.py:83  */
    PC_AF = 3,          /* alignment_fraction               metrics.py:118 */
    PC_AAI = 4,         /* average_aminoacid_identity       metrics.py:178 */
    PC_PEQ = 5          /* proteomic_equivalence_quotient   metrics.py:235 */
};

/* contraction engines for the shared-pham count products (gcs/jc/pocp/af) */
enum {
    PC_ENGINE_AUTO = 0,
    PC_ENGINE_POPC = 1, /* bit-packed AND + POPC tile kernel                */
    PC_ENGINE_I8MMA = 2,/* int8 tcgen05.mma GEMM: persistent CTA pairs (cta_group::2),
                           double-buffered TMEM accumulators, device tile queue          */
    PC_ENGINE_I8MMA_V1 = 3 /* the round-1 kernel: one tile per CTA (comparisons only)    */
};

/* -------- lifetime ------------------------------------------------------ */
/* Replaces: the joblib worker pool set-up of matrix.py:471-472.            */
int pc_init(int device_id, pc_ctx **out);
void pc_destroy(pc_ctx *ctx);
/* ctx may be NULL: returns the last error of a failed pc_init.             */
const char *pc_last_error(const pc_ctx *ctx);
int pc_version(void);
/* external != 0: run all subsequent work of this ctx on the caller's CUDA
 * stream (e.g. torch.cuda.current_stream().cuda_stream; the handle 0 is the
 * legacy default stream); external == 0 restores the ctx's own stream.      */
int pc_set_stream(pc_ctx *ctx, void *cuda_stream, int external);
int pc_set_engine(pc_ctx *ctx, int engine);
/* Which global-alignment DP kernels aai/peq may use.  Both produce the same
 * (matched, columns, score) for every alignment; AUTO picks per alignment.  */
enum {
    PC_DP_AUTO = 0,     /* LEX kernel where its 32-bit word provably fits    */
    PC_DP_CLASSIC = 1   /* always the separate-counter kernel                */
};
int pc_set_dp_mode(pc_ctx *ctx, int mode);
int pc_sync(pc_ctx *ctx);

/* -------- inputs -------------------------------------------------------- */
/*
 * The genome x pham table in CSR form; replaces the list[Genome] argument of
 * matrix_de_novo (matrix.py:432) / Genome.phams (genome.py:30).  Row g of the
 * CSR is genome g IN CALLER ORDER (orientation source < target follows that
 * order, matrix.py:480-486).  Entry e carries: pham_id[e] in [0, n_phams);
 * copies[e] = len(genome[pham]) >= 1 (genome.py:45-48); len_sum[e] = summed
 * translation length of those copies (metrics.py:136-147).  pham ids must be
 * unique within a row; all pointers are host memory.
 */
int pc_load_table(pc_ctx *ctx, int64_t n_genomes, int64_t n_phams,
                  const int64_t *genome_ptr /* n_genomes+1 */,
                  const int32_t *pham_id /* nnz */,
                  const uint32_t *copies /* nnz */,
                  const uint32_t *len_sum /* nnz */);

/*
 * The translations; replaces the list[str] values of Genome.phams consumed by
 * metrics.py:203-221.  Genes are laid out in CSR entry order: entry e owns
 * genes [first(e), first(e)+copies[e]) with first = exclusive cumsum(copies),
 * paralogs in the order they were added (genome.py:45-48; the order decides
 * the "last maximum wins" tie of metrics.py:223).  residues are the raw
 * characters (one byte each, as in the TSV/FASTA); seq_off has n_genes+1
 * entries.  Needs pc_load_table first.  Host memory.
 */
int pc_load_sequences(pc_ctx *ctx, int64_t n_genes, const int64_t *seq_off,
                      const uint8_t *residues);

/*
 * Several GPUs: the same arena, but only the genes of the phams that partition
 * `part` of `n_parts` aligns (the deal of pc_align_accumulate: phams in
 * descending cost order, snake order over the partitions) are shipped, encoded
 * and canonicalised -- 1 / n_parts of the bytes and of the host work.  The
 * partition is fixed until the next load; pc_alignment_identities needs the
 * whole arena (pc_load_sequences).
 */
int pc_load_sequences_part(pc_ctx *ctx, int64_t n_genes, const int64_t *seq_off,
                           const uint8_t *residues, int part, int n_parts);

/* -------- the path ------------------------------------------------------ */
/*
 * matrix_de_novo (matrix.py:432-497) for the whole table: writes
 * N(N+1)/2 int32 micro-units, value = round(weight, 6) * 1e6 exactly as
 * metrics.py:50-53 + matrix.py:323 produce it; diagonal = 1 - as_distance
 * (matrix.py:467-468).  ppos is metrics.py:178's flag (aai only, else 0).
 * out is host memory unless out_on_device != 0.
 */
int pc_matrix(pc_ctx *ctx, int metric, int as_distance, int ppos,
              int32_t *out_micro_upper, int out_on_device);

/*
 * The same for genome rows [row_begin, row_end) only (a contiguous slice of
 * the packed triangle) -- the unit of multi-GPU sharding.  For PC_AAI/PC_PEQ
 * acc_dev must hold the (all-reduced) accumulators filled by
 * pc_align_accumulate; it is ignored for the other metrics.
 */
int pc_matrix_rows(pc_ctx *ctx, int metric, int as_distance, int ppos,
                   int64_t row_begin, int64_t row_end,
                   const uint64_t *acc_dev, int32_t *out, int out_on_device);

/*
 * Stage 1+2 of aai/peq (metrics.py:203-225): align every within-pham member
 * pair owned by partition `part` of `n_parts` (phams are dealt round-robin in
 * descending-cost order) and add each anchor gene's best match into
 * acc_dev[idx(i,j)] as  (sum matched columns << 32) | sum alignment length.
 * acc_dev is device memory of N(N+1)/2 uint64, zeroed by the caller; sums of
 * partitions (ncclAllReduce, integer, order independent) give the full
 * accumulators.
 */
int pc_align_accumulate(pc_ctx *ctx, int ppos, int part, int n_parts,
                        uint64_t *acc_dev);

/*
 * Sparse exchange for several GPUs: instead of a dense
 * N(N+1)/2 accumulator per rank and an all-reduce, every (pham, genome pair)
 * unit of partition `part` becomes one 16-byte record
 *     { index of (i, j) inside the row block of its owner, (matched << 32) | length }
 * in the bucket of the destination that owns genome row i:
 * destination d owns rows [row_bounds[d], row_bounds[d+1]), indices count from
 * idx(row_bounds[d], row_bounds[d]).  records_dev (device, 2 x uint64 per
 * record, pc_align_units() records at most) is filled bucket after bucket;
 * dest_counts[d] (host) = records of bucket d.  The caller exchanges the
 * buckets (ncclSend/Recv all-to-all) and every owner adds what it received with
 * pc_accumulate_records into the accumulators of ITS rows (n_acc uint64, zeroed
 * by the caller), which pc_matrix_rows_owned then turns into the row block.
 * Integer sums: exact, independent of arrival order.
 */
int pc_align_units(pc_ctx *ctx, int part, int n_parts, int64_t *n_units);
int pc_align_contributions(pc_ctx *ctx, int ppos, int part, int n_parts,
                           int n_dest, const int64_t *row_bounds /* n_dest+1 */,
                           uint64_t *records_dev, int64_t *dest_counts /* n_dest */);
int pc_accumulate_records(pc_ctx *ctx, int64_t n_records,
                          const uint64_t *records_dev, int64_t n_acc,
                          uint64_t *acc_rows_dev);
/* pc_matrix_rows with accumulators that hold ONLY rows [row_begin, row_end)
 * (entry (row_begin, row_begin) first): what a rank owns after the exchange. */
int pc_matrix_rows_owned(pc_ctx *ctx, int metric, int as_distance, int ppos,
                         int64_t row_begin, int64_t row_end,
                         const uint64_t *acc_rows_dev, int32_t *out,
                         int out_on_device);

/*
 * The triangle kept in HBM for its first consumer.  What the pipeline does
 * next is thresholded clustering (scripts/phamclust.py:300-317, clustering.py:27); these calls
 * answer its questions without shipping N(N+1)/2 values: pc_matrix_resident computes into a
 * ctx-owned device buffer (valid until the next pc_matrix_resident / pc_load_table);
 * pc_resident_threshold_edges returns the off-diagonal pairs (i < j, storage order) with weight
 * <= max_micro as int32 pairs -- *n_edges is the number FOUND, which may exceed cap (call again
 * with more room); pc_resident_gather reads entries by packed index; pc_resident_fetch copies the
 * whole triangle after all.  Host pointers.
 */
int pc_matrix_resident(pc_ctx *ctx, int metric, int as_distance, int ppos);
int pc_resident_threshold_edges(pc_ctx *ctx, int32_t max_micro, int64_t cap,
                                int32_t *edges_host /* 2 * cap */, int64_t *n_edges);
int pc_resident_gather(pc_ctx *ctx, int64_t n_idx, const int64_t *tri_index_host,
                       int32_t *out_host);
int pc_resident_fetch(pc_ctx *ctx, int32_t *out_host);

/*
 * pc_matrix_rows_owned with the gather of the row blocks FUSED into the kernel that produces them:
 * every value is stored into out_dev (this rank's triangle, slice starting at idx(row_begin,
 * row_begin)) and into the same slice of n_peers other triangles (peer_out_dev[k]: device
 * addresses of the other ranks' buffers mapped into this process -- CUDA IPC / symmetric memory;
 * the stores cross NVLink while the next tile is being multiplied).  After every rank's call and a
 * barrier, every rank holds the whole triangle; no separate all-gather.  gcs/jc/pocp/af need the
 * default tensor engine for this.
 */
int pc_matrix_rows_fused(pc_ctx *ctx, int metric, int as_distance, int ppos,
                         int64_t row_begin, int64_t row_end,
                         const uint64_t *acc_rows_dev, int32_t *out_dev,
                         int n_peers, int32_t *const *peer_out_dev);

/* -------- integer test hooks (bit-exact parity) ------------------------- */
/*
 * Numerators the reference computes as Python ints, packed triangle, host:
 *  PC_GCS/PC_JC: len(source & target)            genome.py:124-135
 *  PC_POCP:      s_conserved + t_conserved       metrics.py:102-104
 *  PC_AF:        s_conserved + t_conserved (aa)  metrics.py:135-149
 *  PC_AAI:       out_num = sum matched columns, out_aux = sum aln length
 * out_aux may be NULL unless metric == PC_AAI.
 */
int pc_numerators(pc_ctx *ctx, int metric, int ppos, int32_t *out_num,
                  int32_t *out_aux);

/*
 * _align (metrics.py:160-175) + the counts of :216-220 for caller-chosen
 * ordered gene pairs: out[4*t..] = identical '|' columns, similar '+'
 * columns, len(traceback.query), alignment score.  Host memory.
 */
int pc_alignment_identities(pc_ctx *ctx, int64_t n_tasks,
                            const int32_t *query_gene, const int32_t *ref_gene,
                            int32_t *out /* n_tasks * 4 */);

/*
 * The device's round(x, 6) (metrics.py:50-53, matrix.py:323) applied to n
 * host doubles -> int32 micro-units; lets the tests pin the epilogue's
 * rounding against CPython's round() on the GPU itself.
 */
int pc_debug_round6(pc_ctx *ctx, int64_t n, const double *x, int32_t *out);

/*
 * The fused epilogue of gcs/jc/pocp/af (metrics.py:45-53, 75-80, 110-115, 152-157 +
 * matrix.py:323) on caller-chosen operands: out = the production path (integer arithmetic, fp64
 * only on exact ties), out_f64 = the fp64 path that restates CPython's arithmetic; the tests
 * compare both with CPython itself.  Host memory.
 */
int pc_debug_metric_micro(pc_ctx *ctx, int metric, int as_distance, int64_t n,
                          const int32_t *numer, const int32_t *den_i,
                          const int32_t *den_j, int32_t *out, int32_t *out_f64);

/* -------- measurement --------------------------------------------------- */
enum {
    PC_T_H2D = 0,       /* host -> device copies                            */
    PC_T_PACK = 1,      /* operand packing kernels                          */
    PC_T_CONTRACT = 2,  /* shared-count contraction (+ fused epilogue)      */
    PC_T_ALIGN = 3,     /* global-alignment DP kernels                      */
    PC_T_GATHER = 4,    /* best-match gather into pair accumulators         */
    PC_T_EPILOGUE = 5,  /* standalone normalise/round kernels               */
    PC_T_D2H = 6,       /* device -> host copy of the result                */
    PC_T_HOST_TABLE = 7,/* host wall-clock inside pc_load_table             */
    PC_T_HOST_SEQS = 8, /* host wall-clock inside pc_load_sequences         */
    PC_T_HOST_PLAN = 9, /* host wall-clock planning the alignment work      */
    PC_T_COUNT = 10
};
/* Milliseconds per stage accumulated since pc_reset_counters: CUDA events on
 * the ctx stream for the device stages, wall clock for the PC_T_HOST_* ones
 * (which CONTAIN the device stages issued from inside them).                */
int pc_timings(pc_ctx *ctx, double *ms, int n);
/* kernels launched / DP cells updated / alignments run since the reset.    */
int pc_counters(pc_ctx *ctx, int64_t *launches, int64_t *dp_cells,
                int64_t *alignments, int64_t *contract_launches,
                int64_t *align_launches);
int pc_reset_counters(pc_ctx *ctx);
enum {
    PC_STAT_LOGICAL_ALIGNMENTS = 0, /* _align calls the reference would make (metrics.py:215)   */
    PC_STAT_UNIQUE_GENES = 1,       /* distinct translations among the loaded genes              */
    PC_STAT_OPERAND_COLUMNS = 2,    /* phams held by >= 2 genomes (contraction K)                */
    PC_STAT_GENES = 3,
    PC_STAT_LEX_ALIGNMENTS = 4,     /* DP tasks run by the LEX kernels since pc_reset_counters   */
    PC_STAT_LEX_CELLS = 5,          /* DP cells of the plain LEX kernel (pc_counters' dp_cells has all) */
    PC_STAT_DP_CONFIGS_USED = 6,    /* distinct DP kernel configs launched since the reset        */
    PC_STAT_DP_CONFIGS = 7,         /* DP kernel configs the chooser can pick for the loaded data */
    PC_STAT_I8_TILE_STEPS = 8,      /* int8 engine: 128x256x128 MMA groups issued since the reset
                                       (the block-sparse K schedule skips the rest)              */
    PC_STAT_I8_TILE_STEPS_DENSE = 9,/* ... and what a dense K schedule would have issued          */
    PC_STAT_LEXR_CELLS = 10,        /* DP cells run by the rebased LEX kernel (10-bit counters)   */
    PC_STAT_LEXC_CELLS = 11,        /* DP cells run by the LEXC kernel (rebased counters, chunks)  */
    PC_STAT_LEXC_FALLBACKS = 12,    /* alignments LEXC handed to the classic kernel (band check)  */
    PC_STAT_H2D_BYTES = 13          /* bytes of the table / sequence loads shipped host -> device  */
};
int pc_stat(pc_ctx *ctx, int which, int64_t *out);

/* -------- host-side text formats (no GPU) -------- */
/*
 * Writers of matrix.py:517-533 (adjacency), :624-651 (squareform / lower
 * triangle) straight from the packed micro-unit triangle; byte-identical
 * output ('%.6f').  order[k] = storage index of the k-th node in traversal
 * order; names = blob + offsets in storage order.  mode 0: squareform,
 * 1: lower triangle (the pipeline's cache, scripts/phamclust.py:254-262),
 * 2: adjacency, 3: adjacency with skip_zero.
 */
int pc_write_matrix(const char *path, int64_t n, const int32_t *micro,
                    const int64_t *order, const char *names,
                    const int64_t *name_off, int mode);
/* Reader of matrix.py:564-621 over a file already in memory. */
int pc_parse_squareform(const char *buf, int64_t len, int64_t n,
                        int32_t *micro_out, int64_t *name_begin,
                        int64_t *name_end);
/* Field scanner for the 2/3-column TSV of scripts/phamclust.py:21-47. */
int64_t pc_scan_tsv(const char *buf, int64_t len, int64_t max_rows,
                    int64_t *fields);
/*
 * TSV -> packed table helpers: intern the name (col 0) or pham (col 1) column of the
 * scanned rows in first-seen order (ids[r], first_row[id]; returns the number of
 * distinct strings or -1), and concatenate byte ranges into the residue arena
 * (start < 0 -> the single byte `fill`, the "M" of 2-column rows).
 */
int64_t pc_tsv_intern(const char *buf, int64_t rows, const int64_t *fields, int col,
                      int32_t *ids, int64_t *first_row);
int pc_gather_fields(const char *buf, int64_t n, const int64_t *start,
                     const int64_t *len, uint8_t fill, uint8_t *out);
/*
 * Dense square view of the packed triangle in traversal order: SymMatrix.to_ndarray
 * (matrix.py:155-167), the array clustering.py:27 hands to scikit-learn.  kind 0:
 * float64 micro / 1e6 (PC_EINVAL if an entry is unset / outside [0, 1e6]); kind 1:
 * the int32 micro-units.  out holds n * n elements.
 */
int pc_dense_from_micro(int64_t n, const int32_t *micro, const int64_t *order,
                        int kind, void *out);
/*
 * SymMatrix.medoid / anti_medoid (matrix.py:81-127) without a dense N x N: per node (storage
 * order) the exact integer sum, in micro-units, of every weight it appears in while the upper
 * half is iterated -- the diagonal twice, as upstream.  PC_EINVAL on an unset weight.
 */
int pc_node_sums(int64_t n, const int32_t *micro, int64_t *sums);
/*
 * One single-genome FASTA file in memory -> (pham, translation) records without building Genome
 * dicts (fasta.py:4-40, genome.py:66-83, scripts/phamclust.py:50-83): pham names appended to
 * name_out, rstrip()ped and concatenated sequence lines to seq_out, fields[4r..4r+3] = name
 * begin/end, sequence begin/end; used[0..1] = bytes written.  Returns the number of records, or
 * -1 (no leading '>'), -2 (header field without exactly one '='), -3 (no pham= field), -4 (out of
 * capacity), -5 (non-ASCII sequence byte).
 */
int64_t pc_scan_fasta(const char *buf, int64_t len, int64_t max_records, int64_t *fields,
                      uint8_t *name_out, int64_t name_cap, uint8_t *seq_out, int64_t seq_cap,
                      int64_t *used);
const char *pc_io_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* PHAMCLUST_B200_H */
