// pc_common.cuh -- shared declarations of libphamclust_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>

#include "../../include/phamclust_b200.h"
#include "nw_rules.h"

#define PC_NEG_INF (-(1 << 28))

// ---------------------------------------------------------------------------
// error plumbing: no exceptions across the C ABI
// ---------------------------------------------------------------------------
struct pc_ctx;
int pc_fail(pc_ctx *ctx, int code, const char *fmt, ...);

#define PC_CUDA(ctx, expr)                                                     \
    do {                                                                       \
        cudaError_t _e = (expr);                                               \
        if (_e != cudaSuccess)                                                 \
            return pc_fail((ctx), PC_ECUDA, "%s failed at %s:%d: %s", #expr,   \
                           __FILE__, __LINE__, cudaGetErrorString(_e));        \
    } while (0)

#define PC_TRY(expr)                                                           \
    do {                                                                       \
        int _rc = (expr);                                                      \
        if (_rc != PC_OK) return _rc;                                          \
    } while (0)

// ---------------------------------------------------------------------------
// a device buffer owned by the ctx (grow-only)
// ---------------------------------------------------------------------------
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    template <typename T> T *as() const { return reinterpret_cast<T *>(p); }
};
int devbuf_reserve(pc_ctx *ctx, DevBuf &b, size_t bytes);
void devbuf_free(DevBuf &b);

// ---------------------------------------------------------------------------
// kernel argument blocks
// ---------------------------------------------------------------------------
#define PC_MAX_PEER_OUT 15        // other ranks of the node whose triangles receive the row block

struct PopcTerm {
    const uint32_t *a;   // bit rows read for the tile's ROW genomes
    const uint32_t *b;   // bit rows read for the tile's COLUMN genomes
    int mult;            // weight of this term (2^plane)
};

struct EpilogueArgs {
    const int32_t *den;  // per-genome denominator part (n_phams / n_genes / total_len)
    int metric;          // PC_GCS .. PC_AF
    int as_distance;
    int raw;             // 1: write the integer numerator instead of micro-units
    int64_t n;           // number of genomes
    int64_t row_begin;   // first genome row covered by `out`
    int64_t row_end;
    int32_t *out;        // packed triangle slice starting at idx(row_begin,row_begin)
    // fused all-gather: the same slice in the triangles of the other ranks (peer-mapped
    // addresses, stores travel over NVLink); honoured by the pair engine and the aai epilogue
    int n_peer_out;
    int32_t *peer_out[PC_MAX_PEER_OUT];
};

struct NwArgs {
    const uint8_t *codes;        // residue codes, genes 16B aligned
    const int64_t *gene_off;     // byte offset of each gene in codes
    const int32_t *gene_len;
    const int32_t *task_q;       // query (s1) gene per task
    const int32_t *task_r;       // ref   (s2) gene per task
    const int64_t *perm;         // processing order -> task id (nullable)
    int64_t n_tasks;             // tasks in this launch
    int64_t task_base;           // added to the processing index
    uint32_t *out_hitlen;        // (matched columns << 16) | alignment length
    int32_t *out_score;          // nullable
    int4 *spill;                 // multi-chunk boundary column, per warp
    int64_t spill_stride;        // int4 elements per warp
    unsigned long long *work_counter;
    unsigned long long *cell_counter;
    const int2 *subtab;          // [32*32] {score, 1 + (hit << 16)}
    int gap_open, gap_extend;
    // LEXC: alignments whose band checks failed are listed here (slots) and re-run by the classic
    // multi-chunk kernel, which then takes its task count from the device (n_tasks_dev)
    unsigned long long *fallback_count;
    int64_t *fallback_slots;
    const unsigned long long *n_tasks_dev;   // nullable; overrides n_tasks (n_tasks still sizes the grid)
};

struct UnitArgs {
    // per pham (processing order of this partition)
    const int64_t *ph_mem_off;   // [n_ph+1] into the member arrays
    const int64_t *ph_unit_off;  // [n_ph+1]
    int64_t n_ph;
    int64_t n_units;
    // per member, grouped by pham, ascending genome index
    const int32_t *mem_genome;
    const int32_t *mem_first_gene;
    const int32_t *mem_copies;
    const int32_t *gene_len;
    const int32_t *gene_hi;      // per-gene score bound; NULL = never use the LEX DP kernel
    int coarse_k;                // 1: the 32-code table is in use (K in multiples of 4 only)
    int no_lexc;                 // 1: LEXC gave up on too many alignments of this data before
    int64_t n_genomes;
    // device-side bucketing of the units by DP kernel config
    int64_t units_per_block;     // contiguous units owned by one classify/scatter block
    int n_blocks;
    unsigned long long *blk_counts;   // [n_blocks][PC_NW_NCFG] tasks, then exclusive offsets
    unsigned long long *cfg_total;    // [PC_NW_NCFG] tasks per config
    unsigned long long *cfg_base;     // [PC_NW_NCFG] first task of each config
    // task table (open addressing): key = canon_q << 32 | canon_r
    const int32_t *canon;             // gene -> canonical gene (same translation)
    unsigned long long *ht_keys;
    unsigned long long *ht_vals;
    uint64_t ht_mask;                 // capacity - 1 (power of two)
};

// destination row blocks of the contribution records (one per rank of the node, at most)
#define PC_MAX_DEST 64
struct DestArgs {
    int n_dest;
    int64_t bounds[PC_MAX_DEST + 1];   // block d owns genome rows [bounds[d], bounds[d+1])
};

// ---------------------------------------------------------------------------
// the context
// ---------------------------------------------------------------------------
struct StageTimer {
    cudaEvent_t start, stop;
    int stage;
};

struct AlignPlan {           // host+device prep for one (part, n_parts)
    int part = -1, n_parts = -1;
    int64_t n_ph = 0, n_units = 0, n_tasks = 0, n_mem = 0;
    int32_t max_gene_len = 0;
    DevBuf ph_mem_off, ph_unit_off;
    DevBuf mem_genome, mem_first_gene, mem_copies;
    DevBuf blk_counts, cfg_info, ht_keys, ht_vals;
};

struct pc_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t copy_stream = nullptr;   // D2H of finished row blocks, overlapped with compute
    cudaStream_t dp_streams[8] = {};             // DP configs run side by side
    cudaStream_t stream = nullptr;
    std::string err;
    int engine = PC_ENGINE_AUTO;

    // ---- table (host copies kept for the host-side planning) ----
    bool have_table = false, have_seqs = false;
    int64_t N = 0, P = 0, nnz = 0;
    int64_t P_eff = 0;                    // phams held by >= 2 genomes (the only ones that can be shared)
    std::vector<int64_t> h_genome_ptr;
    std::vector<int32_t> h_pham_id;
    std::vector<uint32_t> h_copies, h_len_sum;
    std::vector<int32_t> h_entry_genome;
    std::vector<int64_t> h_entry_first_gene;
    std::vector<int64_t> h_pham_ptr;      // CSC
    // pc_load_table scratch, kept between calls (no fresh pages per load)
    std::vector<int64_t> h_first_entry;
    std::vector<int32_t> h_den[3], h_pham_col;
    std::vector<std::vector<int32_t>> h_stamp;   // per host thread: epoch + genome that last listed the pham
    int32_t stamp_epoch = 0;
    std::vector<int32_t> h_pham_part;     // pham -> partition that aligns it (-1: held by < 2 genomes)
    std::vector<int64_t> h_deal_order;    // phams held by >= 2 genomes, descending cost
    int deal_parts = -1;                  // n_parts the deal above was made for
    int seq_part = 0, seq_parts = 1;      // which partition's genes pc_load_sequences[_part] shipped
    void *h_stage = nullptr;              // page-locked staging of the owned residues (partial loads)
    size_t h_stage_cap = 0;
    std::vector<int32_t> h_gene_len_dev;  // partial loads: gene lengths as the device sees them (0 = not shipped)
    std::vector<int64_t> h_raw_off;       // partial loads: offsets into the staged residues
    uint32_t max_copies = 0, max_len_sum = 0;
    int64_t n_genes_total = 0;

    // ---- device: table ----
    DevBuf d_entry_genome, d_pham_id, d_copies, d_len_sum;
    DevBuf d_pham_col;                             // pham id -> operand column, -1 = dropped
    DevBuf d_den_phams, d_den_genes, d_den_len;   // int32 [N]
    int64_t W = 0;                                 // words per bit row (multiple of 32)
    DevBuf d_bits;                                 // presence bit rows [N][W] (popc engine, lazy)
    bool bits_ready = false;
    DevBuf d_planes;                               // weight bit planes
    int planes_kind = -1, planes_count = 0;        // what d_planes currently holds
    DevBuf d_terms;
    DevBuf d_numer;                                // int32 scratch triangle slice
    // int8 operand for the tcgen05 engine
    DevBuf d_i8;                                   // [N_pad][K_pad]
    DevBuf d_i8_occ;                               // [N_pad / 128][occ_words] K-block occupancy bits
    int64_t i8_dense_steps = 0;                    // ... and what the dense K schedule would have issued
    DevBuf d_i8_steps;                             // u64: 128 x 256 x 128 MMA groups issued since the reset
    int i8_kind = -1; int64_t i8_K = 0, i8_rows = 0;
    DevBuf d_p2_tiles, d_p2_masks, d_p2_keys, d_p2_ids, d_p2_temp, d_p2_sync;   // tile queue of the persistent pair kernel

    // ---- sequences ----
    int64_t n_genes = 0;
    std::vector<int32_t> h_gene_len;
    std::vector<int64_t> h_gene_off;               // 16 B aligned offsets in the code arena
    int32_t max_gene_len = 0;
    DevBuf d_codes, d_gene_off, d_gene_len, d_raw, d_raw_off, d_lut;
    DevBuf d_canon;                                // gene -> smallest gene with the same translation
    DevBuf d_gene_hi;                              // per-gene upper bound of any DP score (LEX eligibility)
    std::vector<int32_t> h_gene_hi;                // host copy, fetched on demand (test hook)
    int dp_mode = PC_DP_AUTO;
    DevBuf d_dedup[8];                             // scratch of pc_build_canonical_genes (kept: cudaFree stalls)
    int64_t n_unique_genes = 0;
    DevBuf d_subtab[2];                            // [0] identity, [1] ppos
    int code_of_byte[256];
    int n_codes = 0;
    int code_matrix_index[32];

    // ---- aai work buffers ----
    AlignPlan plan;
    DevBuf d_task_q, d_task_r, d_hitlen, d_score, d_spill, d_counters, d_perm, d_fallback;
    int64_t h2d_bytes = 0;           // bytes handed to H2D copies since pc_reset_counters (the table / sequence loads)
    bool lexc_off = false;           // LEXC fell back too often on the loaded sequences
    int64_t lexc_last_tasks = 0;     // LEXC alignments of the previous align_tasks call
    DevBuf d_acc;                                  // uint64 accumulators (single-GPU path)
    DevBuf d_dest;                                 // [2][PC_MAX_DEST] record counts / write cursors per destination
    DevBuf d_out;                                  // int32 result staging
    DevBuf d_resident, d_res_edges, d_res_idx, d_res_vals;   // the triangle kept in HBM + its query scratch
    int n_peer_out = 0;                            // pc_matrix_rows_fused: peers' slices for the call in flight
    int32_t *peer_out[PC_MAX_PEER_OUT] = {};
    int64_t resident_n = -1;                       // genomes of the resident triangle (-1: none)

    // ---- measurement ----
    std::vector<StageTimer> timers;
    std::vector<cudaEvent_t> event_pool;
    std::vector<cudaEvent_t> deferred_events;      // in flight; recycled after the next stream sync
    double stage_ms[PC_T_COUNT] = {0};
    int64_t launches = 0, contract_launches = 0, align_launches = 0;
    int64_t dp_cells = 0, alignments = 0, logical_alignments = 0, lex_alignments = 0;
    std::vector<char> cfg_used = std::vector<char>(512, 0);   // [config]: launched since the reset
};

// stage timing helpers (events on ctx->stream)
int pc_stage_begin(pc_ctx *ctx, int stage);
int pc_stage_end(pc_ctx *ctx);
int pc_collect_timers(pc_ctx *ctx);

// ---------------------------------------------------------------------------
// launchers implemented in the other translation units
// ---------------------------------------------------------------------------
int pc_launch_pack_bits(pc_ctx *ctx, int source /*0 presence,1 copies,2 len*/,
                        int plane, uint32_t *dst);
int pc_launch_popc_contract(pc_ctx *ctx, const PopcTerm *d_terms, int n_terms,
                            const EpilogueArgs &ep);
int pc_launch_nw(pc_ctx *ctx, const NwArgs &args, int cfg, cudaStream_t stream);
int pc_launch_nw_fallback(pc_ctx *ctx, const NwArgs &args, cudaStream_t stream);   // classic multi-chunk, any length
void pc_launch_add_counter(unsigned long long *total, const unsigned long long *value, cudaStream_t stream);
// ctx->d_counters: [0..PC_NW_NCFG) work counters of the DP launches, then
#define PC_CTR_CELLS PC_NW_NCFG                  // cumulative DP cells: classic, LEX, LEXR, LEXC
#define PC_CTR_FALLBACK (PC_NW_NCFG + 4)         // alignments LEXC listed for the classic kernel (this step)
#define PC_CTR_FALLBACK_WORK (PC_NW_NCFG + 5)    // work counter of the launch that re-runs them
#define PC_CTR_FALLBACK_TOTAL (PC_NW_NCFG + 6)   // cumulative
#define PC_CTR_COUNT (PC_NW_NCFG + 8)
int pc_launch_classify_units(pc_ctx *ctx, const UnitArgs &u);
int pc_launch_scatter_units(pc_ctx *ctx, const UnitArgs &u, int32_t *task_q, int32_t *task_r);
int pc_launch_gather_units(pc_ctx *ctx, const UnitArgs &u, const uint32_t *hitlen,
                           unsigned long long *acc);
int pc_launch_emit_units(pc_ctx *ctx, const UnitArgs &u, const uint32_t *hitlen, const DestArgs &d,
                         unsigned long long *dest_counts /*[2][PC_MAX_DEST]*/, unsigned long long *records);
int pc_launch_accumulate_records(pc_ctx *ctx, int64_t n_records, const unsigned long long *records,
                                 int64_t n_acc, unsigned long long *acc);
struct PeerOuts {
    int n;
    int32_t *p[PC_MAX_PEER_OUT];
};
int pc_launch_aai_epilogue(pc_ctx *ctx, int metric, int as_distance, int raw,
                           int64_t row_begin, int64_t row_end,
                           const unsigned long long *acc, int32_t *out,
                           int32_t *out_aux, const PeerOuts *peers = nullptr);
int pc_launch_encode(pc_ctx *ctx, const uint8_t *raw, const int64_t *raw_off,
                     const int64_t *gene_off, const int32_t *gene_len,
                     const uint8_t *code_lut, int64_t n_genes, uint8_t *codes);
int pc_launch_gene_bound(pc_ctx *ctx, const int *rowmax32, int32_t *gene_hi);
int pc_launch_threshold_edges(pc_ctx *ctx, const int32_t *tri, int64_t n, int32_t max_micro,
                              unsigned long long *count, int2 *edges, int64_t cap);
int pc_launch_gather_entries(pc_ctx *ctx, const int32_t *tri, int64_t n_idx, const int64_t *idx, int32_t *out);
int pc_i8_contract(pc_ctx *ctx, int metric, const EpilogueArgs &ep);
int pc_i8_supported(const pc_ctx *ctx, int metric);
int pc_i8_pair_contract(pc_ctx *ctx, int n_limbs, int64_t p_pad, int64_t row_bytes, int64_t rows_pad,
                        int occ_words, int occ_shift, const EpilogueArgs &ep);
int pc_build_canonical_genes(pc_ctx *ctx);

// DP kernel configs: a subgroup of G lanes per alignment (G in {4,8,16,32}), K ref
// columns per lane in registers (G*K columns per pass), K even.
//   classic single-pass  c = g*13 + (K-8)/2,          K in 8..32   (5 registers per column)
//   multi-chunk          c = 52 + (K-20)/4,  G = 32,  K in {20,24,28,32}: the ref is cut into
//                        ceil(n / 32K) column chunks, any length
//   LEX single-pass      c = 56 + g*21 + (K-8)/2,     K in 8..48   (3 registers per column;
//                        counters inside the score word, for the alignments pc_nw_lex_ok admits)
//   LEXR (rebased LEX)   c = 140 + (K-16)/2, G = 32,  K in 16..32: 10-bit counters and a 10-bit
//                        score field that every lane re-centres every few rows (pc_nw_lexr_ok:
//                        min length < 1024, ref <= 1024 columns, any score range)
//   LEXC (rebased LEX with rebased counters)  c = 149 + (K-16)/2, G = 32, K in 16..24, column
//                        chunks: any length with |m - n| <= PC_NW_LEXC_MAXDIFF; an alignment whose
//                        counters leave the band is re-run by the classic multi-chunk kernel
// The chooser only ever picks K >= 16 for G >= 8 (a smaller G covers the same columns with less
// pipeline fill), and multiples of 4 up to 32 when the 32-code table is in use; only those
// combinations are instantiated (pc_align.cu).
#define PC_NW_NK 13                 // classic K values per G
#define PC_NW_LEX_NK 21             // LEX K values per G
#define PC_NW_MULTI0 (4 * PC_NW_NK) // first multi-chunk config
#define PC_NW_LEX0 (PC_NW_MULTI0 + 4)
#define PC_NW_LEXR0 (PC_NW_LEX0 + 4 * PC_NW_LEX_NK)
#define PC_NW_LEXR_NK 9             // G = 32, K = 16, 18 .. 32
#define PC_NW_LEXC0 (PC_NW_LEXR0 + PC_NW_LEXR_NK)
#define PC_NW_LEXC_NK 5             // G = 32, K = 16, 18 .. 24, ceil(n / 32K) column chunks
#define PC_NW_NCFG (PC_NW_LEXC0 + PC_NW_LEXC_NK)
#define PC_NW_KMAX 32
#define PC_NW_LEX_KMAX 48
#define PC_NW_COST_PER_CELL 18      // relative cost model (the kernels are ALU-pipe bound)
#define PC_NW_COST_PER_STEP 18      // per row-step overhead (shuffles, boundary, loop)
#define PC_NW_LEX_COST_PER_CELL 11
#define PC_NW_LEX_COST_PER_STEP 15

// LEX word: score << SHIFT | prio << 2*BITS | diagonal steps << BITS | matched
#define PC_NW_LEX_BITS 9
#define PC_NW_LEX_SHIFT (2 * PC_NW_LEX_BITS + 2)
#define PC_NW_LEX_FLOOR 40          // H^[0][0] sits FLOOR above the most negative score
#define PC_NW_LEX_ENABLED PC_GAP_TIE_EXTENDS
// LEXR word: the same layout with 10-bit counters and a 10-bit score field kept centred by
// rebasing (pc_align.cu).  DP values are Lipschitz -- neighbouring cells differ by at most
// LIP = max substitution score + open + ext in drifted units -- so a lane that owns K adjacent
// columns and re-centres on its middle column every R row-steps only ever combines values within
// LIP * (K/2 + R) + slack of 0: PC_NW_LEXR_WINDOW_OK is the static guarantee (oracle/lex_model.py
// runs the same schedule on the CPU and reports the largest field value it meets).
#define PC_NW_LEXR_BITS 10
#define PC_NW_LEXR_MAXSUB 11        // largest entry of the substitution matrix (W/W)
#define PC_NW_LEXR_LIP (PC_NW_LEXR_MAXSUB + PC_GAP_OPEN_DEFAULT + PC_GAP_EXTEND_DEFAULT)
#define PC_NW_LEXR_WINDOW_OK(K, R) \
    (PC_NW_LEXR_LIP * ((K) / 2 + (R)) + PC_GAP_OPEN_DEFAULT + PC_NW_LEXR_LIP + 8 < (1 << (32 - (2 * PC_NW_LEXR_BITS + 2) - 1)))
#define PC_NW_LEXR_REBASE(K) ((K) <= 24 ? 8 : 4)   // row-steps between two re-centrings
// LEXC word: the LEXR layout with the two counters stored relative to a per-lane base as well
// (re-centred together with the score every R row-steps), so neither the length nor the score
// range is bounded.  Unlike scores, counters of neighbouring cells have no proven bound on their
// difference (two near-optimal routes may differ in their number of diagonal steps), so the kernel
// CHECKS: after every re-centring all counter fields of a lane must lie within BAND of the centre
// column's, neighbouring lanes' bases within FRAME of each other, every word received from the
// neighbour within BAND after conversion.  While no check has failed no field has borrowed from or
// carried into its neighbour (2 BAND + R < 2^(B-1) and BAND + R + FRAME + 1 <= 2^(B-1), see
// oracle/lexc_model.c for the induction), so the result is exact; a failed check sends the
// alignment to the classic kernel.  Long gaps are what breaks the band, hence MAXDIFF.
#define PC_NW_LEXC_BITS 10
#define PC_NW_LEXC_REBASE 8
#define PC_NW_LEXC_BAND 128
#define PC_NW_LEXC_FRAME 255
#define PC_NW_LEXC_MAXDIFF 128
#define PC_NW_LEXC_FALLBACK_CFG PC_NW_MULTI0    // classic <32, 20, multi-chunk>: any length
static_assert(2 * PC_NW_LEXC_BAND + PC_NW_LEXC_REBASE < (1 << (PC_NW_LEXC_BITS - 1)) &&
              PC_NW_LEXC_BAND + PC_NW_LEXC_REBASE + PC_NW_LEXC_FRAME + 1 <= (1 << (PC_NW_LEXC_BITS - 1)) &&
              PC_NW_LEXC_BITS == PC_NW_LEXR_BITS,
              "LEXC: the counter fields cannot borrow or carry while the checks hold");
// which DP kernel family runs an alignment
#define PC_NW_MODE_CLASSIC 0
#define PC_NW_MODE_LEX 1
#define PC_NW_MODE_LEXR 2
#define PC_NW_MODE_LEXC 3

#ifdef __CUDACC__
#define PC_HD __host__ __device__
#else
#define PC_HD
#endif
PC_HD inline void pc_nw_cfg(int c, int &G, int &K)
{
    if (c >= PC_NW_LEXC0) {
        G = 32;
        K = 16 + 2 * (c - PC_NW_LEXC0);
    } else if (c >= PC_NW_LEXR0) {
        G = 32;
        K = 16 + 2 * (c - PC_NW_LEXR0);
    } else if (c >= PC_NW_LEX0) {
        c -= PC_NW_LEX0;
        G = 4 << (c / PC_NW_LEX_NK);
        K = 8 + 2 * (c % PC_NW_LEX_NK);
    } else if (c >= PC_NW_MULTI0) {
        G = 32;
        K = 20 + 4 * (c - PC_NW_MULTI0);
    } else {
        G = 4 << (c / PC_NW_NK);
        K = 8 + 2 * (c % PC_NW_NK);
    }
}

// May the (m-row query, n-column ref) alignment run on the LEX kernel?  Both
// counters are bounded by min(m, n).  The kernel works on values drifted by
// (i + j) ext: every real H^/E^/F^ and diagonal candidate then lies in
// [-2 open - ext - 4, hi + (m + n) ext] with hi = min over the two sequences of
// the summed per-residue best substitution score (rowmax >= 0 for every
// residue), and H^[0][0] is placed FLOOR above the most negative score.
PC_HD inline bool pc_nw_lex_ok(int m, int n, int hi_q, int hi_r)
{
#if PC_NW_LEX_ENABLED
    const int mn = m < n ? m : n;
    const long long hi = hi_q < hi_r ? hi_q : hi_r;
    return mn < (1 << PC_NW_LEX_BITS) &&
           hi + m + n + PC_NW_LEX_FLOOR + 24 < (1ll << (32 - PC_NW_LEX_SHIFT));
#else
    return false;
#endif
}

// The rebased LEX kernel takes what the plain one cannot: 10-bit counters (min length < 1024),
// no condition on the score range, refs of up to 32 lanes x 32 columns.
PC_HD inline bool pc_nw_lexr_ok(int m, int n)
{
#if PC_NW_LEX_ENABLED
    const int mn = m < n ? m : n;
    return mn < (1 << PC_NW_LEXR_BITS) && n <= 32 * 32;
#else
    return false;
#endif
}

// LEXC takes what is left when the two lengths are close (a global alignment of sequences whose
// lengths differ by d holds gaps of d columns at least, and long gaps break the counter band).
PC_HD inline bool pc_nw_lexc_ok(int m, int n)
{
#if PC_NW_LEX_ENABLED
    const int d = m < n ? n - m : m - n;
    return d <= PC_NW_LEXC_MAXDIFF && n >= 1;
#else
    return false;
#endif
}

// kernel family of an (m-row query, n-column ref) alignment; `allowed`: the 24-code table and
// the per-gene bounds are available and the caller did not ask for the classic kernel only
PC_HD inline int pc_nw_mode(int m, int n, int hi_q, int hi_r, bool allowed, bool lexc = true)
{
    if (!allowed) return PC_NW_MODE_CLASSIC;
    if (n <= 32 * PC_NW_LEX_KMAX && pc_nw_lex_ok(m, n, hi_q, hi_r)) return PC_NW_MODE_LEX;
    if (pc_nw_lexr_ok(m, n)) return PC_NW_MODE_LEXR;
    if (lexc && pc_nw_lexc_ok(m, n)) return PC_NW_MODE_LEXC;
    return PC_NW_MODE_CLASSIC;
}

// cheapest config for an m-row query against n ref columns: pipe slots per
// alignment ~ chunks x (m + G - 1) row-steps x (K cells + overhead) x G/32 of a warp.
// Per G the cost of a single-pass config grows with K, so only the smallest K
// that covers n competes; the LEX grid (when admitted) is cheaper cell for cell
// and covers everything the classic single-pass grid covers.  coarse: K in
// multiples of 4 (the kernels of the 32-code table).
PC_HD inline int pc_nw_choose_cfg(int m, int n, int mode, bool coarse)
{
    if (mode == PC_NW_MODE_LEXR) {
        int K = (n + 31) / 32;
        K = (K + 1) & ~1;
        if (K < 16) K = 16;
        return PC_NW_LEXR0 + (K - 16) / 2;
    }
    if (mode == PC_NW_MODE_LEXC) {
        int best_c = PC_NW_LEXC0;
        long long best_cost = -1;
        for (int c = PC_NW_LEXC0; c < PC_NW_LEXC0 + PC_NW_LEXC_NK; ++c) {
            const int K = 16 + 2 * (c - PC_NW_LEXC0);
            const long long chunks = (n + 32 * K - 1) / (32 * K);
            const long long cost = chunks * (long long)(m + 31) * (K * (long long)PC_NW_LEX_COST_PER_CELL + 22);
            if (best_cost < 0 || cost < best_cost) { best_cost = cost; best_c = c; }
        }
        return best_c;
    }
    const bool lex = mode == PC_NW_MODE_LEX;
    int best = PC_NW_LEX0 - 1;
    long long best_cost = -1;
    const long long per_cell = lex ? PC_NW_LEX_COST_PER_CELL : PC_NW_COST_PER_CELL;
    const long long per_step = lex ? PC_NW_LEX_COST_PER_STEP : PC_NW_COST_PER_STEP;
    const int kmax = lex ? PC_NW_LEX_KMAX : PC_NW_KMAX;
    for (int g = 0; g < 4; ++g) {
        const int G = 4 << g;
        int K = (n + G - 1) / G;
        K = coarse ? (K + 3) & ~3 : (K + 1) & ~1;
        const int kmin = g == 0 ? 8 : 16;
        if (K < kmin) K = kmin;
        if (K > kmax) continue;
        long long cost = (long long)(m + G - 1) * (K * per_cell + per_step) * G;
        // LEX configs with K >= 40 only fit two warps per SM sub-partition: ncu has their ALU
        // pipe at 75-78 % against 84-86 % for the configs with three or more
        if (lex && K >= 40) cost += cost / 10;
        if (best_cost < 0 || cost < best_cost) {
            best_cost = cost;
            best = lex ? PC_NW_LEX0 + g * PC_NW_LEX_NK + (K - 8) / 2 : g * PC_NW_NK + (K - 8) / 2;
        }
    }
    if (best_cost >= 0) return best;      // single-pass configs cover it without the spill
    for (int c = PC_NW_MULTI0; c < PC_NW_LEX0; ++c) {
        const int G = 32, K = 20 + 4 * (c - PC_NW_MULTI0);
        const long long chunks = (n + G * K - 1) / (G * K);
        const long long cost = chunks * (long long)(m + G - 1) *
                               (K * (long long)PC_NW_COST_PER_CELL + PC_NW_COST_PER_STEP) * G;
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = c; }
    }
    return best;
}

// ---------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------
#ifdef __CUDACC__
__host__ __device__ __forceinline__ int64_t pc_tri_index(int64_t n, int64_t i, int64_t j)
{
    // packed upper triangle incl. diagonal, i <= j
    return i * n - (i * (i - 1)) / 2 + (j - i);
}

// Python's round(x, 6) as integer micro-units: round-half-even of the EXACT
// value of the double x times 10^6 (CPython rounds the exact binary value via
// dtoa mode 3; SURVEY.md 7 hard part 3).  x*1e6 is split exactly into hi+lo.
__device__ __forceinline__ int pc_round6_micro(double x)
{
    const double hi = x * 1e6;
    const double lo = fma(x, 1e6, -hi);
    double r = rint(hi);                       // half-even on hi
    if (fabs(hi - r) == 0.5) {                 // hi is an exact half-integer
        const double fl = floor(hi);
        if (lo > 0.0) r = fl + 1.0;
        else if (lo < 0.0) r = fl;
    }
    return (int)r;
}

// The tail shared by every metric (e.g. metrics.py:50-53): similarity -> the
// value stored by SymMatrix.set_weight (matrix.py:323), in micro-units.
__device__ __forceinline__ int pc_finish_micro(double similarity, int as_distance)
{
    return pc_round6_micro(as_distance ? 1.0 - similarity : similarity);
}

// round-half-even(p * 10^6 / q) in integer arithmetic, for 0 <= p <= q < 2^31, q > 0; *tie is set
// when the exact quotient ends in .5 (then the caller must ask the fp64 path, see below).
__device__ __forceinline__ int pc_ratio_micro_int(uint32_t p, uint32_t q, bool *tie)
{
    // estimate from fp32 (relative error < 5e-7 -> at most 1 off), then exact 64-bit remainder
    uint32_t est = (uint32_t)(__fdividef((float)p, (float)q) * 1e6f);
    long long rem = (long long)p * 1000000ll - (long long)est * (long long)q;
    while (rem < 0) { rem += q; --est; }
    while (rem >= (long long)q) { rem -= q; ++est; }
    const long long twice = 2 * rem;
    *tie = false;
    if (twice == (long long)q) {
        // the exact quotient ends in .5.  If p / q in lowest terms has a power-of-two denominator
        // (the odd part of q divides p) every double on CPython's path is exact and round() sees a
        // true tie: half to even.  Otherwise the direction is that of the division's rounding
        // error: the caller asks the fp64 path.
        const uint32_t odd = q >> (__ffs((int)q) - 1);
        if (p % odd == 0) return (int)(est + (est & 1u));
        *tie = true;
    }
    return (int)(twice > (long long)q ? est + 1 : est);
}

// fp64 reference path: exactly CPython's arithmetic (IEEE divide, 1.0 - x, round(x, 6))
static __device__ __noinline__ int pc_cheap_metric_micro_f64(int metric, int as_distance, int numer, int64_t den)
{
    double sim;
    if (metric == PC_GCS)
        sim = (2.0 * (double)numer) / (double)den;            // metrics.py:45-48
    else if (metric == PC_JC)
        sim = (double)numer / (double)(den - numer);          // |S u T| = |S|+|T|-|S n T|
    else
        sim = (double)numer / (double)den;                    // metrics.py:110, :152
    return pc_finish_micro(sim, as_distance);
}

// fused epilogue for one pair of the cheap metrics (both contraction engines).
//
// The reference computes similarity = p / q in doubles and stores round(1.0 - similarity, 6)
// (metrics.py:50-53, matrix.py:323).  On this GPU fp64 issues about one warp instruction per 26
// cycles per SM (ncu: the contraction's epilogue warps sat in math-pipe throttle and a 128 x 256
// half tile took 135 us to normalise -- longer than its 100 us of MMAs at C3), so the value is
// computed in integers instead: the exact rational (q - p) * 10^6 / q rounded half-even.  The two
// agree unless the exact quotient ends in .5: the double path's total error is below 1.2e-10
// micro-units (two roundings of relative size 2^-53, times 10^6) while a non-tie quotient stays at
// least 1 / (2 q) > 2.3e-10 away from the rounding boundary for q < 2^31.  Exact ties need q to be a
// multiple of 128 (about 1 % of pairs); when the reduced fraction has a power-of-two denominator the
// doubles are exact and the tie goes to even, in integers; the remaining ties (q a multiple of 640)
// and out-of-range operands take the fp64 path, which is CPython's arithmetic verbatim.
__device__ __forceinline__ int pc_cheap_metric_micro(int metric, int as_distance, int numer,
                                                     int den_i, int den_j)
{
    // numer: shared phams (gcs, jc) or conserved genes / residues (pocp, af)
    if (numer <= 0) return as_distance ? 1000000 : 0;          // no shared pham: similarity 0.0 exactly
    const int64_t den = (int64_t)den_i + (int64_t)den_j;
    int64_t p = numer, q = den;
    if (metric == PC_GCS) p = 2 * (int64_t)numer;
    else if (metric == PC_JC) q = den - numer;
    if (p <= q && q < (1ll << 31)) {
        bool tie;
        const int micro = pc_ratio_micro_int((uint32_t)(as_distance ? q - p : p), (uint32_t)q, &tie);
        if (!tie) return micro;
    }
    return pc_cheap_metric_micro_f64(metric, as_distance, numer, den);
}

__global__ void pc_debug_metric_kernel(int metric, int as_distance, int64_t n, const int32_t *numer,
                                       const int32_t *den_i, const int32_t *den_j, int32_t *out, int32_t *out_f64);

#endif
