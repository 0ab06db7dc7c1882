// pc_gather.cu -- the kernels either side of the DP:
//   * encode:  raw residue bytes -> 5-bit codes in a 16B-aligned arena
//   * classify / scan / finalize: (pham, member a, member b) units -> one DP task
//     per DISTINCT ordered pair of sequences (hash table over canonical gene
//     ids), bucketed on the device by DP kernel config
//   * gather:  K5, best match per anchor gene -> per-genome-pair accumulators
//   * epilogue: K6 for aai / peq
// Reference semantics: /root/reference/src/phamclust/metrics.py:203-232 (aai),
// :235-253 (peq), statistics.py:22 (length-weighted mean).
#include <algorithm>

#include "pc_common.cuh"

// ---------------------------------------------------------------------------
// encode
// ---------------------------------------------------------------------------
__global__ void encode_kernel(const uint8_t *__restrict__ raw,
                              const int64_t *__restrict__ raw_off,
                              const int64_t *__restrict__ gene_off,
                              const int32_t *__restrict__ gene_len,
                              const uint8_t *__restrict__ lut, int64_t n_genes,
                              uint8_t *__restrict__ codes)
{
    __shared__ uint8_t s_lut[256];
    for (int x = threadIdx.x; x < 256; x += blockDim.x) s_lut[x] = lut[x];
    __syncthreads();
    // one warp per gene, lanes stride the residues
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int lane = threadIdx.x & 31;
    for (int64_t g = warp; g < n_genes; g += n_warps) {
        const uint8_t *src = raw + raw_off[g];
        uint8_t *dst = codes + gene_off[g];
        const int len = gene_len[g];
        for (int x = lane; x < len; x += 32) dst[x] = s_lut[src[x]];
    }
}

int pc_launch_encode(pc_ctx *ctx, const uint8_t *raw, const int64_t *raw_off,
                     const int64_t *gene_off, const int32_t *gene_len,
                     const uint8_t *code_lut, int64_t n_genes, uint8_t *codes)
{
    if (n_genes <= 0) return PC_OK;
    int64_t blocks = (n_genes * 32 + 255) / 256;
    int64_t cap = (int64_t)ctx->sm_count * 8;
    int grid = (int)(blocks < cap ? blocks : cap);
    encode_kernel<<<grid, 256, 0, ctx->stream>>>(raw, raw_off, gene_off, gene_len,
                                                code_lut, n_genes, codes);
    PC_CUDA(ctx, cudaGetLastError());
    ctx->launches++;
    return PC_OK;
}

// per gene: sum over its residues of the best substitution score the residue
// can ever get (rowmax >= 0) -- the upper bound of every DP value of an
// alignment the gene takes part in (pc_nw_lex_ok)
struct RowMax { int v[32]; };

__global__ void gene_bound_kernel(const uint8_t *__restrict__ codes,
                                  const int64_t *__restrict__ gene_off,
                                  const int32_t *__restrict__ gene_len, RowMax rm,
                                  int64_t n_genes, int32_t *__restrict__ gene_hi)
{
    __shared__ int s_rm[32];
    if (threadIdx.x < 32) s_rm[threadIdx.x] = rm.v[threadIdx.x];
    __syncthreads();
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int lane = threadIdx.x & 31;
    for (int64_t g = warp; g < n_genes; g += n_warps) {
        const uint8_t *src = codes + gene_off[g];
        const int len = gene_len[g];
        int sum = 0;
        for (int x = lane; x < len; x += 32) sum += s_rm[src[x] & 31];
        sum = __reduce_add_sync(0xffffffffu, sum);
        if (lane == 0) gene_hi[g] = sum;
    }
}

int pc_launch_gene_bound(pc_ctx *ctx, const int *rowmax32, int32_t *gene_hi)
{
    if (ctx->n_genes <= 0) return PC_OK;
    RowMax rm;
    for (int k = 0; k < 32; ++k) rm.v[k] = rowmax32[k];
    int64_t blocks = (ctx->n_genes * 32 + 255) / 256;
    int64_t cap = (int64_t)ctx->sm_count * 8;
    int grid = (int)(blocks < cap ? blocks : cap);
    gene_bound_kernel<<<grid, 256, 0, ctx->stream>>>(ctx->d_codes.as<uint8_t>(), ctx->d_gene_off.as<int64_t>(),
                                                    ctx->d_gene_len.as<int32_t>(), rm, ctx->n_genes, gene_hi);
    PC_CUDA(ctx, cudaGetLastError());
    ctx->launches++;
    return PC_OK;
}

// ---------------------------------------------------------------------------
// units: (pham, member a, member b) with a < b in ascending genome order
// ---------------------------------------------------------------------------
struct Unit {
    int ga, gb;              // genome indices, ga < gb
    int first_a, first_b;    // first gene of each member
    int ca, cb;              // copies
};

__device__ __forceinline__ Unit decode_unit(const UnitArgs &u, int64_t unit)
{
    // pham = last p with ph_unit_off[p] <= unit
    int64_t lo = 0, hi = u.n_ph;
    while (hi - lo > 1) {
        const int64_t mid = (lo + hi) >> 1;
        if (u.ph_unit_off[mid] <= unit) lo = mid; else hi = mid;
    }
    const int64_t p = lo;
    const int64_t mem0 = u.ph_mem_off[p];
    const int64_t m = u.ph_mem_off[p + 1] - mem0;
    const int64_t lu = unit - u.ph_unit_off[p];
    // row-major strict upper triangle of an m x m grid: row a holds m-1-a units
    const double tm = (double)(2 * m - 1);
    int64_t a = (int64_t)floor((tm - sqrt(tm * tm - 8.0 * (double)lu)) * 0.5);
    if (a < 0) a = 0;
    if (a > m - 2) a = m - 2;
    while (a > 0 && a * (2 * m - a - 1) / 2 > lu) --a;
    while ((a + 1) * (2 * m - (a + 1) - 1) / 2 <= lu) ++a;
    const int64_t b = lu - a * (2 * m - a - 1) / 2 + a + 1;
    Unit r;
    r.ga = u.mem_genome[mem0 + a];
    r.gb = u.mem_genome[mem0 + b];
    r.first_a = u.mem_first_gene[mem0 + a];
    r.first_b = u.mem_first_gene[mem0 + b];
    r.ca = u.mem_copies[mem0 + a];
    r.cb = u.mem_copies[mem0 + b];
    return r;
}

// Anchor = the member with FEWER copies, tie -> the lower genome index
// (metrics.py:208-209; source index < target index by matrix.py:480-486).
struct UnitShape {
    int n_anchor, n_other, anchor0, other0;
};

__device__ __forceinline__ UnitShape unit_shape(const Unit &t)
{
    const bool swap = t.ca > t.cb;
    UnitShape s;
    s.n_anchor = swap ? t.cb : t.ca;
    s.n_other = swap ? t.ca : t.cb;
    s.anchor0 = swap ? t.first_b : t.first_a;
    s.other0 = swap ? t.first_a : t.first_b;
    return s;
}

// ---------------------------------------------------------------------------
// task table: one DP task per DISTINCT ordered pair of canonical genes
// (query sequence, ref sequence).  Open-addressing hash table keyed by
// (canon_q << 32) | canon_r; the value first holds (config, block, rank inside
// the block) of the thread that inserted the key, then the final task index.
// ---------------------------------------------------------------------------
#define HT_EMPTY 0xffffffffffffffffull

__device__ __forceinline__ uint64_t ht_mix(uint64_t k)
{
    k ^= k >> 33; k *= 0xff51afd7ed558ccdull;
    k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull;
    k ^= k >> 33;
    return k;
}

// returns the slot of `key`; *inserted tells whether this call created it
__device__ __forceinline__ uint64_t ht_insert(unsigned long long *keys, uint64_t mask, uint64_t key, bool *inserted)
{
    uint64_t h = ht_mix(key) & mask;
    for (;;) {
        const unsigned long long cur = keys[h];
        if (cur == key) { *inserted = false; return h; }
        if (cur == HT_EMPTY) {
            const unsigned long long prev = atomicCAS(keys + h, HT_EMPTY, (unsigned long long)key);
            if (prev == HT_EMPTY) { *inserted = true; return h; }
            if (prev == key) { *inserted = false; return h; }
        }
        h = (h + 1) & mask;
    }
}

__device__ __forceinline__ uint64_t ht_find(const unsigned long long *keys, uint64_t mask, uint64_t key)
{
    uint64_t h = ht_mix(key) & mask;
    while (keys[h] != key) h = (h + 1) & mask;
    return h;
}

#define UNIT_THREADS 256

// pass A: insert every alignment's (canonical query, canonical ref) pair; the
// inserting thread picks the DP config of the pair and takes a rank inside
// (block, config)
__global__ void __launch_bounds__(UNIT_THREADS) classify_units_kernel(const UnitArgs u)
{
    static_assert(PC_NW_NCFG <= UNIT_THREADS, "one thread per config zeroes and writes the block counts");
    __shared__ unsigned int s_cnt[PC_NW_NCFG];
    if (threadIdx.x < PC_NW_NCFG) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    const int64_t u0 = (int64_t)blockIdx.x * u.units_per_block;
    const int64_t u1 = min(u0 + u.units_per_block, u.n_units);
    for (int64_t unit = u0 + threadIdx.x; unit < u1; unit += UNIT_THREADS) {
        const Unit t = decode_unit(u, unit);
        const UnitShape s = unit_shape(t);
        for (int x = 0; x < s.n_anchor; ++x)
            for (int y = 0; y < s.n_other; ++y) {
                const uint32_t cq = (uint32_t)u.canon[s.anchor0 + x];   // _align(anchor_gene, other_gene)
                const uint32_t cr = (uint32_t)u.canon[s.other0 + y];
                bool inserted;
                const uint64_t h = ht_insert(u.ht_keys, u.ht_mask, ((uint64_t)cq << 32) | cr, &inserted);
                if (inserted) {
                    const int lq = u.gene_len[cq], lr = u.gene_len[cr];
                    const int mode = u.gene_hi ? pc_nw_mode(lq, lr, u.gene_hi[cq], u.gene_hi[cr], true, !u.no_lexc)
                                               : PC_NW_MODE_CLASSIC;
                    const int cfg = pc_nw_choose_cfg(lq, lr, mode, u.coarse_k != 0);
                    const unsigned int rank = atomicAdd(&s_cnt[cfg], 1u);
                    u.ht_vals[h] = ((unsigned long long)cfg << 56) | ((unsigned long long)blockIdx.x << 32) | rank;
                }
            }
    }
    __syncthreads();
    if (threadIdx.x < PC_NW_NCFG)
        u.blk_counts[(int64_t)blockIdx.x * PC_NW_NCFG + threadIdx.x] = s_cnt[threadIdx.x];
}

// pass B: per config, exclusive scan over blocks; config bases in launch order
static_assert(PC_NW_NCFG <= 256, "scan_blocks_kernel runs one thread per config in a 256-thread block");
__global__ void scan_blocks_kernel(const UnitArgs u)
{
    const int c = threadIdx.x;
    if (c < PC_NW_NCFG) {
        unsigned long long run = 0;
        for (int b = 0; b < u.n_blocks; ++b) {
            const unsigned long long v = u.blk_counts[(int64_t)b * PC_NW_NCFG + c];
            u.blk_counts[(int64_t)b * PC_NW_NCFG + c] = run;
            run += v;
        }
        u.cfg_total[c] = run;
    }
    __syncthreads();
    if (c == 0) {
        unsigned long long base = 0;
        for (int k = PC_NW_NCFG - 1; k >= 0; --k) {      // long configs first in memory and in time
            u.cfg_base[k] = base;
            base += u.cfg_total[k];
        }
    }
}

// pass C: every occupied slot becomes a task at config base + block offset + rank
__global__ void finalize_tasks_kernel(const UnitArgs u, int32_t *__restrict__ task_q,
                                      int32_t *__restrict__ task_r)
{
    const uint64_t cap = u.ht_mask + 1;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t h = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; h < cap; h += stride) {
        const unsigned long long key = u.ht_keys[h];
        if (key == HT_EMPTY) continue;
        const unsigned long long v = u.ht_vals[h];
        const int cfg = (int)(v >> 56);
        const int64_t blk = (int64_t)((v >> 32) & 0xffffffull);
        const unsigned long long idx = u.cfg_base[cfg] + u.blk_counts[blk * PC_NW_NCFG + cfg] + (v & 0xffffffffull);
        task_q[idx] = (int32_t)(key >> 32);
        task_r[idx] = (int32_t)(key & 0xffffffffull);
        u.ht_vals[h] = idx;
    }
}

// For each anchor gene keep the candidate with the best matched/length ratio,
// ties -> the LAST candidate (sorted(...)[-1] on a stable sort, metrics.py:223);
// ratios are compared exactly by cross-multiplication.  Returns the unit's sums
// packed as (matched << 32) | length (0: nothing to add).
__device__ __forceinline__ unsigned long long unit_best_matches(const UnitArgs &u, const uint32_t *__restrict__ hitlen,
                                                                const UnitShape &s)
{
    unsigned long long sum_hit = 0, sum_len = 0;
    for (int x = 0; x < s.n_anchor; ++x) {
        unsigned long long best_hit = 0, best_len = 0;
        bool have = false;
        const uint32_t cq = (uint32_t)u.canon[s.anchor0 + x];
        for (int y = 0; y < s.n_other; ++y) {
            const uint32_t cr = (uint32_t)u.canon[s.other0 + y];
            const uint64_t slot = ht_find(u.ht_keys, u.ht_mask, ((uint64_t)cq << 32) | cr);
            const uint32_t v = hitlen[u.ht_vals[slot]];
            const unsigned long long hit = v >> 16, len = v & 0xffffu;
            if (len == 0) continue;           // undefined upstream (0-length alignment)
            if (!have || hit * best_len >= best_hit * len) {
                best_hit = hit; best_len = len; have = true;
            }
        }
        sum_hit += best_hit;
        sum_len += best_len;
    }
    return sum_len ? ((sum_hit << 32) | sum_len) : 0ull;
}

// Single GPU / dense exchange: the unit's sums go to acc[idx(ga, gb)] with ONE 64-bit integer
// atomic: exact and order independent.
__global__ void gather_units_kernel(const UnitArgs u, const uint32_t *__restrict__ hitlen,
                                    unsigned long long *__restrict__ acc)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t unit = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
         unit < u.n_units; unit += stride) {
        const Unit t = decode_unit(u, unit);
        const unsigned long long v = unit_best_matches(u, hitlen, unit_shape(t));
        if (v) atomicAdd(acc + pc_tri_index(u.n_genomes, t.ga, t.gb), v);
    }
}

// ---------------------------------------------------------------------------
// Sparse exchange (several GPUs): instead of adding into a dense N(N+1)/2 accumulator that
// every rank would have to hold and all-reduce, each unit becomes ONE 16-byte record
//   { index of (ga, gb) inside the row block of its owner, (matched << 32) | length }
// written into the bucket of the rank that owns genome row ga.  The buckets are exchanged with
// one all-to-all (only pairs that share a pham exist: < 3 % of the triangle on clustered tables)
// and the owner adds what it receives into the accumulators of ITS rows only.
// Pass 1 counts records per destination, pass 2 writes them behind warp-aggregated cursors.
// ---------------------------------------------------------------------------
__device__ __forceinline__ int dest_of_row(const DestArgs &d, int row)
{
    int b = 0;
    while (b + 1 < d.n_dest && d.bounds[b + 1] <= row) ++b;
    return b;
}

__global__ void count_units_kernel(const UnitArgs u, const DestArgs d, unsigned long long *__restrict__ counts)
{
    __shared__ unsigned int s_cnt[PC_MAX_DEST];
    if (threadIdx.x < PC_MAX_DEST) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t unit = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; unit < u.n_units; unit += stride) {
        const Unit t = decode_unit(u, unit);
        atomicAdd(&s_cnt[dest_of_row(d, t.ga)], 1u);
    }
    __syncthreads();
    if (threadIdx.x < d.n_dest && s_cnt[threadIdx.x]) atomicAdd(counts + threadIdx.x, (unsigned long long)s_cnt[threadIdx.x]);
}

// cursors[k] = first record slot of destination k (exclusive scan of the counts)
__global__ void scan_dest_kernel(int n_dest, const unsigned long long *__restrict__ counts,
                                 unsigned long long *__restrict__ cursors)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        unsigned long long run = 0;
        for (int k = 0; k < n_dest; ++k) { cursors[k] = run; run += counts[k]; }
    }
}

__global__ void emit_units_kernel(const UnitArgs u, const uint32_t *__restrict__ hitlen, const DestArgs d,
                                  unsigned long long *__restrict__ cursors, ulonglong2 *__restrict__ records)
{
    const int lane = threadIdx.x & 31;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    // warp-uniform trip count: every lane of a warp takes part in the match / shuffle below
    for (int64_t first = (int64_t)blockIdx.x * blockDim.x + (threadIdx.x & ~31); first < u.n_units; first += stride) {
        const int64_t unit = first + lane;
        const bool valid = unit < u.n_units;
        int dest = PC_MAX_DEST;                       // lanes past the end form their own group
        ulonglong2 rec = make_ulonglong2(0, 0);
        if (valid) {
            const Unit t = decode_unit(u, unit);
            dest = dest_of_row(d, t.ga);
            rec.x = (unsigned long long)(pc_tri_index(u.n_genomes, t.ga, t.gb) -
                                         pc_tri_index(u.n_genomes, d.bounds[dest], d.bounds[dest]));
            rec.y = unit_best_matches(u, hitlen, unit_shape(t));
        }
        const unsigned peers = __match_any_sync(0xffffffffu, dest);
        const int leader = __ffs((int)peers) - 1;
        unsigned long long base = 0;
        if (valid && lane == leader) base = atomicAdd(cursors + dest, (unsigned long long)__popc(peers));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (valid) records[base + __popc(peers & ((1u << lane) - 1u))] = rec;
    }
}

__global__ void accumulate_records_kernel(int64_t n_records, const ulonglong2 *__restrict__ records,
                                          int64_t n_acc, unsigned long long *__restrict__ acc)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n_records; k += stride) {
        const ulonglong2 r = records[k];
        if (r.y && (int64_t)r.x < n_acc) atomicAdd(acc + r.x, r.y);
    }
}

int pc_launch_emit_units(pc_ctx *ctx, const UnitArgs &u, const uint32_t *hitlen, const DestArgs &d,
                         unsigned long long *dest_counts, unsigned long long *records)
{
    if (u.n_units <= 0) return PC_OK;
    PC_CUDA(ctx, cudaMemsetAsync(dest_counts, 0, sizeof(unsigned long long) * 2 * PC_MAX_DEST, ctx->stream));
    int64_t blocks = (u.n_units + 255) / 256;
    int64_t cap = (int64_t)ctx->sm_count * 16;
    const int grid = (int)std::max<int64_t>(1, blocks < cap ? blocks : cap);
    count_units_kernel<<<grid, 256, 0, ctx->stream>>>(u, d, dest_counts);
    PC_CUDA(ctx, cudaGetLastError());
    scan_dest_kernel<<<1, 32, 0, ctx->stream>>>(d.n_dest, dest_counts, dest_counts + PC_MAX_DEST);
    PC_CUDA(ctx, cudaGetLastError());
    emit_units_kernel<<<grid, 256, 0, ctx->stream>>>(u, hitlen, d, dest_counts + PC_MAX_DEST,
                                                    reinterpret_cast<ulonglong2 *>(records));
    PC_CUDA(ctx, cudaGetLastError());
    ctx->launches += 3;
    return PC_OK;
}

int pc_launch_accumulate_records(pc_ctx *ctx, int64_t n_records, const unsigned long long *records,
                                 int64_t n_acc, unsigned long long *acc)
{
    if (n_records <= 0) return PC_OK;
    int64_t blocks = (n_records + 255) / 256;
    int64_t cap = (int64_t)ctx->sm_count * 16;
    const int grid = (int)std::max<int64_t>(1, blocks < cap ? blocks : cap);
    accumulate_records_kernel<<<grid, 256, 0, ctx->stream>>>(n_records, reinterpret_cast<const ulonglong2 *>(records),
                                                            n_acc, acc);
    PC_CUDA(ctx, cudaGetLastError());
    ctx->launches++;
    return PC_OK;
}

int pc_launch_classify_units(pc_ctx *ctx, const UnitArgs &u)
{
    if (u.n_units <= 0) return PC_OK;
    PC_CUDA(ctx, cudaMemsetAsync(u.ht_keys, 0xff, sizeof(unsigned long long) * (u.ht_mask + 1), ctx->stream));
    classify_units_kernel<<<u.n_blocks, UNIT_THREADS, 0, ctx->stream>>>(u);
    PC_CUDA(ctx, cudaGetLastError());
    scan_blocks_kernel<<<1, 256, 0, ctx->stream>>>(u);
    PC_CUDA(ctx, cudaGetLastError());
    ctx->launches += 2;
    return PC_OK;
}

int pc_launch_scatter_units(pc_ctx *ctx, const UnitArgs &u, int32_t *task_q, int32_t *task_r)
{
    if (u.n_units <= 0) return PC_OK;
    const uint64_t cap = u.ht_mask + 1;
    const int grid = (int)std::min<uint64_t>((cap + 255) / 256, (uint64_t)ctx->sm_count * 16);
    finalize_tasks_kernel<<<grid, 256, 0, ctx->stream>>>(u, task_q, task_r);
    PC_CUDA(ctx, cudaGetLastError());
    ctx->launches++;
    return PC_OK;
}

int pc_launch_gather_units(pc_ctx *ctx, const UnitArgs &u, const uint32_t *hitlen,
                           unsigned long long *acc)
{
    if (u.n_units <= 0) return PC_OK;
    int64_t blocks = (u.n_units + 255) / 256;
    int64_t cap = (int64_t)ctx->sm_count * 16;
    const int grid = (int)(blocks < cap ? blocks : cap);
    gather_units_kernel<<<grid < 1 ? 1 : grid, 256, 0, ctx->stream>>>(u, hitlen, acc);
    PC_CUDA(ctx, cudaGetLastError());
    ctx->launches++;
    return PC_OK;
}

// ---------------------------------------------------------------------------
// K6 for aai and peq.  For PC_PEQ `out` already holds alignment_fraction in
// micro-units AS SIMILARITY (metrics.py:247) and is overwritten in place.
// ---------------------------------------------------------------------------
__global__ void aai_epilogue_kernel(int metric, int as_distance, int raw, int64_t n,
                                    int64_t row_begin, int64_t row_end,
                                    const unsigned long long *__restrict__ acc,
                                    int32_t *__restrict__ out, int32_t *__restrict__ out_aux, const PeerOuts peers)
{
    const int64_t base = pc_tri_index(n, row_begin, row_begin);
    const int64_t total = pc_tri_index(n, row_end, row_end) - base;   // row_end may equal n
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += stride) {
        // recover (i, j) from the packed index: row i starts at pc_tri_index(n, i, i)
        const int64_t g = base + k;
        const double tn = (double)(2 * n + 1);
        int64_t i = (int64_t)floor((tn - sqrt(tn * tn - 8.0 * (double)g)) * 0.5);
        if (i < 0) i = 0;
        if (i > n - 1) i = n - 1;
        while (i > 0 && pc_tri_index(n, i, i) > g) --i;
        while (i + 1 < n && pc_tri_index(n, i + 1, i + 1) <= g) ++i;
        const int64_t j = i + (g - pc_tri_index(n, i, i));
        const unsigned long long v = acc[g];
        const unsigned long long hit = v >> 32, len = v & 0xffffffffull;
        if (raw) {
            out[k] = (i == j) ? 0 : (int32_t)hit;
            if (out_aux) out_aux[k] = (i == j) ? 0 : (int32_t)len;
            continue;
        }
#define PC_PUT(val_)                                                                  \
    do {                                                                              \
        const int32_t v_ = (val_);                                                    \
        out[k] = v_;                                                                  \
        for (int o_ = 0; o_ < peers.n; ++o_) peers.p[o_][k] = v_;                     \
    } while (0)
        if (i == j) {                               // matrix.py:467-468
            PC_PUT(as_distance ? 0 : 1000000);
            continue;
        }
        // aai similarity: length-weighted mean identity = sum(hit) / sum(len)
        // (statistics.py:22); no shared pham -> 0.0 (metrics.py:198-199).
        // Integers first (fp64 issues ~1 warp instruction per 26 cycles per SM on this part, and
        // most pairs of a large table share nothing): the exact rational rounded half-even equals
        // CPython's float result unless the quotient ends in .5 (see pc_cheap_metric_micro); those
        // ties, and lengths beyond 2^31, take the fp64 path below.
        if (len == 0 || hit == 0) {                 // similarity 0.0 exactly; peq = af * 0.0
            PC_PUT(as_distance ? 1000000 : 0);
            continue;
        }
        if (len < (1ull << 31) && hit <= len) {
            bool tie;
            if (metric == PC_AAI) {
                const int micro = pc_ratio_micro_int((uint32_t)(as_distance ? len - hit : hit), (uint32_t)len, &tie);
                if (!tie) { PC_PUT(micro); continue; }
            } else {
                // peq multiplies the two ROUNDED factors (metrics.py:247-248): a = af, b = aai, both in
                // micro-units; round(1.0 - (a / 1e6) * (b / 1e6), 6) is the exact a * b / 1e6 rounded
                // half-even (error of the double path < 5e-10 micro-units, boundary distance 5e-7)
                const int b = pc_ratio_micro_int((uint32_t)hit, (uint32_t)len, &tie);
                if (!tie) {
                    const long long prod = (long long)out[k] * (long long)b;          // <= 1e12
                    const long long num = as_distance ? 1000000000000ll - prod : prod;
                    const long long q = num / 1000000ll, rem = num % 1000000ll;
                    if (rem != 500000ll) { PC_PUT((int32_t)(rem > 500000ll ? q + 1 : q)); continue; }
                }
            }
        }
        const double aai = (double)hit / (double)len;
        if (metric == PC_AAI) {
            PC_PUT(pc_finish_micro(aai, as_distance));
        } else {
            const double af_r = (double)out[k] / 1e6;
            const double aai_r = (double)pc_round6_micro(aai) / 1e6;
            PC_PUT(pc_finish_micro(af_r * aai_r, as_distance));
        }
    }
#undef PC_PUT
}

int pc_launch_aai_epilogue(pc_ctx *ctx, int metric, int as_distance, int raw,
                           int64_t row_begin, int64_t row_end,
                           const unsigned long long *acc, int32_t *out, int32_t *out_aux, const PeerOuts *peers)
{
    PeerOuts po;
    po.n = 0;
    if (peers) po = *peers;
    const int64_t total = pc_tri_index(ctx->N, row_end, row_end) -
                          pc_tri_index(ctx->N, row_begin, row_begin);
    if (total <= 0) return PC_OK;
    int64_t blocks = (total + 255) / 256;
    int64_t cap = (int64_t)ctx->sm_count * 16;
    int grid = (int)(blocks < cap ? blocks : cap);
    aai_epilogue_kernel<<<grid, 256, 0, ctx->stream>>>(metric, as_distance, raw, ctx->N,
                                                      row_begin, row_end, acc, out, out_aux, po);
    PC_CUDA(ctx, cudaGetLastError());
    ctx->launches++;
    return PC_OK;
}
