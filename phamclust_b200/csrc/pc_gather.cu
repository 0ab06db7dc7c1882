// pc_gather.cu -- the kernels either side of the DP:
//   * encode:  raw residue bytes -> 5-bit codes in a 16B-aligned arena
//   * classify / scatter: (pham, member a, member b) units -> ordered gene-pair
//     tasks, bucketed on the device by DP kernel config
//   * gather:  K5, best match per anchor gene -> per-genome-pair accumulators
//   * epilogue: K6 for aai / peq
// Reference semantics: /root/reference/src/phamclust/metrics.py:203-232 (aai),
// :235-253 (peq), statistics.py:22 (length-weighted mean).
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

// DP config of a unit: sized for its longest query (anchor gene) and ref (other gene)
__device__ __forceinline__ int unit_cfg(const UnitArgs &u, const UnitShape &s)
{
    int m = 0, n = 0;
    for (int x = 0; x < s.n_anchor; ++x) m = max(m, u.gene_len[s.anchor0 + x]);
    for (int y = 0; y < s.n_other; ++y) n = max(n, u.gene_len[s.other0 + y]);
    return pc_nw_choose_cfg(m, n);
}

#define UNIT_THREADS 256

// pass A: tasks per (block, config)
__global__ void __launch_bounds__(UNIT_THREADS) classify_units_kernel(const UnitArgs u)
{
    __shared__ unsigned long long s_cnt[PC_NW_NCFG];
    if (threadIdx.x < PC_NW_NCFG) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    const int64_t u0 = (int64_t)blockIdx.x * u.units_per_block;
    const int64_t u1 = min(u0 + u.units_per_block, u.n_units);
    for (int64_t unit = u0 + threadIdx.x; unit < u1; unit += UNIT_THREADS) {
        const Unit t = decode_unit(u, unit);
        const UnitShape s = unit_shape(t);
        atomicAdd(&s_cnt[unit_cfg(u, s)], (unsigned long long)(s.n_anchor * s.n_other));
    }
    __syncthreads();
    if (threadIdx.x < PC_NW_NCFG)
        u.blk_counts[(int64_t)blockIdx.x * PC_NW_NCFG + threadIdx.x] = s_cnt[threadIdx.x];
}

// pass B: per config, exclusive scan over blocks; config bases in launch order
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

// pass C: deterministic slot of every unit inside its config's task range,
// then expansion into ordered (query, ref) gene pairs
__global__ void __launch_bounds__(UNIT_THREADS)
scatter_units_kernel(const UnitArgs u, int32_t *__restrict__ task_q, int32_t *__restrict__ task_r)
{
    __shared__ unsigned long long s_run[PC_NW_NCFG];
    __shared__ unsigned int s_wtot[UNIT_THREADS / 32][PC_NW_NCFG];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x < PC_NW_NCFG)
        s_run[threadIdx.x] = u.cfg_base[threadIdx.x] +
                             u.blk_counts[(int64_t)blockIdx.x * PC_NW_NCFG + threadIdx.x];
    for (int x = threadIdx.x; x < (UNIT_THREADS / 32) * PC_NW_NCFG; x += UNIT_THREADS)
        (&s_wtot[0][0])[x] = 0;
    __syncthreads();
    const int64_t u0 = (int64_t)blockIdx.x * u.units_per_block;
    const int64_t u1 = min(u0 + u.units_per_block, u.n_units);
    for (int64_t base = u0; base < u1; base += UNIT_THREADS) {
        const int64_t unit = base + threadIdx.x;
        const bool valid = unit < u1;
        Unit t;
        UnitShape s;
        int cfg = -1;
        unsigned int nt = 0;
        if (valid) {
            t = decode_unit(u, unit);
            s = unit_shape(t);
            cfg = unit_cfg(u, s);
            nt = (unsigned int)(s.n_anchor * s.n_other);
        }
        // rank among the lanes of this warp that share the config
        unsigned int pre = 0, tot = 0;
        bool leader = true;
        for (int l = 0; l < 32; ++l) {
            const int cl = __shfl_sync(0xffffffffu, cfg, l);
            const unsigned int vl = __shfl_sync(0xffffffffu, nt, l);
            if (cl == cfg) {
                tot += vl;
                if (l < lane) { pre += vl; leader = false; }
            }
        }
        if (valid && leader) s_wtot[warp][cfg] = tot;
        __syncthreads();
        if (valid) {
            unsigned long long off = s_run[cfg] + pre;
            for (int w = 0; w < warp; ++w) off += s_wtot[w][cfg];
            u.unit_task_off[unit] = (int64_t)off;
            for (int x = 0; x < s.n_anchor; ++x)
                for (int y = 0; y < s.n_other; ++y) {
                    const int64_t k = (int64_t)off + (int64_t)x * s.n_other + y;
                    task_q[k] = s.anchor0 + x;      // _align(anchor_gene, other_gene)
                    task_r[k] = s.other0 + y;
                }
        }
        __syncthreads();
        if (threadIdx.x < PC_NW_NCFG) {
            unsigned long long add = 0;
            for (int w = 0; w < UNIT_THREADS / 32; ++w) {
                add += s_wtot[w][threadIdx.x];
                s_wtot[w][threadIdx.x] = 0;
            }
            s_run[threadIdx.x] += add;
        }
        __syncthreads();
    }
}

// For each anchor gene keep the candidate with the best matched/length ratio,
// ties -> the LAST candidate (sorted(...)[-1] on a stable sort, metrics.py:223);
// ratios are compared exactly by cross-multiplication.  The unit's sums go to
// acc[idx(ga, gb)] as (matched << 32) | length with ONE 64-bit integer atomic:
// exact and order independent.
__global__ void gather_units_kernel(const UnitArgs u, const uint32_t *__restrict__ hitlen,
                                    unsigned long long *__restrict__ acc)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t unit = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
         unit < u.n_units; unit += stride) {
        const Unit t = decode_unit(u, unit);
        const UnitShape s = unit_shape(t);
        const int64_t task_off = u.unit_task_off[unit];
        unsigned long long sum_hit = 0, sum_len = 0;
        for (int x = 0; x < s.n_anchor; ++x) {
            unsigned long long best_hit = 0, best_len = 0;
            bool have = false;
            for (int y = 0; y < s.n_other; ++y) {
                const uint32_t v = hitlen[task_off + (int64_t)x * s.n_other + y];
                const unsigned long long hit = v >> 16, len = v & 0xffffu;
                if (len == 0) continue;           // undefined upstream (0-length alignment)
                if (!have || hit * best_len >= best_hit * len) {
                    best_hit = hit; best_len = len; have = true;
                }
            }
            sum_hit += best_hit;
            sum_len += best_len;
        }
        if (sum_len)
            atomicAdd(acc + pc_tri_index(u.n_genomes, t.ga, t.gb), (sum_hit << 32) | sum_len);
    }
}

int pc_launch_classify_units(pc_ctx *ctx, const UnitArgs &u)
{
    if (u.n_units <= 0) return PC_OK;
    classify_units_kernel<<<u.n_blocks, UNIT_THREADS, 0, ctx->stream>>>(u);
    PC_CUDA(ctx, cudaGetLastError());
    scan_blocks_kernel<<<1, 32, 0, ctx->stream>>>(u);
    PC_CUDA(ctx, cudaGetLastError());
    ctx->launches += 2;
    return PC_OK;
}

int pc_launch_scatter_units(pc_ctx *ctx, const UnitArgs &u, int32_t *task_q, int32_t *task_r)
{
    if (u.n_units <= 0) return PC_OK;
    scatter_units_kernel<<<u.n_blocks, UNIT_THREADS, 0, ctx->stream>>>(u, task_q, task_r);
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
                                    int32_t *__restrict__ out, int32_t *__restrict__ out_aux)
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
        if (i == j) {                               // matrix.py:467-468
            out[k] = as_distance ? 0 : 1000000;
            continue;
        }
        // aai similarity: length-weighted mean identity = sum(hit) / sum(len)
        // (statistics.py:22); no shared pham -> 0.0 (metrics.py:198-199)
        const double aai = len ? (double)hit / (double)len : 0.0;
        if (metric == PC_AAI) {
            out[k] = pc_finish_micro(aai, as_distance);
        } else {
            // peq multiplies the two ROUNDED factors (metrics.py:247-248)
            const double af_r = (double)out[k] / 1e6;
            const double aai_r = (double)pc_round6_micro(aai) / 1e6;
            out[k] = pc_finish_micro(af_r * aai_r, as_distance);
        }
    }
}

int pc_launch_aai_epilogue(pc_ctx *ctx, int metric, int as_distance, int raw,
                           int64_t row_begin, int64_t row_end,
                           const unsigned long long *acc, int32_t *out, int32_t *out_aux)
{
    const int64_t total = pc_tri_index(ctx->N, row_end, row_end) -
                          pc_tri_index(ctx->N, row_begin, row_begin);
    if (total <= 0) return PC_OK;
    int64_t blocks = (total + 255) / 256;
    int64_t cap = (int64_t)ctx->sm_count * 16;
    int grid = (int)(blocks < cap ? blocks : cap);
    aai_epilogue_kernel<<<grid, 256, 0, ctx->stream>>>(metric, as_distance, raw, ctx->N,
                                                      row_begin, row_end, acc, out, out_aux);
    PC_CUDA(ctx, cudaGetLastError());
    ctx->launches++;
    return PC_OK;
}
