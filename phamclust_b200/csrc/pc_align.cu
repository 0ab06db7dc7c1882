// pc_align.cu -- K4: global-alignment DP with forward (matched, length)
// counters.  Replaces parasail.nw_trace_diag_16 + get_traceback +
// comp.count("|") of /root/reference/src/phamclust/metrics.py:174-175,216-220.
//
// One SUBGROUP of G lanes per alignment (32/G alignments per warp).  Lane l
// owns K consecutive columns of the ref sequence (s2) in registers and walks
// down the query (s1) rows one step behind lane l-1 (systolic anti-diagonal);
// the right-most column's (H, E, cH, cE) travels to the next lane by
// __shfl_up.  Per cell the recurrences are exactly SURVEY.md Appendix A:
//
//   F = max(H[i-1][j]-open, F[i-1][j]-ext)   open wins only if strictly greater
//   E = max(H[i][j-1]-open, E[i][j-1]-ext)   (PC_GAP_TIE_EXTENDS)
//   H = max(Hd, F, E), source priority diag > F > E (PC_H_PRIORITY_DIAG_DEL_INS)
//
// and a packed counter c = (matched << 16) | columns rides with each state,
// selected by the same predicates that parasail's trace bits record, so
// cH[m][n] equals what the explicit traceback would count -- O(len) state, no
// trace table (the oracle checks this formulation against the explicit
// traceback).  The substitution table lives in shared memory as int2
// {score, 1 + (hit << 16)} so one LDS.64 yields both addends.
//
// Integer-issue bound (~21 ALU instructions per cell); traffic is la+lb bytes
// in, 4-8 bytes out per alignment.
#include "pc_common.cuh"

#define NW_THREADS 128

const int pc_nw_bucket_cap[PC_NW_BUCKETS] = {32, 64, 96, 128, 192, 256, 384, 512, 1 << 30};

int pc_nw_bucket_of(int ref_len)
{
    for (int b = 0; b < PC_NW_BUCKETS; ++b)
        if (ref_len <= pc_nw_bucket_cap[b]) return b;
    return PC_NW_BUCKETS - 1;
}

template <int G, int K, bool MULTI>
__global__ void __launch_bounds__(NW_THREADS) nw_kernel(const NwArgs a)
{
    __shared__ int2 s_tab[32 * 32];
    for (int x = threadIdx.x; x < 32 * 32; x += NW_THREADS) s_tab[x] = a.subtab[x];
    __syncthreads();

    constexpr int SUBS = 32 / G;
    constexpr int CAP = G * K;
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int sub = lane / G;
    const int sl = lane % G;
    const int open = a.gap_open, ext = a.gap_extend;
    const int64_t warp_id = ((int64_t)blockIdx.x * NW_THREADS + threadIdx.x) >> 5;
    int4 *spill = MULTI ? a.spill + warp_id * a.spill_stride : nullptr;

    for (;;) {
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(a.work_counter, (unsigned long long)SUBS);
        base = __shfl_sync(FULL, base, 0);
        if ((int64_t)base >= a.n_tasks) break;

        const int64_t t = (int64_t)base + sub;
        const bool valid = t < a.n_tasks;
        int m = 0, n = 0;
        int64_t slot = 0;
        const uint8_t *q = a.codes, *r = a.codes;
        if (valid) {
            slot = a.task_base + t;
            if (a.perm) slot = a.perm[slot];
            const int qg = a.task_q[slot], rg = a.task_r[slot];
            m = a.gene_len[qg];
            n = a.gene_len[rg];
            q = a.codes + a.gene_off[qg];
            r = a.codes + a.gene_off[rg];
        }
        if (valid && sl == 0 && a.cell_counter)
            atomicAdd(a.cell_counter, (unsigned long long)m * (unsigned long long)n);

        const int nchunks = MULTI ? (n + CAP - 1) / CAP : 1;
        // warp-uniform trip count (subgroups of one warp run in lock step)
        const int steps = __reduce_max_sync(FULL, m) + G - 1;

        int res_h = 0, res_c = 0;   // valid on the owner lane after the last chunk
        for (int chunk = 0; chunk < nchunks; ++chunk) {
            const int col0 = chunk * CAP + sl * K;   // 0-based first column of this lane
            int Hp[K], Fp[K], cHp[K], cFp[K], rc[K];
            // ref residues: K multiple of 4 and genes 16B aligned -> word loads
#pragma unroll
            for (int w = 0; w < K / 4; ++w) {
                const uint32_t word = __ldg(reinterpret_cast<const uint32_t *>(r + col0) + w);
#pragma unroll
                for (int b = 0; b < 4; ++b) rc[4 * w + b] = (int)((word >> (8 * b)) & 31u);
            }
#pragma unroll
            for (int x = 0; x < K; ++x) {
                const int j = col0 + x + 1;          // 1-based column
                Hp[x] = -open - (j - 1) * ext;       // H[0][j]
                Fp[x] = PC_NEG_INF;                  // F[0][j]
                cHp[x] = j;                          // cH[0][j] = (0, j)
                cFp[x] = 0;
            }
            int diagH = (col0 == 0) ? 0 : -open - (col0 - 1) * ext;   // H[0][col0]
            int diagC = col0;
            int outH = 0, outE = PC_NEG_INF, outcH = 0, outcE = 0;

            for (int s = 1; s <= steps; ++s) {
                int inH = __shfl_up_sync(FULL, outH, 1, G);
                int inE = __shfl_up_sync(FULL, outE, 1, G);
                int incH = __shfl_up_sync(FULL, outcH, 1, G);
                int incE = __shfl_up_sync(FULL, outcE, 1, G);
                const int i = s - sl;                // 1-based query row of this lane
                const bool active = (i >= 1) && (i <= m);
                if (sl == 0) {
                    if (!MULTI || chunk == 0) {
                        inH = -open - (i - 1) * ext; // H[i][0]
                        inE = PC_NEG_INF;
                        incH = i;                    // cH[i][0] = (0, i)
                        incE = 0;
                    } else if (active) {
                        const int4 v = __ldcg(spill + i);
                        inH = v.x; inE = v.y; incH = v.z; incE = v.w;
                    }
                }
                if (active) {
                    const int qc = (int)(__ldg(q + (i - 1)) & 31u);
                    const int2 *trow = s_tab + qc * 32;
                    int hl = inH, el = inE, cl = incH, ce = incE;
                    int hd = diagH, cd = diagC;
                    diagH = inH;
                    diagC = incH;
#pragma unroll
                    for (int x = 0; x < K; ++x) {
                        const int2 sc = trow[rc[x]];
                        const int fo = Hp[x] - open, fe = Fp[x] - ext;
                        const int eo = hl - open, ee = el - ext;
#if PC_GAP_TIE_EXTENDS
                        const bool pf = fo > fe, pe = eo > ee;
#else
                        const bool pf = fo >= fe, pe = eo >= ee;
#endif
                        const int f = max(fo, fe);
                        const int e = max(eo, ee);
                        const int cf = (pf ? cHp[x] : cFp[x]) + 1;
                        const int cee = (pe ? cl : ce) + 1;
                        const int hdg = hd + sc.x;
                        const int h = max(hdg, max(f, e));
#if PC_H_PRIORITY_DIAG_DEL_INS
                        const int ch = (h == hdg) ? cd + sc.y : (h == f) ? cf : cee;
#else
                        const int ch = (h == hdg) ? cd + sc.y : (h == e) ? cee : cf;
#endif
                        hd = Hp[x];
                        cd = cHp[x];
                        Hp[x] = h; Fp[x] = f; cHp[x] = ch; cFp[x] = cf;
                        hl = h; el = e; cl = ch; ce = cee;
                    }
                    outH = hl; outE = el; outcH = cl; outcE = ce;
                    if (MULTI && sl == G - 1 && chunk + 1 < nchunks)
                        __stcg(spill + i, make_int4(outH, outE, outcH, outcE));
                }
            }
            if (MULTI) __syncwarp();
            if (chunk == nchunks - 1 && n > 0) {
                const int last = n - 1 - chunk * CAP;    // column index inside the chunk
                if (sl == last / K) {
                    const int xs = last % K;
#pragma unroll
                    for (int x = 0; x < K; ++x)
                        if (x == xs) { res_h = Hp[x]; res_c = cHp[x]; }
                }
            }
        }
        if (valid) {
            bool owner;
            if (n == 0) {            // no DP columns: the answer is a boundary cell
                owner = (sl == 0);
                res_h = (m == 0) ? 0 : -open - (m - 1) * ext;
                res_c = m;
            } else {
                owner = (sl == ((n - 1) % CAP) / K);
            }
            if (owner) {
                a.out_hitlen[slot] = (uint32_t)res_c;
                if (a.out_score) a.out_score[slot] = res_h;
            }
        }
    }
}

template <int G, int K, bool MULTI>
static int launch_one(pc_ctx *ctx, const NwArgs &args)
{
    int blocks_per_sm = 0;
    PC_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(
                     &blocks_per_sm, nw_kernel<G, K, MULTI>, NW_THREADS, 0));
    if (blocks_per_sm < 1) blocks_per_sm = 1;
    int64_t groups = (args.n_tasks + (32 / G) - 1) / (32 / G);
    int64_t want = (groups + (NW_THREADS / 32) - 1) / (NW_THREADS / 32);
    int64_t cap = (int64_t)ctx->sm_count * blocks_per_sm;   // persistent: one resident wave
    int grid = (int)(want < cap ? want : cap);
    if (grid < 1) grid = 1;
    nw_kernel<G, K, MULTI><<<grid, NW_THREADS, 0, ctx->stream>>>(args);
    PC_CUDA(ctx, cudaGetLastError());
    ctx->launches++;
    ctx->align_launches++;
    return PC_OK;
}

// number of warps a launch of the multi-chunk config can have resident
int pc_nw_multichunk_warps(pc_ctx *ctx)
{
    int blocks_per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(
            &blocks_per_sm, nw_kernel<32, 16, true>, NW_THREADS, 0) != cudaSuccess)
        blocks_per_sm = 16;
    if (blocks_per_sm < 1) blocks_per_sm = 1;
    return ctx->sm_count * blocks_per_sm * (NW_THREADS / 32);
}

int pc_launch_nw(pc_ctx *ctx, const NwArgs &args, int bucket)
{
    if (args.n_tasks <= 0) return PC_OK;
    switch (bucket) {
    case 0: return launch_one<4, 8, false>(ctx, args);
    case 1: return launch_one<8, 8, false>(ctx, args);
    case 2: return launch_one<8, 12, false>(ctx, args);
    case 3: return launch_one<8, 16, false>(ctx, args);
    case 4: return launch_one<16, 12, false>(ctx, args);
    case 5: return launch_one<16, 16, false>(ctx, args);
    case 6: return launch_one<32, 12, false>(ctx, args);
    case 7: return launch_one<32, 16, false>(ctx, args);
    case 8: return launch_one<32, 16, true>(ctx, args);
    }
    return pc_fail(ctx, PC_EINVAL, "bad NW bucket %d", bucket);
}
