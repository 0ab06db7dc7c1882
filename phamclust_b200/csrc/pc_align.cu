// pc_align.cu -- K4: global-alignment DP with forward (matched, length)
// counters.  Replaces parasail.nw_trace_diag_16 + get_traceback +
// comp.count("|") of /root/reference/src/phamclust/metrics.py:174-175,216-220.
//
// One SUBGROUP of G lanes per alignment (32/G alignments per warp).  Lane l
// owns K consecutive columns of the ref sequence (s2) in registers and walks
// down the query (s1) rows one step behind lane l-1 (systolic anti-diagonal);
// the right-most column's state travels to the next lane by __shfl_up.  Per
// cell the recurrences are exactly SURVEY.md Appendix A:
//
//   F = max(H[i-1][j]-open, F[i-1][j]-ext)   open wins only if strictly greater
//   E = max(H[i][j-1]-open, E[i][j-1]-ext)   (PC_GAP_TIE_EXTENDS)
//   H = max(Hd, F, E), source priority diag > F > E (PC_H_PRIORITY_DIAG_DEL_INS)
//
// and a packed counter (matched << 16) + columns rides with each state,
// selected by the same predicates that parasail's trace bits record, so the
// counter of H[m][n] equals what the explicit traceback would count -- O(len)
// state, no trace table (the oracle checks this formulation against the
// explicit traceback).
//
// Instruction diet (the kernel is bound by the integer ALU pipe; see DESIGN.md "K4"):
//   * every state is kept DRIFTED by (i + j) * ext:  H^ = H + (i + j) ext, same for
//     E and F.  An extension then leaves a gap state unchanged and an opening adds
//     the constant C = ext - open, so each gap state is ONE instruction with no
//     helper state:  F^ = VIADDMNMX(H^up, C, F^up), and "opened" is simply
//     F^ != F^up (ties extend).  The diagonal adds score + 2 ext (in the table);
//     comparisons inside a cell are unchanged because all its candidates carry
//     the same drift; every boundary cell but (0, 0) is the constant ext - open;
//   * counters are stored RELATIVE to (i + j): gap steps then add nothing and a
//     diagonal step adds (hit << 16) - 1, which comes out of the table; the
//     true counter is recovered once at the end as d + (m + n);
//   * the substitution table is int2 {score + 2 ext, (hit << 16) - 1}: one
//     LDS.64 yields both addends.  In the usual case (<= 24 residue codes) it
//     is replicated once per lane ([24*24][32] int2 = 144 KB) so lane L only
//     ever touches bank pair L: no bank conflicts (r1 lost 2/3 of the LSU
//     wavefronts to conflicts on a shared 8 KB table).
#include "pc_common.cuh"

#define NW_TAB_CODES 24       // codes of the replicated fast-path table

// ncu (r1, profiles/): the DP is bound by the ALU pipe (SEL / ISETP / LOP3 /
// VIMNMX, 16 lanes/clk/SMSP, 90 % busy) while the FMA pipe idles.  An add
// written as a*one+b with `one` a run-time 1 cannot be folded by ptxas and
// issues as IMAD on the FMA pipe (the LEX kernel's diagonal add; PC_NW_LEX_NO_FMA turns it off).
__device__ __forceinline__ int fma_add(int a, int one, int b)
{
    int d;
    asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(one), "r"(b));
    return d;
}


// ref residue codes of this lane's K columns as table offsets (code << SHIFT); genes are
// 16 B aligned and K is even, so lane sl's first column (sl * K) is 2 B aligned at least
template <int K, int SHIFT, bool CLAMP>
__device__ __forceinline__ void load_ref_codes(const uint8_t *p, int (&rc)[K])
{
    if constexpr (K % 4 == 0) {
#pragma unroll
        for (int w = 0; w < K / 4; ++w) {
            const uint32_t word = __ldg(reinterpret_cast<const uint32_t *>(p) + w);
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                int code = (int)((word >> (8 * b)) & 31u);
                if (CLAMP && code >= NW_TAB_CODES) code = 0;   // padding past the gene only
                rc[4 * w + b] = code << SHIFT;
            }
        }
    } else {
#pragma unroll
        for (int w = 0; w < K / 2; ++w) {
            const uint32_t word = __ldg(reinterpret_cast<const uint16_t *>(p) + w);
#pragma unroll
            for (int b = 0; b < 2; ++b) {
                int code = (int)((word >> (8 * b)) & 31u);
                if (CLAMP && code >= NW_TAB_CODES) code = 0;
                rc[2 * w + b] = code << SHIFT;
            }
        }
    }
}

int pc_nw_multichunk_warps(pc_ctx *ctx);

template <int G, int K, bool MULTI, bool REPL, int THREADS>
__global__ void __launch_bounds__(THREADS, 1) nw_kernel(const NwArgs a)
{
    extern __shared__ __align__(16) int2 s_tab[];
    const int lane = threadIdx.x & 31;
    // the reference never changes _align's defaults (metrics.py:160): compile-time constants
    constexpr int open = PC_GAP_OPEN_DEFAULT, ext = PC_GAP_EXTEND_DEFAULT;
    constexpr int C = -open + ext;     // opening a gap, in drifted units
    constexpr int EDGE = -open + ext;  // H^ of every boundary cell but (0, 0)
    if (REPL) {
        // s_tab[(q*24 + r)*32 + lane] = {score(q,r) + 2 ext, (hit << 16) - 1}
        for (int x = threadIdx.x; x < NW_TAB_CODES * NW_TAB_CODES * 32; x += THREADS) {
            const int e = x >> 5;
            const int2 v = a.subtab[(e / NW_TAB_CODES) * 32 + (e % NW_TAB_CODES)];
            s_tab[x] = make_int2(v.x + 2 * ext, v.y - 2);
        }
    } else {
        for (int x = threadIdx.x; x < 32 * 32; x += THREADS) {
            const int2 v = a.subtab[x];
            s_tab[x] = make_int2(v.x + 2 * ext, v.y - 2);
        }
    }
    __syncthreads();

    constexpr int SUBS = 32 / G;
    constexpr int CAP = G * K;
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int QSTRIDE = REPL ? NW_TAB_CODES * 32 : 32;   // int2 elements per query code
    constexpr int RSHIFT = REPL ? 5 : 0;                     // ref code -> element offset
    const int sub = lane / G;
    const int sl = lane % G;
    const int64_t warp_id = ((int64_t)blockIdx.x * THREADS + threadIdx.x) >> 5;
    int4 *spill = MULTI ? a.spill + warp_id * a.spill_stride : nullptr;
    const int2 *tab_lane = REPL ? s_tab + lane : s_tab;
    // the LEXC fallback launch learns its task count from the device
    const int64_t n_tasks = a.n_tasks_dev ? (int64_t)*a.n_tasks_dev : a.n_tasks;

    for (;;) {
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(a.work_counter, (unsigned long long)SUBS);
        base = __shfl_sync(FULL, base, 0);
        if ((int64_t)base >= n_tasks) break;

        const int64_t t = (int64_t)base + sub;
        const bool valid = t < n_tasks;
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

        int res_x = 0, res_c = 0;   // valid on the owner lane after the last chunk
        for (int chunk = 0; chunk < nchunks; ++chunk) {
            const int col0 = chunk * CAP + sl * K;   // 0-based first column of this lane
            // per column: H^ and F^ of the previous row and their relative counters
            int Hp[K], Fp[K], cHp[K], cFp[K], rc[K];
            load_ref_codes<K, RSHIFT, REPL>(r + col0, rc);
#pragma unroll
            for (int x = 0; x < K; ++x) {
                Hp[x] = EDGE;                        // H^[0][j], j >= 1
                Fp[x] = PC_NEG_INF;                  // F[0][j]
                cHp[x] = 0;                          // cH[0][j] = (0, j)  -> relative 0
                cFp[x] = 0;
            }
            int diagH = (col0 == 0) ? 0 : EDGE;      // H^[0][col0]
            int diagC = 0;
            int outH = 0, outE = PC_NEG_INF, outcH = 0, outcE = 0;
            // boundary column of a later chunk: fetched one row ahead so that the L2 round
            // trip hides behind the K cells of the current row
            int4 pre = make_int4(0, 0, 0, 0);
            if (MULTI && chunk > 0 && sl == 0 && m >= 1) pre = __ldcg(spill + 1);
            for (int s = 1; s <= steps; ++s) {
                int inH = __shfl_up_sync(FULL, outH, 1, G);
                int inE = __shfl_up_sync(FULL, outE, 1, G);
                int incH = __shfl_up_sync(FULL, outcH, 1, G);
                int incE = __shfl_up_sync(FULL, outcE, 1, G);
                const int i = s - sl;                // 1-based query row of this lane
                const bool active = (i >= 1) && (i <= m);
                if (sl == 0) {
                    if (!MULTI || chunk == 0) {
                        inH = EDGE;                  // H^[i][0], i >= 1
                        inE = PC_NEG_INF;
                        incH = 0;                    // cH[i][0] = (0, i) -> relative 0
                        incE = 0;
                    } else if (active) {
                        inH = pre.x; inE = pre.y; incH = pre.z; incE = pre.w;
                        if (i < m) pre = __ldcg(spill + i + 1);
                    }
                }
                if (active) {
                    int qc = (int)(__ldg(q + (i - 1)) & 31u);
                    if (REPL && qc >= NW_TAB_CODES) qc = 0;
                    const int2 *trow = tab_lane + qc * QSTRIDE;
                    int hl = inH, el = inE, cl = incH, ce = incE;
                    int hd = diagH, cd = diagC;
                    diagH = inH;
                    diagC = incH;
#pragma unroll
                    for (int x = 0; x < K; ++x) {
                        const int2 sc = trow[rc[x]];         // {score + 2 ext, (hit<<16) - 1}
                        // F^ from above, E^ from the left: max(H^ + C, state), the state itself
                        // when the gap extends
                        const int f = __viaddmax_s32(Hp[x], C, Fp[x]);
                        const int e = __viaddmax_s32(hl, C, el);
#if PC_GAP_TIE_EXTENDS
                        const bool pf = f != Fp[x];          // opened iff H^ + C > F^ (strictly)
                        const bool pe = e != el;
#else
                        const bool pf = Hp[x] + C >= Fp[x];
                        const bool pe = hl + C >= el;
#endif
                        const int cf = pf ? cHp[x] : cFp[x];
                        const int cee = pe ? cl : ce;
                        const int hdg = hd + sc.x;           // H^[i-1][j-1] + score + 2 ext
                        const int cdg = cd + sc.y;
                        const int h = __vimax3_s32(hdg, f, e);
#if PC_H_PRIORITY_DIAG_DEL_INS
                        const int ch = (h == hdg) ? cdg : (h == f) ? cf : cee;
#else
                        const int ch = (h == hdg) ? cdg : (h == e) ? cee : cf;
#endif
                        hd = Hp[x];
                        cd = cHp[x];
                        hl = h;
                        Hp[x] = h; Fp[x] = f; cHp[x] = ch; cFp[x] = cf;
                        el = e; cl = ch; ce = cee;
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
                        if (x == xs) { res_x = Hp[x]; res_c = cHp[x]; }
                }
            }
        }
        if (valid) {
            bool owner;
            int score, counter;
            if (n == 0) {            // no DP columns: the answer is a boundary cell
                owner = (sl == 0);
                score = (m == 0) ? 0 : -open - (m - 1) * ext;
                counter = m;
            } else {
                owner = (sl == ((n - 1) % CAP) / K);
                score = res_x - (m + n) * ext;       // undo the drift
                counter = res_c + (m + n);          // undo the (i + j) normalisation
            }
            if (owner) {
                a.out_hitlen[slot] = (uint32_t)counter;
                if (a.out_score) a.out_score[slot] = score;
            }
        }
    }
}


// ---------------------------------------------------------------------------
// LEX kernel: the same (drifted) recurrences with the counters riding INSIDE
// the score word, for alignments whose scores and counters fit 32 bits
// together (pc_nw_lex_ok: min(m, n) < 2^B and a proven score range < 2^S).
//
//   word = score << (2B+2)  |  prio << 2B  |  diag steps << B  |  matched
//
// A signed max over such words picks the higher score; equal scores are
// decided by the 2-bit priority field, which encodes exactly the tie rules of
// Appendix A: stored H carries 0, E carries 1 and F carries 2 (so an extension
// beats an opening of equal score, and F beats E inside H's max), the diagonal
// candidate gets 3 from the table.  The three candidates of a max always
// differ in (score, prio), so the payload never takes part in a decision: it
// is simply carried by the winner -- no ISETP/SEL per cell, 7 ALU-pipe
// instructions instead of 11, two shuffles per row step instead of four.
// The payload counts diagonal steps d and matches; columns = m + n - d.
//   F = VIADDMNMX(Hup, C, Fup) | PRIO_F         E likewise from the left
//   H = VIMNMX3(Hdiag + T[q][r], F, E) & ~PRIO  T = (s + 2 ext, prio 3, d+1, hit)
// Scores are biased so that the proven range of the drifted values,
// [-2 open - ext, min(rowmax-sum(q), rowmax-sum(r)) + (m + n) ext], sits inside
// S bits; cells of padding columns may wrap, nothing flows from them to real cells.
// ---------------------------------------------------------------------------
//
// REBASE > 0 (the LEXR configs): B = 10-bit counters leave a 10-bit score field, far too narrow
// for absolute scores -- but not for the values one lane ever holds at one time.  DP values are
// Lipschitz (neighbouring cells differ by at most PC_NW_LEXR_LIP, E and F sit within open + LIP
// below their H), so every REBASE row-steps each lane subtracts the score of its centre column
// from all of its state and adds it to a separate register (`base`); words that cross to the
// next lane are converted between the two lanes' frames.  All comparisons stay exact because
// nothing ever leaves the signed window (PC_NW_LEXR_WINDOW_OK, checked at compile time); the
// absolute score is the field plus `base`.  Cost: 2K + 3 IMADs on the FMA pipe per REBASE * K cells.
template <int G, int K, int THREADS, int B, int REBASE>
__global__ void __launch_bounds__(THREADS, 1) nw_lex_kernel(const NwArgs a)
{
    extern __shared__ __align__(16) int s_lex[];
    const int lane = threadIdx.x & 31;
    constexpr int open = PC_GAP_OPEN_DEFAULT, ext = PC_GAP_EXTEND_DEFAULT;
    constexpr int SH = 2 * B + 2, S = 32 - SH;
    constexpr int U = 1 << SH;                       // one score unit
    constexpr int CMASK = (1 << B) - 1;
    constexpr int PRIO_E = (PC_H_PRIORITY_DIAG_DEL_INS ? 1 : 2) << (2 * B);
    constexpr int PRIO_F = (PC_H_PRIORITY_DIAG_DEL_INS ? 2 : 1) << (2 * B);
    constexpr int PRIO_D = 3 << (2 * B);
    constexpr int PMASK = 3 << (2 * B);
    constexpr int BIAS = REBASE ? 0 : -(1 << (S - 1)) + PC_NW_LEX_FLOOR;   // H^[0][0]
    constexpr int CU = (ext - open) * U;             // opening a gap, drifted
    constexpr int ORIGIN = BIAS * U;                 // H^[0][0], no diagonal steps yet
    constexpr int EDGE = (BIAS + ext - open) * U;    // every other boundary cell
    // below every real word; rebased: finite, just under any opening from a boundary cell
    constexpr int NEG = REBASE ? EDGE + CU - 4 * U : (-(1 << (S - 1)) + 2) * U;
    static_assert(PC_GAP_TIE_EXTENDS, "the LEX priorities encode 'ties extend'");
    static_assert(REBASE == 0 || ((REBASE & (REBASE - 1)) == 0 && PC_NW_LEXR_WINDOW_OK(K, REBASE)),
                  "rebased LEX: the lane's values must stay inside the score field");
    // the diagonal add issues as IMAD on the otherwise idle FMA pipe (measured +3.5 %)
    const int one = a.gap_extend;                    // == 1 at run time, opaque to ptxas

    // s_lex[(q*24 + r)*32 + lane] = (score + 2 ext, prio 3, one diagonal step, hit)
    for (int x = threadIdx.x; x < NW_TAB_CODES * NW_TAB_CODES * 32; x += THREADS) {
        const int e = x >> 5;
        const int2 v = a.subtab[(e / NW_TAB_CODES) * 32 + (e % NW_TAB_CODES)];
        s_lex[x] = (v.x + 2 * ext) * U + PRIO_D + (1 << B) + (v.y >> 16);
    }
    __syncthreads();

    constexpr int SUBS = 32 / G;
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int QSTRIDE = NW_TAB_CODES * 32;
    const int sub = lane / G;
    const int sl = lane % G;
    const int *tab_lane = s_lex + lane;

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
            atomicAdd(a.cell_counter + (REBASE ? 2 : 1), (unsigned long long)m * (unsigned long long)n);
        const int steps = __reduce_max_sync(FULL, m) + G - 1;

        const int col0 = sl * K;
        int Hp[K], Fp[K], rc[K];
        load_ref_codes<K, 5, true>(r + col0, rc);
#pragma unroll
        for (int x = 0; x < K; ++x) {
            Hp[x] = EDGE;                                // H^[0][j], j >= 1
            Fp[x] = NEG;
        }
        int diagH = (col0 == 0) ? ORIGIN : EDGE;         // H^[0][col0]
        int outH = 0, outE = NEG;
        int rb = 0, rbU = 0;                         // REBASE: score taken out of this lane's words

        for (int s = 1; s <= steps; ++s) {
            if (REBASE && (s & (REBASE - 1)) == 0) {
                // re-centre on the middle column (warp-uniform branch; every lane has its own frame)
                const int d = Hp[K / 2] & ~(U - 1);      // its score field, in place
                rb += Hp[K / 2] >> SH;
                rbU += d;
                const int nd = -d;
#pragma unroll
                for (int x = 0; x < K; ++x) {
                    Hp[x] = fma_add(Hp[x], one, nd);
                    Fp[x] = fma_add(Fp[x], one, nd);
                }
                diagH = fma_add(diagH, one, nd);
                outH = fma_add(outH, one, nd);
                outE = fma_add(outE, one, nd);
            }
            int inH = __shfl_up_sync(FULL, outH, 1, G);
            int inE = __shfl_up_sync(FULL, outE, 1, G);
            const int i = s - sl;
            const bool active = (i >= 1) && (i <= m);
            if (REBASE) {
                int inB = __shfl_up_sync(FULL, rbU, 1, G);
                if (sl == 0) { inH = EDGE; inE = NEG; inB = 0; }
                const int conv = inB - rbU;            // sender's frame -> this lane's (mod 2^32)
                inH += conv;
                inE += conv;
            } else if (sl == 0) {
                inH = EDGE;                              // H^[i][0], i >= 1
                inE = NEG;
            }
            if (active) {
                int qc = (int)(__ldg(q + (i - 1)) & 31u);
                if (qc >= NW_TAB_CODES) qc = 0;
                const int *trow = tab_lane + qc * QSTRIDE;
                int hl = inH, el = inE;
                int hd = diagH;
                diagH = inH;
#pragma unroll
                for (int x = 0; x < K; ++x) {
                    const int tw = trow[rc[x]];
                    const int f = __viaddmax_s32(Hp[x], CU, Fp[x]) | PRIO_F;
                    const int e = __viaddmax_s32(hl, CU, el) | PRIO_E;
                    const int hdg = fma_add(hd, one, tw);
                    const int h = __vimax3_s32(hdg, f, e) & ~PMASK;
                    hd = Hp[x];
                    hl = h;
                    Hp[x] = h; Fp[x] = f;
                    el = e;
                }
                outH = hl; outE = el;
            }
        }
        if (valid) {
            bool owner;
            int score, matched, columns;
            if (n == 0) {            // no DP columns: the answer is a boundary cell
                owner = (sl == 0);
                score = (m == 0) ? 0 : -open - (m - 1) * ext;
                matched = 0;
                columns = m;
            } else {
                const int last = n - 1;
                owner = (sl == last / K);
                const int xs = last % K;
                int word = 0;
#pragma unroll
                for (int x = 0; x < K; ++x)
                    if (x == xs) word = Hp[x];
                score = (word >> SH) + rb - BIAS - (m + n) * ext;     // undo rebasing, bias and drift
                matched = word & CMASK;
                columns = m + n - ((word >> B) & CMASK);
            }
            if (owner) {
                a.out_hitlen[slot] = ((uint32_t)matched << 16) | (uint32_t)columns;
                if (a.out_score) a.out_score[slot] = score;
            }
        }
    }
}

template <int G, int K, int B = PC_NW_LEX_BITS, int REBASE = 0>
static int launch_lex(pc_ctx *ctx, const NwArgs &args, cudaStream_t stream)
{
    // 3 words of state per column: more warps fit than in the classic kernel, and K goes to 48
    // (registers per thread by ptxas: 57 / 71 / 86 / 101 / 117 / 133 / 148 for K = 8, 12 .. 32)
    // warps per SM sub-partition (16384 registers each): 8 / 7 / 6 / 5 / 4 / 3 / 2
    // (the rebased variant carries two more registers: one block size earlier from K = 28)
    constexpr int THREADS = K <= 10 ? 1024 : K <= 12 ? 896 : K <= 14 ? 768 : K <= 18 ? 640
                          : K <= (REBASE ? 26 : 28) ? 512 : K <= 38 ? 384 : 256;
    const size_t smem = (size_t)NW_TAB_CODES * NW_TAB_CODES * 32 * sizeof(int);
    auto kern = nw_lex_kernel<G, K, THREADS, B, REBASE>;
    PC_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int blocks_per_sm = 0;
    PC_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, kern, THREADS, smem));
    if (blocks_per_sm < 1) blocks_per_sm = 1;
    const int64_t groups = (args.n_tasks + (32 / G) - 1) / (32 / G);
    const int64_t want = (groups + (THREADS / 32) - 1) / (THREADS / 32);
    const int64_t cap = (int64_t)ctx->sm_count * blocks_per_sm;
    int grid = (int)(want < cap ? want : cap);
    if (grid < 1) grid = 1;
    kern<<<grid, THREADS, smem, stream>>>(args);
    PC_CUDA(ctx, cudaGetLastError());
    ctx->launches++;
    ctx->align_launches++;
    return PC_OK;
}


// ---------------------------------------------------------------------------
// LEXC kernel: the rebased LEX word with the two COUNTERS rebased as well, in column chunks --
// any length, any score range (pc_common.cuh, PC_NW_LEXC_*; executable model with the soundness
// argument: oracle/lexc_model.c).  One warp per alignment.
//
//   word  = score << 22 | prio << 20 | (diagonal steps - base_d) << 10 | (matches - base_m)
//   frame = (score base rb, base_d, base_m) per lane; rbU is the packed frame (mod 2^32) used to
//           convert words between lanes, rbDM = (base_d + 512) << 16 | (base_m + 512) the centre
//           column's true counters at the last re-centring
// Every R = 8 row-steps a lane subtracts (centre word - C0) from all of its state, C0 = both
// counter fields at 512, and CHECKS that every counter field then lies in [384, 640): one IMAD
// (FMA pipe) + one LOP3 per word.  Every row-step the word pair received from the left lane (or
// from the spilled boundary column of the previous chunk) is converted by the difference of the
// two frames, the two lanes' counter bases are checked to be within 255 of each other and the
// converted words to be inside the band.  While no check has failed no field has borrowed from
// or carried into another, so every comparison and every counter is exact; a failed check lists
// the alignment for the classic kernel (a.fallback_slots) instead of writing a result.
// ---------------------------------------------------------------------------
template <int K, int THREADS>
__global__ void __launch_bounds__(THREADS, 1) nw_lexc_kernel(const NwArgs a)
{
    extern __shared__ __align__(16) int s_lex[];
    const int lane = threadIdx.x & 31;
    constexpr int open = PC_GAP_OPEN_DEFAULT, ext = PC_GAP_EXTEND_DEFAULT;
    constexpr int G = 32, B = PC_NW_LEXC_BITS, R = PC_NW_LEXC_REBASE, BAND = PC_NW_LEXC_BAND;
    constexpr int SH = 2 * B + 2;
    constexpr int U = 1 << SH;
    constexpr int CMASK = (1 << B) - 1;
    constexpr int PRIO_E = (PC_H_PRIORITY_DIAG_DEL_INS ? 1 : 2) << (2 * B);
    constexpr int PRIO_F = (PC_H_PRIORITY_DIAG_DEL_INS ? 2 : 1) << (2 * B);
    constexpr int PRIO_D = 3 << (2 * B);
    constexpr int PMASK = 3 << (2 * B);
    constexpr int c0 = 1 << (B - 1);
    constexpr int C0 = (c0 << B) | c0;                       // both counter fields centred
    constexpr int LO = ((c0 - BAND) << B) | (c0 - BAND);     // lower edge of the band, both fields
    constexpr int FTOP = CMASK & ~(2 * BAND - 1);            // bits a field inside the band never has (after - LO)
    constexpr int TOPMASK = (FTOP << B) | FTOP;
    constexpr int DBIAS = (PC_NW_LEXC_FRAME << 16) | PC_NW_LEXC_FRAME;
    constexpr unsigned DMASK = 0xFE00FE00u;                  // frame difference outside [-255, 256]
    constexpr int CU = (ext - open) * U;                     // opening a gap, drifted
    constexpr int EDGE = (ext - open) * U + C0;              // boundary cells but (0, 0), initial frame
    constexpr int ORIGIN = C0;                               // H^[0][0]
    constexpr int NEG = EDGE + CU - 4 * U;                   // just under any opening from a boundary cell
    constexpr int CAP = G * K;
    static_assert(PC_GAP_TIE_EXTENDS, "the LEX priorities encode 'ties extend'");
    static_assert((R & (R - 1)) == 0 && PC_NW_LEXR_WINDOW_OK(K, R), "LEXC: the score window of LEXR");
    static_assert(2 * BAND == 256 && PC_NW_LEXC_FRAME == 255, "LEXC: the masks above");
    const int one = a.gap_extend;                            // == 1 at run time, opaque to ptxas

    for (int x = threadIdx.x; x < NW_TAB_CODES * NW_TAB_CODES * 32; x += THREADS) {
        const int e = x >> 5;
        const int2 v = a.subtab[(e / NW_TAB_CODES) * 32 + (e % NW_TAB_CODES)];
        s_lex[x] = (v.x + 2 * ext) * U + PRIO_D + (1 << B) + (v.y >> 16);
    }
    __syncthreads();

    constexpr unsigned FULL = 0xffffffffu;
    constexpr int QSTRIDE = NW_TAB_CODES * 32;
    const int64_t warp_id = ((int64_t)blockIdx.x * THREADS + threadIdx.x) >> 5;
    int4 *spill = a.spill + warp_id * a.spill_stride;
    const int *tab_lane = s_lex + lane;

    for (;;) {
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(a.work_counter, 1ull);
        base = __shfl_sync(FULL, base, 0);
        if ((int64_t)base >= a.n_tasks) break;

        int64_t slot = a.task_base + (int64_t)base;
        if (a.perm) slot = a.perm[slot];
        const int qg = a.task_q[slot], rg = a.task_r[slot];
        const int m = a.gene_len[qg], n = a.gene_len[rg];
        const uint8_t *q = a.codes + a.gene_off[qg], *r = a.codes + a.gene_off[rg];
        const int nchunks = (n + CAP - 1) / CAP;
        const int steps = m + G - 1;

        unsigned chk_all = 0;
        bool flagged = false;
        int res_word = 0, res_rb = 0, res_dm = 0;
        for (int chunk = 0; chunk < nchunks && !flagged; ++chunk) {
            const int col0 = chunk * CAP + lane * K;
            int Hp[K], Fp[K], rc[K];
            load_ref_codes<K, 5, true>(r + col0, rc);
#pragma unroll
            for (int x = 0; x < K; ++x) {
                Hp[x] = EDGE;
                Fp[x] = NEG;
            }
            int diagH = (col0 == 0) ? ORIGIN : EDGE;
            int outH = C0, outE = NEG;
            int rb = 0, rbDM = 0;
            unsigned rbU = 0u - (unsigned)C0;
            unsigned chk = 0;
            int4 pre = make_int4(0, 0, 0, 0);
            if (chunk > 0 && lane == 0 && m >= 1) pre = __ldcg(spill + 1);

            for (int s = 1; s <= steps; ++s) {
                if ((s & (R - 1)) == 0) {
                    // re-centre on the middle column, then check the band (warp-uniform branch)
                    const int wc = Hp[K / 2];
                    const int nd = C0 - wc;                  // minus the shift of the frame
                    const int ndl = nd - LO;
                    rb += wc >> SH;
                    rbU -= (unsigned)nd;
                    rbDM += ((wc >> B) & CMASK) * 65536 + (wc & CMASK) - (c0 * 65536 + c0);
#pragma unroll
                    for (int x = 0; x < K; ++x) {
                        chk |= (unsigned)(fma_add(Hp[x], one, ndl) & TOPMASK);
                        Hp[x] = fma_add(Hp[x], one, nd);
                        chk |= (unsigned)(fma_add(Fp[x], one, ndl) & TOPMASK);
                        Fp[x] = fma_add(Fp[x], one, nd);
                    }
                    chk |= (unsigned)(fma_add(diagH, one, ndl) & TOPMASK);
                    diagH = fma_add(diagH, one, nd);
                    chk |= (unsigned)(fma_add(outH, one, ndl) & TOPMASK);
                    outH = fma_add(outH, one, nd);
                    chk |= (unsigned)(fma_add(outE, one, ndl) & TOPMASK);
                    outE = fma_add(outE, one, nd);
                    // lanes that hold padding columns only do not count; give up early
                    if (__any_sync(FULL, chk != 0 && col0 < n)) { flagged = true; break; }
                }
                int inH = __shfl_up_sync(FULL, outH, 1);
                int inE = __shfl_up_sync(FULL, outE, 1);
                unsigned inB = __shfl_up_sync(FULL, rbU, 1);
                int inDM = __shfl_up_sync(FULL, rbDM, 1);
                const int i = s - lane;
                const bool active = (i >= 1) && (i <= m);
                if (lane == 0) {
                    if (chunk == 0 || !active) {
                        inH = EDGE; inE = NEG; inB = 0u - (unsigned)C0; inDM = 0;
                    } else {
                        inH = pre.x; inE = pre.y; inB = (unsigned)pre.z; inDM = pre.w;
                        if (i < m) pre = __ldcg(spill + i + 1);
                    }
                }
                const int conv = (int)(inB - rbU);           // sender's frame -> this lane's (mod 2^32)
                inH += conv;
                inE += conv;
                if (active) {
                    chk |= (unsigned)(inDM - rbDM + DBIAS) & DMASK;
                    chk |= (unsigned)((inH - LO) & TOPMASK);
                    chk |= (unsigned)((inE - LO) & TOPMASK);
                    int qc = (int)(__ldg(q + (i - 1)) & 31u);
                    if (qc >= NW_TAB_CODES) qc = 0;
                    const int *trow = tab_lane + qc * QSTRIDE;
                    int hl = inH, el = inE;
                    int hd = diagH;
                    diagH = inH;
#pragma unroll
                    for (int x = 0; x < K; ++x) {
                        const int tw = trow[rc[x]];
                        const int f = __viaddmax_s32(Hp[x], CU, Fp[x]) | PRIO_F;
                        const int e = __viaddmax_s32(hl, CU, el) | PRIO_E;
                        const int hdg = fma_add(hd, one, tw);
                        const int h = __vimax3_s32(hdg, f, e) & ~PMASK;
                        hd = Hp[x];
                        hl = h;
                        Hp[x] = h; Fp[x] = f;
                        el = e;
                    }
                    outH = hl; outE = el;
                    if (lane == G - 1 && chunk + 1 < nchunks)
                        __stcg(spill + i, make_int4(outH, outE, (int)rbU, rbDM));
                }
            }
            __syncwarp();
            if (col0 < n) chk_all |= chk;
            if (chunk == nchunks - 1) {
                const int last = n - 1 - chunk * CAP;        // column index inside the chunk
                if (lane == last / K) {
                    const int xs = last % K;
#pragma unroll
                    for (int x = 0; x < K; ++x)
                        if (x == xs) res_word = Hp[x];
                    res_rb = rb;
                    res_dm = rbDM;
                }
            }
        }
        flagged = flagged || __any_sync(FULL, chk_all != 0);
        if (flagged) {
            if (lane == 0) a.fallback_slots[atomicAdd(a.fallback_count, 1ull)] = slot;
        } else {
            if (lane == 0 && a.cell_counter)
                atomicAdd(a.cell_counter + 3, (unsigned long long)m * (unsigned long long)n);
            if (n == 0) {            // no DP columns: the answer is a boundary cell
                if (lane == 0) {
                    a.out_hitlen[slot] = (uint32_t)m;
                    if (a.out_score) a.out_score[slot] = (m == 0) ? 0 : -open - (m - 1) * ext;
                }
            } else if (lane == ((n - 1) % CAP) / K) {
                const int matched = (res_word & CMASK) + (res_dm & 0xffff) - c0;
                const int columns = m + n - (((res_word >> B) & CMASK) + (res_dm >> 16) - c0);
                a.out_hitlen[slot] = ((uint32_t)matched << 16) | (uint32_t)columns;
                if (a.out_score) a.out_score[slot] = (res_word >> SH) + res_rb - (m + n) * ext;
            }
        }
    }
}

template <int K>
static int launch_lexc(pc_ctx *ctx, const NwArgs &args, cudaStream_t stream)
{
    // a few registers more than the LEXR variant of the same K (frame, check word, prefetch)
    constexpr int THREADS = K <= 16 ? 640 : K <= 22 ? 512 : 384;
    const size_t smem = (size_t)NW_TAB_CODES * NW_TAB_CODES * 32 * sizeof(int);
    auto kern = nw_lexc_kernel<K, THREADS>;
    PC_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int blocks_per_sm = 0;
    PC_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, kern, THREADS, smem));
    if (blocks_per_sm < 1) blocks_per_sm = 1;
    const int64_t want = (args.n_tasks + (THREADS / 32) - 1) / (THREADS / 32);
    const int64_t cap = (int64_t)ctx->sm_count * blocks_per_sm;
    int grid = (int)(want < cap ? want : cap);
    if (grid < 1) grid = 1;
    if ((int64_t)grid * (THREADS / 32) > pc_nw_multichunk_warps(ctx))
        return pc_fail(ctx, PC_EINVAL, "LEXC launch of %d warps exceeds the spill region", grid * (THREADS / 32));
    kern<<<grid, THREADS, smem, stream>>>(args);
    PC_CUDA(ctx, cudaGetLastError());
    ctx->launches++;
    ctx->align_launches++;
    return PC_OK;
}

// the G x K grid the chooser can pick (pc_nw_choose_cfg): K >= 16 for G >= 8
#define PC_K_CASE(K_, CALL) case K_: if constexpr (G == 4 || K_ >= 16) { return CALL; } break;

template <int G>
static int launch_lex_k(pc_ctx *ctx, const NwArgs &args, int K, cudaStream_t stream)
{
#define PC_LEX_CASE(K_) PC_K_CASE(K_, (launch_lex<G, K_>(ctx, args, stream)))
    switch (K) {
    PC_LEX_CASE(8) PC_LEX_CASE(10) PC_LEX_CASE(12) PC_LEX_CASE(14) PC_LEX_CASE(16) PC_LEX_CASE(18)
    PC_LEX_CASE(20) PC_LEX_CASE(22) PC_LEX_CASE(24) PC_LEX_CASE(26) PC_LEX_CASE(28) PC_LEX_CASE(30)
    PC_LEX_CASE(32) PC_LEX_CASE(34) PC_LEX_CASE(36) PC_LEX_CASE(38) PC_LEX_CASE(40) PC_LEX_CASE(42)
    PC_LEX_CASE(44) PC_LEX_CASE(46) PC_LEX_CASE(48)
    }
#undef PC_LEX_CASE
    return pc_fail(ctx, PC_EINVAL, "no LEX DP kernel for G=%d K=%d", G, K);
}

template <int G, int K, bool MULTI, bool REPL, int THREADS>
static int launch_cfg(pc_ctx *ctx, const NwArgs &args, cudaStream_t stream)
{
    const size_t smem = REPL ? (size_t)NW_TAB_CODES * NW_TAB_CODES * 32 * sizeof(int2)
                             : (size_t)32 * 32 * sizeof(int2);
    auto kern = nw_kernel<G, K, MULTI, REPL, THREADS>;
    PC_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int blocks_per_sm = 0;
    PC_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, kern, THREADS, smem));
    if (blocks_per_sm < 1) blocks_per_sm = 1;
    const int64_t groups = (args.n_tasks + (32 / G) - 1) / (32 / G);
    const int64_t want = (groups + (THREADS / 32) - 1) / (THREADS / 32);
    const int64_t cap = (int64_t)ctx->sm_count * blocks_per_sm;   // persistent: one resident wave
    int grid = (int)(want < cap ? want : cap);
    if (grid < 1) grid = 1;
    kern<<<grid, THREADS, smem, stream>>>(args);
    PC_CUDA(ctx, cudaGetLastError());
    ctx->launches++;
    ctx->align_launches++;
    return PC_OK;
}

template <int G, int K, bool MULTI, int THREADS>
static int launch_one(pc_ctx *ctx, const NwArgs &args, bool repl, cudaStream_t stream)
{
    if (repl) return launch_cfg<G, K, MULTI, true, THREADS>(ctx, args, stream);
    // the 32-code table (inputs with residue symbols beyond the 24 matrix letters) only
    // comes with K in multiples of 4 (pc_nw_choose_cfg's coarse grid)
    if constexpr (K % 4 == 0) return launch_cfg<G, K, MULTI, false, THREADS>(ctx, args, stream);
    return pc_fail(ctx, PC_EINVAL, "no 32-code DP kernel for G=%d K=%d", G, K);
}

// upper bound on the warps a launch of the multi-chunk config keeps resident
int pc_nw_multichunk_warps(pc_ctx *ctx)
{
    return ctx->sm_count * 12 * 2;     // <= 384-thread blocks, at most 2 per SM (non-replicated table)
}

// one CTA per SM (the replicated table takes 144 KB); as many warps as the
// per-thread state of K columns leaves room for in the register file
template <int G, int K, bool MULTI>
static int launch_gk(pc_ctx *ctx, const NwArgs &args, bool repl, cudaStream_t stream)
{
    // (registers are split per SM sub-partition: 6/5/4/3/2 warps each -> 80/96/128/168/255 per thread)
    constexpr int THREADS = MULTI ? (K <= 20 ? 384 : 256)
                          : K <= 8 ? 768 : K <= 10 ? 640 : K <= 14 ? 512 : K <= 20 ? 384 : 256;
    return launch_one<G, K, MULTI, THREADS>(ctx, args, repl, stream);
}

template <int G>
static int launch_k(pc_ctx *ctx, const NwArgs &args, bool repl, int K, cudaStream_t stream)
{
#define PC_CLS_CASE(K_) PC_K_CASE(K_, (launch_gk<G, K_, false>(ctx, args, repl, stream)))
    switch (K) {
    PC_CLS_CASE(8) PC_CLS_CASE(10) PC_CLS_CASE(12) PC_CLS_CASE(14) PC_CLS_CASE(16) PC_CLS_CASE(18)
    PC_CLS_CASE(20) PC_CLS_CASE(22) PC_CLS_CASE(24) PC_CLS_CASE(26) PC_CLS_CASE(28) PC_CLS_CASE(30)
    PC_CLS_CASE(32)
    }
#undef PC_CLS_CASE
    return pc_fail(ctx, PC_EINVAL, "no DP kernel for G=%d K=%d", G, K);
}

// the launch that re-runs what LEXC gave up on: classic <32, 20, multi-chunk> takes any length
int pc_launch_nw_fallback(pc_ctx *ctx, const NwArgs &args, cudaStream_t stream)
{
    if (args.n_tasks <= 0) return PC_OK;
    return launch_gk<32, 20, true>(ctx, args, ctx->n_codes <= NW_TAB_CODES, stream);
}

__global__ void add_counter_kernel(unsigned long long *total, const unsigned long long *value)
{
    *total += *value;
}

void pc_launch_add_counter(unsigned long long *total, const unsigned long long *value, cudaStream_t stream)
{
    add_counter_kernel<<<1, 1, 0, stream>>>(total, value);
}

int pc_launch_nw(pc_ctx *ctx, const NwArgs &args, int cfg, cudaStream_t stream)
{
    if (args.n_tasks <= 0) return PC_OK;
    if (cfg < 0 || cfg >= PC_NW_NCFG) return pc_fail(ctx, PC_EINVAL, "bad DP config %d", cfg);
    ctx->cfg_used[cfg] = 1;
    const bool repl = ctx->n_codes <= NW_TAB_CODES;
    int G, K;
    pc_nw_cfg(cfg, G, K);
    if (cfg >= PC_NW_LEXC0) {
        if (!repl) return pc_fail(ctx, PC_EINVAL, "LEX DP config %d needs the 24-code table", cfg);
        if (!args.spill || !args.fallback_count || !args.fallback_slots)
            return pc_fail(ctx, PC_EINVAL, "LEXC DP config %d needs a spill region and a fallback list", cfg);
        switch (K) {
        case 16: return launch_lexc<16>(ctx, args, stream);
        case 18: return launch_lexc<18>(ctx, args, stream);
        case 20: return launch_lexc<20>(ctx, args, stream);
        case 22: return launch_lexc<22>(ctx, args, stream);
        case 24: return launch_lexc<24>(ctx, args, stream);
        }
        return pc_fail(ctx, PC_EINVAL, "no LEXC DP kernel for K=%d", K);
    }
    if (cfg >= PC_NW_LEXR0) {
        if (!repl) return pc_fail(ctx, PC_EINVAL, "LEX DP config %d needs the 24-code table", cfg);
#define PC_LEXR_CASE(K_) case K_: return launch_lex<32, K_, PC_NW_LEXR_BITS, PC_NW_LEXR_REBASE(K_)>(ctx, args, stream);
        switch (K) {
        PC_LEXR_CASE(16) PC_LEXR_CASE(18) PC_LEXR_CASE(20) PC_LEXR_CASE(22) PC_LEXR_CASE(24)
        PC_LEXR_CASE(26) PC_LEXR_CASE(28) PC_LEXR_CASE(30) PC_LEXR_CASE(32)
        }
#undef PC_LEXR_CASE
        return pc_fail(ctx, PC_EINVAL, "no rebased LEX DP kernel for K=%d", K);
    }
    if (cfg >= PC_NW_LEX0) {
        if (!repl) return pc_fail(ctx, PC_EINVAL, "LEX DP config %d needs the 24-code table", cfg);
        switch (G) {
        case 4: return launch_lex_k<4>(ctx, args, K, stream);
        case 8: return launch_lex_k<8>(ctx, args, K, stream);
        case 16: return launch_lex_k<16>(ctx, args, K, stream);
        default: return launch_lex_k<32>(ctx, args, K, stream);
        }
    }
    if (cfg >= PC_NW_MULTI0) {
        switch (K) {
        case 20: return launch_gk<32, 20, true>(ctx, args, repl, stream);
        case 24: return launch_gk<32, 24, true>(ctx, args, repl, stream);
        case 28: return launch_gk<32, 28, true>(ctx, args, repl, stream);
        default: return launch_gk<32, 32, true>(ctx, args, repl, stream);
        }
    }
    switch (G) {
    case 4: return launch_k<4>(ctx, args, repl, K, stream);
    case 8: return launch_k<8>(ctx, args, repl, K, stream);
    case 16: return launch_k<16>(ctx, args, repl, K, stream);
    case 32: return launch_k<32>(ctx, args, repl, K, stream);
    }
    return pc_fail(ctx, PC_EINVAL, "bad DP config %d", cfg);
}
