/*
 * oracle/nw_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the one third-party call on PhamClust's aai/peq path:
 *
 *     parasail.nw_trace_diag_16(seq_a, seq_b, 11, 1, blosum62)
 *         .get_traceback(mch="|", sim="+", neg=" ")
 *
 * Reference call sites: /root/reference/src/phamclust/metrics.py:20 (import),
 * :160 (defaults open=11, extend=1, blosum62), :174 (nw_trace_diag_16),
 * :175 (get_traceback), consumers :216-220 (len(query), comp.count("|"),
 * comp.count("+")).
 *
 * The arithmetic lives in the third-party module `parasail` (PyPI
 * parasail~=1.3.0, requirements.txt:5 / setup.cfg:38; it wraps the parasail C
 * library 2.6.x).  That source is NOT under /root/reference and the wheel is
 * not installable offline, and the reference ships no tests or golden vectors,
 * so THIS FILE RESTATES parasail's PUBLISHED ALGORITHM:
 *
 *     *** PARITY UNPINNED AT THE PARASAIL BOUNDARY ***
 *
 * What can be pinned is pinned in tests/: (i) score optimality against the
 * independent textbook Gotoh DP below (nw_gotoh_score), (ii) the emitted
 * traceback re-scores to the DP score and consumes both sequences,
 * (iii) forward counters == explicit traceback, (iv) the CUDA kernel ==
 * this file, bit for bit, on (ident, similar, length).
 *
 * Every tie-break rule that parasail's result depends on sits behind ONE
 * switch below (U1..U6 of the survey's uncertainty register) so it can be
 * flipped if a live parasail ever becomes available; the CUDA kernel reads
 * the same rules (phamclust_b200/csrc/nw_rules.h mirrors them).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <ctype.h>

/* ---- rule switches (SURVEY.md Appendix A uncertainty register) ---------- */
/* U1: a gap of length k costs open + (k-1)*extend                           */
/* U2: H tie priority diag > F (DEL, gap in ref) > E (INS, gap in query)     */
#define NW_H_PRIORITY_DIAG_DEL_INS 1
/* U3: open-vs-extend tie resolves to EXTEND (open wins only if strictly >)  */
#define NW_GAP_TIE_EXTENDS 1
/* U5: traceback match test compares characters case-insensitively           */
#define NW_CASE_SENSITIVE 0
/* U6: characters outside the matrix alphabet map to the '*' column          */

#define NEG_INF (-(1 << 28))

/* trace bits, one byte per cell */
#define T_DIAG   1
#define T_DEL    2   /* H came from F: consumes s1[i], gap in ref   */
#define T_INS    4   /* H came from E: consumes s2[j], gap in query */
#define T_DIAG_F 8   /* F opened from H[i-1][j]                      */
#define T_DIAG_E 16  /* E opened from H[i][j-1]                      */

static const char ALPHABET[] = "ARNDCQEGHILKMFPSTWYVBZX*";

/* NCBI BLOSUM62, rows/cols in ALPHABET order. */
static const int8_t BLOSUM62[24][24] = {
    { 4,-1,-2,-2, 0,-1,-1, 0,-2,-1,-1,-1,-1,-2,-1, 1, 0,-3,-2, 0,-2,-1, 0,-4},
    {-1, 5, 0,-2,-3, 1, 0,-2, 0,-3,-2, 2,-1,-3,-2,-1,-1,-3,-2,-3,-1, 0,-1,-4},
    {-2, 0, 6, 1,-3, 0, 0, 0, 1,-3,-3, 0,-2,-3,-2, 1, 0,-4,-2,-3, 3, 0,-1,-4},
    {-2,-2, 1, 6,-3, 0, 2,-1,-1,-3,-4,-1,-3,-3,-1, 0,-1,-4,-3,-3, 4, 1,-1,-4},
    { 0,-3,-3,-3, 9,-3,-4,-3,-3,-1,-1,-3,-1,-2,-3,-1,-1,-2,-2,-1,-3,-3,-2,-4},
    {-1, 1, 0, 0,-3, 5, 2,-2, 0,-3,-2, 1, 0,-3,-1, 0,-1,-2,-1,-2, 0, 3,-1,-4},
    {-1, 0, 0, 2,-4, 2, 5,-2, 0,-3,-3, 1,-2,-3,-1, 0,-1,-3,-2,-2, 1, 4,-1,-4},
    { 0,-2, 0,-1,-3,-2,-2, 6,-2,-4,-4,-2,-3,-3,-2, 0,-2,-2,-3,-3,-1,-2,-1,-4},
    {-2, 0, 1,-1,-3, 0, 0,-2, 8,-3,-3,-1,-2,-1,-2,-1,-2,-2, 2,-3, 0, 0,-1,-4},
    {-1,-3,-3,-3,-1,-3,-3,-4,-3, 4, 2,-3, 1, 0,-3,-2,-1,-3,-1, 3,-3,-3,-1,-4},
    {-1,-2,-3,-4,-1,-2,-3,-4,-3, 2, 4,-2, 2, 0,-3,-2,-1,-2,-1, 1,-4,-3,-1,-4},
    {-1, 2, 0,-1,-3, 1, 1,-2,-1,-3,-2, 5,-1,-3,-1, 0,-1,-3,-2,-2, 0, 1,-1,-4},
    {-1,-1,-2,-3,-1, 0,-2,-3,-2, 1, 2,-1, 5, 0,-2,-1,-1,-1,-1, 1,-3,-1,-1,-4},
    {-2,-3,-3,-3,-2,-3,-3,-3,-1, 0, 0,-3, 0, 6,-4,-2,-2, 1, 3,-1,-3,-3,-1,-4},
    {-1,-2,-2,-1,-3,-1,-1,-2,-2,-3,-3,-1,-2,-4, 7,-1,-1,-4,-3,-2,-2,-1,-2,-4},
    { 1,-1, 1, 0,-1, 0, 0, 0,-1,-2,-2, 0,-1,-2,-1, 4, 1,-3,-2,-2, 0, 0, 0,-4},
    { 0,-1, 0,-1,-1,-1,-1,-2,-2,-1,-1,-1,-1,-2,-1, 1, 5,-2,-2, 0,-1,-1, 0,-4},
    {-3,-3,-4,-4,-2,-2,-3,-2,-2,-3,-2,-3,-1, 1,-4,-3,-2,11, 2,-3,-4,-3,-2,-4},
    {-2,-2,-2,-3,-2,-1,-2,-3, 2,-1,-1,-2,-1, 3,-3,-2,-2, 2, 7,-1,-3,-2,-1,-4},
    { 0,-3,-3,-3,-1,-2,-2,-3,-3, 3, 1,-2, 1,-1,-2,-2, 0,-3,-1, 4,-3,-2,-1,-4},
    {-2,-1, 3, 4,-3, 0, 1,-1, 0,-3,-4, 0,-3,-3,-2, 0,-1,-4,-3,-3, 4, 1,-1,-4},
    {-1, 0, 0, 1,-3, 3, 4,-2, 0,-3,-3, 1,-1,-3,-1, 0,-1,-3,-2,-2, 1, 4,-1,-4},
    { 0,-1,-1,-1,-2,-1,-1,-1,-1,-1,-1,-1,-1,-1,-2, 0, 0,-2,-1,-1,-1,-1,-1,-4},
    {-4,-4,-4,-4,-4,-4,-4,-4,-4,-4,-4,-4,-4,-4,-4,-4,-4,-4,-4,-4,-4,-4,-4, 1},
};

static int mapper_ready = 0;
static uint8_t MAPPER[256];

static void init_mapper(void)
{
    if (mapper_ready) return;
    memset(MAPPER, 23, sizeof MAPPER);            /* U6: unknown -> '*' */
    for (int k = 0; k < 24; ++k) {
        MAPPER[(unsigned char)ALPHABET[k]] = (uint8_t)k;
        MAPPER[(unsigned char)tolower((unsigned char)ALPHABET[k])] = (uint8_t)k;
    }
    mapper_ready = 1;
}

/* exported so tests can read the exact table the oracle used */
int nw_blosum62(int row, int col) { return BLOSUM62[row][col]; }
int nw_map_char(int c) { init_mapper(); return MAPPER[c & 255]; }

static inline int chars_match(unsigned char a, unsigned char b)
{
#if NW_CASE_SENSITIVE
    return a == b;
#else
    return toupper(a) == toupper(b);
#endif
}

static inline int imax(int a, int b) { return a > b ? a : b; }

/*
 * Fill the DP with trace bits exactly as SURVEY.md Appendix A states
 * parasail's nw_trace does; returns the global score H[m][n].
 * trace is m*n bytes, row-major over (i-1, j-1).
 */
static int fill_trace(const unsigned char *s1, int m, const unsigned char *s2,
                      int n, int open, int ext, uint8_t *trace)
{
    int *H = (int *)malloc(sizeof(int) * (size_t)(n + 1));
    int *F = (int *)malloc(sizeof(int) * (size_t)(n + 1));
    H[0] = 0;
    F[0] = NEG_INF;
    for (int j = 1; j <= n; ++j) {
        H[j] = -open - (j - 1) * ext;             /* U1 boundary */
        F[j] = NEG_INF;
    }
    for (int i = 1; i <= m; ++i) {
        const int8_t *row = BLOSUM62[MAPPER[s1[i - 1]]];
        int NH = H[0];                            /* H[i-1][0] */
        int WH = -open - (i - 1) * ext;           /* H[i][0]   */
        int E = NEG_INF;
        H[0] = WH;
        for (int j = 1; j <= n; ++j) {
            int NWH = NH;                         /* H[i-1][j-1] */
            NH = H[j];                            /* H[i-1][j]   */
            int F_opn = NH - open, F_ext = F[j] - ext;
            int E_opn = WH - open, E_ext = E - ext;
            uint8_t t;
            F[j] = imax(F_opn, F_ext);
            E = imax(E_opn, E_ext);
#if NW_GAP_TIE_EXTENDS
            t = (F_opn > F_ext) ? T_DIAG_F : 0;
            t |= (E_opn > E_ext) ? T_DIAG_E : 0;
#else
            t = (F_opn >= F_ext) ? T_DIAG_F : 0;
            t |= (E_opn >= E_ext) ? T_DIAG_E : 0;
#endif
            int H_dag = NWH + row[MAPPER[s2[j - 1]]];
            WH = imax(imax(H_dag, F[j]), E);
#if NW_H_PRIORITY_DIAG_DEL_INS
            t |= (WH == H_dag) ? T_DIAG : (WH == F[j]) ? T_DEL : T_INS;
#else
            t |= (WH == H_dag) ? T_DIAG : (WH == E) ? T_INS : T_DEL;
#endif
            H[j] = WH;
            trace[(size_t)(i - 1) * (size_t)n + (size_t)(j - 1)] = t;
        }
    }
    int score = H[n];
    free(H);
    free(F);
    return score;
}

/*
 * Explicit traceback state machine (Appendix A).  If q/c/r are non-NULL the
 * three traceback strings are written REVERSED-then-fixed (i.e. in reading
 * order) and NUL-terminated; buffers must hold m+n+1 bytes.
 * out[0]=identical ('|') columns, out[1]=similar ('+') columns,
 * out[2]=alignment length, out[3]=score.
 */
int nw_trace_stats(const char *a, int m, const char *b, int n, int open,
                   int ext, int *out, char *q, char *c, char *r)
{
    init_mapper();
    const unsigned char *s1 = (const unsigned char *)a;
    const unsigned char *s2 = (const unsigned char *)b;
    uint8_t *trace = NULL;
    int score;
    if (m > 0 && n > 0) {
        trace = (uint8_t *)malloc((size_t)m * (size_t)n);
        if (!trace) return -1;
        score = fill_trace(s1, m, s2, n, open, ext, trace);
    } else if (m == 0 && n == 0) {
        score = 0;
    } else {
        int k = m > n ? m : n;
        score = -open - (k - 1) * ext;
    }
    int i = m - 1, j = n - 1, where = T_DIAG;
    int ident = 0, similar = 0, len = 0;
    while (i >= 0 || j >= 0) {
        char qc, cc, rc;
        if (i < 0) {
            qc = '-'; rc = (char)s2[j]; cc = ' '; --j;
        } else if (j < 0) {
            qc = (char)s1[i]; rc = '-'; cc = ' '; --i;
        } else {
            uint8_t t = trace[(size_t)i * (size_t)n + (size_t)j];
            if (where == T_DIAG) {
                if (t & T_DIAG) {
                    qc = (char)s1[i]; rc = (char)s2[j];
                    if (chars_match(s1[i], s2[j])) { cc = '|'; ++ident; }
                    else if (BLOSUM62[MAPPER[s1[i]]][MAPPER[s2[j]]] > 0) { cc = '+'; ++similar; }
                    else cc = ' ';
                    --i; --j;
                } else {
                    where = (t & T_INS) ? T_INS : T_DEL;
                    continue;                     /* no column emitted */
                }
            } else if (where == T_INS) {
                qc = '-'; rc = (char)s2[j]; cc = ' ';
                --j;
                where = (t & T_DIAG_E) ? T_DIAG : T_INS;
            } else {
                qc = (char)s1[i]; rc = '-'; cc = ' ';
                --i;
                where = (t & T_DIAG_F) ? T_DIAG : T_DEL;
            }
        }
        if (q) { q[len] = qc; c[len] = cc; r[len] = rc; }
        ++len;
    }
    if (q) {
        for (int x = 0, y = len - 1; x < y; ++x, --y) {
            char t;
            t = q[x]; q[x] = q[y]; q[y] = t;
            t = c[x]; c[x] = c[y]; c[y] = t;
            t = r[x]; r[x] = r[y]; r[y] = t;
        }
        q[len] = c[len] = r[len] = 0;
    }
    out[0] = ident; out[1] = similar; out[2] = len; out[3] = score;
    free(trace);
    return 0;
}

/*
 * Forward-counter formulation (Appendix A, last paragraph): carry
 * (ident, similar, len) through the H/E/F states with the SAME predicates that
 * write the trace bits; O(n) memory, no trace table.  This is the formulation
 * the CUDA kernel implements; tests assert it equals nw_trace_stats.
 */
typedef struct { int id, sm, ln; } ctr_t;

int nw_forward_stats(const char *a, int m, const char *b, int n, int open,
                     int ext, int *out)
{
    init_mapper();
    const unsigned char *s1 = (const unsigned char *)a;
    const unsigned char *s2 = (const unsigned char *)b;
    int *H = (int *)malloc(sizeof(int) * (size_t)(n + 1));
    int *F = (int *)malloc(sizeof(int) * (size_t)(n + 1));
    ctr_t *cH = (ctr_t *)malloc(sizeof(ctr_t) * (size_t)(n + 1));
    ctr_t *cF = (ctr_t *)malloc(sizeof(ctr_t) * (size_t)(n + 1));
    H[0] = 0; F[0] = NEG_INF;
    cH[0] = (ctr_t){0, 0, 0}; cF[0] = (ctr_t){0, 0, 0};
    for (int j = 1; j <= n; ++j) {
        H[j] = -open - (j - 1) * ext; F[j] = NEG_INF;
        cH[j] = (ctr_t){0, 0, j}; cF[j] = (ctr_t){0, 0, 0};
    }
    for (int i = 1; i <= m; ++i) {
        const int8_t *row = BLOSUM62[MAPPER[s1[i - 1]]];
        int NH = H[0]; ctr_t cNH = cH[0];
        int WH = -open - (i - 1) * ext; ctr_t cWH = {0, 0, i};
        int E = NEG_INF; ctr_t cE = {0, 0, 0};
        H[0] = WH; cH[0] = cWH;
        for (int j = 1; j <= n; ++j) {
            int NWH = NH; ctr_t cNWH = cNH;
            NH = H[j]; cNH = cH[j];
            int F_opn = NH - open, F_ext = F[j] - ext;
            int E_opn = WH - open, E_ext = E - ext;
#if NW_GAP_TIE_EXTENDS
            int fo = F_opn > F_ext, eo = E_opn > E_ext;
#else
            int fo = F_opn >= F_ext, eo = E_opn >= E_ext;
#endif
            ctr_t f = fo ? cNH : cF[j]; f.ln += 1;
            ctr_t e = eo ? cWH : cE;    e.ln += 1;
            F[j] = imax(F_opn, F_ext); cF[j] = f;
            E = imax(E_opn, E_ext); cE = e;
            int sub = row[MAPPER[s2[j - 1]]];
            int H_dag = NWH + sub;
            WH = imax(imax(H_dag, F[j]), E);
            if (WH == H_dag) {
                cWH = cNWH; cWH.ln += 1;
                if (chars_match(s1[i - 1], s2[j - 1])) cWH.id += 1;
                else if (sub > 0) cWH.sm += 1;
            }
#if NW_H_PRIORITY_DIAG_DEL_INS
            else if (WH == F[j]) cWH = f;
            else cWH = e;
#else
            else if (WH == E) cWH = e;
            else cWH = f;
#endif
            H[j] = WH; cH[j] = cWH;
        }
    }
    if (m == 0) { /* H row 0 already holds the answer */ }
    out[0] = cH[n].id; out[1] = cH[n].sm; out[2] = cH[n].ln; out[3] = H[n];
    free(H); free(F); free(cH); free(cF);
    return 0;
}

/*
 * Independent textbook Gotoh global alignment score (three full matrices
 * M / X / Y, different recurrence shape from fill_trace) -- only used to pin
 * score optimality of the restatement above.
 */
int nw_gotoh_score(const char *a, int m, const char *b, int n, int open, int ext)
{
    init_mapper();
    const unsigned char *s1 = (const unsigned char *)a;
    const unsigned char *s2 = (const unsigned char *)b;
    size_t W = (size_t)n + 1;
    int *M = (int *)malloc(sizeof(int) * (size_t)(m + 1) * W);
    int *X = (int *)malloc(sizeof(int) * (size_t)(m + 1) * W);
    int *Y = (int *)malloc(sizeof(int) * (size_t)(m + 1) * W);
    /* M: ends in a residue pair; X: ends with gap in b (consumes a);
       Y: ends with gap in a (consumes b) */
    M[0] = 0; X[0] = NEG_INF; Y[0] = NEG_INF;
    for (int j = 1; j <= n; ++j) {
        M[j] = NEG_INF; X[j] = NEG_INF; Y[j] = -open - (j - 1) * ext;
    }
    for (int i = 1; i <= m; ++i) {
        M[i * W] = NEG_INF; Y[i * W] = NEG_INF; X[i * W] = -open - (i - 1) * ext;
        for (int j = 1; j <= n; ++j) {
            size_t c = i * W + j, u = (i - 1) * W + j, l = i * W + j - 1,
                   d = (i - 1) * W + j - 1;
            int best_d = imax(M[d], imax(X[d], Y[d]));
            M[c] = best_d + BLOSUM62[MAPPER[s1[i - 1]]][MAPPER[s2[j - 1]]];
            X[c] = imax(imax(M[u], Y[u]) - open, X[u] - ext);
            Y[c] = imax(imax(M[l], X[l]) - open, Y[l] - ext);
        }
    }
    size_t e = (size_t)m * W + n;
    int s = imax(M[e], imax(X[e], Y[e]));
    free(M); free(X); free(Y);
    return s;
}

/*
 * Batch entry point used by the parity tests and by bench.py's cpu_baseline:
 * alignments t = 0..n_tasks-1 of gene q_idx[t] (query, s1) against gene
 * r_idx[t] (ref, s2); genes live in one arena: residues[seq_off[g]..seq_off[g+1]).
 * out is n_tasks x 4 ints (ident, similar, length, score).
 * mode 0 = explicit trace table + traceback, 1 = forward counters.
 */
int nw_batch(const char *residues, const int64_t *seq_off, const int32_t *q_idx,
             const int32_t *r_idx, int64_t n_tasks, int open, int ext, int mode,
             int32_t *out)
{
    for (int64_t t = 0; t < n_tasks; ++t) {
        int q = q_idx[t], r = r_idx[t];
        const char *a = residues + seq_off[q];
        const char *b = residues + seq_off[r];
        int m = (int)(seq_off[q + 1] - seq_off[q]);
        int n = (int)(seq_off[r + 1] - seq_off[r]);
        int o[4];
        int rc = mode ? nw_forward_stats(a, m, b, n, open, ext, o)
                      : nw_trace_stats(a, m, b, n, open, ext, o, NULL, NULL, NULL);
        if (rc) return rc;
        out[4 * t + 0] = o[0]; out[4 * t + 1] = o[1];
        out[4 * t + 2] = o[2]; out[4 * t + 3] = o[3];
    }
    return 0;
}

/*
 * How much of (ident, similar, length) depends on the UNPINNED tie rules (U2-U4)?
 * Over ALL co-optimal alignments of the affine-gap score model (U1: any chain of
 * tying candidates through the H/E/F state graph, whatever a traceback would
 * prefer) compute the min and max of each count.  min == max for ident and
 * length means the pair's contribution to aai/peq is the same under ANY tie
 * rule, i.e. pinned by the score model alone (tests/test_oracle_nw.py checks
 * that the restated rules always land inside the range).
 * out: id_min, id_max, sm_min, sm_max, ln_min, ln_max, score.
 */
typedef struct { int lo[3], hi[3]; } rng_t;      /* 0 = ident, 1 = similar, 2 = length */

static inline void rng_join(rng_t *acc, const rng_t *x, int did, int dsm, int dln, int *have)
{
    const int d[3] = {did, dsm, dln};
    for (int k = 0; k < 3; ++k) {
        const int lo = x->lo[k] + d[k], hi = x->hi[k] + d[k];
        if (!*have || lo < acc->lo[k]) acc->lo[k] = lo;
        if (!*have || hi > acc->hi[k]) acc->hi[k] = hi;
    }
    *have = 1;
}

int nw_tie_range(const char *a, int m, const char *b, int n, int open, int ext, int *out)
{
    init_mapper();
    const unsigned char *s1 = (const unsigned char *)a;
    const unsigned char *s2 = (const unsigned char *)b;
    int *H = (int *)malloc(sizeof(int) * (size_t)(n + 1));
    int *F = (int *)malloc(sizeof(int) * (size_t)(n + 1));
    rng_t *rH = (rng_t *)calloc((size_t)(n + 1), sizeof(rng_t));
    rng_t *rF = (rng_t *)calloc((size_t)(n + 1), sizeof(rng_t));
    if (!H || !F || !rH || !rF) return -1;
    H[0] = 0; F[0] = NEG_INF;
    for (int j = 1; j <= n; ++j) {
        H[j] = -open - (j - 1) * ext; F[j] = NEG_INF;
        rH[j].lo[2] = rH[j].hi[2] = j;
    }
    for (int i = 1; i <= m; ++i) {
        const int8_t *row = BLOSUM62[MAPPER[s1[i - 1]]];
        int NH = H[0]; rng_t rNH = rH[0];
        int WH = -open - (i - 1) * ext;
        rng_t rWH; memset(&rWH, 0, sizeof rWH); rWH.lo[2] = rWH.hi[2] = i;
        int E = NEG_INF; rng_t rE; memset(&rE, 0, sizeof rE);
        H[0] = WH; rH[0] = rWH;
        for (int j = 1; j <= n; ++j) {
            int NWH = NH; rng_t rNWH = rNH;
            NH = H[j]; rNH = rH[j];
            const int F_opn = NH - open, F_ext = F[j] - ext;
            const int E_opn = WH - open, E_ext = E - ext;
            const int f = imax(F_opn, F_ext), e = imax(E_opn, E_ext);
            rng_t nf, ne, nh;
            int have = 0;
            if (F_opn == f) rng_join(&nf, &rNH, 0, 0, 1, &have);
            if (F_ext == f) rng_join(&nf, &rF[j], 0, 0, 1, &have);
            have = 0;
            if (E_opn == e) rng_join(&ne, &rWH, 0, 0, 1, &have);
            if (E_ext == e) rng_join(&ne, &rE, 0, 0, 1, &have);
            const int sub = row[MAPPER[s2[j - 1]]];
            const int H_dag = NWH + sub;
            const int h = imax(imax(H_dag, f), e);
            have = 0;
            if (H_dag == h) {
                const int hit = chars_match(s1[i - 1], s2[j - 1]);
                rng_join(&nh, &rNWH, hit, (!hit && sub > 0) ? 1 : 0, 1, &have);
            }
            if (f == h) rng_join(&nh, &nf, 0, 0, 0, &have);
            if (e == h) rng_join(&nh, &ne, 0, 0, 0, &have);
            F[j] = f; rF[j] = nf;
            E = e; rE = ne;
            WH = h; rWH = nh;
            H[j] = h; rH[j] = nh;
        }
    }
    out[0] = rH[n].lo[0]; out[1] = rH[n].hi[0];
    out[2] = rH[n].lo[1]; out[3] = rH[n].hi[1];
    out[4] = rH[n].lo[2]; out[5] = rH[n].hi[2];
    out[6] = H[n];
    free(H); free(F); free(rH); free(rF);
    return 0;
}

/* batch form of nw_tie_range over a residue arena; out is n_tasks x 7 ints */
int nw_tie_batch(const char *residues, const int64_t *seq_off, const int32_t *q_idx,
                 const int32_t *r_idx, int64_t n_tasks, int open, int ext, int32_t *out)
{
    for (int64_t t = 0; t < n_tasks; ++t) {
        int q = q_idx[t], r = r_idx[t];
        int o[7];
        int rc = nw_tie_range(residues + seq_off[q], (int)(seq_off[q + 1] - seq_off[q]),
                              residues + seq_off[r], (int)(seq_off[r + 1] - seq_off[r]), open, ext, o);
        if (rc) return rc;
        for (int k = 0; k < 7; ++k) out[7 * t + k] = o[k];
    }
    return 0;
}
