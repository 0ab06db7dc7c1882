/*
 * oracle/lexc_model.c -- TEST INFRASTRUCTURE: executable CPU model of the LEXC DP kernel
 * (nw_lex_kernel<..., CTR = true> in phamclust_b200/csrc/pc_align.cu), the rebased LEX word whose
 * COUNTERS are re-centred per lane as well, so that alignments of any length run on the
 * max-plus recurrence.  The DP is parasail's nw_trace as SURVEY.md Appendix A states it
 * (reference call site /root/reference/src/phamclust/metrics.py:160-175, 216-220).
 *
 * The model executes the kernel's arithmetic word for word -- 32-bit wrap-around, lane
 * schedule, frame conversions between lanes, column chunks with the spilled boundary column,
 * the band checks that raise the fallback flag -- so that the formulation can be checked
 * against nw_forward_stats (oracle/nw_oracle.c) on thousands of long alignments without a GPU,
 * and so that the band parameters can be chosen from measured deviations.
 *
 * word  = score << (2B+2) | prio << 2B | (diag steps - base_d) << B | (matches - base_m)
 * frame = (score base, base_d, base_m) per lane; rbU is the packed frame (mod 2^32), rb the
 *         score base, rbDM = (base_d + c0) << 16 | (base_m + c0): the centre column's true
 *         counters at the last re-centring.
 * Every R row-steps a lane subtracts (centre word - C0) from all of its state, C0 = c0 << B | c0,
 * c0 = 2^(B-1).  Soundness of the counter fields (no borrow or carry between fields) is by
 * induction while no check has failed:
 *   - after a re-centring every state word's counter fields are checked to lie in
 *     [c0 - band, c0 + band); a row-step adds at most one to a counter, so until the next
 *     re-centring they stay in [c0 - band, c0 + band + R];
 *   - the centre word then shifts the frame by at most band + R: new fields lie within
 *     2 band + R of c0, inside the field when 2 band + R < c0;
 *   - a word crossing to the next lane is converted by the difference of the two frames; the
 *     counter bases of neighbouring lanes are checked to differ by at most 128, so the
 *     converted fields stay inside the field when band + R + 128 < c0; the converted word
 *     is then checked against the band like the lane's own state.
 * A failed check only raises the flag: the caller re-runs that alignment on the classic kernel.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <ctype.h>

int nw_blosum62(int row, int col);
int nw_map_char(int c);

#define OPEN 11
#define EXT 1
#define KMAX 64
#define GLANES 32

typedef struct {
    int32_t Hp[KMAX], Fp[KMAX];
    int32_t diagH, outH, outE;
    int32_t rb;         /* score base */
    uint32_t rbU;       /* packed frame */
    int32_t rbDM;       /* centre's true counters at the last re-centring, 16 | 16 */
    uint32_t chk;       /* sticky: non-zero = a band / frame check failed */
    int rc[KMAX];       /* ref residue index per column */
    unsigned char rch[KMAX];
} lane_t;

static inline int32_t add32(int32_t a, int32_t b) { return (int32_t)((uint32_t)a + (uint32_t)b); }
static inline int32_t imax2(int32_t a, int32_t b) { return a > b ? a : b; }

/*
 * out[0] matched, out[1] columns, out[2] score, out[3] flagged, out[4] largest |field - c0| any
 * checked word had (diagnostic), out[5] largest |frame difference| between neighbours, out[6]
 * 1 if an UNCHECKED invariant broke (a field actually left [0, 2^B) -- must never happen while
 * the flag is clear), out[7] largest |score field|, out[8] largest spread (max - min) of a counter
 * field over one lane's state at a re-centring: what a band centred on the lane's mid-range instead
 * of its centre column would have to hold (diagnostic for DESIGN.md "what comes next").
 */
int lexc_align(const char *q, int m, const char *r, int n, int K, int R, int B, int band,
               int ppos, int *out)
{
    const int G = GLANES;
    const int SH = 2 * B + 2;
    const int32_t U = (int32_t)1 << SH;
    const int32_t CMASK = ((int32_t)1 << B) - 1;
    const int32_t PRIO_E = (int32_t)1 << (2 * B), PRIO_F = (int32_t)2 << (2 * B);
    const int32_t PRIO_D = (int32_t)3 << (2 * B), PMASK = (int32_t)3 << (2 * B);
    const int32_t c0 = (int32_t)1 << (B - 1);
    const int32_t C0 = (c0 << B) | c0;
    const int32_t lo = c0 - band;
    const int32_t LO = (lo << B) | lo;
    /* band = 2^k: a field f is inside [lo, lo + 2 band) iff (f - lo) has no bit >= k + 1 */
    int kb = 0;
    while ((1 << kb) < 2 * band) ++kb;
    if ((1 << kb) != 2 * band || K > KMAX || K < 2 || (R & (R - 1)) || 2 * band + R >= c0 ||
        band + R + (B >= 10 ? 256 : 128) >= c0)
        return -1;
    const int32_t FTOP = CMASK & ~(((int32_t)1 << kb) - 1);
    const int32_t TOPMASK = (FTOP << B) | FTOP;
    const int32_t CU = (EXT - OPEN) * U;
    const int32_t EDGE0 = (EXT - OPEN) * U;
    const int32_t NEG0 = EDGE0 + CU - 4 * U;
    const int CAP = G * K;
    memset(out, 0, 9 * sizeof(int));
    if (n == 0) {
        out[0] = 0; out[1] = m; out[2] = m == 0 ? 0 : -OPEN - (m - 1) * EXT;
        return 0;
    }
    const int nchunks = (n + CAP - 1) / CAP;
    typedef struct { int32_t h, e; uint32_t u; int32_t dm; } spill_t;
    spill_t *spill = (spill_t *)calloc((size_t)m + 2, sizeof(spill_t));
    lane_t *L = (lane_t *)calloc(G, sizeof(lane_t));
    int dev_max = 0, delta_max = 0, broke = 0, score_max = 0, spread_max = 0;
    int res_word = 0, res_rb = 0, res_dm = 0;
    uint32_t flag = 0;

#define TRACK(w_) do { \
        const int32_t w__ = (w_); \
        const int fd__ = (w__ >> B) & CMASK, fm__ = w__ & CMASK; \
        if (abs(fd__ - c0) > dev_max) dev_max = abs(fd__ - c0); \
        if (abs(fm__ - c0) > dev_max) dev_max = abs(fm__ - c0); \
        if (abs(w__ >> SH) > score_max) score_max = abs(w__ >> SH); \
    } while (0)
#define CHECK(l_, w_) do { (l_)->chk |= (uint32_t)(add32((w_), -LO) & TOPMASK); TRACK(w_); } while (0)

    for (int chunk = 0; chunk < nchunks; ++chunk) {
        for (int l = 0; l < G; ++l) {
            lane_t *p = &L[l];
            const int col0 = chunk * CAP + l * K;
            for (int x = 0; x < K; ++x) {
                const int j = col0 + x;
                p->rc[x] = j < n ? nw_map_char((unsigned char)r[j]) : 0;
                p->rch[x] = j < n ? (unsigned char)toupper((unsigned char)r[j]) : 'A';
                p->Hp[x] = EDGE0 + C0;
                p->Fp[x] = NEG0 + C0;
            }
            p->diagH = (col0 == 0 ? 0 : EDGE0) + C0;
            p->outH = C0;
            p->outE = NEG0 + C0;
            p->rb = 0;
            p->rbU = 0u - (uint32_t)C0;
            p->rbDM = 0;
            p->chk = 0;
        }
        const int steps = m + G - 1;
        for (int s = 1; s <= steps; ++s) {
            if ((s & (R - 1)) == 0) {
                for (int l = 0; l < G; ++l) {
                    lane_t *p = &L[l];
                    const int32_t wc = p->Hp[K / 2];
                    const int32_t d = add32(wc, -C0);
                    p->rb += wc >> SH;
                    p->rbU += (uint32_t)d;
                    p->rbDM += (((wc >> B) & CMASK) - c0) * 65536 + ((wc & CMASK) - c0);
                    for (int x = 0; x < K; ++x) {
                        /* would the subtraction borrow across a field?  (diagnostic only) */
                        const int fd = ((p->Hp[x] >> B) & CMASK) - (((wc >> B) & CMASK) - c0);
                        const int fm = (p->Hp[x] & CMASK) - ((wc & CMASK) - c0);
                        if (!p->chk && (fd < 0 || fd > CMASK || fm < 0 || fm > CMASK)) broke = 1;
                        p->Hp[x] = add32(p->Hp[x], -d);
                        p->Fp[x] = add32(p->Fp[x], -d);
                        CHECK(p, p->Hp[x]);
                        CHECK(p, p->Fp[x]);
                    }
                    if (!p->chk && chunk * CAP + l * K < n) {
                        int lo_d = CMASK, hi_d = 0, lo_m = CMASK, hi_m = 0;
                        for (int x = 0; x < K; ++x) {
                            const int32_t w2[2] = {p->Hp[x], p->Fp[x]};
                            for (int z = 0; z < 2; ++z) {
                                const int fd = (w2[z] >> B) & CMASK, fm = w2[z] & CMASK;
                                if (fd < lo_d) lo_d = fd;
                                if (fd > hi_d) hi_d = fd;
                                if (fm < lo_m) lo_m = fm;
                                if (fm > hi_m) hi_m = fm;
                            }
                        }
                        if (hi_d - lo_d > spread_max) spread_max = hi_d - lo_d;
                        if (hi_m - lo_m > spread_max) spread_max = hi_m - lo_m;
                    }
                    p->diagH = add32(p->diagH, -d);
                    p->outH = add32(p->outH, -d);
                    p->outE = add32(p->outE, -d);
                    CHECK(p, p->diagH);
                    CHECK(p, p->outH);
                    CHECK(p, p->outE);
                }
            }
            int32_t shH[GLANES], shE[GLANES], shDM[GLANES];
            uint32_t shU[GLANES];
            for (int l = 0; l < G; ++l) {
                shH[l] = L[l].outH; shE[l] = L[l].outE; shU[l] = L[l].rbU; shDM[l] = L[l].rbDM;
            }
            for (int l = 0; l < G; ++l) {
                lane_t *p = &L[l];
                const int i = s - l;
                const int active = i >= 1 && i <= m;
                int32_t inH, inE, inDM;
                uint32_t inB;
                if (l > 0) {
                    inH = shH[l - 1]; inE = shE[l - 1]; inB = shU[l - 1]; inDM = shDM[l - 1];
                } else if (chunk == 0 || !active) {
                    inH = EDGE0 + C0; inE = NEG0 + C0; inB = 0u - (uint32_t)C0; inDM = 0;
                } else {
                    inH = spill[i].h; inE = spill[i].e; inB = spill[i].u; inDM = spill[i].dm;
                }
                const int32_t conv = (int32_t)(inB - p->rbU);
                if (active) {
                    /* neighbouring frames: counter bases within [-127, 128] of each other */
                    const int32_t t = add32(add32(inDM, -p->rbDM), B >= 10 ? (255 << 16) | 255 : (127 << 16) | 127);
                    p->chk |= (uint32_t)t & (B >= 10 ? 0xFE00FE00u : 0xFF00FF00u);
                    const int dd = abs((inDM >> 16) - (p->rbDM >> 16)), dmm = abs((inDM & 0xffff) - (p->rbDM & 0xffff));
                    if (dd > delta_max) delta_max = dd;
                    if (dmm > delta_max) delta_max = dmm;
                    if (!p->chk) {
                        const int fd = ((inH >> B) & CMASK) + ((inDM >> 16) - (p->rbDM >> 16));
                        const int fm = (inH & CMASK) + ((inDM & 0xffff) - (p->rbDM & 0xffff));
                        if (fd < 0 || fd > CMASK || fm < 0 || fm > CMASK) broke = 1;
                    }
                }
                inH = add32(inH, conv);
                inE = add32(inE, conv);
                if (!active) continue;
                CHECK(p, inH);
                CHECK(p, inE);
                const int qi = nw_map_char((unsigned char)q[i - 1]);
                const unsigned char qch = (unsigned char)toupper((unsigned char)q[i - 1]);
                int32_t hl = inH, el = inE, hd = p->diagH;
                p->diagH = inH;
                for (int x = 0; x < K; ++x) {
                    const int sc = nw_blosum62(qi, p->rc[x]);
                    const int hit = (qch == p->rch[x]) || (ppos && sc > 0);
                    const int32_t tw = (sc + 2 * EXT) * U + PRIO_D + ((int32_t)1 << B) + hit;
                    const int32_t f = imax2(add32(p->Hp[x], CU), p->Fp[x]) | PRIO_F;
                    const int32_t e = imax2(add32(hl, CU), el) | PRIO_E;
                    const int32_t hdg = add32(hd, tw);
                    const int32_t h = imax2(imax2(hdg, f), e) & ~PMASK;
                    hd = p->Hp[x];
                    hl = h; el = e;
                    p->Hp[x] = h; p->Fp[x] = f;
                }
                p->outH = hl; p->outE = el;
                if (l == G - 1 && chunk + 1 < nchunks) {
                    spill[i].h = hl; spill[i].e = el; spill[i].u = p->rbU; spill[i].dm = p->rbDM;
                }
            }
        }
        for (int l = 0; l < G; ++l)
            if (chunk * CAP + l * K < n) flag |= L[l].chk;      /* lanes of padding only do not count */
        if (chunk == nchunks - 1) {
            const int last = n - 1 - chunk * CAP;
            const lane_t *p = &L[last / K];
            res_word = p->Hp[last % K];
            res_rb = p->rb;
            res_dm = p->rbDM;
        }
    }
    out[0] = (res_word & CMASK) + (res_dm & 0xffff) - c0;
    out[1] = m + n - (((res_word >> B) & CMASK) + (res_dm >> 16) - c0);
    out[2] = (res_word >> SH) + res_rb - (m + n) * EXT;
    out[3] = flag != 0;
    out[4] = dev_max;
    out[5] = delta_max;
    out[6] = broke;
    out[7] = score_max;
    out[8] = spread_max;
    free(spill);
    free(L);
    return 0;
}
