// pc_contract_i8p.cu -- K1/K2/K3 as a PERSISTENT CTA-PAIR int8 tcgen05 GEMM (the default
// tensor engine; pc_contract_i8.cu keeps the one-tile-per-CTA kernel of round 1 as engine
// PC_ENGINE_I8MMA_V1 and owns the operand packing).
//
// Replaces, per genome pair, Genome.__and__ + len()
// (/root/reference/src/phamclust/genome.py:124-135) and the arithmetic of
// metrics.py:40-53 (gcs), :70-80 (jc), :102-115 (pocp), :135-157 (af).
//
// Why a pair: a 256 x 256 output tile is what keeps the operand bytes per MAC low enough for L2
// (round 1: the 128 x 256 tile ran the tensor pipe at 55 %, L2-bound), but one CTA needs all 512
// TMEM columns for it, so its epilogue (fp64 normalise + Python rounding of 65 536 pairs) cannot
// overlap the next tile's MMAs -- the tensor pipe idled through every epilogue (73 %).  Two CTAs
// of one TPC (`cta_group::2`) split the tile's 256 rows: each SM holds a 128-row x 256-column
// accumulator = 256 TMEM columns, so TWO accumulators fit and the epilogue of tile t runs under
// the main loop of tile t+1.  Each CTA loads its own 128 A rows and HALF of the 256 B rows per K
// block (32 KB per stage and SM instead of 64 KB), the leader's MMA thread issues
// `tcgen05.mma.cta_group::2.kind::i8` with M = 256 across the pair.
//
// Layout of one CTA (384 threads): warp 0 = TMA producer, warp 1 = MMA issuer (leader CTA only),
// warp 2 = TMEM allocator, warp 3 idle, warps 4..11 = epilogue (TMEM lane quadrant = warp % 4,
// two warps per quadrant split the 8 column chunks).
// Pipelines: smem ring of 5 stages (full: the leader's producer arrives and expects the 64 KB of
// BOTH CTAs' TMA loads on the LEADER's barrier; empty: tcgen05.commit multicast to both CTAs), accumulator ring of 2
// (tmem_full: commit multicast to both CTAs; tmem_empty: every epilogue warp of BOTH CTAs
// arrives on the leader's barrier).
// Scheduling: a tile queue built on the device per launch -- upper-triangle tiles of the launch's
// row range with their K-block occupancy mask (block-sparse K schedule of round 1: only the K
// blocks both genome groups occupy are streamed; a tile with none never touches the tensor core)
// ordered in locality blocks of the tile grid (late rows first, longest first inside a block);
// round r of the static schedule hands entries [r C, (r + 1) C) to the C clusters in snake order.
// Epilogue stores go through a 32 x 33 smem transpose per warp so that every store instruction
// writes 32 consecutive int32 of ONE row of the packed triangle (round 1 wrote 32 rows x 4 B).
#include <cub/cub.cuh>
#include <cuda.h>

#include "pc_common.cuh"
#include "pc_tcgen05.cuh"

#define P2_BM 128                 // rows per CTA; 256 per pair
#define P2_BN 256                 // columns per tile
#define P2_BK 128                 // bytes (= int8 elements) of K per stage
#define P2_STAGES 5
#define P2_A_BYTES (P2_BM * P2_BK)
#define P2_B_BYTES ((P2_BN / 2) * P2_BK)
#define P2_STAGE_BYTES (P2_A_BYTES + P2_B_BYTES)            // 32 KB per CTA and stage
#define P2_EPI_WARPS 8
#define P2_FIRST_EPI_WARP 4
#define P2_THREADS (32 * (P2_FIRST_EPI_WARP + P2_EPI_WARPS))
#define P2_ORDER_ROWS 8           // tile-queue locality blocks: 8 row tiles x 9 column tiles ~ one round of 74 pairs
#define P2_ORDER_COLS 9
#define P2_BAR_BYTES 256
#define P2_DEN_BYTES (P2_BN * 4)
#define P2_XPOSE_BYTES (P2_EPI_WARPS * 32 * 33 * 4)
#define P2_SMEM_BYTES (P2_STAGES * P2_STAGE_BYTES + P2_BAR_BYTES + 2 * P2_DEN_BYTES + P2_XPOSE_BYTES + 1024)

struct P2Tile {
    int32_t bi, bj;        // 256-row tile, 256-column tile
    int32_t n_steps;       // K blocks to stream over all passes (0: the tensor core is not used)
    int32_t mask_slot;     // occupancy mask at masks + (int64)mask_slot * occ_words
};

struct P2Args {
    int64_t p_pad;         // bytes per operand segment (multiple of 128)
    int n_limbs;           // 0: gcs/jc (single segment); else weight limbs before the B segment
    int occ_words, occ_shift;
    const P2Tile *tiles;   // unsorted entries
    const int32_t *order;  // queue position -> entry (longest first)
    const uint32_t *masks;
    int n_tiles;
    unsigned long long *tile_steps;
    // rounds in lock step (L2 reuse): sync[0] = sum of n_steps over the queue, sync[1] = the longest
    // tile (both written by the build kernel), sync[2 + r] = producers that have reached round r
    unsigned int *sync;
    int lock_mode;         // 0: lock step when the launch is tensor-heavy, 1: always, 2: never
    EpilogueArgs ep;
};

// instruction descriptor: D = s32, A = B = unsigned 8 bit, K-major, N = 256, M = 256 (pair)
__device__ __forceinline__ constexpr uint32_t p2_idesc()
{
    return (2u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(P2_BN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
}

// ---------------------------------------------------------------------------
// tile queue: one entry + occupancy mask per upper-triangle tile of the launch
// ---------------------------------------------------------------------------
__global__ void p2_build_tiles_kernel(const uint32_t *__restrict__ occ, int occ_words, int occ_shift,
                                      int64_t kb_per_seg, int n_pass, int pairs, int n_groups, int tile_row0,
                                      int n_row_tiles, int n_col_tiles, P2Tile *__restrict__ tiles,
                                      uint32_t *__restrict__ masks,
                                      uint32_t *__restrict__ keys, int32_t *__restrict__ ids, int *__restrict__ counter,
                                      unsigned int *__restrict__ total_steps)
{
    // one warp per candidate tile (bi, bj) of 256 * pairs rows x 256 columns; tiles strictly below
    // the diagonal do not exist
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (warp >= (int64_t)n_row_tiles * n_col_tiles) return;
    const int bi = tile_row0 + (int)(warp / n_col_tiles), bj = (int)(warp % n_col_tiles);
    if (bj < bi * pairs) return;
    int slot = 0;
    if (lane == 0) slot = atomicAdd(counter, 1);
    slot = __shfl_sync(0xffffffffu, slot, 0);
    long long steps = 0;
    for (int w = lane; w < occ_words; w += 32) {
        uint32_t rows = 0;
        for (int u = 0; u < 2 * pairs; ++u) {
            const int g = 2 * pairs * bi + u;                 // 128-genome group; past the table: empty
            if (g < n_groups) rows |= occ[(int64_t)g * occ_words + w];
        }
        const uint32_t cols = occ[(int64_t)(2 * bj) * occ_words + w] | occ[(int64_t)(2 * bj + 1) * occ_words + w];
        uint32_t bits = rows & cols;
        masks[(int64_t)slot * occ_words + w] = bits;
        while (bits) {
            const int64_t b = (int64_t)w * 32 + (__ffs((int)bits) - 1);
            bits &= bits - 1;
            const int64_t lo = b << occ_shift, hi = min((b + 1) << occ_shift, kb_per_seg);
            if (hi > lo) steps += hi - lo;
        }
    }
    for (int o = 16; o; o >>= 1) steps += __shfl_xor_sync(0xffffffffu, steps, o);
    if (lane == 0) {
        P2Tile t;
        t.bi = bi; t.bj = bj; t.n_steps = (int32_t)(steps * n_pass); t.mask_slot = slot;
        tiles[slot] = t;
        ids[slot] = slot;
        if (t.n_steps) atomicMax(total_steps + 1, (unsigned int)t.n_steps);    // longest tile of the launch
        if (t.n_steps) atomicAdd(total_steps, (unsigned int)min(t.n_steps, 1 << 16));
    }
}

// Queue order (descending key).  One round of the static schedule is a window of #clusters
// consecutive entries that run in lock step, so a round lasts as long as its LONGEST tile, and its
// tiles share operand strips in L2 only if they sit close together in the tile grid (a 256-genome
// strip is 15 MB at C3, 51 MB at C5).  Primary key: work bucket (visited K blocks in ~16-32 coarse
// classes) -- rounds of equal tiles; inside a bucket: locality blocks of 8 row tiles x 9 column
// tiles, late rows first; inside a block: longest first.  Simulated on the C3 / C5 occupancy
// (round-max efficiency, distinct strips per round): pure locality order 0.85 / 22 and 0.94 / 24,
// pure work order 0.94 / 29 and 0.99 / 64 (DRAM-bound at C5), this order 0.94 / 30 and 0.98 / 36.
__global__ void p2_keys_kernel(const P2Tile *__restrict__ tiles, int n_tiles, int pairs,
                               const unsigned int *__restrict__ stats, uint32_t *__restrict__ keys)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_tiles) return;
    const P2Tile t = tiles[k];
    int shift = 0;
    while ((stats[1] >> shift) > 31u) ++shift;                 // <= 32 buckets over the launch's range
    const uint32_t bucket = (uint32_t)t.n_steps >> shift;
    const uint32_t blk = (uint32_t)(t.bi * pairs / P2_ORDER_ROWS) * 4096u + (uint32_t)(4095 - min(t.bj / P2_ORDER_COLS, 4095));
    const uint32_t fine = shift >= 8 ? ((uint32_t)t.n_steps >> (shift - 8)) & 255u
                                     : ((uint32_t)t.n_steps << (8 - shift)) & 255u;
    keys[k] = (bucket << 26) | ((blk & 0x3ffffu) << 8) | fine;
}

// K schedule of one tile, identical for producer and MMA issuer: n_limbs == 0 -> one pass over
// the visited K blocks; else per weight limb two passes (row tile reads the limb while the column
// tile reads B, then swapped), limb t accumulating into accumulator t of the tile.
// static schedule: round r hands queue entries [r C, (r + 1) C) to the C clusters, alternating
// direction (snake) so that no cluster always gets the longest tile of a round
__device__ __forceinline__ int p2_queue_slot(int round, int cluster, int n_clusters)
{
    return round * n_clusters + ((round & 1) ? n_clusters - 1 - cluster : cluster);
}

template <typename F>
__device__ __forceinline__ void p2_for_each_block(const P2Args &a, const uint32_t *__restrict__ mask, F &&body)
{
    const int64_t kb_per_seg = a.p_pad / P2_BK;
    const int n_pass = a.n_limbs == 0 ? 1 : 2 * a.n_limbs;
    for (int pass = 0; pass < n_pass; ++pass) {
        const int limb = pass >> 1;
        const bool swapped = (pass & 1) != 0;
        uint32_t next = a.occ_words ? __ldg(mask) : 0u;
        for (int w = 0; w < a.occ_words; ++w) {
            uint32_t bits = next;
            if (w + 1 < a.occ_words) next = __ldg(mask + w + 1);      // the next word is in flight
            while (bits) {
                const int64_t b = (int64_t)w * 32 + (__ffs((int)bits) - 1);
                bits &= bits - 1;
                const int64_t kb_end = min((b + 1) << a.occ_shift, kb_per_seg);
                for (int64_t kb = b << a.occ_shift; kb < kb_end; ++kb) {
                    int64_t a_col, b_col;
                    if (a.n_limbs == 0) {
                        a_col = b_col = kb * P2_BK;
                    } else {
                        const int64_t wcol = (int64_t)limb * a.p_pad + kb * P2_BK;
                        const int64_t bcol = (int64_t)a.n_limbs * a.p_pad + kb * P2_BK;
                        a_col = swapped ? bcol : wcol;
                        b_col = swapped ? wcol : bcol;
                    }
                    body(a_col, b_col);
                }
            }
        }
    }
}

// PAIRS = 2: a cluster of FOUR CTAs = two CTA pairs working on two vertically adjacent tiles
// (512 rows x 256 columns).  Both pairs need the same B tile, so each CTA fetches only a QUARTER
// of it (64 rows) and the TMA multicasts that quarter into the smem of its counterpart in the
// other pair: 24 KB per CTA and K block cross the L2 -> SM fabric instead of 32 KB.  ncu had the
// single-pair kernel at the fabric's cap (11.9 TB/s into the SMs, tensor pipe 68 %).  The two
// pairs stream the union of their K blocks in lock step: a stage is refilled only after BOTH
// pairs' MMAs have released it (empty barriers count 2 commits, multicast to all four CTAs).
template <int PAIRS>
__global__ void __launch_bounds__(P2_THREADS, 1)
i8_pair_kernel(const __grid_constant__ CUtensorMap map, const __grid_constant__ CUtensorMap map_quarter,
               const P2Args a)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t *tiles_smem = smem;                                            // STAGES x (A | half B)
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem + P2_STAGES * P2_STAGE_BYTES);
    uint64_t *empty_bar = full_bar + P2_STAGES;
    uint64_t *tmem_full_bar = empty_bar + P2_STAGES;                       // [2]
    uint64_t *tmem_empty_bar = tmem_full_bar + 2;                          // [2], used in the leader
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(tmem_empty_bar + 2);
    int32_t *s_den = reinterpret_cast<int32_t *>(smem + P2_STAGES * P2_STAGE_BYTES + P2_BAR_BYTES);   // [2][256]
    int32_t *s_xpose = s_den + 2 * P2_BN;                                  // [EPI_WARPS][32][33]

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t cta_rank = cluster_ctarank();
    const uint32_t pair = cta_rank >> 1, half_rank = cta_rank & 1u;       // which pair, which half of its tile
    const uint32_t leader_rank = pair * 2;                                 // the CTA that issues this pair's MMAs
    const bool leader = half_rank == 0;
    const int n_clusters = (int)cluster_count_x();
    const int cluster = (int)cluster_id_x();
    const int acc_per_tile = a.n_limbs <= 1 ? 1 : a.n_limbs;            // host guarantees <= 2
    const int acc_stages = 2 / acc_per_tile;                              // accumulator ring depth

    if (warp == 0 && lane == 0) asm volatile("prefetch.tensormap [%0];" ::"l"(&map) : "memory");
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < P2_STAGES; ++s) {
            mbar_init(full_bar + s, 1);                // the leader's producer arrives (+ 64 KB of tx from BOTH CTAs)
            mbar_init(empty_bar + s, PAIRS);           // one multicast commit per pair of the cluster
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(tmem_full_bar + s, 1);
            mbar_init(tmem_empty_bar + s, 2 * P2_EPI_WARPS);   // every epilogue warp of both CTAs
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)),
                     "r"(512u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cluster_sync_all();                                // barriers initialised and TMEM allocated in both CTAs
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        // ===== TMA producer (one lane per CTA): own 128 A rows + own half of the 256 B rows =====
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0;
            const uint32_t full0 = mapa_u32(smem_u32(full_bar), leader_rank);   // this pair's LEADER's full barriers
            // Rounds in lock step.  The tiles of one round form a compact block of the tile grid
            // and share operand strips, but a strip only stays in L2 for a window of K (a
            // 256-genome strip is 15 MB at C3, 51 MB at C5): the sharing is real only while the
            // pairs stream the same K range at the same time.  Left alone the pairs drift apart
            // by whole tiles within a few rounds (C5, 8 GPUs: DRAM-bound at 45 % tensor), so every
            // producer waits at the start of a round until all producers of the grid have issued
            // the last load of the previous one.  All CTAs are co-resident (one per SM), so the
            // spin cannot deadlock.  Launches that hardly use the tensor core (clustered tables:
            // most tiles are empty) skip it -- there a round lasts microseconds.
            const bool lock_step = a.lock_mode == 1 || (a.lock_mode == 0 && a.sync[0] >= (unsigned int)a.n_tiles * 64u);
            const unsigned int n_producers = 2u * PAIRS * (unsigned int)n_clusters;
            bool in_step = lock_step;
            for (int round = 0; round * n_clusters < a.n_tiles; ++round) {
                if (in_step) {
                    unsigned int *ctr = a.sync + 2 + round;
                    atomicAdd(ctr, 1u);
                    unsigned int seen, spins = 0;
                    do {
                        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(ctr) : "memory");
                        if (seen < n_producers) __nanosleep(200);
                        // lock step is an optimisation, never a dependency: if the grid is not
                        // co-resident after all (a foreign kernel holds SMs) stop waiting for it
                        if (++spins > (1u << 21)) { in_step = false; break; }
                    } while (seen < n_producers);
                }
                const int q = p2_queue_slot(round, cluster, n_clusters);
                if (q >= a.n_tiles) continue;
                const P2Tile t = a.tiles[a.order[q]];
                if (t.n_steps == 0) continue;
                const uint32_t *mask = a.masks + (int64_t)t.mask_slot * a.occ_words;
                const int32_t a_row = t.bi * (256 * PAIRS) + (int)pair * 256 + (int)half_rank * P2_BM;
                const int32_t b_row = t.bj * P2_BN + (int)half_rank * (P2_BN / 2);
                p2_for_each_block(a, mask, [&](int64_t a_col, int64_t b_col) {
                    mbar_wait(empty_bar + s, ph ^ 1);
                    const uint32_t bar = full0 + (uint32_t)s * 8u;
                    uint8_t *dst = tiles_smem + (size_t)s * P2_STAGE_BYTES;
                    // only the leader arrives, expecting the bytes of both CTAs; the peer's TMA just
                    // completes transactions on the leader's barrier (a remote release-arrive here
                    // costs a MEMBAR.ALL.GPU per K block: ncu had the peer's producer 75 % stalled
                    // in it).  The peer cannot run a phase ahead: its empty barrier is only
                    // signalled after the MMAs that consumed the previous phase.
                    if (leader)
                        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(
                                         smem_u32(full_bar + s)),
                                     "r"((uint32_t)(2 * P2_STAGE_BYTES))
                                     : "memory");
                    tma_load_2d_pair(dst, &map, bar, (int32_t)a_col, a_row);
                    if (PAIRS == 1) {
                        tma_load_2d_pair(dst + P2_A_BYTES, &map, bar, (int32_t)b_col, b_row);
                    } else {
                        // this CTA's quarter of the B tile, to itself and to the CTA of the other
                        // pair that holds the same half; each destination's 8 KB complete on the
                        // full barrier of ITS pair's leader
                        const uint16_t dests = (uint16_t)((1u << half_rank) | (1u << (2 + half_rank)));
                        tma_load_2d_pair_multicast(dst + P2_A_BYTES + pair * (P2_B_BYTES / 2), &map_quarter, bar,
                                                   (int32_t)b_col, b_row + (int)pair * (P2_BN / 4), dests);
                    }
                    if (++s == P2_STAGES) { s = 0; ph ^= 1; }
                });
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: one lane of the leader CTA drives both tensor cores =====
        if (leader && lane == 0) {
            const uint32_t idesc = p2_idesc();
            int s = 0;
            uint32_t ph = 0;
            int acc_stage = 0;
            uint32_t acc_ph = 0;
            unsigned long long issued = 0;
            for (int round = 0; round * n_clusters < a.n_tiles; ++round) {
                const int q = p2_queue_slot(round, cluster, n_clusters);
                if (q >= a.n_tiles) continue;
                const P2Tile t = a.tiles[a.order[q]];
                if (t.n_steps == 0) continue;
                const int n_pass = a.n_limbs == 0 ? 1 : 2 * a.n_limbs;
                const int per_pass = t.n_steps / n_pass;
                mbar_wait(tmem_empty_bar + acc_stage, acc_ph ^ 1);     // the epilogues drained this accumulator
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                for (int pass = 0; pass < n_pass; ++pass) {
                    const int acc = acc_stage * acc_per_tile + (a.n_limbs == 0 ? 0 : (pass >> 1));
                    const uint32_t tmem_d = tmem_base + (uint32_t)acc * P2_BN;
                    for (int k = 0; k < per_pass; ++k) {
                        mbar_wait(full_bar + s, ph);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        const uint32_t sa = smem_u32(tiles_smem + (size_t)s * P2_STAGE_BYTES);
                        const uint64_t adesc = umma_desc_sw128_kmajor(sa);
                        const uint64_t bdesc = umma_desc_sw128_kmajor(sa + P2_A_BYTES);
                        const bool first = k == 0 && (pass & 1) == 0;
#pragma unroll
                        for (int kk = 0; kk < P2_BK / 32; ++kk)
                            umma_i8_pair(tmem_d, adesc + (uint64_t)(2 * kk), bdesc + (uint64_t)(2 * kk), idesc,
                                         (first && kk == 0) ? 0u : 1u);
                        umma_commit_pair(empty_bar + s, PAIRS == 1 ? 3 : 15);   // frees the slot in every CTA of the cluster
                        if (++s == P2_STAGES) { s = 0; ph ^= 1; }
                    }
                }
                umma_commit_pair(tmem_full_bar + acc_stage, (uint16_t)(3u << leader_rank));   // accumulators complete, both CTAs of the pair
                issued += (unsigned long long)t.n_steps;
                if (++acc_stage == acc_stages) { acc_stage = 0; acc_ph ^= 1; }
            }
            if (issued) atomicAdd(a.tile_steps, issued * 2ull);       // 128 x 256 x 128 MMA groups
        }
    } else if (warp >= P2_FIRST_EPI_WARP) {
        // ===== epilogue: this CTA's 128 rows x 256 columns of every tile of the cluster =====
        const int ew = warp - P2_FIRST_EPI_WARP;
        const int quad = warp & 3;                         // TMEM lane quadrant this warp may touch
        const int half = ew >> 2;                          // which 4 of the 8 column chunks
        const EpilogueArgs &ep = a.ep;
        const int64_t n = ep.n;
        const int64_t out_base = pc_tri_index(n, ep.row_begin, ep.row_begin);
        const uint32_t lane_addr = tmem_base + ((uint32_t)(quad * 32) << 16);
        const uint32_t tmem_empty0 = mapa_u32(smem_u32(tmem_empty_bar), leader_rank);   // this pair's leader's barriers
        int32_t *xp = s_xpose + ew * 32 * 33;
        int acc_stage = 0;
        uint32_t acc_ph = 0;
        int it = -1;
        for (int round = 0; round * n_clusters < a.n_tiles; ++round) {
            const int q = p2_queue_slot(round, cluster, n_clusters);
            if (q >= a.n_tiles) continue;
            ++it;
            const P2Tile t = a.tiles[a.order[q]];
            const bool empty_tile = t.n_steps == 0;
            const int64_t col0 = (int64_t)t.bj * P2_BN;
            int32_t *den = s_den + (it & 1) * P2_BN;       // double buffered: no barrier against the previous tile's readers
            if (!ep.raw)
                for (int x = threadIdx.x - 32 * P2_FIRST_EPI_WARP; x < P2_BN; x += 32 * P2_EPI_WARPS)
                    den[x] = (col0 + x < n) ? ep.den[col0 + x] : 0;
            asm volatile("bar.sync 1, %0;" ::"n"(32 * P2_EPI_WARPS) : "memory");      // epilogue warps only
            if (!empty_tile) {
                mbar_wait(tmem_full_bar + acc_stage, acc_ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            }
            const int64_t row0 = (int64_t)t.bi * (256 * PAIRS) + (int64_t)pair * 256 + (int64_t)half_rank * P2_BM +
                                 quad * 32;                                                  // first row of this warp
            const int64_t i = row0 + lane;
            const bool row_ok = i < n && i >= ep.row_begin && i < ep.row_end;
            const int den_i = (row_ok && !ep.raw) ? ep.den[i] : 0;
            const int64_t row_base = row_ok ? pc_tri_index(n, i, i) - out_base - i : 0;    // out index of column 0 (may be < 0)
            const unsigned rows_ok = __ballot_sync(0xffffffffu, row_ok);
            for (int c = half * 4; c < half * 4 + 4; ++c) {
                const int64_t cbase = col0 + c * 32;
                const bool skip = cbase + 31 < row0 || cbase >= n;      // nothing of this chunk is in the triangle
                uint32_t v[32];
                if (!empty_tile && !skip) {
                    tmem_ld32(lane_addr + (uint32_t)(acc_stage * acc_per_tile * P2_BN + c * 32), v);
                    if (acc_per_tile == 2) {
                        uint32_t hi[32];
                        tmem_ld32(lane_addr + (uint32_t)((acc_stage * acc_per_tile + 1) * P2_BN + c * 32), hi);
#pragma unroll
                        for (int x = 0; x < 32; ++x) v[x] += hi[x] << 8;
                    }
                } else {
#pragma unroll
                    for (int x = 0; x < 32; ++x) v[x] = 0;
                }
                if (skip) continue;                                     // warp-uniform
                // values of row `lane`, then transposed through smem: store instruction r writes
                // 32 consecutive columns of row r
#pragma unroll
                for (int x = 0; x < 32; ++x) {
                    const int64_t j = cbase + x;
                    int val;
                    if (ep.raw)
                        val = (i == j) ? 0 : (int)v[x];
                    else if (i == j)
                        val = ep.as_distance ? 0 : 1000000;                  // matrix.py:467-468
                    else
                        val = pc_cheap_metric_micro(ep.metric, ep.as_distance, (int)v[x], den_i, den[c * 32 + x]);
                    xp[lane * 33 + x] = val;
                }
                __syncwarp();
                const int64_t j = cbase + lane;
#pragma unroll 4
                for (int r = 0; r < 32; ++r) {
                    const int64_t rb = __shfl_sync(0xffffffffu, row_base, r);
                    const int val = xp[r * 33 + lane];
                    if (((rows_ok >> r) & 1u) && j < n && j >= row0 + r) {
                        ep.out[rb + j] = val;
                        // fused all-gather: the same 128-byte row segment into every peer's triangle
                        for (int o = 0; o < ep.n_peer_out; ++o) ep.peer_out[o][rb + j] = val;
                    }
                }
                __syncwarp();
            }
            if (!empty_tile) {
                // this warp is done reading the accumulator(s): tell the leader's MMA issuer
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(tmem_empty0 + (uint32_t)acc_stage * 8u);
                if (++acc_stage == acc_stages) { acc_stage = 0; acc_ph ^= 1; }
            }
        }
    }
    // nobody leaves (and frees TMEM / lets smem die) while the peer may still signal or read
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cluster_sync_all();
    if (warp == 2) {
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled2)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *,
                                     const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                     const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                     CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int p2_make_map(pc_ctx *ctx, CUtensorMap *map, void *base, int64_t row_bytes, int64_t rows, int box_rows)
{
    static PFN_encodeTiled2 encode = nullptr;
    if (!encode) {
        cudaDriverEntryPointQueryResult q;
        void *fn = nullptr;
        PC_CUDA(ctx, cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
        if (!fn || q != cudaDriverEntryPointSuccess)
            return pc_fail(ctx, PC_ECUDA, "cuTensorMapEncodeTiled is not available from the driver");
        encode = (PFN_encodeTiled2)fn;
    }
    const cuuint64_t dims[2] = {(cuuint64_t)row_bytes, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)row_bytes};
    const cuuint32_t box[2] = {(cuuint32_t)P2_BK, (cuuint32_t)box_rows};     // 128 B x 128 rows (A, half B) or 64 (quarter B)
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, base, dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return pc_fail(ctx, PC_ECUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return PC_OK;
}

// Launch the persistent pair kernel over the operand pc_i8_prepare_operand left in ctx->d_i8.
int pc_i8_pair_contract(pc_ctx *ctx, int n_limbs, int64_t p_pad, int64_t row_bytes, int64_t rows_pad,
                        int occ_words, int occ_shift, const EpilogueArgs &ep)
{
    const int64_t kb_per_seg = p_pad / P2_BK;
    const int n_pass = n_limbs == 0 ? 1 : 2 * n_limbs;
    // two pairs per cluster (B tile multicast) whenever the device can keep (nearly) all SMs busy
    // with clusters of four; PC_I8_PAIRS=1|2 overrides
    static int pairs_choice = 0, clusters_of[3] = {0, 0, 0};
    if (!pairs_choice) {
        for (int pr = 1; pr <= 2; ++pr) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(2 * pr * 64);
            cfg.blockDim = dim3(P2_THREADS);
            cfg.dynamicSmemBytes = P2_SMEM_BYTES;
            cudaLaunchAttribute attr;
            attr.id = cudaLaunchAttributeClusterDimension;
            attr.val.clusterDim.x = 2 * pr; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
            cfg.attrs = &attr;
            cfg.numAttrs = 1;
            int n = 0;
            if (pr == 1) {
                PC_CUDA(ctx, cudaFuncSetAttribute(i8_pair_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, P2_SMEM_BYTES));
                PC_CUDA(ctx, cudaOccupancyMaxActiveClusters(&n, i8_pair_kernel<1>, &cfg));
            } else {
                PC_CUDA(ctx, cudaFuncSetAttribute(i8_pair_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, P2_SMEM_BYTES));
                PC_CUDA(ctx, cudaOccupancyMaxActiveClusters(&n, i8_pair_kernel<2>, &cfg));
            }
            clusters_of[pr] = n;
        }
        pairs_choice = (clusters_of[2] * 4 * 100 >= ctx->sm_count * 94) ? 2 : 1;
        if (const char *env = getenv("PC_I8_PAIRS")) {
            const int v = atoi(env);
            if ((v == 1 || v == 2) && clusters_of[v] > 0) pairs_choice = v;
        }
    }
    const int pairs = pairs_choice;
    const int rows_per_tile = 256 * pairs;
    const int n_groups = (int)(rows_pad / 128);
    const int n_col_tiles = (int)((ctx->N + P2_BN - 1) / P2_BN);
    const int tile_row0 = (int)(ep.row_begin / rows_per_tile);
    const int tile_row1 = (int)((ep.row_end + rows_per_tile - 1) / rows_per_tile);
    const int n_row_tiles = tile_row1 - tile_row0;
    const int64_t cand = (int64_t)n_row_tiles * n_col_tiles;
    // the real tiles: those with bj >= bi * pairs
    int64_t max_tiles = 0;
    for (int bi = tile_row0; bi < tile_row1; ++bi) max_tiles += std::max(0, n_col_tiles - bi * pairs);
    if (max_tiles <= 0) return PC_OK;
    for (int bi = tile_row0; bi < tile_row1; ++bi)          // statistics: what a dense K schedule would issue
        ctx->i8_dense_steps += (int64_t)std::max(0, n_col_tiles - bi * pairs) * kb_per_seg * n_pass * 2 * pairs;

    PC_TRY(devbuf_reserve(ctx, ctx->d_p2_tiles, (size_t)max_tiles * sizeof(P2Tile)));
    PC_TRY(devbuf_reserve(ctx, ctx->d_p2_masks, (size_t)max_tiles * occ_words * sizeof(uint32_t)));
    PC_TRY(devbuf_reserve(ctx, ctx->d_p2_keys, (size_t)max_tiles * 2 * sizeof(uint32_t) + 16));
    PC_TRY(devbuf_reserve(ctx, ctx->d_p2_ids, (size_t)max_tiles * 2 * sizeof(int32_t)));
    const size_t sync_words = (size_t)max_tiles + 8;              // >= 1 + number of rounds
    PC_TRY(devbuf_reserve(ctx, ctx->d_p2_sync, sync_words * sizeof(unsigned int)));
    uint32_t *keys_in = ctx->d_p2_keys.as<uint32_t>(), *keys_out = keys_in + max_tiles;
    int *counter = reinterpret_cast<int *>(keys_in + 2 * max_tiles);
    int32_t *ids_in = ctx->d_p2_ids.as<int32_t>(), *ids_out = ids_in + max_tiles;
    size_t temp_bytes = 0;
    cub::DeviceRadixSort::SortPairsDescending(nullptr, temp_bytes, keys_in, keys_out, ids_in, ids_out, (int)max_tiles,
                                              0, 32, ctx->stream);
    PC_TRY(devbuf_reserve(ctx, ctx->d_p2_temp, temp_bytes));
    if (!ctx->d_i8_steps.p) {
        PC_TRY(devbuf_reserve(ctx, ctx->d_i8_steps, sizeof(unsigned long long)));
        PC_CUDA(ctx, cudaMemsetAsync(ctx->d_i8_steps.p, 0, sizeof(unsigned long long), ctx->stream));
    }
    CUtensorMap map, map_quarter;
    PC_TRY(p2_make_map(ctx, &map, ctx->d_i8.p, row_bytes, rows_pad, P2_BM));
    PC_TRY(p2_make_map(ctx, &map_quarter, ctx->d_i8.p, row_bytes, rows_pad, P2_BN / 4));

    PC_TRY(pc_stage_begin(ctx, PC_T_CONTRACT));
    PC_CUDA(ctx, cudaMemsetAsync(counter, 0, sizeof(int), ctx->stream));
    PC_CUDA(ctx, cudaMemsetAsync(ctx->d_p2_sync.p, 0, sync_words * sizeof(unsigned int), ctx->stream));
    {
        const int64_t threads = cand * 32;
        const int grid = (int)((threads + 255) / 256);
        p2_build_tiles_kernel<<<grid, 256, 0, ctx->stream>>>(ctx->d_i8_occ.as<uint32_t>(), occ_words, occ_shift,
                                                            kb_per_seg, n_pass, pairs, n_groups, tile_row0, n_row_tiles,
                                                            n_col_tiles,
                                                            ctx->d_p2_tiles.as<P2Tile>(), ctx->d_p2_masks.as<uint32_t>(),
                                                            keys_in, ids_in, counter,
                                                            ctx->d_p2_sync.as<unsigned int>());
        PC_CUDA(ctx, cudaGetLastError());
        p2_keys_kernel<<<(int)((max_tiles + 255) / 256), 256, 0, ctx->stream>>>(
            ctx->d_p2_tiles.as<P2Tile>(), (int)max_tiles, pairs, ctx->d_p2_sync.as<unsigned int>(), keys_in);
        PC_CUDA(ctx, cudaGetLastError());
    }
    PC_CUDA(ctx, cub::DeviceRadixSort::SortPairsDescending(ctx->d_p2_temp.p, temp_bytes, keys_in, keys_out, ids_in,
                                                           ids_out, (int)max_tiles, 0, 32, ctx->stream));
    P2Args args;
    args.p_pad = p_pad;
    args.n_limbs = n_limbs;
    args.occ_words = occ_words;
    args.occ_shift = occ_shift;
    args.tiles = ctx->d_p2_tiles.as<P2Tile>();
    args.order = ids_out;
    args.masks = ctx->d_p2_masks.as<uint32_t>();
    args.n_tiles = (int)max_tiles;
    args.tile_steps = ctx->d_i8_steps.as<unsigned long long>();
    args.sync = ctx->d_p2_sync.as<unsigned int>();
    args.ep = ep;
    // one cluster (CTA pair, or two pairs) per TPC group that can host it; fewer when there are fewer tiles
    int clusters = clusters_of[pairs];
    if ((int64_t)clusters > max_tiles) clusters = (int)max_tiles;
    // Lock step pays once the pairs have rounds enough to drift apart: measured on one GPU
    // (tools/probe_lock_threshold.py, profiles/r2_lock_probe.txt) 10 000 genomes x 58k-95k columns
    // (11 rounds) run 3-4 % FASTER without it, 16 000 x 93k (27 rounds) 7 % slower, 20 000 x 154k
    // (43 rounds) 35 % slower.  PC_I8_LOCK = 1 / 2 forces it on / off.
    args.lock_mode = getenv("PC_I8_LOCK") ? atoi(getenv("PC_I8_LOCK")) : (max_tiles > 16 * (int64_t)clusters ? 0 : 2);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(2 * pairs * clusters));
    cfg.blockDim = dim3(P2_THREADS);
    cfg.dynamicSmemBytes = P2_SMEM_BYTES;
    cfg.stream = ctx->stream;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = 2 * pairs; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    if (pairs == 1)
        PC_CUDA(ctx, cudaLaunchKernelEx(&cfg, i8_pair_kernel<1>, map, map_quarter, args));
    else
        PC_CUDA(ctx, cudaLaunchKernelEx(&cfg, i8_pair_kernel<2>, map, map_quarter, args));
    PC_CUDA(ctx, cudaGetLastError());
    ctx->launches += 4;
    ctx->contract_launches++;
    PC_TRY(pc_stage_end(ctx));
    return PC_OK;
}
