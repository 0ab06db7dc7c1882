// pc_contract_i8.cu -- operand packing of the int8 tensor engine + the round-1 kernel (one tile
// per CTA, engine PC_ENGINE_I8MMA_V1; the default engine is the persistent CTA-pair kernel of
// pc_contract_i8p.cu, which this file's pc_i8_contract dispatches to).
// K1/K2/K3 as an int8 tcgen05 GEMM: S = M_a . M_b^T over
// the dense genome x pham operand, s32 accumulators in TMEM, TMA-fed, with the
// per-metric normalisation + Python-exact rounding (K6) fused into the epilogue
// so that only the upper-triangle tiles of the result are ever written.
//
// Replaces, per genome pair, Genome.__and__ + len()
// (/root/reference/src/phamclust/genome.py:124-135) and the arithmetic of
// metrics.py:40-53 (gcs), :70-80 (jc), :102-115 (pocp), :135-157 (af).
//
// Operand (row-major uint8, one row per genome, K contiguous = "K-major"):
//   gcs/jc : [ B ]                    B = 0/1 presence          K = P_pad
//   pocp   : [ C | B ]                C = copies (1 limb)       K = 2 P_pad
//   af     : [ L0 | L1 | B ]          L = L0 + 256 L1 (aa)      K = 3 P_pad
// shared(i,j) = B_i.B_j ; conserved(i,j) = sum_t 256^t (Lt_i.B_j + B_i.Lt_j):
// the row tile reads segment t while the column tile reads segment B and vice
// versa, limb t accumulating into its own TMEM accumulator.
//
// V1 CTA = one (128*MT) x 256 output tile (UMMA M=128, N=256, K=32 per
// instruction, cta_group::1; MT = 2 row tiles sharing one column tile whenever a
// single accumulator per row tile suffices -- r1 ncu showed the 128x256 tile at
// 55 % tensor / 55 % L2 throughput, i.e. limited by operand traffic from L2, and
// the taller tile moves 1.5x fewer bytes per MAC), 448 threads: warp 0 = TMA producer,
// warp 1 = MMA issuer (one elected lane each), warps 2-13 = epilogue (tcgen05.ld 32x32b; a warp
// may only touch TMEM lane quadrant warp%4, three warps share a quadrant and split the columns).
// 3-stage (MT = 2, 64 KB per stage) or 4-stage (MT = 1, 48 KB) smem ring of 128B-swizzled
// 128-byte K slabs, mbarrier full/empty pairs, tcgen05.commit releases the slots.
// Tensor-pipe bound: 2*K*128*256 int8 ops per tile.
//
// Block sparsity: pack_i8_kernel also records which (128-genome group, 128-column
// K block) pairs hold any presence bit.  A CTA ANDs the occupancy of its row
// groups with that of its column groups and only streams the K blocks both sides
// occupy; a tile whose genomes share no K block at all (e.g. genomes of
// different phage clusters, whose phams sit in disjoint column ranges) never
// touches the tensor core: its shared counts are 0 by construction and the
// epilogue values are written directly.  Unclustered tables (C3/C5) still skip the K
// blocks of rare phams that a tile's genomes do not hold (22 % / 13 % of the blocks).
#include <cuda.h>

#include "pc_common.cuh"
#include "pc_tcgen05.cuh"

#define I8_BM 128
#define I8_BN 256
#define I8_BK 128                 // bytes (= int8 elements) per stage
#ifndef I8_EPI_WARPS
#define I8_EPI_WARPS 12           // three warps per TMEM lane quadrant, each drains a share of the columns
#endif
#define I8_THREADS (64 + 32 * I8_EPI_WARPS)
#define I8_A_BYTES (I8_BM * I8_BK)
#define I8_B_BYTES (I8_BN * I8_BK)
#define I8_MAX_STAGES 4
// occupancy words per genome group: bit b covers K blocks [b << occ_shift, (b + 1) << occ_shift);
// occ_shift is the smallest shift that fits the segment's K blocks into this many words
#define I8_MAX_OCC_WORDS 1024
template <int MT> struct I8Cfg {
    static constexpr int STAGE_BYTES = MT * I8_A_BYTES + I8_B_BYTES;     // 48 KB / 64 KB
    static constexpr int STAGES = MT == 1 ? 4 : 3;                        // 192 KB ring either way
    static constexpr int MASK_OFF = STAGES * STAGE_BYTES + 256 /*barriers*/ + 2 * I8_BN * 4 /*denominators*/;
    static constexpr int SMEM_BYTES = MASK_OFF + I8_MAX_OCC_WORDS * 4 + 1024 /*align*/;
};

struct I8Args {
    int64_t p_pad;        // bytes per segment (multiple of 128)
    int n_limbs;          // 0: gcs/jc (single segment); else weight limbs before the B segment
    int tile_row0;        // first 128-row tile of the launch
    const uint32_t *occ;  // [rows_pad / 128][occ_words]: bit b set <=> the group holds a pham of K blocks b << occ_shift ..
    int occ_words;
    int occ_shift;
    unsigned long long *tile_steps;   // += MT per K block visited (one 128 x 256 x 128 MMA group each)
    EpilogueArgs ep;
};

// instruction descriptor: D = s32, A = B = unsigned 8 bit, K-major, N = 256, M = 128
__device__ __forceinline__ constexpr uint32_t i8_idesc()
{
    return (2u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(I8_BN >> 3) << 17) |
           ((uint32_t)(I8_BM >> 4) << 24);
}

// K schedule: n_limbs == 0 -> one pass over the K blocks of the single segment;
// else per weight limb two passes (row tile reads the limb while the column tile
// reads B, then swapped), limb t accumulating into accumulator t.  Within a pass
// only the K blocks of the CTA's occupancy mask are visited, in ascending order
// (producer and MMA issuer walk the same mask, so they agree on the sequence).
struct I8Step {
    int64_t a_col, b_col;
    int acc;
    bool first;
};

template <typename F>
__device__ __forceinline__ void i8_for_each_step(const I8Args &a, const uint32_t *mask, F &&body)
{
    const int64_t kb_per_seg = a.p_pad / I8_BK;
    const int n_pass = a.n_limbs == 0 ? 1 : 2 * a.n_limbs;
    for (int pass = 0; pass < n_pass; ++pass) {
        const int limb = pass >> 1;
        const bool swapped = (pass & 1) != 0;
        bool first = !swapped;                       // first step of this accumulator
        for (int w = 0; w < a.occ_words; ++w) {
            uint32_t bits = mask[w];
            while (bits) {
                const int64_t b = (int64_t)w * 32 + (__ffs((int)bits) - 1);
                bits &= bits - 1;
                const int64_t kb_end = min((b + 1) << a.occ_shift, kb_per_seg);
                for (int64_t kb = b << a.occ_shift; kb < kb_end; ++kb) {
                    I8Step st;
                    if (a.n_limbs == 0) {
                        st.a_col = st.b_col = kb * I8_BK;
                        st.acc = 0;
                    } else {
                        const int64_t wcol = (int64_t)limb * a.p_pad + kb * I8_BK;
                        const int64_t bcol = (int64_t)a.n_limbs * a.p_pad + kb * I8_BK;
                        st.a_col = swapped ? bcol : wcol;
                        st.b_col = swapped ? wcol : bcol;
                        st.acc = limb;
                    }
                    st.first = first;
                    first = false;
                    body(st);
                }
            }
        }
    }
}

template <int MT>
__global__ void __launch_bounds__(I8_THREADS, 1)
i8_contract_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                   const I8Args a)
{
    constexpr int I8_STAGES = I8Cfg<MT>::STAGES;
    constexpr int I8_STAGE_BYTES = I8Cfg<MT>::STAGE_BYTES;
    // last row tiles first: with the columns in first-genome order the late genomes occupy the
    // most K blocks, so their tiles are the long ones and must not form the tail of the launch.
    // (A banded launch order -- 8 row tiles column-major -- cut the DRAM reads of C3 jc from 3.4
    // to 2.7 GB but lost 4-9 % of time to a longer tail: DRAM is not what binds this kernel.)
    const int bi = (int)(gridDim.y - 1 - blockIdx.y) + a.tile_row0;      // (128*MT)-row tile
    const int bj = blockIdx.x;                                           // 256-column tile
    if ((int64_t)bj * I8_BN + (I8_BN - 1) < (int64_t)bi * I8_BM * MT) return;   // strictly lower triangle

    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t *tiles = smem;                                            // STAGES x (A | B)
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem + I8_STAGES * I8_STAGE_BYTES);
    uint64_t *empty_bar = full_bar + I8_STAGES;
    uint64_t *tmem_full_bar = empty_bar + I8_STAGES;
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(tmem_full_bar + 1);
    int32_t *s_den = reinterpret_cast<int32_t *>(smem + I8_STAGES * I8_STAGE_BYTES + 256);
    uint32_t *s_mask = reinterpret_cast<uint32_t *>(smem + I8Cfg<MT>::MASK_OFF);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int n_acc = (a.n_limbs == 0 ? 1 : a.n_limbs) * MT;       // host guarantees <= 2
    const uint32_t tmem_cols = n_acc * I8_BN <= 256 ? 256 : 512;

    // K blocks this tile has to visit: occupied by one of its row groups AND one of its column groups
    int any_block = 0;
    for (int w = threadIdx.x; w < a.occ_words; w += I8_THREADS) {
        uint32_t rows = 0;
#pragma unroll
        for (int u = 0; u < MT; ++u) rows |= a.occ[(int64_t)(bi * MT + u) * a.occ_words + w];
        uint32_t cols = 0;
#pragma unroll
        for (int u = 0; u < I8_BN / I8_BM; ++u)
            cols |= a.occ[(int64_t)(bj * (I8_BN / I8_BM) + u) * a.occ_words + w];
        s_mask[w] = rows & cols;
        any_block |= (rows & cols) != 0;
    }

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < I8_STAGES; ++s) {
            mbar_init(full_bar + s, 1);
            mbar_init(empty_bar + s, 1);
        }
        mbar_init(tmem_full_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                         smem_u32(tmem_ptr)),
                     "r"(tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    // no common K block: every shared count of the tile is 0, the tensor core is never started
    const bool empty_tile = __syncthreads_or(any_block) == 0;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0;
            i8_for_each_step(a, s_mask, [&](const I8Step &st) {
                mbar_wait(empty_bar + s, ph ^ 1);
                mbar_expect_tx(full_bar + s, I8_STAGE_BYTES);
                uint8_t *sa = tiles + (size_t)s * I8_STAGE_BYTES;
#pragma unroll
                for (int u = 0; u < MT; ++u)
                    tma_load_2d(sa + u * I8_A_BYTES, &map_a, full_bar + s, (int32_t)st.a_col,
                                (bi * MT + u) * I8_BM);
                tma_load_2d(sa + MT * I8_A_BYTES, &map_b, full_bar + s, (int32_t)st.b_col, bj * I8_BN);
                if (++s == I8_STAGES) { s = 0; ph ^= 1; }
            });
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            const uint32_t idesc = i8_idesc();
            int s = 0, n_steps = 0;
            uint32_t ph = 0;
            i8_for_each_step(a, s_mask, [&](const I8Step &st) {
                const int acc = st.acc;
                const bool first = st.first;
                mbar_wait(full_bar + s, ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t sa = smem_u32(tiles + (size_t)s * I8_STAGE_BYTES);
                const uint64_t bdesc = umma_desc_sw128_kmajor(sa + MT * I8_A_BYTES);
#pragma unroll
                for (int u = 0; u < MT; ++u) {
                    const uint64_t adesc = umma_desc_sw128_kmajor(sa + u * I8_A_BYTES);
                    const uint32_t tmem_d = tmem_base + (uint32_t)(MT == 1 ? acc : u) * I8_BN;
#pragma unroll
                    for (int k = 0; k < I8_BK / 32; ++k) {
                        // advance 32 bytes of K inside the 128B swizzle atom: +2 in 16 B units
                        umma_i8(tmem_d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc,
                                (first && k == 0) ? 0u : 1u);
                    }
                }
                umma_commit(empty_bar + s);          // frees the smem slot when the MMAs retire
                ++n_steps;
                if (++s == I8_STAGES) { s = 0; ph ^= 1; }
            });
            if (!empty_tile) umma_commit(tmem_full_bar);   // accumulators complete
            if (n_steps) atomicAdd(a.tile_steps, (unsigned long long)n_steps * MT);
        }
    } else {
        // ===== epilogue: warps 2..13, TMEM lane quadrant = warp % 4 =====
        const int quad = warp & 3;
        const int et = threadIdx.x - 64;              // 0 .. 32*I8_EPI_WARPS-1
        constexpr int EPI_GROUPS = I8_EPI_WARPS / 4;  // warps sharing one TMEM lane quadrant
        const int grp = (warp - 2) >> 2;              // this warp's share of the 8 column chunks
        const int c_begin = grp * (I8_BN / 32) / EPI_GROUPS, c_end = (grp + 1) * (I8_BN / 32) / EPI_GROUPS;
        const EpilogueArgs &ep = a.ep;
        const int64_t n = ep.n;
        const int64_t col0 = (int64_t)bj * I8_BN;
        if (!ep.raw)
            for (int x = et; x < I8_BN; x += 32 * I8_EPI_WARPS) s_den[x] = (col0 + x < n) ? ep.den[col0 + x] : 0;
        asm volatile("bar.sync 1, %0;" ::"n"(32 * I8_EPI_WARPS) : "memory");      // epilogue warps only
        if (!empty_tile) {
            mbar_wait(tmem_full_bar, 0);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }

        const int64_t out_base = pc_tri_index(n, ep.row_begin, ep.row_begin);
        const uint32_t lane_addr = tmem_base + ((uint32_t)(quad * 32) << 16);
#pragma unroll 1
        for (int u = 0; u < MT; ++u) {
            const int64_t i = ((int64_t)bi * MT + u) * I8_BM + quad * 32 + lane;
            const bool row_ok = i < n && i >= ep.row_begin && i < ep.row_end;
            const int den_i = (row_ok && !ep.raw) ? ep.den[i] : 0;
            const int64_t row_base = row_ok ? pc_tri_index(n, i, i) - out_base - i : 0;
            // column chunks entirely left of this warp's rows hold nothing to write
            const int64_t warp_row0 = ((int64_t)bi * MT + u) * I8_BM + quad * 32;
            for (int c = c_begin; c < c_end; ++c) {
                if (col0 + c * 32 + 31 < warp_row0) continue;          // warp-uniform
                uint32_t v[32];
                if (empty_tile) {
#pragma unroll
                    for (int x = 0; x < 32; ++x) v[x] = 0;
                } else {
                    tmem_ld32(lane_addr + (uint32_t)((MT == 1 ? 0 : u) * I8_BN + c * 32), v);
                }
                if (!empty_tile && MT == 1 && n_acc == 2) {
                    uint32_t hi[32];
                    tmem_ld32(lane_addr + (uint32_t)(I8_BN + c * 32), hi);
#pragma unroll
                    for (int x = 0; x < 32; ++x) v[x] += hi[x] << 8;
                }
                if (!row_ok) continue;
#pragma unroll
                for (int x = 0; x < 32; ++x) {
                    const int64_t j = col0 + c * 32 + x;
                    if (j >= n || j < i) continue;
                    int val;
                    if (ep.raw)
                        val = (i == j) ? 0 : (int)v[x];
                    else if (i == j)
                        val = ep.as_distance ? 0 : 1000000;                  // matrix.py:467-468
                    else
                        val = pc_cheap_metric_micro(ep.metric, ep.as_distance, (int)v[x], den_i,
                                                    s_den[c * 32 + x]);
                    ep.out[row_base + j] = val;
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols)
                     : "memory");
    }
}

// ---------------------------------------------------------------------------
// operand packing: CSR entries -> dense uint8 rows
// ---------------------------------------------------------------------------
__global__ void pack_i8_kernel(const int32_t *__restrict__ entry_genome,
                               const int32_t *__restrict__ pham_id,
                               const int32_t *__restrict__ pham_col,
                               const uint32_t *__restrict__ weight, int n_limbs, int64_t p_pad,
                               int64_t row_bytes, int64_t nnz, uint8_t *__restrict__ dst,
                               uint32_t *__restrict__ occ, int occ_words, int occ_shift)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < nnz; e += stride) {
        uint8_t *row = dst + (int64_t)entry_genome[e] * row_bytes;
        const int32_t p = pham_col[pham_id[e]];
        if (p < 0) continue;                     // held by one genome only: never shared
        const uint32_t w = weight ? weight[e] : 0;
        for (int t = 0; t < n_limbs; ++t) row[(int64_t)t * p_pad + p] = (uint8_t)((w >> (8 * t)) & 0xffu);
        row[(int64_t)n_limbs * p_pad + p] = 1;
        // occupancy of (128-genome group, K block); most entries find their bit already set
        const int b = (p / I8_BK) >> occ_shift;
        uint32_t *word = occ + (int64_t)(entry_genome[e] / I8_BM) * occ_words + (b >> 5);
        const uint32_t bit = 1u << (b & 31);
        if (!(*reinterpret_cast<volatile uint32_t *>(word) & bit)) atomicOr(word, bit);
    }
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *,
                                    const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                    const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int make_map(pc_ctx *ctx, CUtensorMap *map, void *base, int64_t row_bytes, int64_t rows,
                    int box_rows)
{
    static PFN_encodeTiled encode = nullptr;
    if (!encode) {
        cudaDriverEntryPointQueryResult q;
        void *fn = nullptr;
        PC_CUDA(ctx, cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
        if (!fn || q != cudaDriverEntryPointSuccess)
            return pc_fail(ctx, PC_ECUDA, "cuTensorMapEncodeTiled is not available from the driver");
        encode = (PFN_encodeTiled)fn;
    }
    const cuuint64_t dims[2] = {(cuuint64_t)row_bytes, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)row_bytes};
    const cuuint32_t box[2] = {(cuuint32_t)I8_BK, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, base, dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return pc_fail(ctx, PC_ECUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return PC_OK;
}

// 0 if the int8 engine can serve `metric` for the loaded table
int pc_i8_supported(const pc_ctx *ctx, int metric)
{
    if (metric == PC_POCP && ctx->max_copies > 255) return 0;
    if (metric == PC_AF && ctx->max_len_sum > 65535) return 0;
    return 1;
}

int pc_i8_contract(pc_ctx *ctx, int metric, const EpilogueArgs &ep)
{
    if (ep.row_end <= ep.row_begin) return PC_OK;
    if (!pc_i8_supported(ctx, metric))
        return pc_fail(ctx, PC_EINVAL, "int8 engine: weights of metric %d do not fit the limb layout", metric);
    const int kind = (metric == PC_GCS || metric == PC_JC) ? 0 : (metric == PC_POCP ? 1 : 2);
    int n_limbs = 0;
    if (kind == 1) n_limbs = 1;
    if (kind == 2) n_limbs = ctx->max_len_sum > 255 ? 2 : 1;
    const int64_t p_pad = std::max<int64_t>(I8_BK, (ctx->P_eff + I8_BK - 1) / I8_BK * I8_BK);
    const int64_t row_bytes = p_pad * (n_limbs + 1);
    const int64_t rows_pad = (ctx->N + I8_BN - 1) / I8_BN * I8_BN;
    const int64_t kb_per_seg = p_pad / I8_BK;
    auto occ_words_at = [&](int shift) { return (((kb_per_seg - 1) >> shift) + 32) / 32; };
    int occ_shift = 0;
    while (occ_words_at(occ_shift) > I8_MAX_OCC_WORDS) ++occ_shift;
    const int occ_words = (int)occ_words_at(occ_shift);
    const size_t occ_bytes = (size_t)(rows_pad / I8_BM) * occ_words * sizeof(uint32_t);
    if (ctx->i8_kind != kind || ctx->i8_K != row_bytes) {
        const size_t bytes = (size_t)rows_pad * row_bytes;
        PC_TRY(devbuf_reserve(ctx, ctx->d_i8, bytes));
        PC_TRY(devbuf_reserve(ctx, ctx->d_i8_occ, occ_bytes));
        PC_TRY(pc_stage_begin(ctx, PC_T_PACK));
        PC_CUDA(ctx, cudaMemsetAsync(ctx->d_i8.p, 0, bytes, ctx->stream));
        PC_CUDA(ctx, cudaMemsetAsync(ctx->d_i8_occ.p, 0, occ_bytes, ctx->stream));
        if (ctx->nnz) {
            const uint32_t *w = kind == 0 ? nullptr
                              : kind == 1 ? ctx->d_copies.as<uint32_t>() : ctx->d_len_sum.as<uint32_t>();
            int64_t blocks = (ctx->nnz + 255) / 256;
            int64_t cap = (int64_t)ctx->sm_count * 16;
            pack_i8_kernel<<<(int)(blocks < cap ? blocks : cap), 256, 0, ctx->stream>>>(
                ctx->d_entry_genome.as<int32_t>(), ctx->d_pham_id.as<int32_t>(), ctx->d_pham_col.as<int32_t>(), w,
                n_limbs, p_pad,
                row_bytes, ctx->nnz, ctx->d_i8.as<uint8_t>(), ctx->d_i8_occ.as<uint32_t>(), occ_words,
                occ_shift);
            PC_CUDA(ctx, cudaGetLastError());
            ctx->launches++;
        }
        PC_TRY(pc_stage_end(ctx));
        ctx->i8_kind = kind;
        ctx->i8_K = row_bytes;
    }
    if (ctx->engine != PC_ENGINE_I8MMA_V1)       // default: the persistent CTA-pair kernel
        return pc_i8_pair_contract(ctx, n_limbs, p_pad, row_bytes, rows_pad, occ_words, occ_shift, ep);
    CUtensorMap map_a, map_b;
    PC_TRY(make_map(ctx, &map_a, ctx->d_i8.p, row_bytes, rows_pad, I8_BM));
    PC_TRY(make_map(ctx, &map_b, ctx->d_i8.p, row_bytes, rows_pad, I8_BN));
    I8Args args;
    args.p_pad = p_pad;
    args.n_limbs = n_limbs;
    args.occ = ctx->d_i8_occ.as<uint32_t>();
    args.occ_words = occ_words;
    args.occ_shift = occ_shift;
    if (!ctx->d_i8_steps.p) {
        PC_TRY(devbuf_reserve(ctx, ctx->d_i8_steps, sizeof(unsigned long long)));
        PC_CUDA(ctx, cudaMemsetAsync(ctx->d_i8_steps.p, 0, sizeof(unsigned long long), ctx->stream));
    }
    args.tile_steps = ctx->d_i8_steps.as<unsigned long long>();
    args.ep = ep;
    // two row tiles per CTA whenever one accumulator per row tile is enough
    const int mt = n_limbs <= 1 ? 2 : 1;
    const int64_t rows_per_tile = (int64_t)I8_BM * mt;
    args.tile_row0 = (int)(ep.row_begin / rows_per_tile);
    const int tile_row1 = (int)((ep.row_end + rows_per_tile - 1) / rows_per_tile);
    dim3 grid((unsigned)((ctx->N + I8_BN - 1) / I8_BN), (unsigned)(tile_row1 - args.tile_row0));
    // MMA groups the dense schedule would issue for this launch (statistics only)
    for (unsigned by = 0; by < grid.y; ++by) {
        const int64_t first_col_tile = ((int64_t)(by + args.tile_row0) * rows_per_tile) / I8_BN;
        if (first_col_tile < (int64_t)grid.x)
            ctx->i8_dense_steps += ((int64_t)grid.x - first_col_tile) * kb_per_seg * (n_limbs == 0 ? 1 : 2 * n_limbs) * mt;
    }
    PC_TRY(pc_stage_begin(ctx, PC_T_CONTRACT));
    if (mt == 2) {
        PC_CUDA(ctx, cudaFuncSetAttribute(i8_contract_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          I8Cfg<2>::SMEM_BYTES));
        i8_contract_kernel<2><<<grid, I8_THREADS, I8Cfg<2>::SMEM_BYTES, ctx->stream>>>(map_a, map_b, args);
    } else {
        PC_CUDA(ctx, cudaFuncSetAttribute(i8_contract_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          I8Cfg<1>::SMEM_BYTES));
        i8_contract_kernel<1><<<grid, I8_THREADS, I8Cfg<1>::SMEM_BYTES, ctx->stream>>>(map_a, map_b, args);
    }
    PC_CUDA(ctx, cudaGetLastError());
    ctx->launches++;
    ctx->contract_launches++;
    PC_TRY(pc_stage_end(ctx));
    return PC_OK;
}
