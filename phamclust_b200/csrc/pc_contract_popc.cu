// pc_contract_popc.cu -- K1/K2/K3 as a bit-packed AND+POPC tile kernel, with
// the per-metric normalisation + Python-exact 6-dp rounding (K6) fused into
// the epilogue so only the upper-triangle result tile is written.
//
// Replaces, per genome pair, Genome.__and__ + len()
// (/root/reference/src/phamclust/genome.py:124-135) and the arithmetic of
// metrics.py:40-53 (gcs), :70-80 (jc), :102-115 (pocp), :135-157 (af).
//
//   shared(i,j)      = sum_w popc(B[i][w] & B[j][w])
//   conserved(i,j)   = sum_planes 2^b * ( popc(Wb[i] & B[j]) + popc(B[i] & Wb[j]) )
//
// where B is the genome x pham presence bit matrix and Wb the b-th bit plane
// of the per-(genome,pham) weight (copies for pocp, summed aa length for af).
// Every product is a sum of `terms` (a-rows, b-rows, multiplier).
//
// Tile: 128 x 128 genome pairs per CTA, 256 threads, 8 x 8 pairs per thread,
// K staged 32 words (1024 phams) at a time through shared memory with a
// register prefetch of the next stage.  POPC-issue bound (16 lanes/clk/SM).
#include "pc_common.cuh"

#define PB_M 128
#define PB_N 128
#define PB_K 32            // 32-bit words per stage
#define PB_LD (PB_K + 4)   // padded row stride (words): conflict-free LDS.128
#define PB_THREADS 256

// ---------------------------------------------------------------------------
// packing: CSR entries -> bit rows
// ---------------------------------------------------------------------------
__global__ void pack_bits_kernel(const int32_t *__restrict__ entry_genome,
                                 const int32_t *__restrict__ pham_id,
                                 const int32_t *__restrict__ pham_col,
                                 const uint32_t *__restrict__ weight, int plane,
                                 int64_t nnz, int64_t words_per_row,
                                 uint32_t *__restrict__ dst)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < nnz; e += stride) {
        if (weight && !((weight[e] >> plane) & 1u)) continue;
        const int32_t p = pham_col[pham_id[e]];
        if (p < 0) continue;                     // held by one genome only: never shared
        atomicOr(dst + (int64_t)entry_genome[e] * words_per_row + (p >> 5), 1u << (p & 31));
    }
}

int pc_launch_pack_bits(pc_ctx *ctx, int source, int plane, uint32_t *dst)
{
    PC_CUDA(ctx, cudaMemsetAsync(dst, 0, (size_t)ctx->N * ctx->W * sizeof(uint32_t), ctx->stream));
    if (ctx->nnz == 0) return PC_OK;
    const uint32_t *w = source == 0 ? nullptr
                      : source == 1 ? ctx->d_copies.as<uint32_t>()
                                    : ctx->d_len_sum.as<uint32_t>();
    int64_t blocks = (ctx->nnz + 255) / 256;
    int64_t cap = (int64_t)ctx->sm_count * 16;
    int grid = (int)(blocks < cap ? blocks : cap);
    pack_bits_kernel<<<grid, 256, 0, ctx->stream>>>(ctx->d_entry_genome.as<int32_t>(),
                                                   ctx->d_pham_id.as<int32_t>(), ctx->d_pham_col.as<int32_t>(), w, plane,
                                                   ctx->nnz, ctx->W, dst);
    PC_CUDA(ctx, cudaGetLastError());
    ctx->launches++;
    return PC_OK;
}

// ---------------------------------------------------------------------------
// the tile kernel
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(PB_THREADS)
popc_contract_kernel(const PopcTerm *__restrict__ terms, int n_terms, int64_t words_per_row,
                     int tile_row0, const EpilogueArgs ep)
{
    const int bi = blockIdx.y + tile_row0;
    const int bj = blockIdx.x;
    if (bj < bi) return;                      // lower triangle: nothing to do

    __shared__ __align__(16) uint32_t As[PB_M * PB_LD];
    __shared__ __align__(16) uint32_t Bs[PB_N * PB_LD];

    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int64_t row0 = (int64_t)bi * PB_M, col0 = (int64_t)bj * PB_N;
    const int64_t n = ep.n;

    int acc[8][8];
#pragma unroll
    for (int u = 0; u < 8; ++u)
#pragma unroll
        for (int v = 0; v < 8; ++v) acc[u][v] = 0;

    // global -> register staging: 128 rows x 32 words = 1024 uint4 per operand
    uint4 ra[4], rb[4];
    const int ld_row = tid >> 3;              // + 32*s
    const int ld_c4 = tid & 7;
    const int64_t n_k = words_per_row / PB_K;

    for (int t = 0; t < n_terms; ++t) {
        const uint32_t *__restrict__ A = terms[t].a;
        const uint32_t *__restrict__ B = terms[t].b;
        const int mult = terms[t].mult;

        auto load_stage = [&](int64_t kb) {
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                const int r = ld_row + 32 * s;
                const int64_t ga = row0 + r, gb = col0 + r;
                const int64_t off = kb * PB_K + 4 * ld_c4;
                ra[s] = (ga < n) ? __ldg(reinterpret_cast<const uint4 *>(A + ga * words_per_row + off))
                                 : make_uint4(0, 0, 0, 0);
                rb[s] = (gb < n) ? __ldg(reinterpret_cast<const uint4 *>(B + gb * words_per_row + off))
                                 : make_uint4(0, 0, 0, 0);
            }
        };
        load_stage(0);
        for (int64_t kb = 0; kb < n_k; ++kb) {
            __syncthreads();                  // previous stage fully consumed
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                const int r = ld_row + 32 * s;
                *reinterpret_cast<uint4 *>(As + r * PB_LD + 4 * ld_c4) = ra[s];
                *reinterpret_cast<uint4 *>(Bs + r * PB_LD + 4 * ld_c4) = rb[s];
            }
            __syncthreads();
            if (kb + 1 < n_k) load_stage(kb + 1);
#pragma unroll 4
            for (int k2 = 0; k2 < PB_K / 2; ++k2) {
                // two words of K at a time: 32 operand registers instead of 64 (no spills)
                uint2 a[8], b[8];
#pragma unroll
                for (int u = 0; u < 8; ++u)
                    a[u] = *reinterpret_cast<const uint2 *>(As + (ty + 16 * u) * PB_LD + 2 * k2);
#pragma unroll
                for (int v = 0; v < 8; ++v)
                    b[v] = *reinterpret_cast<const uint2 *>(Bs + (tx + 16 * v) * PB_LD + 2 * k2);
#pragma unroll
                for (int u = 0; u < 8; ++u)
#pragma unroll
                    for (int v = 0; v < 8; ++v) {
                        const int c = __popc(a[u].x & b[v].x) + __popc(a[u].y & b[v].y);
                        acc[u][v] += c * mult;
                    }
            }
        }
    }

    // fused epilogue: only pairs i <= j inside [row_begin, row_end) are written
    const int64_t out_base = pc_tri_index(n, ep.row_begin, ep.row_begin);
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        const int64_t i = row0 + ty + 16 * u;
        if (i >= n || i < ep.row_begin || i >= ep.row_end) continue;
        const int den_i = ep.raw ? 0 : ep.den[i];
        const int64_t row_base = pc_tri_index(n, i, i) - out_base - i;
#pragma unroll
        for (int v = 0; v < 8; ++v) {
            const int64_t j = col0 + tx + 16 * v;
            if (j >= n || j < i) continue;
            int val;
            if (ep.raw)
                val = (i == j) ? 0 : acc[u][v];
            else if (i == j)
                val = ep.as_distance ? 0 : 1000000;           // matrix.py:467-468
            else
                val = pc_cheap_metric_micro(ep.metric, ep.as_distance, acc[u][v], den_i, ep.den[j]);
            ep.out[row_base + j] = val;
        }
    }
}

int pc_launch_popc_contract(pc_ctx *ctx, const PopcTerm *d_terms, int n_terms,
                            const EpilogueArgs &ep)
{
    if (ep.row_end <= ep.row_begin) return PC_OK;
    const int tiles = (int)((ctx->N + PB_N - 1) / PB_N);
    const int tile_row0 = (int)(ep.row_begin / PB_M);
    const int tile_row1 = (int)((ep.row_end + PB_M - 1) / PB_M);
    dim3 grid(tiles, tile_row1 - tile_row0);
    popc_contract_kernel<<<grid, PB_THREADS, 0, ctx->stream>>>(d_terms, n_terms, ctx->W,
                                                              tile_row0, ep);
    PC_CUDA(ctx, cudaGetLastError());
    ctx->launches++;
    ctx->contract_launches++;
    return PC_OK;
}
