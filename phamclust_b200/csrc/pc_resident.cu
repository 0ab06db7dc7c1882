// pc_resident.cu -- the finished triangle kept in HBM for the FIRST consumer of the matrix.
//
// What the reference does next with a distance matrix is thresholded clustering
// (/root/reference/src/phamclust/scripts/phamclust.py:300-317: nr pre-grouping at a small
// distance, medoids, clustering of the representatives; clustering.py:27 exports the dense
// matrix).  At 50 000 genomes the packed triangle is 5 GB; shipping it to the host just to ask
// "which pairs are closer than eps" is the dominant cost of that stage.  These entry points
// answer that question where the triangle already is:
//   pc_matrix_resident            whole triangle into a ctx-owned device buffer (no D2H)
//   pc_resident_threshold_edges   the off-diagonal pairs with weight <= max_micro, as an edge
//                                 list (what SymMatrix.threshold_components scans for on the host)
//   pc_resident_gather            arbitrary entries by packed index (the small dense blocks of
//                                 each connected component that scikit-learn then clusters)
//   pc_resident_fetch             the whole triangle after all (D2H)
#include "pc_common.cuh"

__global__ void threshold_edges_kernel(const int32_t *__restrict__ tri, int64_t n, int64_t total,
                                       int32_t max_micro, unsigned long long *__restrict__ count,
                                       int2 *__restrict__ edges, int64_t cap)
{
    const int lane = threadIdx.x & 31;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t first = (int64_t)blockIdx.x * blockDim.x + (threadIdx.x & ~31); first < total; first += stride) {
        const int64_t k = first + lane;
        bool hit = false;
        int64_t i = 0, j = 0;
        if (k < total && tri[k] <= max_micro) {
            // recover (i, j) from the packed index: row i starts at pc_tri_index(n, i, i)
            const double tn = (double)(2 * n + 1);
            i = (int64_t)floor((tn - sqrt(tn * tn - 8.0 * (double)k)) * 0.5);
            if (i < 0) i = 0;
            if (i > n - 1) i = n - 1;
            while (i > 0 && pc_tri_index(n, i, i) > k) --i;
            while (i + 1 < n && pc_tri_index(n, i + 1, i + 1) <= k) ++i;
            j = i + (k - pc_tri_index(n, i, i));
            hit = j != i;
        }
        const unsigned mask = __ballot_sync(0xffffffffu, hit);
        if (!mask) continue;
        unsigned long long base = 0;
        const int leader = __ffs((int)mask) - 1;
        if (lane == leader) base = atomicAdd(count, (unsigned long long)__popc(mask));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (hit) {
            const unsigned long long slot = base + __popc(mask & ((1u << lane) - 1u));
            if ((int64_t)slot < cap) edges[slot] = make_int2((int)i, (int)j);
        }
    }
}

__global__ void gather_entries_kernel(const int32_t *__restrict__ tri, int64_t n_idx,
                                      const int64_t *__restrict__ idx, int32_t *__restrict__ out)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n_idx; k += stride) out[k] = tri[idx[k]];
}

int pc_launch_threshold_edges(pc_ctx *ctx, const int32_t *tri, int64_t n, int32_t max_micro,
                              unsigned long long *count, int2 *edges, int64_t cap)
{
    const int64_t total = pc_tri_index(n, n, n);
    if (total <= 0) return PC_OK;
    int64_t blocks = (total + 255) / 256;
    int64_t limit = (int64_t)ctx->sm_count * 16;
    threshold_edges_kernel<<<(int)(blocks < limit ? blocks : limit), 256, 0, ctx->stream>>>(tri, n, total, max_micro,
                                                                                         count, edges, cap);
    PC_CUDA(ctx, cudaGetLastError());
    ctx->launches++;
    return PC_OK;
}

int pc_launch_gather_entries(pc_ctx *ctx, const int32_t *tri, int64_t n_idx, const int64_t *idx, int32_t *out)
{
    if (n_idx <= 0) return PC_OK;
    int64_t blocks = (n_idx + 255) / 256;
    int64_t limit = (int64_t)ctx->sm_count * 16;
    gather_entries_kernel<<<(int)(blocks < limit ? blocks : limit), 256, 0, ctx->stream>>>(tri, n_idx, idx, out);
    PC_CUDA(ctx, cudaGetLastError());
    ctx->launches++;
    return PC_OK;
}
