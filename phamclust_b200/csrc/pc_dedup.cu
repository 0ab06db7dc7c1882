// pc_dedup.cu -- canonical gene ids: genes with byte-identical translations
// share one id, so that an ordered (query, ref) pair of SEQUENCES is aligned
// once no matter how many genomes carry those sequences (phage tables are
// duplicate-rich; valid because _align is a pure function of its two strings,
// /root/reference/src/phamclust/metrics.py:160-175; SURVEY.md 7 hard part 8).
//
//   hash every gene (one warp per gene) -> stable radix sort of (hash, gene)
//   -> a run of equal hashes whose bytes equal their predecessor's is one group
//   -> canon[gene] = smallest gene index of its group.
// Hash collisions can only SPLIT a group (bytes are compared), never merge two
// different sequences.  CUB (part of the CUDA toolkit) does the sort and scan.
#include <cub/cub.cuh>

#include "pc_common.cuh"

__global__ void gene_hash_kernel(const uint8_t *__restrict__ codes, const int64_t *__restrict__ gene_off,
                                 const int32_t *__restrict__ gene_len, int64_t n_genes,
                                 unsigned long long *__restrict__ hash, int32_t *__restrict__ index)
{
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int lane = threadIdx.x & 31;
    for (int64_t g = warp; g < n_genes; g += n_warps) {
        const uint8_t *s = codes + gene_off[g];
        const int len = gene_len[g];
        // per-lane FNV-1a over the lane's strided bytes, then a fixed-order fold
        unsigned long long h = 1469598103934665603ull ^ (unsigned long long)(lane * 0x9E3779B97F4A7C15ull);
        for (int x = lane; x < len; x += 32) {
            h ^= (unsigned long long)s[x] + ((unsigned long long)x << 8);
            h *= 1099511628211ull;
        }
        for (int d = 16; d >= 1; d >>= 1) {
            const unsigned long long o = __shfl_xor_sync(0xffffffffu, h, d);
            h = (h ^ (o + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2))) * 1099511628211ull;
            // symmetric in (h, o) would lose order; mix the partner's lane parity in
            h ^= (unsigned long long)((lane & d) ? 0x51ed27ull : 0x2545F4914F6CDD1Dull);
        }
        // lanes end with different values; take lane 0's after folding in the length
        h = __shfl_sync(0xffffffffu, h, 0) ^ ((unsigned long long)len * 0xD6E8FEB86659FD93ull);
        if (lane == 0) { hash[g] = h; index[g] = (int32_t)g; }
    }
}

__global__ void mark_heads_kernel(const uint8_t *__restrict__ codes, const int64_t *__restrict__ gene_off,
                                  const int32_t *__restrict__ gene_len,
                                  const unsigned long long *__restrict__ hash_sorted,
                                  const int32_t *__restrict__ index_sorted, int64_t n_genes,
                                  int32_t *__restrict__ head_pos)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_genes) return;
    bool head = true;
    if (k > 0 && hash_sorted[k] == hash_sorted[k - 1]) {
        const int a = index_sorted[k], b = index_sorted[k - 1];
        if (gene_len[a] == gene_len[b]) {
            // arenas are 16 B aligned and zero padded: compare whole 16 B words
            const uint4 *pa = reinterpret_cast<const uint4 *>(codes + gene_off[a]);
            const uint4 *pb = reinterpret_cast<const uint4 *>(codes + gene_off[b]);
            const int words = (gene_len[a] + 15) / 16;
            bool same = true;
            for (int w = 0; w < words && same; ++w) {
                const uint4 x = pa[w], y = pb[w];
                same = (x.x == y.x) && (x.y == y.y) && (x.z == y.z) && (x.w == y.w);
            }
            head = !same;
        }
    }
    head_pos[k] = head ? (int32_t)k : 0;
}

__global__ void assign_canon_kernel(const int32_t *__restrict__ index_sorted,
                                    const int32_t *__restrict__ head_scan, int64_t n_genes,
                                    int32_t *__restrict__ canon, unsigned long long *__restrict__ n_unique)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_genes) return;
    const int32_t head = head_scan[k];
    canon[index_sorted[k]] = index_sorted[head];   // stable sort: the head has the smallest gene index
    if (head == (int32_t)k) atomicAdd(n_unique, 1ull);
}

// Fills ctx->d_canon (int32 per gene) and ctx->n_unique_genes.
int pc_build_canonical_genes(pc_ctx *ctx)
{
    const int64_t n = ctx->n_genes;
    PC_TRY(devbuf_reserve(ctx, ctx->d_canon, sizeof(int32_t) * (size_t)std::max<int64_t>(n, 1)));
    ctx->n_unique_genes = 0;
    if (n == 0) return PC_OK;
    if (n >= (1ll << 31) - 1) return pc_fail(ctx, PC_EINVAL, "too many genes for canonical ids");
    DevBuf &hash_a = ctx->d_dedup[0], &hash_b = ctx->d_dedup[1], &idx_a = ctx->d_dedup[2],
           &idx_b = ctx->d_dedup[3], &head = ctx->d_dedup[4], &head_scan = ctx->d_dedup[5],
           &temp = ctx->d_dedup[6], &counter = ctx->d_dedup[7];
    int rc = PC_OK;
    do {
        if ((rc = devbuf_reserve(ctx, hash_a, sizeof(unsigned long long) * n))) break;
        if ((rc = devbuf_reserve(ctx, hash_b, sizeof(unsigned long long) * n))) break;
        if ((rc = devbuf_reserve(ctx, idx_a, sizeof(int32_t) * n))) break;
        if ((rc = devbuf_reserve(ctx, idx_b, sizeof(int32_t) * n))) break;
        if ((rc = devbuf_reserve(ctx, head, sizeof(int32_t) * n))) break;
        if ((rc = devbuf_reserve(ctx, head_scan, sizeof(int32_t) * n))) break;
        if ((rc = devbuf_reserve(ctx, counter, sizeof(unsigned long long)))) break;
        const int grid_w = (int)std::min<int64_t>((n * 32 + 255) / 256, (int64_t)ctx->sm_count * 8);
        const int grid_t = (int)((n + 255) / 256);
        gene_hash_kernel<<<grid_w, 256, 0, ctx->stream>>>(ctx->d_codes.as<uint8_t>(), ctx->d_gene_off.as<int64_t>(),
                                                          ctx->d_gene_len.as<int32_t>(), n,
                                                          hash_a.as<unsigned long long>(), idx_a.as<int32_t>());
        size_t sort_bytes = 0, scan_bytes = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, hash_a.as<unsigned long long>(),
                                        hash_b.as<unsigned long long>(), idx_a.as<int32_t>(), idx_b.as<int32_t>(),
                                        (int)n, 0, 64, ctx->stream);
        cub::DeviceScan::InclusiveScan(nullptr, scan_bytes, head.as<int32_t>(), head_scan.as<int32_t>(),
                                       cub::Max(), (int)n, ctx->stream);
        if ((rc = devbuf_reserve(ctx, temp, std::max(sort_bytes, scan_bytes)))) break;
        cub::DeviceRadixSort::SortPairs(temp.p, sort_bytes, hash_a.as<unsigned long long>(),
                                        hash_b.as<unsigned long long>(), idx_a.as<int32_t>(), idx_b.as<int32_t>(),
                                        (int)n, 0, 64, ctx->stream);
        mark_heads_kernel<<<grid_t, 256, 0, ctx->stream>>>(ctx->d_codes.as<uint8_t>(), ctx->d_gene_off.as<int64_t>(),
                                                           ctx->d_gene_len.as<int32_t>(),
                                                           hash_b.as<unsigned long long>(), idx_b.as<int32_t>(), n,
                                                           head.as<int32_t>());
        cub::DeviceScan::InclusiveScan(temp.p, scan_bytes, head.as<int32_t>(), head_scan.as<int32_t>(),
                                       cub::Max(), (int)n, ctx->stream);
        if (cudaMemsetAsync(counter.p, 0, sizeof(unsigned long long), ctx->stream) != cudaSuccess) {
            rc = pc_fail(ctx, PC_ECUDA, "memset failed");
            break;
        }
        assign_canon_kernel<<<grid_t, 256, 0, ctx->stream>>>(idx_b.as<int32_t>(), head_scan.as<int32_t>(), n,
                                                             ctx->d_canon.as<int32_t>(),
                                                             counter.as<unsigned long long>());
        unsigned long long uniq = 0;
        if (cudaMemcpyAsync(&uniq, counter.p, sizeof uniq, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
            cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
            rc = pc_fail(ctx, PC_ECUDA, "canonical gene ids failed: %s", cudaGetErrorString(cudaGetLastError()));
            break;
        }
        ctx->n_unique_genes = (int64_t)uniq;
        ctx->launches += 5;
    } while (0);
    return rc;
}
