// pc_api.cu -- the C ABI of libphamclust_b200 (include/phamclust_b200.h):
// context, input loading, host-side planning and the per-metric flows that
// stand in for matrix_de_novo (/root/reference/src/phamclust/matrix.py:432-497).
#include <algorithm>
#include <cctype>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <numeric>
#include <chrono>
#include <climits>
#include <thread>
#include <unistd.h>
#include <mutex>
#include <condition_variable>
#include <functional>

#include "pc_common.cuh"

int pc_nw_multichunk_warps(pc_ctx *ctx);

static thread_local std::string g_init_error;

static void dbg_mark(const char *what)
{
    static const bool on = getenv("PC_DEBUG") != nullptr;
    static auto last = std::chrono::steady_clock::now();
    if (!on) return;
    auto now = std::chrono::steady_clock::now();
    fprintf(stderr, "[pc] %-28s +%.2f ms\n", what, std::chrono::duration<double, std::milli>(now - last).count());
    last = now;
}

struct HostTimer {
    pc_ctx *ctx; int stage;
    std::chrono::steady_clock::time_point t0;
    HostTimer(pc_ctx *c, int s) : ctx(c), stage(s), t0(std::chrono::steady_clock::now()) {}
    ~HostTimer() {
        ctx->stage_ms[stage] += std::chrono::duration<double, std::milli>(
            std::chrono::steady_clock::now() - t0).count();
    }
};

int pc_fail(pc_ctx *ctx, int code, const char *fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf; else g_init_error = buf;
    return code;
}

int devbuf_reserve(pc_ctx *ctx, DevBuf &b, size_t bytes)
{
    if (bytes == 0) bytes = 16;
    if (b.cap >= bytes) return PC_OK;
    if (b.p) {
        // buffers may still be in use by queued work on the ctx stream
        PC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        cudaFree(b.p);
        b.p = nullptr;
        b.cap = 0;
    }
    cudaError_t e = cudaMalloc(&b.p, bytes);
    if (e != cudaSuccess) {
        b.p = nullptr;
        cudaGetLastError();
        return pc_fail(ctx, PC_ENOMEM, "cudaMalloc(%zu bytes) failed: %s", bytes,
                       cudaGetErrorString(e));
    }
    b.cap = bytes;
    return PC_OK;
}

void devbuf_free(DevBuf &b)
{
    if (b.p) cudaFree(b.p);
    b.p = nullptr;
    b.cap = 0;
}

// ---------------------------------------------------------------------------
// stage timers
// ---------------------------------------------------------------------------
static int take_event(pc_ctx *ctx, cudaEvent_t *ev)
{
    if (!ctx->event_pool.empty()) {
        *ev = ctx->event_pool.back();
        ctx->event_pool.pop_back();
        return PC_OK;
    }
    PC_CUDA(ctx, cudaEventCreate(ev));
    return PC_OK;
}

int pc_stage_begin(pc_ctx *ctx, int stage)
{
    StageTimer t;
    t.stage = stage;
    PC_TRY(take_event(ctx, &t.start));
    PC_TRY(take_event(ctx, &t.stop));
    PC_CUDA(ctx, cudaEventRecord(t.start, ctx->stream));
    ctx->timers.push_back(t);
    return PC_OK;
}

int pc_stage_end(pc_ctx *ctx)
{
    PC_CUDA(ctx, cudaEventRecord(ctx->timers.back().stop, ctx->stream));
    return PC_OK;
}

int pc_collect_timers(pc_ctx *ctx)
{
    if (ctx->timers.empty()) return PC_OK;
    PC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (auto &t : ctx->timers) {
        float ms = 0.f;
        PC_CUDA(ctx, cudaEventElapsedTime(&ms, t.start, t.stop));
        ctx->stage_ms[t.stage] += ms;
        ctx->event_pool.push_back(t.start);
        ctx->event_pool.push_back(t.stop);
    }
    ctx->timers.clear();
    for (cudaEvent_t ev : ctx->deferred_events) ctx->event_pool.push_back(ev);
    ctx->deferred_events.clear();
    return PC_OK;
}

// keep the number of pending events bounded on long runs
static int maybe_collect(pc_ctx *ctx)
{
    if (ctx->timers.size() > 512) return pc_collect_timers(ctx);
    return PC_OK;
}

// ---------------------------------------------------------------------------
// lifetime
// ---------------------------------------------------------------------------
extern "C" int pc_version(void) { return 100; }

extern "C" const char *pc_last_error(const pc_ctx *ctx)
{
    return ctx ? ctx->err.c_str() : g_init_error.c_str();
}

extern "C" int pc_init(int device_id, pc_ctx **out)
{
    if (!out) return pc_fail(nullptr, PC_EINVAL, "pc_init: out is NULL");
    *out = nullptr;
    int n_dev = 0;
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    if (e != cudaSuccess || n_dev == 0) {
        cudaGetLastError();
        return pc_fail(nullptr, PC_ENOGPU,
                       "no CUDA device (%s); libphamclust_b200 has no CPU fallback",
                       e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    }
    if (device_id < 0 || device_id >= n_dev)
        return pc_fail(nullptr, PC_EINVAL, "device %d out of range [0,%d)", device_id, n_dev);
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device_id) != cudaSuccess)
        return pc_fail(nullptr, PC_ECUDA, "cudaGetDeviceProperties failed");
    if (prop.major != 10)
        return pc_fail(nullptr, PC_ENOGPU,
                       "device %d (%s) is sm_%d%d; this library is built for sm_100a only",
                       device_id, prop.name, prop.major, prop.minor);
    if (cudaSetDevice(device_id) != cudaSuccess)
        return pc_fail(nullptr, PC_ECUDA, "cudaSetDevice(%d) failed", device_id);
    pc_ctx *ctx = new pc_ctx();
    ctx->device = device_id;
    ctx->sm_count = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete ctx;
        return pc_fail(nullptr, PC_ECUDA, "cudaStreamCreate failed");
    }
    ctx->stream = ctx->own_stream;
    if (const char *mode = getenv("PC_DP_MODE"))        // measurement aid: "classic" = never the LEX DP kernel
        ctx->dp_mode = strcmp(mode, "classic") == 0 ? PC_DP_CLASSIC : PC_DP_AUTO;
    *out = ctx;
    return PC_OK;
}

static void free_plan(AlignPlan &p)
{
    devbuf_free(p.ph_mem_off); devbuf_free(p.ph_unit_off);
    devbuf_free(p.mem_genome); devbuf_free(p.mem_first_gene); devbuf_free(p.mem_copies);
    devbuf_free(p.blk_counts); devbuf_free(p.cfg_info); devbuf_free(p.ht_keys); devbuf_free(p.ht_vals);
    p.part = p.n_parts = -1;
}

extern "C" void pc_destroy(pc_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    DevBuf *bufs[] = {&ctx->d_entry_genome, &ctx->d_pham_id, &ctx->d_pham_col, &ctx->d_copies, &ctx->d_len_sum,
                      &ctx->d_den_phams, &ctx->d_den_genes, &ctx->d_den_len, &ctx->d_bits,
                      &ctx->d_planes, &ctx->d_terms, &ctx->d_numer, &ctx->d_i8, &ctx->d_i8_occ, &ctx->d_i8_steps, &ctx->d_codes,
                      &ctx->d_p2_tiles, &ctx->d_p2_masks, &ctx->d_p2_keys, &ctx->d_p2_ids, &ctx->d_p2_temp, &ctx->d_p2_sync, &ctx->d_dest,
                      &ctx->d_resident, &ctx->d_res_edges, &ctx->d_res_idx, &ctx->d_res_vals,
                      &ctx->d_gene_off, &ctx->d_gene_len, &ctx->d_raw, &ctx->d_raw_off, &ctx->d_lut, &ctx->d_canon, &ctx->d_gene_hi, &ctx->d_subtab[0],
                      &ctx->d_subtab[1], &ctx->d_task_q, &ctx->d_task_r, &ctx->d_hitlen,
                      &ctx->d_score, &ctx->d_spill, &ctx->d_counters, &ctx->d_perm,
                      &ctx->d_acc, &ctx->d_out};
    for (DevBuf *b : bufs) devbuf_free(*b);
    for (DevBuf &b : ctx->d_dedup) devbuf_free(b);
    if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
    free_plan(ctx->plan);
    for (auto &t : ctx->timers) { cudaEventDestroy(t.start); cudaEventDestroy(t.stop); }
    for (auto ev : ctx->event_pool) cudaEventDestroy(ev);
    cudaStreamDestroy(ctx->own_stream);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    for (cudaStream_t st : ctx->dp_streams) if (st) cudaStreamDestroy(st);
    for (auto ev : ctx->deferred_events) cudaEventDestroy(ev);
    delete ctx;
}

#define PC_ENTER(ctx)                                                          \
    do {                                                                       \
        if (!(ctx)) return PC_EINVAL;                                          \
        PC_CUDA((ctx), cudaSetDevice((ctx)->device));                          \
    } while (0)

extern "C" int pc_set_stream(pc_ctx *ctx, void *cuda_stream, int external)
{
    PC_ENTER(ctx);
    PC_TRY(pc_collect_timers(ctx));
    PC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->stream = external ? (cudaStream_t)cuda_stream : ctx->own_stream;
    return PC_OK;
}

extern "C" int pc_set_engine(pc_ctx *ctx, int engine)
{
    PC_ENTER(ctx);
    if (engine < PC_ENGINE_AUTO || engine > PC_ENGINE_I8MMA_V1)
        return pc_fail(ctx, PC_EINVAL, "unknown engine %d", engine);
    ctx->engine = engine;
    return PC_OK;
}

extern "C" int pc_set_dp_mode(pc_ctx *ctx, int mode)
{
    PC_ENTER(ctx);
    if (mode != PC_DP_AUTO && mode != PC_DP_CLASSIC)
        return pc_fail(ctx, PC_EINVAL, "unknown DP mode %d", mode);
    ctx->dp_mode = mode;
    return PC_OK;
}

extern "C" int pc_sync(pc_ctx *ctx)
{
    PC_ENTER(ctx);
    PC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PC_OK;
}

template <typename T>
static int upload(pc_ctx *ctx, DevBuf &b, const T *src, size_t count)
{
    PC_TRY(devbuf_reserve(ctx, b, count * sizeof(T)));
    if (count)
        PC_CUDA(ctx, cudaMemcpyAsync(b.p, src, count * sizeof(T), cudaMemcpyHostToDevice,
                                     ctx->stream));
    ctx->h2d_bytes += (int64_t)(count * sizeof(T));
    return PC_OK;
}

// is this host buffer page-locked (a DMA engine can read it at full PCIe rate, asynchronously)?
static bool is_pinned(const void *p)
{
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return attr.type == cudaMemoryTypeHost;
}

// host threads for the O(nnz) passes over the caller's table (PC_HOST_THREADS overrides)
static int host_threads(int64_t work)
{
    if (const char *env = getenv("PC_HOST_THREADS")) {
        const int t = atoi(env);
        if (t >= 1) return std::min(t, 64);
    }
    if (work < 100000) return 1;
    unsigned hw = std::thread::hardware_concurrency();
    // one process per GPU: the ranks of a node share its cores (torchrun exports LOCAL_WORLD_SIZE)
    if (const char *lws = getenv("LOCAL_WORLD_SIZE")) {
        const int ranks = atoi(lws);
        if (ranks > 1) hw = std::max(1u, hw / (unsigned)ranks);
        return (int)std::max(1u, std::min(8u, hw));
    }
    return (int)std::max(1u, std::min(8u, hw / 2));
}

// A small persistent pool: the host passes over the table are short (fractions of a millisecond),
// and a step runs a dozen of them -- spawning and joining T threads for each cost more than the
// passes themselves (about 0.1 ms per pass).  One job at a time; a second caller (another context
// on another host thread) simply spawns its own threads as before.
namespace {
class HostPool {
public:
    static HostPool &get() { static HostPool pool; return pool; }

    template <typename F>
    bool run(int T, F &&body)
    {
        if (getpid() != pid_) return false;          // a forked child has no workers
        std::unique_lock<std::mutex> job_lock(job_mutex_, std::try_to_lock);
        if (!job_lock.owns_lock()) return false;
        {
            std::lock_guard<std::mutex> lk(m_);
            while ((int)workers_->size() < T - 1) {
                const int id = (int)workers_->size();
                workers_->emplace_back([this, id] { work(id); });
            }
            fn_ = [&body](int t) { body(t); };
            active_ = T - 1;
            pending_ = T - 1;
            ++generation_;
        }
        cv_.notify_all();
        body(0);
        std::unique_lock<std::mutex> lk(m_);
        done_.wait(lk, [this] { return pending_ == 0; });
        fn_ = nullptr;
        return true;
    }

private:
    HostPool() : pid_(getpid()) {}
    ~HostPool()
    {
        if (getpid() != pid_) return;               // forked child: the thread objects are not ours, leak them
        {
            std::lock_guard<std::mutex> lk(m_);
            stop_ = true;
        }
        cv_.notify_all();
        for (auto &w : *workers_) w.join();
        delete workers_;
    }

    void work(int id)
    {
        unsigned long long seen = 0;
        std::unique_lock<std::mutex> lk(m_);
        for (;;) {
            cv_.wait(lk, [&] { return stop_ || (generation_ != seen && id < active_); });
            if (stop_) return;
            seen = generation_;
            auto fn = fn_;
            lk.unlock();
            fn(id + 1);
            lk.lock();
            if (--pending_ == 0) done_.notify_one();
        }
    }

    std::mutex job_mutex_, m_;
    std::condition_variable cv_, done_;
    std::vector<std::thread> *workers_ = new std::vector<std::thread>();
    std::function<void(int)> fn_;
    unsigned long long generation_ = 0;
    int active_ = 0, pending_ = 0;
    bool stop_ = false;
    const pid_t pid_;
};
}  // namespace

template <typename F>
static void run_threads(int T, F &&body)
{
    if (T <= 1) {
        body(0);
        return;
    }
    static const bool use_pool = getenv("PC_HOST_POOL") == nullptr || atoi(getenv("PC_HOST_POOL")) != 0;
    if (use_pool && HostPool::get().run(T, body)) return;
    std::vector<std::thread> pool;
    pool.reserve(T - 1);
    for (int t = 1; t < T; ++t) pool.emplace_back([&body, t] { body(t); });
    body(0);
    for (auto &th : pool) th.join();
}

// ---------------------------------------------------------------------------
// pc_load_table
// ---------------------------------------------------------------------------
extern "C" int pc_load_table(pc_ctx *ctx, int64_t n_genomes, int64_t n_phams,
                             const int64_t *genome_ptr, const int32_t *pham_id,
                             const uint32_t *copies, const uint32_t *len_sum)
{
    PC_ENTER(ctx);
    HostTimer host_timer(ctx, PC_T_HOST_TABLE);
    if (n_genomes <= 0)   // matrix.py:451-452
        return pc_fail(ctx, PC_EINVAL, "need at least 1 genome to construct matrix de novo");
    if (n_phams < 0 || !genome_ptr)
        return pc_fail(ctx, PC_EINVAL, "pc_load_table: bad arguments");
    if (n_genomes > 2000000)
        return pc_fail(ctx, PC_EINVAL, "too many genomes (%lld)", (long long)n_genomes);
    if (genome_ptr[0] != 0) return pc_fail(ctx, PC_EINVAL, "genome_ptr[0] must be 0");
    const int64_t nnz = genome_ptr[n_genomes];
    if (nnz < 0 || (nnz > 0 && (!pham_id || !copies || !len_sum)))
        return pc_fail(ctx, PC_EINVAL, "pc_load_table: missing entry arrays");

    for (int64_t g = 0; g < n_genomes; ++g)
        if (genome_ptr[g + 1] < genome_ptr[g] || genome_ptr[g + 1] > nnz)
            return pc_fail(ctx, PC_EINVAL, "genome_ptr is not monotone at %lld", (long long)g);

    // the entry arrays go to the device as they are, while the host validates them
    PC_TRY(pc_stage_begin(ctx, PC_T_H2D));
    PC_TRY(upload(ctx, ctx->d_pham_id, pham_id, nnz));
    PC_TRY(upload(ctx, ctx->d_copies, copies, nnz));
    PC_TRY(upload(ctx, ctx->d_len_sum, len_sum, nnz));
    PC_TRY(pc_stage_end(ctx));

    // validate + derive per-genome denominators, CSC counts and the first entry of every
    // pham, on a few host threads over contiguous genome ranges of equal entry counts
    // (every buffer lives in the ctx: a reload of a table of the same size touches no fresh page)
    ctx->have_table = false;
    for (auto &d : ctx->h_den) d.resize(n_genomes);
    std::vector<int32_t> &den_phams = ctx->h_den[0], &den_genes = ctx->h_den[1], &den_len = ctx->h_den[2];
    std::vector<int32_t> &entry_genome = ctx->h_entry_genome;
    std::vector<int64_t> &entry_first_gene = ctx->h_entry_first_gene;
    std::vector<int64_t> &pham_count = ctx->h_pham_ptr;
    std::vector<int64_t> &first_entry = ctx->h_first_entry;
    entry_genome.resize(nnz);
    entry_first_gene.resize(nnz + 1);
    pham_count.assign(n_phams + 1, 0);
    first_entry.assign(n_phams, INT64_MAX);
    ctx->h_genome_ptr.assign(genome_ptr, genome_ptr + n_genomes + 1);
    ctx->h_pham_id.resize(nnz);
    ctx->h_copies.resize(nnz);
    ctx->h_len_sum.resize(nnz);
    const int T = host_threads(nnz);
    std::vector<int64_t> cut(T + 1, n_genomes);       // thread t owns genomes [cut[t], cut[t+1])
    cut[0] = 0;
    for (int t = 1; t < T; ++t)
        cut[t] = std::lower_bound(genome_ptr, genome_ptr + n_genomes, nnz * t / T) - genome_ptr;
    struct Part {
        int64_t genes = 0, err_at = INT64_MAX;
        uint32_t max_copies = 0, max_len = 0;
        std::string err;
        std::vector<int32_t> first_seen;          // phams whose first entry lies in this range, in order
    };
    std::vector<Part> part(T);
    // duplicate check: stamp[p] = epoch + genome that last listed the pham; a new epoch per load
    // instead of clearing the per-thread arrays
    if ((int64_t)ctx->stamp_epoch + 2 * (n_genomes + 1) >= INT32_MAX) {
        ctx->stamp_epoch = 0;
        for (auto &v : ctx->h_stamp) std::fill(v.begin(), v.end(), 0);
    }
    const int32_t epoch = ctx->stamp_epoch + 1;
    ctx->stamp_epoch += (int32_t)n_genomes + 1;
    if ((int)ctx->h_stamp.size() < T) ctx->h_stamp.resize(T);
    run_threads(T, [&](int t) {
        Part &me = part[t];
        std::vector<int32_t> &stamp = ctx->h_stamp[t];
        if ((int64_t)stamp.size() < n_phams) stamp.resize(n_phams, 0);
        auto fail = [&](int64_t at, const char *fmt, auto... args) {
            char buf[160];
            snprintf(buf, sizeof buf, fmt, args...);
            me.err_at = at;                        // the lowest entry index wins across threads
            me.err = buf;
        };
        for (int64_t g = cut[t]; g < cut[t + 1]; ++g) {
            const int64_t b = genome_ptr[g], e = genome_ptr[g + 1];
            int64_t genes = 0, len = 0;
            for (int64_t k = b; k < e; ++k) {
                const int32_t p = pham_id[k];
                if (p < 0 || p >= n_phams)
                    return fail(k, "pham id %d out of range at entry %lld", p, (long long)k);
                if (stamp[p] == epoch + (int32_t)g)
                    return fail(k, "pham %d listed twice for genome %lld", p, (long long)g);
                stamp[p] = epoch + (int32_t)g;
                if (copies[k] == 0) return fail(k, "entry %lld has 0 copies", (long long)k);
                entry_genome[k] = (int32_t)g;
                genes += copies[k];
                len += len_sum[k];
                me.max_copies = std::max(me.max_copies, copies[k]);
                me.max_len = std::max(me.max_len, len_sum[k]);
                __atomic_fetch_add(&pham_count[p + 1], (int64_t)1, __ATOMIC_RELAXED);
                int64_t cur = __atomic_load_n(&first_entry[p], __ATOMIC_RELAXED);
                while (k < cur && !__atomic_compare_exchange_n(&first_entry[p], &cur, k, true, __ATOMIC_RELAXED,
                                                               __ATOMIC_RELAXED)) {
                }
            }
            if (genes > (1 << 30) || len > (1 << 30))
                return fail(b, "genome %lld too large for int32 numerators", (long long)g);
            den_phams[g] = (int32_t)(e - b);
            den_genes[g] = (int32_t)genes;
            den_len[g] = (int32_t)len;
            me.genes += genes;
        }
        const int64_t b = genome_ptr[cut[t]], e = genome_ptr[cut[t + 1]];
        if (e > b) {
            memcpy(ctx->h_pham_id.data() + b, pham_id + b, (size_t)(e - b) * sizeof(int32_t));
            memcpy(ctx->h_copies.data() + b, copies + b, (size_t)(e - b) * sizeof(uint32_t));
            memcpy(ctx->h_len_sum.data() + b, len_sum + b, (size_t)(e - b) * sizeof(uint32_t));
        }
    });
    uint32_t max_copies = 0, max_len = 0;
    int64_t gene_cursor = 0;
    {
        const Part *bad = nullptr;
        for (const Part &q : part) {
            if (q.err_at != INT64_MAX && (!bad || q.err_at < bad->err_at)) bad = &q;
            max_copies = std::max(max_copies, q.max_copies);
            max_len = std::max(max_len, q.max_len);
            gene_cursor += q.genes;
        }
        if (bad) {
            cudaStreamSynchronize(ctx->stream);        // the uploads read the caller's buffers
            return pc_fail(ctx, PC_EINVAL, "%s", bad->err.c_str());
        }
    }
    if (gene_cursor >= (1ll << 31)) {
        cudaStreamSynchronize(ctx->stream);
        return pc_fail(ctx, PC_EINVAL, "too many genes (%lld)", (long long)gene_cursor);
    }
    // second pass: first gene of every entry, and the phams in order of their first entry
    std::vector<int64_t> gene_base(T + 1, 0);
    for (int t = 0; t < T; ++t) gene_base[t + 1] = gene_base[t] + part[t].genes;
    run_threads(T, [&](int t) {
        int64_t cursor = gene_base[t];
        Part &me = part[t];
        for (int64_t k = genome_ptr[cut[t]]; k < genome_ptr[cut[t + 1]]; ++k) {
            entry_first_gene[k] = cursor;
            cursor += copies[k];
            if (first_entry[pham_id[k]] == k) me.first_seen.push_back(pham_id[k]);
        }
    });
    entry_first_gene[nnz] = gene_cursor;

    // per-pham member counts -> CSC pointers; the member lists themselves are only
    // needed by the alignment planner and are built there on demand
    for (int64_t p = 0; p < n_phams; ++p) pham_count[p + 1] += pham_count[p];

    ctx->N = n_genomes; ctx->P = n_phams; ctx->nnz = nnz;
    ctx->deal_parts = -1;
    ctx->resident_n = -1;
    ctx->bits_ready = false;
    ctx->max_copies = max_copies; ctx->max_len_sum = max_len;
    ctx->n_genes_total = gene_cursor;
    ctx->have_seqs = false;
    ctx->planes_kind = -1; ctx->i8_kind = -1;
    ctx->plan.part = ctx->plan.n_parts = -1;
    // operand columns: a pham held by a single genome can never be shared, so it
    // gets no column (exact; orphams are common in phage tables).  Columns are
    // handed out in order of the first genome holding the pham (the result does
    // not depend on the column order): the phams of a run of related genomes land
    // in neighbouring K blocks whatever their ids, which is what the block-sparse
    // K schedule of the int8 engine feeds on.
    std::vector<int32_t> &pham_col = ctx->h_pham_col;
    pham_col.assign(std::max<int64_t>(n_phams, 1), -1);
    int64_t p_eff = 0;
    for (const auto &q : part)
        for (const int32_t p : q.first_seen)
            if (ctx->h_pham_ptr[p + 1] - ctx->h_pham_ptr[p] >= 2) pham_col[p] = (int32_t)p_eff++;
    ctx->P_eff = p_eff;
    const int64_t words = (p_eff + 31) / 32;
    ctx->W = std::max<int64_t>(32, (words + 31) / 32 * 32);

    PC_TRY(pc_stage_begin(ctx, PC_T_H2D));
    PC_TRY(upload(ctx, ctx->d_entry_genome, ctx->h_entry_genome.data(), nnz));
    PC_TRY(upload(ctx, ctx->d_pham_col, pham_col.data(), pham_col.size()));
    PC_TRY(upload(ctx, ctx->d_den_phams, den_phams.data(), n_genomes));
    PC_TRY(upload(ctx, ctx->d_den_genes, den_genes.data(), n_genomes));
    PC_TRY(upload(ctx, ctx->d_den_len, den_len.data(), n_genomes));
    PC_TRY(pc_stage_end(ctx));
    // the copies read the caller's buffers and scratch the next load rewrites: finish them
    PC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->have_table = true;
    return PC_OK;
}

// ---------------------------------------------------------------------------
// pc_load_sequences
// ---------------------------------------------------------------------------
__global__ void byte_presence_kernel(const uint8_t *__restrict__ raw, int64_t total,
                                     unsigned int *__restrict__ present)
{
    __shared__ unsigned int s_flag[256];
    s_flag[threadIdx.x] = 0;
    __syncthreads();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += stride)
        s_flag[raw[k]] = 1;                    // benign race: every writer stores 1
    __syncthreads();
    if (s_flag[threadIdx.x]) present[threadIdx.x] = 1;
}

static void build_subtab(const pc_ctx *ctx, int ppos, std::vector<int2> &tab)
{
    tab.assign(32 * 32, make_int2(-4, 1));
    for (int a = 0; a < 32; ++a)
        for (int b = 0; b < 32; ++b) {
            const int ia = a < ctx->n_codes ? ctx->code_matrix_index[a] : 23;
            const int ib = b < ctx->n_codes ? ctx->code_matrix_index[b] : 23;
            const int s = PC_BLOSUM62[ia][ib];
            // '|' when the characters match, '+' when the score is positive
            // (Appendix A traceback); codes are per distinct (case-folded) character
            const bool hit = (a == b) || (ppos && s > 0);
            tab[a * 32 + b] = make_int2(s, 1 + (hit ? (1 << 16) : 0));
        }
}

static void deal_phams(pc_ctx *ctx, int n_parts);

// part / n_parts: ship, encode and canonicalise only the genes of the phams partition `part`
// aligns (deal_phams); n_parts == 1 ships everything straight from the caller's arena.
static int load_sequences_impl(pc_ctx *ctx, int64_t n_genes, const int64_t *seq_off,
                               const uint8_t *residues, int my_part, int n_parts)
{
    HostTimer host_timer(ctx, PC_T_HOST_SEQS);
    if (!ctx->have_table) return pc_fail(ctx, PC_ESTATE, "pc_load_sequences needs pc_load_table first");
    if (n_parts < 1 || my_part < 0 || my_part >= n_parts)
        return pc_fail(ctx, PC_EINVAL, "pc_load_sequences_part: bad partition %d/%d", my_part, n_parts);
    const bool partial = n_parts > 1;
    if (n_genes != ctx->n_genes_total)
        return pc_fail(ctx, PC_EINVAL, "got %lld genes but the table's copies sum to %lld",
                       (long long)n_genes, (long long)ctx->n_genes_total);
    if (!seq_off || seq_off[0] != 0) return pc_fail(ctx, PC_EINVAL, "seq_off[0] must be 0");
    const int64_t total = seq_off[n_genes];
    if (total < 0) return pc_fail(ctx, PC_EINVAL, "seq_off is not monotone (negative total)");
    if (total > 0 && !residues) return pc_fail(ctx, PC_EINVAL, "residues is NULL");

    dbg_mark("seqs: enter");
    std::vector<int32_t> &gene_len = ctx->h_gene_len;
    std::vector<int64_t> &gene_off = ctx->h_gene_off;
    ctx->have_seqs = false;
    gene_len.resize(n_genes);
    gene_off.resize(n_genes + 1);
    struct SeqPart {
        int64_t padded = 0, err_at = INT64_MAX, raw = 0, sum_err_at = INT64_MAX;
        int32_t max_len = 0;
        std::string err, sum_err;
    };
    // a gene's length, or -1 with the reason in `me`
    auto checked_len = [&](int64_t g, SeqPart &me) -> int64_t {
        const int64_t len = seq_off[g + 1] - seq_off[g];
        char buf[200];
        if (len < 0) {
            snprintf(buf, sizeof buf, "seq_off is not monotone at gene %lld", (long long)g);
            me.err_at = g; me.err = buf;
            return -1;
        }
        if (len > 32767) {
            snprintf(buf, sizeof buf,
                     "gene %lld has %lld residues; the alignment counters hold at most 32767 "
                     "(parasail's 16-bit scores are undefined there anyway)", (long long)g, (long long)len);
            me.err_at = g; me.err = buf;
            return -1;
        }
        return len;
    };
    auto len_sum_error = [&](int64_t e, int64_t sum, SeqPart &me) {
        char buf[160];
        snprintf(buf, sizeof buf, "len_sum[%lld]=%u but its genes sum to %lld", (long long)e, ctx->h_len_sum[e], (long long)sum);
        me.sum_err_at = e; me.sum_err = buf;
    };
    // the error a serial scan would meet first: a bad offset / length before any len_sum mismatch
    auto first_error = [&](const std::vector<SeqPart> &parts) -> const std::string * {
        const SeqPart *bad = nullptr;
        for (const SeqPart &sp : parts)
            if (sp.err_at != INT64_MAX && (!bad || sp.err_at < bad->err_at)) bad = &sp;
        if (bad) return &bad->err;
        for (const SeqPart &sp : parts)
            if (sp.sum_err_at != INT64_MAX && (!bad || sp.sum_err_at < bad->sum_err_at)) bad = &sp;
        return bad ? &bad->sum_err : nullptr;
    };
    int64_t cursor = 0;
    if (partial) {
        // This rank needs the genes of the entries whose pham it aligns (deal_phams).  Two passes
        // over the ENTRIES (an entry's genes are consecutive): lengths, checks and sizes, then --
        // with every thread's share of the arena and of the staging known -- offsets and the
        // gather of the owned residues into page-locked staging (1 / n_parts of the arena); the
        // device sees length 0 for every other gene.
        // Gathering costs the host about as much per byte as 7 bytes cost the DMA engine (one cache
        // miss per gene, ~1.2 GB/s per thread against ~24 GB/s per PCIe link, measured with
        // tools/probe_partial_load.py): with fewer than 8 partitions and a page-locked caller buffer
        // it is cheaper to ship the whole arena asynchronously and let the device skip the genes of
        // other ranks (their length is 0 there) -- every GPU has its own link.  PC_SHIP_ALL=0/1 forces.
        const int ship_env = getenv("PC_SHIP_ALL") ? atoi(getenv("PC_SHIP_ALL")) : -1;
        const bool ship_all = total > 0 && (ship_env >= 0 ? ship_env != 0 : (n_parts < 8 && is_pinned(residues)));
        if (ship_all) {
            PC_TRY(pc_stage_begin(ctx, PC_T_H2D));
            PC_TRY(upload(ctx, ctx->d_raw, residues, (size_t)total));
            PC_TRY(upload(ctx, ctx->d_raw_off, seq_off, (size_t)n_genes + 1));
            PC_TRY(pc_stage_end(ctx));
        }
        deal_phams(ctx, n_parts);
        dbg_mark("seqs: phams dealt");
        const int T = host_threads(ctx->nnz);
        std::vector<SeqPart> part(T);
        auto elo_of = [&](int t) { return ctx->nnz * t / T; };
        ctx->h_gene_len_dev.resize((size_t)n_genes);       // every element is written below
        ctx->h_raw_off.resize((size_t)n_genes + 1);
        run_threads(T, [&](int t) {
            SeqPart &me = part[t];
            for (int64_t e = elo_of(t); e < elo_of(t + 1); ++e) {
                const bool own = ctx->h_pham_part[ctx->h_pham_id[e]] == my_part;
                int64_t sum = 0;
                for (int64_t g = ctx->h_entry_first_gene[e]; g < ctx->h_entry_first_gene[e + 1]; ++g) {
                    const int64_t len = checked_len(g, me);
                    if (len < 0) return;
                    gene_len[g] = (int32_t)len;
                    sum += len;
                    if (!own) continue;
                    me.max_len = std::max(me.max_len, (int32_t)len);
                    me.padded += (len + 15) / 16 * 16;
                    me.raw += len;
                }
                // the table's len_sum must describe these genes (metrics.py:136-147 sums them)
                if (sum != (int64_t)ctx->h_len_sum[e] && me.sum_err_at == INT64_MAX) len_sum_error(e, sum, me);
            }
        });
        dbg_mark("seqs: lengths");
        if (const std::string *msg = first_error(part)) {
            if (ship_all) cudaStreamSynchronize(ctx->stream);      // the uploads read the caller's buffers
            return pc_fail(ctx, PC_EINVAL, "%s", msg->c_str());
        }
        std::vector<int64_t> base(T + 1, 0), raw_base(T + 1, 0);
        int32_t max_len = 0;
        for (int t = 0; t < T; ++t) {
            base[t + 1] = base[t] + part[t].padded;
            raw_base[t + 1] = raw_base[t] + part[t].raw;
            max_len = std::max(max_len, part[t].max_len);
        }
        cursor = base[T];
        ctx->max_gene_len = max_len;
        const size_t need = ship_all ? 0 : (size_t)raw_base[T] + 16;
        if (ctx->h_stage_cap < need) {
            if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
            ctx->h_stage = nullptr;
            ctx->h_stage_cap = 0;
            const size_t cap = need + need / 4;
            if (cudaHostAlloc(&ctx->h_stage, cap, cudaHostAllocDefault) != cudaSuccess) {
                cudaGetLastError();
                return pc_fail(ctx, PC_ENOMEM, "no page-locked staging for %zu residue bytes", cap);
            }
            ctx->h_stage_cap = cap;
        }
        run_threads(T, [&](int t) {
            int64_t at = base[t], raw_at = raw_base[t];
            uint8_t *stage = (uint8_t *)ctx->h_stage;
            const int64_t e_end = elo_of(t + 1);
            constexpr int64_t AHEAD = 12;                  // entries; one in n_parts is owned
            for (int64_t e = elo_of(t); e < e_end; ++e) {
                // the owned genes sit ~n_parts genes apart in the caller's arena: every copy would
                // start with a cache miss, so ask for the lines of an owned entry a few entries early
                if (!ship_all && e + AHEAD < e_end && ctx->h_pham_part[ctx->h_pham_id[e + AHEAD]] == my_part) {
                    const uint8_t *src = residues + seq_off[ctx->h_entry_first_gene[e + AHEAD]];
                    for (int64_t o = 0; o < (int64_t)ctx->h_len_sum[e + AHEAD]; o += 64) __builtin_prefetch(src + o, 0, 0);
                }
                const bool own = ctx->h_pham_part[ctx->h_pham_id[e]] == my_part;
                for (int64_t g = ctx->h_entry_first_gene[e]; g < ctx->h_entry_first_gene[e + 1]; ++g) {
                    gene_off[g] = at;
                    ctx->h_raw_off[g] = raw_at;
                    if (!own) {
                        ctx->h_gene_len_dev[g] = 0;
                        continue;
                    }
                    const int32_t len = gene_len[g];
                    ctx->h_gene_len_dev[g] = len;
                    if (!ship_all) memcpy(stage + raw_at, residues + seq_off[g], (size_t)len);
                    raw_at += len;
                    at += ((int64_t)len + 15) / 16 * 16;
                }
            }
        });
        gene_off[n_genes] = cursor;
        ctx->h_raw_off[n_genes] = ship_all ? total : raw_base[T];
        dbg_mark("seqs: offsets + staging");
        if (!ship_all) {
            PC_TRY(pc_stage_begin(ctx, PC_T_H2D));
            PC_TRY(upload(ctx, ctx->d_raw, (const uint8_t *)ctx->h_stage, (size_t)raw_base[T]));
            PC_TRY(upload(ctx, ctx->d_raw_off, ctx->h_raw_off.data(), (size_t)n_genes + 1));
            PC_TRY(pc_stage_end(ctx));
        }
    } else {
        // the residues go first: the copy runs while the host validates the offsets
        PC_TRY(pc_stage_begin(ctx, PC_T_H2D));
        PC_TRY(upload(ctx, ctx->d_raw, residues, (size_t)total));
        PC_TRY(upload(ctx, ctx->d_raw_off, seq_off, (size_t)n_genes + 1));
        PC_TRY(pc_stage_end(ctx));
        // gene lengths and their 16 B aligned offsets in the code arena, on a few host threads
        const int T = host_threads(n_genes);
        std::vector<SeqPart> part(T);
        auto lo_of = [&](int t) { return n_genes * t / T; };
        auto elo_of = [&](int t) { return ctx->nnz * t / T; };
        run_threads(T, [&](int t) {
            SeqPart &me = part[t];
            for (int64_t g = lo_of(t); g < lo_of(t + 1); ++g) {
                const int64_t len = checked_len(g, me);
                if (len < 0) return;
                gene_len[g] = (int32_t)len;
                me.max_len = std::max(me.max_len, (int32_t)len);
                me.padded += (len + 15) / 16 * 16;
            }
        });
        dbg_mark("seqs: lengths");
        if (const std::string *msg = first_error(part)) {
            cudaStreamSynchronize(ctx->stream);        // the uploads read the caller's buffers
            return pc_fail(ctx, PC_EINVAL, "%s", msg->c_str());
        }
        std::vector<int64_t> base(T + 1, 0);
        int32_t max_len = 0;
        for (int t = 0; t < T; ++t) {
            base[t + 1] = base[t] + part[t].padded;
            max_len = std::max(max_len, part[t].max_len);
        }
        cursor = base[T];
        ctx->max_gene_len = max_len;
        // second pass: arena offsets; and the table's len_sum must describe these genes
        // (metrics.py:136-147 sums them)
        run_threads(T, [&](int t) {
            int64_t at = base[t];
            for (int64_t g = lo_of(t); g < lo_of(t + 1); ++g) {
                gene_off[g] = at;
                at += ((int64_t)gene_len[g] + 15) / 16 * 16;
            }
            for (int64_t e = elo_of(t); e < elo_of(t + 1); ++e) {
                int64_t sum = 0;
                for (int64_t g = ctx->h_entry_first_gene[e]; g < ctx->h_entry_first_gene[e + 1]; ++g)
                    sum += gene_len[g];
                if (sum != (int64_t)ctx->h_len_sum[e]) {
                    len_sum_error(e, sum, part[t]);
                    return;
                }
            }
        });
        gene_off[n_genes] = cursor;
        dbg_mark("seqs: offsets");
        if (const std::string *msg = first_error(part)) {
            cudaStreamSynchronize(ctx->stream);
            return pc_fail(ctx, PC_EINVAL, "%s", msg->c_str());
        }
    }
    const int64_t raw_total = partial ? ctx->h_raw_off[n_genes] : total;

    dbg_mark("seqs: lengths validated");
    // residue codes: one per distinct character after case folding (U5); the
    // 24 matrix letters keep their matrix index, anything else maps to '*' (U6)
    // which byte values occur: the residues are shipped first and scanned on the
    // device (a host scan of the arena cost 11 ms of a 16 ms call at 40 MB)
    unsigned int present[256];
    {
        PC_TRY(devbuf_reserve(ctx, ctx->d_lut, 256 * sizeof(unsigned int)));
        PC_CUDA(ctx, cudaMemsetAsync(ctx->d_lut.p, 0, 256 * sizeof(unsigned int), ctx->stream));
        if (raw_total > 0) {
            const int grid = (int)std::min<int64_t>((raw_total / 16 + 255) / 256 + 1, (int64_t)ctx->sm_count * 8);
            byte_presence_kernel<<<grid, 256, 0, ctx->stream>>>(ctx->d_raw.as<uint8_t>(), raw_total,
                                                                ctx->d_lut.as<unsigned int>());
            PC_CUDA(ctx, cudaGetLastError());
            ctx->launches++;
        }
        PC_CUDA(ctx, cudaMemcpyAsync(present, ctx->d_lut.p, sizeof present, cudaMemcpyDeviceToHost, ctx->stream));
        PC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    for (int c = 0; c < 256; ++c) ctx->code_of_byte[c] = -1;
    ctx->n_codes = 24;
    for (int k = 0; k < 24; ++k) {
        ctx->code_matrix_index[k] = k;
        ctx->code_of_byte[(unsigned char)PC_ALPHABET[k]] = k;
#if !PC_CASE_SENSITIVE
        ctx->code_of_byte[(unsigned char)tolower((unsigned char)PC_ALPHABET[k])] = k;
#endif
    }
    for (int c = 0; c < 256; ++c) {
        if (!present[c] || ctx->code_of_byte[c] >= 0) continue;
#if !PC_CASE_SENSITIVE
        const int folded = toupper(c);
        if (folded != c && ctx->code_of_byte[folded] >= 0 && folded < 256) {
            ctx->code_of_byte[c] = ctx->code_of_byte[folded];
            continue;
        }
#else
        const int folded = c;
#endif
        if (ctx->n_codes >= 32)
            return pc_fail(ctx, PC_EINVAL,
                           "more than 32 distinct residue symbols in the input (byte 0x%02x)", c);
        int midx = 23;
        const char *pos = strchr(PC_ALPHABET, toupper(c));
        if (pos && toupper(c) != 0) midx = (int)(pos - PC_ALPHABET);
        ctx->code_matrix_index[ctx->n_codes] = midx;
        ctx->code_of_byte[c] = ctx->n_codes;
        if (folded >= 0 && folded < 256 && ctx->code_of_byte[folded] < 0)
            ctx->code_of_byte[folded] = ctx->n_codes;
        ctx->n_codes++;
    }
    dbg_mark("seqs: alphabet scanned");
    uint8_t lut[256];
    for (int c = 0; c < 256; ++c) lut[c] = (uint8_t)(ctx->code_of_byte[c] < 0 ? 23 : ctx->code_of_byte[c]);

    ctx->n_genes = n_genes;

    DevBuf &d_raw_off = ctx->d_raw_off, &d_lut = ctx->d_lut;
    std::vector<int2> tabs[2];           // alive until the sync below
    int rc = PC_OK;
    do {
        if ((rc = pc_stage_begin(ctx, PC_T_H2D))) break;
        if ((rc = upload(ctx, ctx->d_gene_off, gene_off.data(), (size_t)n_genes + 1))) break;
        if ((rc = upload(ctx, ctx->d_gene_len, partial ? ctx->h_gene_len_dev.data() : ctx->h_gene_len.data(),
                         (size_t)n_genes))) break;
        if ((rc = upload(ctx, d_lut, lut, 256))) break;
        for (int ppos = 0; ppos < 2 && rc == PC_OK; ++ppos) {
            build_subtab(ctx, ppos, tabs[ppos]);
            rc = upload(ctx, ctx->d_subtab[ppos], tabs[ppos].data(), tabs[ppos].size());
        }
        if (rc) break;
        if ((rc = pc_stage_end(ctx))) break;
        dbg_mark("seqs: uploads queued");
        // + tail padding: lanes read whole K-column words past the last gene
        if ((rc = devbuf_reserve(ctx, ctx->d_codes, (size_t)cursor + 2048))) break;
        if (cudaMemsetAsync(ctx->d_codes.p, 0, (size_t)cursor + 2048, ctx->stream) != cudaSuccess) {
            rc = pc_fail(ctx, PC_ECUDA, "memset of the code arena failed");
            break;
        }
        if ((rc = pc_stage_begin(ctx, PC_T_PACK))) break;
        if ((rc = pc_launch_encode(ctx, ctx->d_raw.as<uint8_t>(), d_raw_off.as<int64_t>(),
                                   ctx->d_gene_off.as<int64_t>(), ctx->d_gene_len.as<int32_t>(),
                                   d_lut.as<uint8_t>(), n_genes, ctx->d_codes.as<uint8_t>()))) break;
        if ((rc = pc_stage_end(ctx))) break;
        dbg_mark("seqs: encode queued");
        if (cudaStreamSynchronize(ctx->stream) != cudaSuccess)
            rc = pc_fail(ctx, PC_ECUDA, "encode kernel failed: %s", cudaGetErrorString(cudaGetLastError()));
    } while (0);
    if (rc) return rc;
    dbg_mark("seqs: synced");
    PC_TRY(pc_stage_begin(ctx, PC_T_PACK));
    PC_TRY(pc_build_canonical_genes(ctx));
    ctx->h_gene_hi.clear();
    if (ctx->n_codes <= 24) {
        // per-residue best substitution score, >= 0 (the 24-code table only)
        int rowmax[32];
        for (int a = 0; a < 32; ++a) {
            int best = 0;
            for (int b = 0; b < 24; ++b)
                best = std::max(best, (int)PC_BLOSUM62[ctx->code_matrix_index[a < 24 ? a : 23]][ctx->code_matrix_index[b]]);
            rowmax[a] = best;
        }
        PC_TRY(devbuf_reserve(ctx, ctx->d_gene_hi, (size_t)std::max<int64_t>(n_genes, 1) * sizeof(int32_t)));
        PC_TRY(pc_launch_gene_bound(ctx, rowmax, ctx->d_gene_hi.as<int32_t>()));
    }
    PC_TRY(pc_stage_end(ctx));
    dbg_mark("seqs: canonical ids");
    ctx->plan.part = ctx->plan.n_parts = -1;
    ctx->lexc_off = false;
    ctx->lexc_last_tasks = 0;
    ctx->seq_part = my_part;
    ctx->seq_parts = n_parts;
    ctx->have_seqs = true;
    return PC_OK;
}

extern "C" int pc_load_sequences(pc_ctx *ctx, int64_t n_genes, const int64_t *seq_off,
                                 const uint8_t *residues)
{
    PC_ENTER(ctx);
    return load_sequences_impl(ctx, n_genes, seq_off, residues, 0, 1);
}

extern "C" int pc_load_sequences_part(pc_ctx *ctx, int64_t n_genes, const int64_t *seq_off,
                                      const uint8_t *residues, int part, int n_parts)
{
    PC_ENTER(ctx);
    return load_sequences_impl(ctx, n_genes, seq_off, residues, part, n_parts);
}

// ---------------------------------------------------------------------------
// contraction operands
// ---------------------------------------------------------------------------
static int bit_length(uint32_t v)
{
    int n = 0;
    while (v) { ++n; v >>= 1; }
    return n;
}

// Build the term list of `metric` on the device (packing weight planes on demand).
static int prepare_terms(pc_ctx *ctx, int metric, int *n_terms_out)
{
    std::vector<PopcTerm> terms;
    if (!ctx->bits_ready) {          // presence bit rows are only packed when this engine runs
        PC_TRY(devbuf_reserve(ctx, ctx->d_bits, (size_t)ctx->N * ctx->W * sizeof(uint32_t)));
        PC_TRY(pc_stage_begin(ctx, PC_T_PACK));
        PC_TRY(pc_launch_pack_bits(ctx, 0, 0, ctx->d_bits.as<uint32_t>()));
        PC_TRY(pc_stage_end(ctx));
        ctx->bits_ready = true;
    }
    const uint32_t *bits = ctx->d_bits.as<uint32_t>();
    if (metric == PC_GCS || metric == PC_JC) {
        terms.push_back({bits, bits, 1});
    } else {
        const int kind = metric == PC_POCP ? 1 : 2;
        const int planes = bit_length(kind == 1 ? ctx->max_copies : ctx->max_len_sum);
        const size_t plane_words = (size_t)ctx->N * ctx->W;
        if (ctx->planes_kind != kind) {
            PC_TRY(devbuf_reserve(ctx, ctx->d_planes, plane_words * sizeof(uint32_t) * std::max(planes, 1)));
            PC_TRY(pc_stage_begin(ctx, PC_T_PACK));
            for (int b = 0; b < planes; ++b)
                PC_TRY(pc_launch_pack_bits(ctx, kind, b, ctx->d_planes.as<uint32_t>() + b * plane_words));
            PC_TRY(pc_stage_end(ctx));
            ctx->planes_kind = kind;
            ctx->planes_count = planes;
        }
        for (int b = 0; b < planes; ++b) {
            const uint32_t *pl = ctx->d_planes.as<uint32_t>() + b * plane_words;
            terms.push_back({pl, bits, 1 << b});   // source's conserved part
            terms.push_back({bits, pl, 1 << b});   // target's conserved part
        }
    }
    *n_terms_out = (int)terms.size();
    PC_TRY(devbuf_reserve(ctx, ctx->d_terms, sizeof(PopcTerm) * std::max<size_t>(terms.size(), 1)));
    if (!terms.empty()) {
        PC_CUDA(ctx, cudaMemcpyAsync(ctx->d_terms.p, terms.data(), sizeof(PopcTerm) * terms.size(),
                                     cudaMemcpyHostToDevice, ctx->stream));
        PC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));   // `terms` is a local
    }
    return PC_OK;
}

static const int32_t *den_of(pc_ctx *ctx, int metric)
{
    switch (metric) {
    case PC_GCS: case PC_JC: return ctx->d_den_phams.as<int32_t>();
    case PC_POCP: return ctx->d_den_genes.as<int32_t>();
    default: return ctx->d_den_len.as<int32_t>();
    }
}

// cheap metric for rows [r0, r1) into out_dev (device)
static int contract_rows(pc_ctx *ctx, int metric, int as_distance, int raw, int64_t r0,
                         int64_t r1, int32_t *out_dev, bool to_peers = false)
{
    EpilogueArgs ep;
    ep.den = den_of(ctx, metric);
    ep.metric = metric;
    ep.as_distance = as_distance;
    ep.raw = raw;
    ep.n = ctx->N;
    ep.row_begin = r0;
    ep.row_end = r1;
    ep.out = out_dev;
    ep.n_peer_out = 0;
    // AUTO: the tensor-core engine whenever the weights fit its limb layout
    int engine = ctx->engine;
    if (engine == PC_ENGINE_AUTO) engine = pc_i8_supported(ctx, metric) ? PC_ENGINE_I8MMA : PC_ENGINE_POPC;
    if (to_peers && ctx->n_peer_out) {
        if (engine != PC_ENGINE_I8MMA)
            return pc_fail(ctx, PC_EINVAL, "the fused all-gather needs the CTA-pair tensor engine (metric %d, engine %d)",
                           metric, engine);
        ep.n_peer_out = ctx->n_peer_out;
        for (int o = 0; o < ctx->n_peer_out; ++o) ep.peer_out[o] = ctx->peer_out[o];
    }
    if (engine == PC_ENGINE_I8MMA || engine == PC_ENGINE_I8MMA_V1) {
        PC_TRY(pc_i8_contract(ctx, metric, ep));
        return PC_OK;
    }
    int n_terms = 0;
    PC_TRY(prepare_terms(ctx, metric, &n_terms));
    PC_TRY(pc_stage_begin(ctx, PC_T_CONTRACT));
    PC_TRY(pc_launch_popc_contract(ctx, ctx->d_terms.as<PopcTerm>(), n_terms, ep));
    PC_TRY(pc_stage_end(ctx));
    return PC_OK;
}

// ---------------------------------------------------------------------------
// alignment planning (host) -- which member pairs exist, in which order
// ---------------------------------------------------------------------------
// ascending LSD radix sort of 64-bit keys, bytes that are the same in every key skipped
static void radix_sort_u64(std::vector<uint64_t> &keys)
{
    if (keys.size() < 2) return;
    uint64_t all_or = 0, all_and = ~0ull;
    for (const uint64_t k : keys) { all_or |= k; all_and &= k; }
    const uint64_t varying = all_or ^ all_and;
    std::vector<uint64_t> tmp(keys.size());
    for (int byte = 0; byte < 8; ++byte) {
        if (((varying >> (8 * byte)) & 0xff) == 0) continue;
        size_t hist[257] = {0};
        for (const uint64_t k : keys) hist[((k >> (8 * byte)) & 0xff) + 1]++;
        for (int d = 0; d < 256; ++d) hist[d + 1] += hist[d];
        for (const uint64_t k : keys) tmp[hist[(k >> (8 * byte)) & 0xff]++] = k;
        keys.swap(tmp);
    }
}

// Which partition aligns which pham.  The DP cells of a pham are the sum over its member pairs of
// len_a x len_b = (S^2 - Q) / 2 with S, Q the sum and the sum of squares of the members' summed
// translation lengths -- one pass over the entries, no member lists.  Phams held by >= 2 genomes
// are taken in descending order of S (ties: lower pham first; one radix sort) and each goes to the
// partition with the least cells so far (greedy LPT; r2 first dealt them in snake order of S^2,
// which overrates small phams up to 2x and left the ranks 10 % apart on C2).  A function of the
// TABLE only -- every rank computes the same deal, and pc_load_sequences_part can decide which
// genes a rank needs before any plan exists.  ctx->h_deal_order: the phams in deal order;
// ctx->h_pham_part[p]: owner of pham p or -1.
static void deal_phams(pc_ctx *ctx, int n_parts)
{
    if (ctx->deal_parts == n_parts && ctx->h_pham_part.size() == (size_t)std::max<int64_t>(ctx->P, 1)) return;
    const int T = host_threads(ctx->nnz);
    const int64_t P = ctx->P;
    auto entries_of = [&](int t) { return ctx->nnz * t / T; };
    auto phams_of = [&](int t) { return P * t / T; };
    std::vector<std::vector<uint64_t>> tot(T), sq(T);
    run_threads(T, [&](int t) {
        tot[t].assign((size_t)P, 0);
        sq[t].assign((size_t)P, 0);
        for (int64_t e = entries_of(t); e < entries_of(t + 1); ++e) {
            const uint64_t len = ctx->h_len_sum[e];
            tot[t][ctx->h_pham_id[e]] += len;
            sq[t][ctx->h_pham_id[e]] += len * len;
        }
    });
    constexpr uint64_t TOT_MAX = (1ull << 40) - 1;
    std::vector<std::vector<uint64_t>> local(T);
    std::vector<char> wide(T, 0);
    run_threads(T, [&](int t) {
        for (int64_t p = phams_of(t); p < phams_of(t + 1); ++p) {
            uint64_t sum = 0, sum_sq = 0;
            for (int u = 0; u < T; ++u) { sum += tot[u][p]; sum_sq += sq[u][p]; }
            tot[0][p] = sum;
            sq[0][p] = sum_sq;
            if (ctx->h_pham_ptr[p + 1] - ctx->h_pham_ptr[p] < 2) continue;   // cannot be shared by two genomes
            if (sum > TOT_MAX) wide[t] = 1;
            local[t].push_back(((TOT_MAX - std::min(sum, TOT_MAX)) << 24) | (uint64_t)(p & 0xffffff));
        }
    });
    size_t n = 0;
    bool packed = P < (1ll << 24);
    for (int t = 0; t < T; ++t) { n += local[t].size(); packed = packed && !wide[t]; }
    ctx->h_pham_part.assign(std::max<int64_t>(P, 1), -1);
    ctx->h_deal_order.resize(n);
    if (packed) {
        std::vector<uint64_t> keys;
        keys.reserve(n);
        for (auto &v : local) keys.insert(keys.end(), v.begin(), v.end());
        radix_sort_u64(keys);                  // (TOT_MAX - sum, pham) ascending = sum descending, pham ascending
        for (size_t pos = 0; pos < n; ++pos) ctx->h_deal_order[pos] = (int64_t)(keys[pos] & 0xffffff);
    } else {                                   // more than 16 M phams or sums beyond 40 bits: comparison sort
        size_t pos = 0;
        for (int64_t p = 0; p < P; ++p)
            if (ctx->h_pham_ptr[p + 1] - ctx->h_pham_ptr[p] >= 2) ctx->h_deal_order[pos++] = p;
        const std::vector<uint64_t> &sum = tot[0];
        std::sort(ctx->h_deal_order.begin(), ctx->h_deal_order.end(), [&](int64_t x, int64_t y) {
            return sum[x] != sum[y] ? sum[x] > sum[y] : x < y;
        });
    }
    std::vector<double> load((size_t)n_parts, 0.0);
    for (size_t pos = 0; pos < n; ++pos) {
        const int64_t p = ctx->h_deal_order[pos];
        int best = 0;
        for (int k = 1; k < n_parts; ++k)
            if (load[k] < load[best]) best = k;
        const double s1 = (double)tot[0][p], s2 = (double)sq[0][p];
        // + one row-step's worth per member pair: short alignments are not free
        const double members = (double)(ctx->h_pham_ptr[p + 1] - ctx->h_pham_ptr[p]);
        load[best] += 0.5 * (s1 * s1 - s2) + 2000.0 * 0.5 * members * (members - 1.0);
        ctx->h_pham_part[p] = (int32_t)best;
    }
    ctx->deal_parts = n_parts;
}

static int build_plan(pc_ctx *ctx, int part, int n_parts)
{
    AlignPlan &pl = ctx->plan;
    if (pl.part == part && pl.n_parts == n_parts) return PC_OK;
    if (ctx->seq_parts > 1 && (ctx->seq_parts != n_parts || ctx->seq_part != part))
        return pc_fail(ctx, PC_ESTATE, "sequences were loaded for partition %d/%d, the plan asks for %d/%d",
                       ctx->seq_part, ctx->seq_parts, part, n_parts);
    HostTimer host_timer(ctx, PC_T_HOST_PLAN);
    const int T = host_threads(ctx->nnz);
    deal_phams(ctx, n_parts);
    // this partition's phams, in deal (descending cost) order
    std::vector<int64_t> mine;
    for (const int64_t p : ctx->h_deal_order)
        if (ctx->h_pham_part[p] == part) mine.push_back(p);
    std::vector<int64_t> ph_mem_off(mine.size() + 1, 0), ph_unit_off(mine.size() + 1, 0);
    for (size_t x = 0; x < mine.size(); ++x) {
        const int64_t m = ctx->h_pham_ptr[mine[x] + 1] - ctx->h_pham_ptr[mine[x]];
        ph_mem_off[x + 1] = ph_mem_off[x] + m;
        ph_unit_off[x + 1] = ph_unit_off[x] + m * (m - 1) / 2;
    }
    std::vector<int32_t> mem_genome(ph_mem_off.back()), mem_first_gene(ph_mem_off.back()), mem_copies(ph_mem_off.back());
    // member lists of THIS partition's phams (ascending genome order) by a stable counting sort
    // over T entry ranges -- the other partitions' entries are only looked at
    const size_t nm = mine.size();
    std::vector<int32_t> slot_of((size_t)std::max<int64_t>(ctx->P, 1), -1);
    for (size_t x = 0; x < nm; ++x) slot_of[mine[x]] = (int32_t)x;
    auto entries_of = [&](int t) { return ctx->nnz * t / T; };
    std::vector<std::vector<int32_t>> cnt(T);
    run_threads(T, [&](int t) {
        cnt[t].assign(nm, 0);
        for (int64_t e = entries_of(t); e < entries_of(t + 1); ++e) {
            const int32_t x = slot_of[ctx->h_pham_id[e]];
            if (x >= 0) cnt[t][x]++;
        }
    });
    std::vector<int64_t> tasks_of(T, 0);
    run_threads(T, [&](int t) {
        for (size_t x = nm * (size_t)t / (size_t)T; x < nm * (size_t)(t + 1) / (size_t)T; ++x) {
            int32_t run = 0;
            for (int u = 0; u < T; ++u) {
                const int32_t c = cnt[u][x];
                cnt[u][x] = run;
                run += c;
            }
        }
    });
    run_threads(T, [&](int t) {
        for (int64_t e = entries_of(t); e < entries_of(t + 1); ++e) {
            const int32_t x = slot_of[ctx->h_pham_id[e]];
            if (x < 0) continue;
            const int64_t at = ph_mem_off[x] + cnt[t][x]++;
            mem_genome[at] = ctx->h_entry_genome[e];
            mem_first_gene[at] = (int32_t)ctx->h_entry_first_gene[e];
            mem_copies[at] = (int32_t)ctx->h_copies[e];
        }
    });
    run_threads(T, [&](int t) {
        int64_t tasks = 0;
        for (size_t x = nm * (size_t)t / (size_t)T; x < nm * (size_t)(t + 1) / (size_t)T; ++x) {
            int64_t sum_c = 0, sum_c2 = 0;
            for (int64_t at = ph_mem_off[x]; at < ph_mem_off[x + 1]; ++at) {
                const int64_t c = mem_copies[at];
                sum_c += c;
                sum_c2 += c * c;
            }
            tasks += (sum_c * sum_c - sum_c2) / 2;        // sum over member pairs of ca * cb
        }
        tasks_of[t] = tasks;
    });
    int64_t n_tasks = 0;
    for (int64_t v : tasks_of) n_tasks += v;
    pl.n_ph = (int64_t)ph_mem_off.size() - 1;
    pl.n_units = ph_unit_off.back();
    pl.n_tasks = n_tasks;
    pl.n_mem = (int64_t)mem_genome.size();
    pl.max_gene_len = ctx->max_gene_len;
    PC_TRY(upload(ctx, pl.ph_mem_off, ph_mem_off.data(), ph_mem_off.size()));
    PC_TRY(upload(ctx, pl.ph_unit_off, ph_unit_off.data(), ph_unit_off.size()));
    PC_TRY(upload(ctx, pl.mem_genome, mem_genome.data(), mem_genome.size()));
    PC_TRY(upload(ctx, pl.mem_first_gene, mem_first_gene.data(), mem_first_gene.size()));
    PC_TRY(upload(ctx, pl.mem_copies, mem_copies.data(), mem_copies.size()));
    PC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));   // sources are locals
    pl.part = part;
    pl.n_parts = n_parts;
    return PC_OK;
}

static void fill_nw_common(pc_ctx *ctx, NwArgs &a, int ppos)
{
    a.codes = ctx->d_codes.as<uint8_t>();
    a.gene_off = ctx->d_gene_off.as<int64_t>();
    a.gene_len = ctx->d_gene_len.as<int32_t>();
    a.subtab = ctx->d_subtab[ppos ? 1 : 0].as<int2>();
    a.gap_open = PC_GAP_OPEN_DEFAULT;
    a.gap_extend = PC_GAP_EXTEND_DEFAULT;
    a.cell_counter = ctx->d_counters.as<unsigned long long>() + PC_CTR_CELLS;
}

// the LEX DP kernel needs the 24-code replicated table and the per-gene bounds
static bool lex_allowed(const pc_ctx *ctx)
{
    return PC_NW_LEX_ENABLED && ctx->dp_mode == PC_DP_AUTO && ctx->n_codes <= 24 && ctx->d_gene_hi.p;
}

// counters layout: [0..PC_NW_NCFG) work counters, then cumulative cells of the classic DP kernel,
// the LEX kernel, the rebased LEX kernel and the LEXC kernel, then the LEXC fallback: number of
// listed alignments and the work counter of the launch that re-runs them, then the cumulative
// number of alignments that fell back
static int ensure_counters(pc_ctx *ctx)
{
    if (!ctx->d_counters.p) {
        PC_TRY(devbuf_reserve(ctx, ctx->d_counters, sizeof(unsigned long long) * PC_CTR_COUNT));
        PC_CUDA(ctx, cudaMemsetAsync(ctx->d_counters.p, 0, sizeof(unsigned long long) * PC_CTR_COUNT, ctx->stream));
    }
    return PC_OK;
}

// every chunked config gets its own spill region (they may run side by side): the 4 classic
// multi-chunk configs, the LEXC configs, and the launch that re-runs what LEXC gave up on
#define PC_SPILL_REGIONS ((PC_NW_LEX0 - PC_NW_MULTI0) + PC_NW_LEXC_NK + 1)
static int ensure_spill(pc_ctx *ctx, int max_qlen, NwArgs &a, int region)
{
    a.spill_stride = (int64_t)max_qlen + 2;
    const int64_t warps = pc_nw_multichunk_warps(ctx);
    PC_TRY(devbuf_reserve(ctx, ctx->d_spill, (size_t)PC_SPILL_REGIONS * warps * a.spill_stride * sizeof(int4)));
    a.spill = ctx->d_spill.as<int4>() + (size_t)region * warps * a.spill_stride;
    return PC_OK;
}

static bool cfg_spills(int cfg) { return (cfg >= PC_NW_MULTI0 && cfg < PC_NW_LEX0) || cfg >= PC_NW_LEXC0; }
static int spill_region(int cfg)
{
    return cfg >= PC_NW_LEXC0 ? (PC_NW_LEX0 - PC_NW_MULTI0) + (cfg - PC_NW_LEXC0) : cfg - PC_NW_MULTI0;
}

// LEXC launches list the alignments whose band checks failed; `n_lexc` bounds the list
static int prepare_fallback(pc_ctx *ctx, NwArgs &a, int64_t n_lexc)
{
    PC_TRY(devbuf_reserve(ctx, ctx->d_fallback, sizeof(int64_t) * (size_t)std::max<int64_t>(n_lexc, 1)));
    unsigned long long *ctr = ctx->d_counters.as<unsigned long long>();
    PC_CUDA(ctx, cudaMemsetAsync(ctr + PC_CTR_FALLBACK, 0, 2 * sizeof(unsigned long long), ctx->stream));
    a.fallback_count = ctr + PC_CTR_FALLBACK;
    a.fallback_slots = ctx->d_fallback.as<int64_t>();
    return PC_OK;
}

// ... and the classic multi-chunk kernel re-runs them (task count read on the device; an empty
// list costs one launch of blocks that exit at once).  `stream` must be ordered after every LEXC launch.
static int launch_fallback(pc_ctx *ctx, NwArgs a, int max_qlen, int64_t n_lexc, cudaStream_t stream)
{
    unsigned long long *ctr = ctx->d_counters.as<unsigned long long>();
    PC_TRY(ensure_spill(ctx, max_qlen, a, PC_SPILL_REGIONS - 1));
    a.perm = ctx->d_fallback.as<int64_t>();
    a.task_base = 0;
    a.n_tasks = n_lexc;
    a.n_tasks_dev = ctr + PC_CTR_FALLBACK;
    a.work_counter = ctr + PC_CTR_FALLBACK_WORK;
    a.fallback_count = nullptr;
    a.fallback_slots = nullptr;
    PC_TRY(pc_launch_nw_fallback(ctx, a, stream));
    pc_launch_add_counter(ctr + PC_CTR_FALLBACK_TOTAL, ctr + PC_CTR_FALLBACK, stream);
    return PC_OK;
}

// Plan partition part/n_parts, bucket its units by DP config and run every alignment: on return
// `u` describes the units and d_hitlen holds (matched << 16 | columns) per task, ready for the
// best-match pass (dense gather or contribution records).  *empty: nothing to align.
static int align_tasks(pc_ctx *ctx, int ppos, int part, int n_parts, UnitArgs &u, bool *empty)
{
    if (!ctx->have_seqs) return pc_fail(ctx, PC_ESTATE, "aai/peq need pc_load_sequences first");
    if (n_parts < 1 || part < 0 || part >= n_parts)
        return pc_fail(ctx, PC_EINVAL, "alignment partition %d/%d does not exist", part, n_parts);
    PC_TRY(build_plan(ctx, part, n_parts));
    AlignPlan &pl = ctx->plan;
    *empty = pl.n_tasks == 0;
    if (pl.n_tasks == 0) return PC_OK;
    if (pl.n_tasks >= (1ll << 40)) return pc_fail(ctx, PC_EINVAL, "too many alignments");
    PC_TRY(devbuf_reserve(ctx, ctx->d_task_q, (size_t)pl.n_tasks * sizeof(int32_t)));
    PC_TRY(devbuf_reserve(ctx, ctx->d_task_r, (size_t)pl.n_tasks * sizeof(int32_t)));
    PC_TRY(devbuf_reserve(ctx, ctx->d_hitlen, (size_t)pl.n_tasks * sizeof(uint32_t)));
    PC_TRY(ensure_counters(ctx));
    PC_CUDA(ctx, cudaMemsetAsync(ctx->d_counters.p, 0, sizeof(unsigned long long) * PC_NW_NCFG, ctx->stream));

    u.ph_mem_off = pl.ph_mem_off.as<int64_t>();
    u.ph_unit_off = pl.ph_unit_off.as<int64_t>();
    u.n_ph = pl.n_ph;
    u.n_units = pl.n_units;
    u.mem_genome = pl.mem_genome.as<int32_t>();
    u.mem_first_gene = pl.mem_first_gene.as<int32_t>();
    u.mem_copies = pl.mem_copies.as<int32_t>();
    u.gene_len = ctx->d_gene_len.as<int32_t>();
    u.gene_hi = lex_allowed(ctx) ? ctx->d_gene_hi.as<int32_t>() : nullptr;
    u.coarse_k = ctx->n_codes > 24;
    u.no_lexc = ctx->lexc_off ? 1 : 0;
    u.n_genomes = ctx->N;
    {
        int64_t blocks = (pl.n_units + 255) / 256;
        blocks = std::max<int64_t>(1, std::min<int64_t>(blocks, (int64_t)ctx->sm_count * 8));
        u.units_per_block = ((pl.n_units + blocks - 1) / blocks + 255) / 256 * 256;
        u.n_blocks = (int)((pl.n_units + u.units_per_block - 1) / u.units_per_block);
    }
    PC_TRY(devbuf_reserve(ctx, pl.blk_counts, sizeof(unsigned long long) * (size_t)u.n_blocks * PC_NW_NCFG));
    PC_TRY(devbuf_reserve(ctx, pl.cfg_info, sizeof(unsigned long long) * 2 * PC_NW_NCFG));
    uint64_t ht_cap = 1024;
    while (ht_cap < 2 * (uint64_t)pl.n_tasks) ht_cap <<= 1;
    PC_TRY(devbuf_reserve(ctx, pl.ht_keys, sizeof(unsigned long long) * ht_cap));
    PC_TRY(devbuf_reserve(ctx, pl.ht_vals, sizeof(unsigned long long) * ht_cap));
    u.blk_counts = pl.blk_counts.as<unsigned long long>();
    u.cfg_total = pl.cfg_info.as<unsigned long long>();
    u.cfg_base = u.cfg_total + PC_NW_NCFG;
    u.canon = ctx->d_canon.as<int32_t>();
    u.ht_keys = pl.ht_keys.as<unsigned long long>();
    u.ht_vals = pl.ht_vals.as<unsigned long long>();
    u.ht_mask = ht_cap - 1;

    // bucket the units by DP config on the device, then expand them into tasks
    PC_TRY(pc_stage_begin(ctx, PC_T_GATHER));
    PC_TRY(pc_launch_classify_units(ctx, u));
    unsigned long long cfg_info[2 * PC_NW_NCFG];
    PC_CUDA(ctx, cudaMemcpyAsync(cfg_info, pl.cfg_info.p, sizeof cfg_info, cudaMemcpyDeviceToHost, ctx->stream));
    // how the previous call's LEXC launches fared on this data (read at the sync that is needed anyway)
    unsigned long long prev_fallbacks = 0;
    if (ctx->lexc_last_tasks > 0)
        PC_CUDA(ctx, cudaMemcpyAsync(&prev_fallbacks, ctx->d_counters.as<unsigned long long>() + PC_CTR_FALLBACK,
                                     sizeof prev_fallbacks, cudaMemcpyDeviceToHost, ctx->stream));
    PC_TRY(pc_launch_scatter_units(ctx, u, ctx->d_task_q.as<int32_t>(), ctx->d_task_r.as<int32_t>()));
    PC_TRY(pc_stage_end(ctx));
    PC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));    // per-config task counts size the launches
    // more than a third of the LEXC alignments went to the classic kernel as well: this data has
    // long gaps, stop trying (takes effect at the next call; reset by pc_load_sequences)
    if (ctx->lexc_last_tasks >= 64 && 3 * (int64_t)prev_fallbacks > ctx->lexc_last_tasks) ctx->lexc_off = true;

    NwArgs a;
    memset(&a, 0, sizeof a);
    fill_nw_common(ctx, a, ppos);
    a.task_q = ctx->d_task_q.as<int32_t>();
    a.task_r = ctx->d_task_r.as<int32_t>();
    a.perm = nullptr;
    a.out_hitlen = ctx->d_hitlen.as<uint32_t>();
    a.out_score = nullptr;
    // the configs are independent: fork them over a few streams (long ones first) so
    // that one config's tail (its last, longest alignments) overlaps the next config
    PC_TRY(pc_stage_begin(ctx, PC_T_ALIGN));
    constexpr int MAX_DP_STREAMS = 8;
    static const int N_DP_STREAMS = [] {
        const char *env = getenv("PC_DP_STREAMS");
        const int v = env ? atoi(env) : 4;
        return v >= 1 && v <= MAX_DP_STREAMS ? v : 4;
    }();
    int64_t n_lexc = 0;
    for (int c = PC_NW_LEXC0; c < PC_NW_NCFG; ++c) n_lexc += (int64_t)cfg_info[c];
    ctx->lexc_last_tasks = n_lexc;
    if (n_lexc) PC_TRY(prepare_fallback(ctx, a, n_lexc));   // zeroes its counters on ctx->stream: before the fork
    cudaEvent_t fork = nullptr, join[MAX_DP_STREAMS] = {};
    PC_TRY(take_event(ctx, &fork));
    PC_CUDA(ctx, cudaEventRecord(fork, ctx->stream));
    for (int k = 0; k < N_DP_STREAMS; ++k) {
        if (!ctx->dp_streams[k])
            PC_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->dp_streams[k], cudaStreamNonBlocking));
        PC_TRY(take_event(ctx, &join[k]));
        PC_CUDA(ctx, cudaStreamWaitEvent(ctx->dp_streams[k], fork, 0));
    }
    for (int c = 0; c < PC_NW_NCFG; ++c)                  // size the shared spill buffer once
        if (cfg_info[c] && cfg_spills(c)) PC_TRY(ensure_spill(ctx, pl.max_gene_len, a, spill_region(c)));
    // launch order: the configs whose single alignments take longest go first (their tail -- a few
    // warps finishing their last alignment -- is then covered by the kernels behind them), the
    // configs of short alignments last (the step ends with a short tail).  Time of one alignment
    // ~ rows x columns per lane x instructions per cell; rows ~ columns for members of one pham.
    std::vector<std::pair<double, int>> order;
    for (int c = 0; c < PC_NW_NCFG; ++c) {
        if (cfg_info[c] == 0) continue;
        int G, K;
        pc_nw_cfg(c, G, K);
        const bool multi = cfg_spills(c);
        const double cols = (double)G * K * (multi ? 2.0 : 1.0);
        const double per_cell = c >= PC_NW_LEX0 ? 1.0 : 1.7;
        order.push_back({cols * K * (multi ? 2.0 : 1.0) * per_cell, c});
    }
    std::sort(order.begin(), order.end(), [](const std::pair<double, int> &x, const std::pair<double, int> &y) {
        return x.first != y.first ? x.first > y.first : x.second > y.second;
    });
    int next = 0;
    std::vector<cudaEvent_t> lexc_done;
    for (const auto &entry : order) {
        const int c = entry.second;
        const int64_t count = (int64_t)cfg_info[c];
        a.task_base = (int64_t)cfg_info[PC_NW_NCFG + c];
        a.n_tasks = count;
        a.work_counter = ctx->d_counters.as<unsigned long long>() + c;
        if (cfg_spills(c)) PC_TRY(ensure_spill(ctx, pl.max_gene_len, a, spill_region(c)));
        cudaStream_t st = ctx->dp_streams[next++ % N_DP_STREAMS];
        PC_TRY(pc_launch_nw(ctx, a, c, st));
        if (c >= PC_NW_LEXC0) {
            cudaEvent_t ev = nullptr;
            PC_TRY(take_event(ctx, &ev));
            PC_CUDA(ctx, cudaEventRecord(ev, st));
            lexc_done.push_back(ev);
        }
    }
    if (n_lexc) {
        // what LEXC gave up on, behind every LEXC launch (the stream that got the first -- longest --
        // config: the LEXC configs are at the head of the order)
        cudaStream_t st = ctx->dp_streams[0];
        for (cudaEvent_t ev : lexc_done) PC_CUDA(ctx, cudaStreamWaitEvent(st, ev, 0));
        PC_TRY(launch_fallback(ctx, a, pl.max_gene_len, n_lexc, st));
        for (cudaEvent_t ev : lexc_done) ctx->deferred_events.push_back(ev);
    }
    for (int k = 0; k < N_DP_STREAMS; ++k) {
        PC_CUDA(ctx, cudaEventRecord(join[k], ctx->dp_streams[k]));
        PC_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, join[k], 0));
    }
    PC_TRY(pc_stage_end(ctx));
    ctx->deferred_events.push_back(fork);
    for (int k = 0; k < N_DP_STREAMS; ++k) ctx->deferred_events.push_back(join[k]);
    for (int c = 0; c < PC_NW_NCFG; ++c) ctx->alignments += (int64_t)cfg_info[c];   // distinct sequence pairs
    for (int c = PC_NW_LEX0; c < PC_NW_NCFG; ++c) ctx->lex_alignments += (int64_t)cfg_info[c];
    ctx->logical_alignments += pl.n_tasks;
    return PC_OK;
}

extern "C" int pc_align_accumulate(pc_ctx *ctx, int ppos, int part, int n_parts,
                                   uint64_t *acc_dev)
{
    PC_ENTER(ctx);
    if (!acc_dev) return pc_fail(ctx, PC_EINVAL, "pc_align_accumulate: acc_dev is NULL");
    UnitArgs u;
    bool empty = false;
    PC_TRY(align_tasks(ctx, ppos, part, n_parts, u, &empty));
    if (empty) return PC_OK;
    PC_TRY(pc_stage_begin(ctx, PC_T_GATHER));
    PC_TRY(pc_launch_gather_units(ctx, u, ctx->d_hitlen.as<uint32_t>(),
                                  reinterpret_cast<unsigned long long *>(acc_dev)));
    PC_TRY(pc_stage_end(ctx));
    return maybe_collect(ctx);
}

extern "C" int pc_align_units(pc_ctx *ctx, int part, int n_parts, int64_t *n_units)
{
    PC_ENTER(ctx);
    if (!n_units) return pc_fail(ctx, PC_EINVAL, "pc_align_units: n_units is NULL");
    if (!ctx->have_seqs) return pc_fail(ctx, PC_ESTATE, "aai/peq need pc_load_sequences first");
    if (n_parts < 1 || part < 0 || part >= n_parts)
        return pc_fail(ctx, PC_EINVAL, "alignment partition %d/%d does not exist", part, n_parts);
    PC_TRY(build_plan(ctx, part, n_parts));
    *n_units = ctx->plan.n_units;
    return PC_OK;
}

extern "C" int pc_align_contributions(pc_ctx *ctx, int ppos, int part, int n_parts, int n_dest,
                                      const int64_t *row_bounds, uint64_t *records_dev,
                                      int64_t *dest_counts)
{
    PC_ENTER(ctx);
    if (n_dest < 1 || n_dest > PC_MAX_DEST || !row_bounds || !dest_counts)
        return pc_fail(ctx, PC_EINVAL, "pc_align_contributions: 1..%d destination row blocks", PC_MAX_DEST);
    if (row_bounds[0] != 0 || row_bounds[n_dest] != ctx->N)
        return pc_fail(ctx, PC_EINVAL, "pc_align_contributions: row_bounds must run from 0 to n_genomes");
    for (int d = 0; d < n_dest; ++d) {
        dest_counts[d] = 0;
        if (row_bounds[d] > row_bounds[d + 1]) return pc_fail(ctx, PC_EINVAL, "row_bounds is not monotone");
    }
    UnitArgs u;
    bool empty = false;
    PC_TRY(align_tasks(ctx, ppos, part, n_parts, u, &empty));
    if (empty || ctx->plan.n_units == 0) return PC_OK;
    if (!records_dev) return pc_fail(ctx, PC_EINVAL, "pc_align_contributions: records_dev is NULL");
    DestArgs d;
    d.n_dest = n_dest;
    for (int k = 0; k <= n_dest; ++k) d.bounds[k] = row_bounds[k];
    PC_TRY(devbuf_reserve(ctx, ctx->d_dest, sizeof(unsigned long long) * 2 * PC_MAX_DEST));
    unsigned long long counts[PC_MAX_DEST];
    PC_TRY(pc_stage_begin(ctx, PC_T_GATHER));
    PC_TRY(pc_launch_emit_units(ctx, u, ctx->d_hitlen.as<uint32_t>(), d, ctx->d_dest.as<unsigned long long>(),
                                reinterpret_cast<unsigned long long *>(records_dev)));
    PC_TRY(pc_stage_end(ctx));
    PC_CUDA(ctx, cudaMemcpyAsync(counts, ctx->d_dest.p, sizeof(unsigned long long) * n_dest,
                                 cudaMemcpyDeviceToHost, ctx->stream));
    PC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (int k = 0; k < n_dest; ++k) dest_counts[k] = (int64_t)counts[k];
    return maybe_collect(ctx);
}

extern "C" int pc_accumulate_records(pc_ctx *ctx, int64_t n_records, const uint64_t *records_dev,
                                     int64_t n_acc, uint64_t *acc_rows_dev)
{
    PC_ENTER(ctx);
    if (n_records < 0 || n_acc < 0 || (n_records > 0 && (!records_dev || !acc_rows_dev)))
        return pc_fail(ctx, PC_EINVAL, "pc_accumulate_records: bad arguments");
    if (n_records == 0) return PC_OK;
    PC_TRY(pc_stage_begin(ctx, PC_T_GATHER));
    PC_TRY(pc_launch_accumulate_records(ctx, n_records, reinterpret_cast<const unsigned long long *>(records_dev),
                                        n_acc, reinterpret_cast<unsigned long long *>(acc_rows_dev)));
    PC_TRY(pc_stage_end(ctx));
    return maybe_collect(ctx);
}

// ---------------------------------------------------------------------------
// pc_matrix_rows / pc_matrix / pc_numerators
// ---------------------------------------------------------------------------
static int64_t slice_len(const pc_ctx *ctx, int64_t r0, int64_t r1)
{
    return pc_tri_index(ctx->N, r1, r1) - pc_tri_index(ctx->N, r0, r0);
}

static int check_metric(pc_ctx *ctx, int metric)
{
    if (!ctx->have_table) return pc_fail(ctx, PC_ESTATE, "no table loaded (pc_load_table)");
    if (metric < PC_GCS || metric > PC_PEQ) return pc_fail(ctx, PC_EINVAL, "unknown metric %d", metric);
    if ((metric == PC_AAI || metric == PC_PEQ) && !ctx->have_seqs)
        return pc_fail(ctx, PC_ESTATE, "aai/peq need pc_load_sequences first");
    return PC_OK;
}

static int download(pc_ctx *ctx, void *host, const void *dev, size_t bytes)
{
    PC_TRY(pc_stage_begin(ctx, PC_T_D2H));
    PC_CUDA(ctx, cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    PC_TRY(pc_stage_end(ctx));
    PC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PC_OK;
}

extern "C" int pc_matrix_rows(pc_ctx *ctx, int metric, int as_distance, int ppos,
                              int64_t row_begin, int64_t row_end, const uint64_t *acc_dev,
                              int32_t *out, int out_on_device)
{
    PC_ENTER(ctx);
    PC_TRY(check_metric(ctx, metric));
    if (row_begin < 0 || row_end > ctx->N || row_begin > row_end || !out)
        return pc_fail(ctx, PC_EINVAL, "bad row range [%lld,%lld)", (long long)row_begin, (long long)row_end);
    if (ppos && metric != PC_AAI) return pc_fail(ctx, PC_EINVAL, "ppos only applies to aai");
    as_distance = as_distance ? 1 : 0;
    const int64_t count = slice_len(ctx, row_begin, row_end);
    if (count == 0) return PC_OK;
    int32_t *out_dev = out;
    if (!out_on_device) {
        PC_TRY(devbuf_reserve(ctx, ctx->d_out, (size_t)count * sizeof(int32_t)));
        out_dev = ctx->d_out.as<int32_t>();
    }
    if (metric <= PC_AF && !out_on_device && count > (int64_t(1) << 22)) {
        // host output of a large slice: compute row blocks (whole 256-row tiles, about
        // equal pair counts) and ship each finished block on a second stream while
        // the next one is computed -- the D2H of 4 B/pair otherwise dominates
        if (!ctx->copy_stream)
            PC_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
        const int n_blocks = (int)std::min<int64_t>(16, std::max<int64_t>(2, count >> 22));
        std::vector<int64_t> bounds{row_begin};
        for (int b = 1; b < n_blocks; ++b) {
            const int64_t want = pc_tri_index(ctx->N, row_begin, row_begin) + count * b / n_blocks;
            int64_t lo = bounds.back(), hi = row_end;
            while (lo < hi) {                       // smallest row whose block start >= want
                const int64_t mid = (lo + hi) / 2;
                if (pc_tri_index(ctx->N, mid, mid) >= want) hi = mid; else lo = mid + 1;
            }
            lo = std::min(row_end, (lo + 255) / 256 * 256);
            if (lo > bounds.back() && lo < row_end) bounds.push_back(lo);
        }
        bounds.push_back(row_end);
        cudaEvent_t ev = nullptr, c0 = nullptr, c1 = nullptr;
        PC_TRY(take_event(ctx, &ev));
        PC_TRY(take_event(ctx, &c0));
        PC_TRY(take_event(ctx, &c1));
        for (size_t b = 0; b + 1 < bounds.size(); ++b) {
            const int64_t off = slice_len(ctx, row_begin, bounds[b]);
            const int64_t len = slice_len(ctx, bounds[b], bounds[b + 1]);
            PC_TRY(contract_rows(ctx, metric, as_distance, 0, bounds[b], bounds[b + 1], out_dev + off));
            PC_CUDA(ctx, cudaEventRecord(ev, ctx->stream));
            PC_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ev, 0));
            if (b == 0) PC_CUDA(ctx, cudaEventRecord(c0, ctx->copy_stream));
            PC_CUDA(ctx, cudaMemcpyAsync(out + off, out_dev + off, (size_t)len * sizeof(int32_t),
                                         cudaMemcpyDeviceToHost, ctx->copy_stream));
        }
        PC_CUDA(ctx, cudaEventRecord(c1, ctx->copy_stream));
        PC_CUDA(ctx, cudaStreamSynchronize(ctx->copy_stream));
        PC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        float ms = 0.f;
        PC_CUDA(ctx, cudaEventElapsedTime(&ms, c0, c1));
        ctx->stage_ms[PC_T_D2H] += ms;
        ctx->event_pool.push_back(ev);
        ctx->event_pool.push_back(c0);
        ctx->event_pool.push_back(c1);
        return maybe_collect(ctx);
    }
    if (metric <= PC_AF) {
        PC_TRY(contract_rows(ctx, metric, as_distance, 0, row_begin, row_end, out_dev, true));
    } else {
        if (!acc_dev) return pc_fail(ctx, PC_EINVAL, "aai/peq rows need the accumulators");
        if (metric == PC_PEQ)   // alignment_fraction(source, target) as SIMILARITY, metrics.py:247
            PC_TRY(contract_rows(ctx, PC_AF, 0, 0, row_begin, row_end, out_dev));
        PeerOuts peers;
        peers.n = ctx->n_peer_out;
        for (int o = 0; o < ctx->n_peer_out; ++o) peers.p[o] = ctx->peer_out[o];
        PC_TRY(pc_stage_begin(ctx, PC_T_EPILOGUE));
        PC_TRY(pc_launch_aai_epilogue(ctx, metric, as_distance, 0, row_begin, row_end,
                                      reinterpret_cast<const unsigned long long *>(acc_dev),
                                      out_dev, nullptr, &peers));
        PC_TRY(pc_stage_end(ctx));
    }
    if (!out_on_device) PC_TRY(download(ctx, out, out_dev, (size_t)count * sizeof(int32_t)));
    return maybe_collect(ctx);
}

extern "C" int pc_matrix_rows_owned(pc_ctx *ctx, int metric, int as_distance, int ppos,
                                    int64_t row_begin, int64_t row_end, const uint64_t *acc_rows_dev,
                                    int32_t *out, int out_on_device)
{
    PC_ENTER(ctx);
    if (row_begin < 0 || row_end > ctx->N || row_begin > row_end)
        return pc_fail(ctx, PC_EINVAL, "bad row range [%lld,%lld)", (long long)row_begin, (long long)row_end);
    // the kernels index the accumulators by the packed-triangle index of the pair: hand them the
    // address entry (0, 0) WOULD have if the owned slice were part of a whole triangle (only the
    // entries of rows [row_begin, row_end) are ever read)
    const uint64_t *base = acc_rows_dev ? acc_rows_dev - pc_tri_index(ctx->N, row_begin, row_begin) : nullptr;
    return pc_matrix_rows(ctx, metric, as_distance, ppos, row_begin, row_end, base, out, out_on_device);
}

extern "C" int pc_matrix_rows_fused(pc_ctx *ctx, int metric, int as_distance, int ppos,
                                    int64_t row_begin, int64_t row_end, const uint64_t *acc_rows_dev,
                                    int32_t *out_dev, int n_peers, int32_t *const *peer_out_dev)
{
    PC_ENTER(ctx);
    if (n_peers < 0 || n_peers > PC_MAX_PEER_OUT || (n_peers > 0 && !peer_out_dev) || !out_dev)
        return pc_fail(ctx, PC_EINVAL, "pc_matrix_rows_fused: 0..%d peers", PC_MAX_PEER_OUT);
    ctx->n_peer_out = n_peers;
    for (int o = 0; o < n_peers; ++o) {
        if (!peer_out_dev[o]) { ctx->n_peer_out = 0; return pc_fail(ctx, PC_EINVAL, "pc_matrix_rows_fused: peer %d is NULL", o); }
        ctx->peer_out[o] = peer_out_dev[o];
    }
    const int rc = pc_matrix_rows_owned(ctx, metric, as_distance, ppos, row_begin, row_end, acc_rows_dev, out_dev, 1);
    ctx->n_peer_out = 0;
    return rc;
}

static int full_accumulate(pc_ctx *ctx, int ppos)
{
    const size_t bytes = (size_t)slice_len(ctx, 0, ctx->N) * sizeof(uint64_t);
    PC_TRY(devbuf_reserve(ctx, ctx->d_acc, bytes));
    PC_CUDA(ctx, cudaMemsetAsync(ctx->d_acc.p, 0, bytes, ctx->stream));
    return pc_align_accumulate(ctx, ppos, 0, 1, ctx->d_acc.as<uint64_t>());
}

extern "C" int pc_matrix(pc_ctx *ctx, int metric, int as_distance, int ppos,
                         int32_t *out_micro_upper, int out_on_device)
{
    PC_ENTER(ctx);
    PC_TRY(check_metric(ctx, metric));
    const uint64_t *acc = nullptr;
    if (metric == PC_AAI || metric == PC_PEQ) {
        PC_TRY(full_accumulate(ctx, ppos));
        acc = ctx->d_acc.as<uint64_t>();
    }
    return pc_matrix_rows(ctx, metric, as_distance, ppos, 0, ctx->N, acc, out_micro_upper,
                          out_on_device);
}

// ---------------------------------------------------------------------------
// the resident triangle (pc_resident.cu)
// ---------------------------------------------------------------------------
extern "C" int pc_matrix_resident(pc_ctx *ctx, int metric, int as_distance, int ppos)
{
    PC_ENTER(ctx);
    PC_TRY(check_metric(ctx, metric));
    ctx->resident_n = -1;
    PC_TRY(devbuf_reserve(ctx, ctx->d_resident, (size_t)slice_len(ctx, 0, ctx->N) * sizeof(int32_t)));
    PC_TRY(pc_matrix(ctx, metric, as_distance, ppos, ctx->d_resident.as<int32_t>(), 1));
    ctx->resident_n = ctx->N;
    return PC_OK;
}

static int need_resident(pc_ctx *ctx)
{
    if (ctx->resident_n < 0 || ctx->resident_n != ctx->N)
        return pc_fail(ctx, PC_ESTATE, "no resident matrix (pc_matrix_resident, and no table load since)");
    return PC_OK;
}

extern "C" int pc_resident_threshold_edges(pc_ctx *ctx, int32_t max_micro, int64_t cap, int32_t *edges_host,
                                           int64_t *n_edges)
{
    PC_ENTER(ctx);
    PC_TRY(need_resident(ctx));
    if (cap < 0 || !n_edges || (cap > 0 && !edges_host))
        return pc_fail(ctx, PC_EINVAL, "pc_resident_threshold_edges: bad arguments");
    PC_TRY(devbuf_reserve(ctx, ctx->d_res_edges, (size_t)std::max<int64_t>(cap, 1) * sizeof(int2) + 16));
    unsigned long long *count = reinterpret_cast<unsigned long long *>(ctx->d_res_edges.as<int2>() + std::max<int64_t>(cap, 1));
    PC_CUDA(ctx, cudaMemsetAsync(count, 0, sizeof(unsigned long long), ctx->stream));
    PC_TRY(pc_stage_begin(ctx, PC_T_EPILOGUE));
    PC_TRY(pc_launch_threshold_edges(ctx, ctx->d_resident.as<int32_t>(), ctx->resident_n, max_micro, count,
                                     ctx->d_res_edges.as<int2>(), cap));
    PC_TRY(pc_stage_end(ctx));
    unsigned long long found = 0;
    PC_CUDA(ctx, cudaMemcpyAsync(&found, count, sizeof found, cudaMemcpyDeviceToHost, ctx->stream));
    PC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *n_edges = (int64_t)found;                      // may exceed cap: the caller retries with more room
    const int64_t have = std::min<int64_t>((int64_t)found, cap);
    if (have > 0) PC_TRY(download(ctx, edges_host, ctx->d_res_edges.p, (size_t)have * sizeof(int2)));
    return maybe_collect(ctx);
}

extern "C" int pc_resident_gather(pc_ctx *ctx, int64_t n_idx, const int64_t *tri_index_host, int32_t *out_host)
{
    PC_ENTER(ctx);
    PC_TRY(need_resident(ctx));
    if (n_idx < 0 || (n_idx > 0 && (!tri_index_host || !out_host)))
        return pc_fail(ctx, PC_EINVAL, "pc_resident_gather: bad arguments");
    if (n_idx == 0) return PC_OK;
    const int64_t total = slice_len(ctx, 0, ctx->N);
    for (int64_t k = 0; k < n_idx; ++k)
        if (tri_index_host[k] < 0 || tri_index_host[k] >= total)
            return pc_fail(ctx, PC_EINVAL, "pc_resident_gather: index %lld outside the triangle", (long long)tri_index_host[k]);
    PC_TRY(upload(ctx, ctx->d_res_idx, tri_index_host, (size_t)n_idx));
    PC_TRY(devbuf_reserve(ctx, ctx->d_res_vals, (size_t)n_idx * sizeof(int32_t)));
    PC_TRY(pc_launch_gather_entries(ctx, ctx->d_resident.as<int32_t>(), n_idx, ctx->d_res_idx.as<int64_t>(),
                                    ctx->d_res_vals.as<int32_t>()));
    return download(ctx, out_host, ctx->d_res_vals.p, (size_t)n_idx * sizeof(int32_t));
}

extern "C" int pc_resident_fetch(pc_ctx *ctx, int32_t *out_host)
{
    PC_ENTER(ctx);
    PC_TRY(need_resident(ctx));
    if (!out_host) return pc_fail(ctx, PC_EINVAL, "pc_resident_fetch: out is NULL");
    return download(ctx, out_host, ctx->d_resident.p, (size_t)slice_len(ctx, 0, ctx->N) * sizeof(int32_t));
}

extern "C" int pc_numerators(pc_ctx *ctx, int metric, int ppos, int32_t *out_num,
                             int32_t *out_aux)
{
    PC_ENTER(ctx);
    PC_TRY(check_metric(ctx, metric));
    if (!out_num || metric == PC_PEQ || (metric == PC_AAI && !out_aux))
        return pc_fail(ctx, PC_EINVAL, "pc_numerators: bad arguments");
    const int64_t count = slice_len(ctx, 0, ctx->N);
    PC_TRY(devbuf_reserve(ctx, ctx->d_out, (size_t)count * sizeof(int32_t)));
    if (metric <= PC_AF) {
        PC_TRY(contract_rows(ctx, metric, 0, 1, 0, ctx->N, ctx->d_out.as<int32_t>()));
        return download(ctx, out_num, ctx->d_out.p, (size_t)count * sizeof(int32_t));
    }
    PC_TRY(full_accumulate(ctx, ppos));
    PC_TRY(devbuf_reserve(ctx, ctx->d_numer, (size_t)count * sizeof(int32_t)));
    PC_TRY(pc_stage_begin(ctx, PC_T_EPILOGUE));
    PC_TRY(pc_launch_aai_epilogue(ctx, PC_AAI, 0, 1, 0, ctx->N,
                                  ctx->d_acc.as<unsigned long long>(), ctx->d_out.as<int32_t>(),
                                  ctx->d_numer.as<int32_t>()));
    PC_TRY(pc_stage_end(ctx));
    PC_TRY(download(ctx, out_num, ctx->d_out.p, (size_t)count * sizeof(int32_t)));
    return download(ctx, out_aux, ctx->d_numer.p, (size_t)count * sizeof(int32_t));
}

// ---------------------------------------------------------------------------
// pc_alignment_identities -- test hook over caller-chosen gene pairs
// ---------------------------------------------------------------------------
extern "C" int pc_alignment_identities(pc_ctx *ctx, int64_t n_tasks, const int32_t *query_gene,
                                       const int32_t *ref_gene, int32_t *out)
{
    PC_ENTER(ctx);
    if (!ctx->have_seqs) return pc_fail(ctx, PC_ESTATE, "pc_alignment_identities needs pc_load_sequences first");
    if (ctx->seq_parts > 1)
        return pc_fail(ctx, PC_ESTATE, "pc_alignment_identities needs ALL genes (pc_load_sequences, not _part)");
    if (n_tasks < 0 || (n_tasks > 0 && (!query_gene || !ref_gene || !out)))
        return pc_fail(ctx, PC_EINVAL, "pc_alignment_identities: bad arguments");
    if (n_tasks == 0) return PC_OK;
    const bool use_lex = lex_allowed(ctx);
    if (use_lex && (int64_t)ctx->h_gene_hi.size() != ctx->n_genes) {
        ctx->h_gene_hi.resize((size_t)ctx->n_genes);
        PC_CUDA(ctx, cudaMemcpyAsync(ctx->h_gene_hi.data(), ctx->d_gene_hi.p, (size_t)ctx->n_genes * sizeof(int32_t),
                                     cudaMemcpyDeviceToHost, ctx->stream));
        PC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    // processing order: by kernel config of the ref length, largest first
    std::vector<int64_t> bucket_count(PC_NW_NCFG, 0);
    std::vector<int> bucket(n_tasks);
    int32_t max_q_multi = 0;
    for (int64_t t = 0; t < n_tasks; ++t) {
        if (query_gene[t] < 0 || query_gene[t] >= ctx->n_genes || ref_gene[t] < 0 || ref_gene[t] >= ctx->n_genes)
            return pc_fail(ctx, PC_EINVAL, "task %lld names a gene outside [0,%lld)", (long long)t, (long long)ctx->n_genes);
        const int lq = ctx->h_gene_len[query_gene[t]], lr = ctx->h_gene_len[ref_gene[t]];
        const int mode = use_lex ? pc_nw_mode(lq, lr, ctx->h_gene_hi[query_gene[t]], ctx->h_gene_hi[ref_gene[t]], true)
                                 : PC_NW_MODE_CLASSIC;
        bucket[t] = pc_nw_choose_cfg(lq, lr, mode, ctx->n_codes > 24);
        bucket_count[bucket[t]]++;
        if (cfg_spills(bucket[t]))
            max_q_multi = std::max(max_q_multi, ctx->h_gene_len[query_gene[t]]);
    }
    std::vector<int64_t> bucket_off(PC_NW_NCFG + 1, 0);
    for (int b = 0; b < PC_NW_NCFG; ++b) bucket_off[b + 1] = bucket_off[b] + bucket_count[b];
    std::vector<int64_t> perm(n_tasks);
    {
        std::vector<int64_t> cursor(bucket_off.begin(), bucket_off.end() - 1);
        for (int64_t t = 0; t < n_tasks; ++t) perm[cursor[bucket[t]]++] = t;
    }
    PC_TRY(upload(ctx, ctx->d_task_q, query_gene, (size_t)n_tasks));
    PC_TRY(upload(ctx, ctx->d_task_r, ref_gene, (size_t)n_tasks));
    PC_TRY(upload(ctx, ctx->d_perm, perm.data(), (size_t)n_tasks));
    PC_TRY(devbuf_reserve(ctx, ctx->d_hitlen, (size_t)n_tasks * sizeof(uint32_t) * 2));
    PC_TRY(devbuf_reserve(ctx, ctx->d_score, (size_t)n_tasks * sizeof(int32_t)));
    PC_TRY(ensure_counters(ctx));

    NwArgs a;
    memset(&a, 0, sizeof a);
    a.task_q = ctx->d_task_q.as<int32_t>();
    a.task_r = ctx->d_task_r.as<int32_t>();
    a.perm = ctx->d_perm.as<int64_t>();
    a.out_score = ctx->d_score.as<int32_t>();
    for (int ppos = 0; ppos < 2; ++ppos) {
        fill_nw_common(ctx, a, ppos);
        a.out_hitlen = ctx->d_hitlen.as<uint32_t>() + (size_t)ppos * n_tasks;
        PC_CUDA(ctx, cudaMemsetAsync(ctx->d_counters.p, 0, sizeof(unsigned long long) * PC_NW_NCFG, ctx->stream));
        int64_t n_lexc = 0;
        for (int b = PC_NW_LEXC0; b < PC_NW_NCFG; ++b) n_lexc += bucket_count[b];
        if (n_lexc) PC_TRY(prepare_fallback(ctx, a, n_lexc));
        ctx->lexc_last_tasks = 0;             // the fallback counter no longer describes an align_tasks call
        PC_TRY(pc_stage_begin(ctx, PC_T_ALIGN));
        for (int b = PC_NW_NCFG - 1; b >= 0; --b) {
            if (bucket_count[b] == 0) continue;
            a.task_base = bucket_off[b];
            a.n_tasks = bucket_count[b];
            a.work_counter = ctx->d_counters.as<unsigned long long>() + b;
            if (cfg_spills(b)) PC_TRY(ensure_spill(ctx, max_q_multi, a, spill_region(b)));
            PC_TRY(pc_launch_nw(ctx, a, b, ctx->stream));
        }
        if (n_lexc) {
            NwArgs fb = a;
            PC_TRY(launch_fallback(ctx, fb, max_q_multi, n_lexc, ctx->stream));
        }
        PC_TRY(pc_stage_end(ctx));
        ctx->alignments += n_tasks;
        for (int b = PC_NW_LEX0; b < PC_NW_NCFG; ++b) ctx->lex_alignments += bucket_count[b];
    }
    std::vector<uint32_t> hl((size_t)n_tasks * 2);
    std::vector<int32_t> score(n_tasks);
    PC_CUDA(ctx, cudaMemcpyAsync(hl.data(), ctx->d_hitlen.p, hl.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    PC_CUDA(ctx, cudaMemcpyAsync(score.data(), ctx->d_score.p, score.size() * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    PC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (int64_t t = 0; t < n_tasks; ++t) {
        const uint32_t id = hl[t], pp = hl[(size_t)n_tasks + t];
        out[4 * t + 0] = (int32_t)(id >> 16);
        out[4 * t + 1] = (int32_t)(pp >> 16) - (int32_t)(id >> 16);
        out[4 * t + 2] = (int32_t)(id & 0xffffu);
        out[4 * t + 3] = score[t];
    }
    return maybe_collect(ctx);
}

// ---------------------------------------------------------------------------
// pc_debug_round6
// ---------------------------------------------------------------------------
__global__ void round6_kernel(const double *__restrict__ x, int64_t n, int32_t *__restrict__ out)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) out[k] = pc_round6_micro(x[k]);
}

extern "C" int pc_debug_round6(pc_ctx *ctx, int64_t n, const double *x, int32_t *out)
{
    PC_ENTER(ctx);
    if (n < 0 || (n > 0 && (!x || !out))) return pc_fail(ctx, PC_EINVAL, "pc_debug_round6: bad arguments");
    if (n == 0) return PC_OK;
    DevBuf dx, dout;
    int rc = upload(ctx, dx, x, (size_t)n);
    if (rc == PC_OK) rc = devbuf_reserve(ctx, dout, (size_t)n * sizeof(int32_t));
    if (rc == PC_OK) {
        round6_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(dx.as<double>(), n, dout.as<int32_t>());
        ctx->launches++;
        if (cudaMemcpyAsync(out, dout.p, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
            cudaStreamSynchronize(ctx->stream) != cudaSuccess)
            rc = pc_fail(ctx, PC_ECUDA, "round6 kernel failed: %s", cudaGetErrorString(cudaGetLastError()));
    }
    devbuf_free(dx);
    devbuf_free(dout);
    return rc;
}

// both arithmetic paths of the cheap-metric epilogue on caller-chosen operands (test hook)
__global__ void pc_debug_metric_kernel(int metric, int as_distance, int64_t n, const int32_t *numer,
                                       const int32_t *den_i, const int32_t *den_j, int32_t *out, int32_t *out_f64)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    out[k] = pc_cheap_metric_micro(metric, as_distance, numer[k], den_i[k], den_j[k]);
    out_f64[k] = numer[k] > 0 ? pc_cheap_metric_micro_f64(metric, as_distance, numer[k], (int64_t)den_i[k] + den_j[k])
                              : (as_distance ? 1000000 : 0);
}

extern "C" int pc_debug_metric_micro(pc_ctx *ctx, int metric, int as_distance, int64_t n, const int32_t *numer,
                                     const int32_t *den_i, const int32_t *den_j, int32_t *out, int32_t *out_f64)
{
    PC_ENTER(ctx);
    if (n < 0 || metric < PC_GCS || metric > PC_AF || (n > 0 && (!numer || !den_i || !den_j || !out || !out_f64)))
        return pc_fail(ctx, PC_EINVAL, "pc_debug_metric_micro: bad arguments");
    if (n == 0) return PC_OK;
    DevBuf a, b, c, d, e;
    int rc = PC_OK;
    do {
        if ((rc = upload(ctx, a, numer, (size_t)n))) break;
        if ((rc = upload(ctx, b, den_i, (size_t)n))) break;
        if ((rc = upload(ctx, c, den_j, (size_t)n))) break;
        if ((rc = devbuf_reserve(ctx, d, (size_t)n * sizeof(int32_t)))) break;
        if ((rc = devbuf_reserve(ctx, e, (size_t)n * sizeof(int32_t)))) break;
        pc_debug_metric_kernel<<<(int)((n + 255) / 256), 256, 0, ctx->stream>>>(metric, as_distance ? 1 : 0, n,
            a.as<int32_t>(), b.as<int32_t>(), c.as<int32_t>(), d.as<int32_t>(), e.as<int32_t>());
        ctx->launches++;
        if (cudaMemcpyAsync(out, d.p, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
            cudaMemcpyAsync(out_f64, e.p, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
            cudaStreamSynchronize(ctx->stream) != cudaSuccess)
            rc = pc_fail(ctx, PC_ECUDA, "pc_debug_metric_micro failed: %s", cudaGetErrorString(cudaGetLastError()));
    } while (0);
    cudaStreamSynchronize(ctx->stream);
    devbuf_free(a); devbuf_free(b); devbuf_free(c); devbuf_free(d); devbuf_free(e);
    return rc;
}

// ---------------------------------------------------------------------------
// measurement
// ---------------------------------------------------------------------------
extern "C" int pc_timings(pc_ctx *ctx, double *ms, int n)
{
    PC_ENTER(ctx);
    PC_TRY(pc_collect_timers(ctx));
    for (int k = 0; k < n && k < PC_T_COUNT; ++k) ms[k] = ctx->stage_ms[k];
    return PC_OK;
}

extern "C" int pc_counters(pc_ctx *ctx, int64_t *launches, int64_t *dp_cells,
                           int64_t *alignments, int64_t *contract_launches,
                           int64_t *align_launches)
{
    PC_ENTER(ctx);
    unsigned long long cells = 0;
    if (ctx->d_counters.p) {
        unsigned long long all[4] = {0, 0, 0, 0};     // classic, LEX, rebased LEX, LEXC
        PC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        PC_CUDA(ctx, cudaMemcpy(all, ctx->d_counters.as<unsigned long long>() + PC_CTR_CELLS,
                                sizeof all, cudaMemcpyDeviceToHost));
        cells = all[0] + all[1] + all[2] + all[3];
    }
    if (launches) *launches = ctx->launches;
    if (dp_cells) *dp_cells = (int64_t)cells;
    if (alignments) *alignments = ctx->alignments;
    if (contract_launches) *contract_launches = ctx->contract_launches;
    if (align_launches) *align_launches = ctx->align_launches;
    return PC_OK;
}

extern "C" int pc_stat(pc_ctx *ctx, int which, int64_t *out)
{
    PC_ENTER(ctx);
    if (!out) return pc_fail(ctx, PC_EINVAL, "pc_stat: out is NULL");
    switch (which) {
    case PC_STAT_LOGICAL_ALIGNMENTS: *out = ctx->logical_alignments; break;
    case PC_STAT_UNIQUE_GENES: *out = ctx->n_unique_genes; break;
    case PC_STAT_OPERAND_COLUMNS: *out = ctx->P_eff; break;
    case PC_STAT_GENES: *out = ctx->n_genes; break;
    case PC_STAT_LEX_ALIGNMENTS: *out = ctx->lex_alignments; break;
    case PC_STAT_LEX_CELLS:
    case PC_STAT_LEXR_CELLS:
    case PC_STAT_LEXC_CELLS:
    case PC_STAT_LEXC_FALLBACKS: {
        unsigned long long cells = 0;
        if (ctx->d_counters.p) {
            const int at = which == PC_STAT_LEX_CELLS ? PC_CTR_CELLS + 1 : which == PC_STAT_LEXR_CELLS ? PC_CTR_CELLS + 2
                         : which == PC_STAT_LEXC_CELLS ? PC_CTR_CELLS + 3 : PC_CTR_FALLBACK_TOTAL;
            PC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            PC_CUDA(ctx, cudaMemcpy(&cells, ctx->d_counters.as<unsigned long long>() + at,
                                    sizeof cells, cudaMemcpyDeviceToHost));
        }
        *out = (int64_t)cells;
        break;
    }
    case PC_STAT_I8_TILE_STEPS: {
        unsigned long long steps = 0;
        if (ctx->d_i8_steps.p) {
            PC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            PC_CUDA(ctx, cudaMemcpy(&steps, ctx->d_i8_steps.p, sizeof steps, cudaMemcpyDeviceToHost));
        }
        *out = (int64_t)steps;
        break;
    }
    case PC_STAT_I8_TILE_STEPS_DENSE: *out = ctx->i8_dense_steps; break;
    case PC_STAT_H2D_BYTES: *out = ctx->h2d_bytes; break;
    case PC_STAT_DP_CONFIGS_USED: {
        int64_t used = 0;
        for (char c : ctx->cfg_used) used += c;
        *out = used;
        break;
    }
    case PC_STAT_DP_CONFIGS: {
        // every config pc_nw_choose_cfg returns for some (query rows, ref columns) with the
        // loaded alphabet, by enumeration; the classic kernel only gets what the LEX kernel
        // cannot take for ANY residues (score bound of an all-W sequence: 11 per residue)
        const bool coarse = ctx->n_codes > 24, lex = lex_allowed(ctx);
        std::vector<char> reach(PC_NW_NCFG, 0);
        for (int n = 1; n <= 4200; ++n) {
            for (int m = 1; m <= 4096; m *= 4) {
                // cheapest and dearest residues (score bounds 0 and 11 per residue) span the modes
                reach[pc_nw_choose_cfg(m, n, pc_nw_mode(m, n, 0, 0, lex), coarse)] = 1;
                reach[pc_nw_choose_cfg(m, n, pc_nw_mode(m, n, 11 * m, 11 * n, lex), coarse)] = 1;
            }
            // queries about as long as the ref (LEXC only takes those)
            reach[pc_nw_choose_cfg(n, n, pc_nw_mode(n, n, 11 * n, 11 * n, lex), coarse)] = 1;
        }
        int64_t count = 0;
        for (char c : reach) count += c;
        *out = count;
        break;
    }
    default: return pc_fail(ctx, PC_EINVAL, "pc_stat: unknown statistic %d", which);
    }
    return PC_OK;
}

extern "C" int pc_reset_counters(pc_ctx *ctx)
{
    PC_ENTER(ctx);
    PC_TRY(pc_collect_timers(ctx));
    for (double &v : ctx->stage_ms) v = 0.0;
    ctx->launches = ctx->contract_launches = ctx->align_launches = 0;
    ctx->alignments = 0;
    ctx->logical_alignments = 0;
    ctx->lex_alignments = 0;
    ctx->h2d_bytes = 0;
    std::fill(ctx->cfg_used.begin(), ctx->cfg_used.end(), 0);
    ctx->i8_dense_steps = 0;
    if (ctx->d_i8_steps.p)
        PC_CUDA(ctx, cudaMemsetAsync(ctx->d_i8_steps.p, 0, sizeof(unsigned long long), ctx->stream));
    if (ctx->d_counters.p) {
        PC_CUDA(ctx, cudaMemsetAsync(ctx->d_counters.as<unsigned long long>() + PC_CTR_CELLS, 0,
                                     (PC_CTR_COUNT - PC_CTR_CELLS) * sizeof(unsigned long long), ctx->stream));
        PC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return PC_OK;
}
