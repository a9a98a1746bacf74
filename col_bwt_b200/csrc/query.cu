// query.cu -- host driver of the query path: read packing, device batches, the pinned multi-slot
// streaming pipeline and multi-GPU read sharding.
//
// Replaces the per-read loop of src/pml_query.cpp:74-86 (`while (patterns.read()) tbl.query_pml(...)`).
// Reads are independent (state is re-initialised per read, include/col_bwt.hpp:503-508), so the batch is cut into
// chunks that are packed by a host thread pool, copied with cudaMemcpyAsync from pinned staging buffers, traversed,
// and copied back in input order; with several GPUs the chunks are dealt round-robin to replicas of the table.
// No inter-GPU communication exists on this path.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <thread>

#include "internal.h"
#include "tasks.h"

namespace colbwt {

// ---------------------------------------------------------------------------------------------------------
// Minimal persistent thread pool: parallel_for over task indices.
// ---------------------------------------------------------------------------------------------------------
class Pool {
public:
    static Pool &get()
    {
        static Pool p;
        return p;
    }
    int size() const { return (int)workers_.size() + 1; }
    // One job at a time: the job lives in shared fields (fn_, n_tasks_, next_, pending_, epoch_), and the pool is shared by
    // every index and every calling thread of the process (colbwt_query on two indexes, colbwt_batch_upload next to a query,
    // the per-device feeder threads), so callers queue on caller_m_ for the whole parallel_for.
    // Workers sleep on a condition variable between jobs (optionally after spinning on the epoch for a while, see SPINS).
    void parallel_for(int n_tasks, const std::function<void(int)> &fn)
    {
        if (n_tasks <= 1 || workers_.empty()) {
            for (int i = 0; i < n_tasks; ++i) fn(i);
            return;
        }
        std::lock_guard<std::mutex> one_job(caller_m_);
        {
            std::lock_guard<std::mutex> g(m_);
            fn_ = &fn;
            n_tasks_ = n_tasks;
            next_.store(0);
            pending_.store((int)workers_.size());
            epoch_.fetch_add(1, std::memory_order_release);
        }
        cv_.notify_all();
        run_tasks();
        for (int spin = 0; spin < SPINS && pending_.load(std::memory_order_acquire) != 0; ++spin) cpu_relax();
        if (pending_.load(std::memory_order_acquire) != 0) {
            std::unique_lock<std::mutex> g(m_);
            done_cv_.wait(g, [&] { return pending_.load(std::memory_order_acquire) == 0; });
        }
        fn_ = nullptr;
    }

private:
    // Spinning between jobs (COLBWT_POOL_SPIN = iterations of `pause`, ~50 ns each) is OFF by default: on one GPU with 16 cores
    // it changed nothing (29.1 / 28.9 against 29.5 / 28.9 Gbases/s end to end, profiles/r2/r2_spin*.json), and with one rank
    // per GPU on a fully subscribed host (8 ranks x 4 threads + their CUDA and Python threads on 32 vCPUs) the spinners take
    // the cores the other ranks' packers need: every host-heavy mode ran at half its rate in the 8-rank run that had them
    // (profiles/r2/SUMMARY.md section 6).
    const int SPINS = spin_budget();
    static int spin_budget()
    {
        if (const char *e = getenv("COLBWT_POOL_SPIN")) return std::max(0, atoi(e));
        return 0;
    }
    static inline void cpu_relax()
    {
#if defined(__x86_64__)
        __builtin_ia32_pause();
#endif
    }
    Pool()
    {
        int n = (int)std::thread::hardware_concurrency();
        if (const char *e = getenv("LOCAL_WORLD_SIZE")) n = std::max(1, n / std::max(1, atoi(e)));   // one process per GPU: share the cores
        if (const char *e = getenv("COLBWT_HOST_THREADS")) n = atoi(e);
        n = std::max(1, std::min(n, 64));
        for (int i = 1; i < n; ++i) workers_.emplace_back([this] { worker(); });
    }
    ~Pool()
    {
        {
            std::lock_guard<std::mutex> g(m_);
            stop_ = true;
            epoch_.fetch_add(1, std::memory_order_release);
        }
        cv_.notify_all();
        for (auto &t : workers_) t.join();
    }
    void run_tasks()
    {
        for (;;) {
            int i = next_.fetch_add(1);
            if (i >= n_tasks_) break;
            (*fn_)(i);
        }
    }
    void worker()
    {
        uint64_t seen = 0;
        for (;;) {
            for (int spin = 0; spin < SPINS && epoch_.load(std::memory_order_acquire) == seen; ++spin) cpu_relax();
            if (epoch_.load(std::memory_order_acquire) == seen) {
                std::unique_lock<std::mutex> g(m_);
                cv_.wait(g, [&] { return epoch_.load(std::memory_order_acquire) != seen; });
            }
            seen = epoch_.load(std::memory_order_acquire);
            if (stop_) return;
            run_tasks();
            if (pending_.fetch_sub(1, std::memory_order_acq_rel) == 1) {
                std::lock_guard<std::mutex> g(m_);
                done_cv_.notify_one();
            }
        }
    }
    std::vector<std::thread> workers_;
    std::mutex caller_m_;   // serialises parallel_for callers
    std::mutex m_;
    std::condition_variable cv_, done_cv_;
    const std::function<void(int)> *fn_ = nullptr;
    std::atomic<int> next_{0}, pending_{0};
    int n_tasks_ = 0;
    std::atomic<uint64_t> epoch_{0};
    std::atomic<bool> stop_{false};
};

// ---------------------------------------------------------------------------------------------------------
// Host-side batch preparation into caller-provided (pinned) staging memory.
// ---------------------------------------------------------------------------------------------------------
struct Staging {
    ReadMeta *meta = nullptr;     // n_reads entries in input order; irregular reads have len = 0 here
    ReadMeta *meta_b = nullptr;   // irregular (byte) reads
    uint32_t *words = nullptr;
    uint8_t *bytes = nullptr;
    uint64_t n_reads = 0, n_words = 0, n_irregular = 0, n_byte_bases = 0, n_bases = 0;
    uint32_t max_len = 0;
    bool sorted = false;
    TaskPlan plan;                // chunk tasks of the long reads that were split (empty for short-read batches)
};

static inline uint64_t words_of(uint64_t len) { return (len + 15) >> 4; }

// Reads [r0, r1) of (seqs, off).  Output offsets are relative to off[r0].  Staging arrays must hold
// (r1-r0) metas (x2), sum(words_of(len)) words and up to n_bases bytes.
static void prepare_reads(const uint8_t *seqs, const uint64_t *off, uint64_t r0, uint64_t r1, uint64_t seq_end, uint64_t lanes, Staging &st)
{
    Pool &pool = Pool::get();
    const uint64_t n_reads = r1 - r0, base0 = off[r0], n_bases = off[r1] - base0;
    st.n_reads = n_reads;
    st.n_bases = n_bases;
    // four slices per thread: the threads take them from a shared counter, so one that is held up (the copy engines and the
    // other ranks share the memory system) does not hold up the chunk
    const int T = (int)std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)pool.size() * 4, n_bases / 65536 + 1));
    std::vector<uint64_t> cut(T + 1), wcount(T + 1, 0), icount(T + 1, 0), ibases(T + 1, 0);
    std::vector<uint32_t> tmax(T, 0);
    for (int t = 0; t <= T; ++t) {   // split by bases
        const uint64_t target = base0 + n_bases * (uint64_t)t / (uint64_t)T;
        cut[t] = (t == T) ? r1 : (uint64_t)(std::lower_bound(off + r0, off + r1, target) - off);
    }
    cut[0] = r0;
    // pass A: word counts per slice (lengths only)
    pool.parallel_for(T, [&](int t) {
        uint64_t w = 0;
        uint32_t mx = 0;
        for (uint64_t i = cut[t]; i < cut[t + 1]; ++i) {
            const uint64_t len = off[i + 1] - off[i];
            w += words_of(len);
            mx = std::max<uint32_t>(mx, (uint32_t)len);
        }
        wcount[t + 1] = w;
        tmax[t] = mx;
    });
    for (int t = 0; t < T; ++t) wcount[t + 1] += wcount[t];
    st.n_words = wcount[T];
    st.max_len = *std::max_element(tmax.begin(), tmax.end());
    // pass B: pack; remember irregular reads per slice
    std::vector<std::vector<uint64_t>> irr(T);
    pool.parallel_for(T, [&](int t) {
        pack_slice(seqs, off, cut[t], cut[t + 1], r0, base0, seq_end, st.words, wcount[t], st.meta, irr[t]);
        uint64_t b = 0;
        for (uint64_t i : irr[t]) b += off[i + 1] - off[i];
        icount[t + 1] = irr[t].size();
        ibases[t + 1] = b;
    });
    for (int t = 0; t < T; ++t) {
        icount[t + 1] += icount[t];
        ibases[t + 1] += ibases[t];
    }
    st.n_irregular = icount[T];
    st.n_byte_bases = ibases[T];
    if (st.n_irregular) {
        pool.parallel_for(T, [&](int t) {
            uint64_t k = icount[t], b = ibases[t];
            for (uint64_t i : irr[t]) {
                const uint64_t len = off[i + 1] - off[i];
                memcpy(st.bytes + b, seqs + off[i], len);
                st.meta_b[k++] = ReadMeta{off[i] - base0, (uint32_t)len, (uint32_t)b};
                b += len;
            }
        });
    }
    // Long reads: cut into chunk tasks when one read's serial chain would dominate the batch (tasks.h).
    st.plan.clear();
    const SplitParams sp0 = SplitParams::from_env();   // read per call: tests switch it with the environment
    if (sp0.wanted(st.max_len, n_bases, lanes)) {
        const SplitParams sp = sp0.adapted(n_bases, lanes);   // chunk length follows a lane's share of the batch (tasks.h)
        for (uint64_t i = 0; i < n_reads; ++i)
            if (st.plan.add(st.meta[i].out_off, st.meta[i].in_off, st.meta[i].len, true, sp)) st.meta[i].len = 0;
        for (uint64_t i = 0; i < st.n_irregular; ++i)
            if (st.plan.add(st.meta_b[i].out_off, st.meta_b[i].in_off, st.meta_b[i].len, false, sp)) st.meta_b[i].len = 0;
        st.plan.finish();
    }
    // Lanes take reads in meta order (global cursor): with reads of very different lengths, start the longest first so
    // that the last lanes to finish are working on short reads, not on a 100 kbp one.
    if (st.max_len >= 1024 && n_reads > 1 && (uint64_t)st.max_len * n_reads > 2 * n_bases) {
        std::sort(st.meta, st.meta + n_reads, [](const ReadMeta &a, const ReadMeta &b) { return a.len > b.len; });
        st.sorted = true;
    }
}

// Device-pack variant of prepare_reads: the host only derives the per-read records from the offsets (no sequence
// byte is read); ReadMeta.out_off doubles as the read's byte offset in the chunk's raw buffer.
static void prepare_offsets_only(const uint64_t *off, uint64_t r0, uint64_t r1, Staging &st)
{
    Pool &pool = Pool::get();
    const uint64_t n_reads = r1 - r0, base0 = off[r0];
    st.n_reads = n_reads;
    st.n_bases = off[r1] - base0;
    st.n_irregular = st.n_byte_bases = 0;
    st.plan.clear();
    const int T = (int)std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)pool.size(), n_reads / 65536 + 1));
    std::vector<uint64_t> wcount(T + 1, 0);
    std::vector<uint32_t> tmax(T, 0);
    pool.parallel_for(T, [&](int t) {
        uint64_t w = 0;
        uint32_t mx = 0;
        for (uint64_t i = r0 + n_reads * t / T; i < r0 + n_reads * (t + 1) / T; ++i) {
            const uint64_t len = off[i + 1] - off[i];
            w += words_of(len);
            mx = std::max<uint32_t>(mx, (uint32_t)len);
        }
        wcount[t + 1] = w;
        tmax[t] = mx;
    });
    for (int t = 0; t < T; ++t) wcount[t + 1] += wcount[t];
    st.n_words = wcount[T];
    st.max_len = *std::max_element(tmax.begin(), tmax.end());
    pool.parallel_for(T, [&](int t) {
        uint64_t w = wcount[t];
        for (uint64_t i = r0 + n_reads * t / T; i < r0 + n_reads * (t + 1) / T; ++i) {
            const uint64_t len = off[i + 1] - off[i];
            st.meta[i - r0] = ReadMeta{off[i] - base0, (uint32_t)len, (uint32_t)w};
            w += words_of(len);
        }
    });
}

// Device copies of a TaskPlan (grown on demand, reused between chunks / calls).
struct DevicePlan {
    ChunkTask *tasks = nullptr, *by_slot = nullptr;
    ChainState *start_state = nullptr, *end_state = nullptr;
    ChainDesc *chains = nullptr;
    size_t cap_tasks = 0, cap_chains = 0;
    void release()
    {
        cudaFree(tasks);
        cudaFree(by_slot);
        cudaFree(start_state);
        cudaFree(end_state);
        cudaFree(chains);
        *this = DevicePlan{};
    }
};

static int upload_plan(const TaskPlan &plan, DevicePlan &dp, BatchView &bv, cudaStream_t stream)
{
    bv.n_tasks = plan.n_tasks;
    bv.n_tasks_b = plan.n_tasks_b;
    bv.n_chains = (uint32_t)plan.chains.size();
    if (plan.tasks.empty()) {
        bv.tasks = bv.by_slot = nullptr;
        bv.start_state = bv.end_state = nullptr;
        bv.chains = nullptr;
        return COLBWT_OK;
    }
    // the scheduling list also holds the whole reads that take part in the longest-first order (no slot, no chain)
    const size_t nt = plan.tasks.size(), ns = plan.by_slot.size(), nc = plan.chains.size();
    if (nt > dp.cap_tasks || nc > dp.cap_chains) {
        CB_CUDA(cudaStreamSynchronize(stream));
        dp.release();
        dp.cap_tasks = nt + nt / 4 + 64;
        dp.cap_chains = nc + nc / 4 + 64;
        CB_CUDA(cudaMalloc(&dp.tasks, dp.cap_tasks * sizeof(ChunkTask)));
        CB_CUDA(cudaMalloc(&dp.by_slot, dp.cap_tasks * sizeof(ChunkTask)));
        CB_CUDA(cudaMalloc(&dp.start_state, dp.cap_tasks * sizeof(ChainState)));
        CB_CUDA(cudaMalloc(&dp.end_state, dp.cap_tasks * sizeof(ChainState)));
        CB_CUDA(cudaMalloc(&dp.chains, dp.cap_chains * sizeof(ChainDesc)));
    }
    CB_CUDA(cudaMemcpyAsync(dp.tasks, plan.tasks.data(), nt * sizeof(ChunkTask), cudaMemcpyHostToDevice, stream));
    if (ns) CB_CUDA(cudaMemcpyAsync(dp.by_slot, plan.by_slot.data(), ns * sizeof(ChunkTask), cudaMemcpyHostToDevice, stream));
    if (nc) CB_CUDA(cudaMemcpyAsync(dp.chains, plan.chains.data(), nc * sizeof(ChainDesc), cudaMemcpyHostToDevice, stream));
    bv.tasks = dp.tasks;
    bv.by_slot = dp.by_slot;
    bv.start_state = dp.start_state;
    bv.end_state = dp.end_state;
    bv.chains = dp.chains;
    return COLBWT_OK;
}

} // namespace colbwt

using namespace colbwt;

// ---------------------------------------------------------------------------------------------------------
// Device-resident batch
// ---------------------------------------------------------------------------------------------------------
struct colbwt_batch {
    colbwt_index *idx = nullptr;
    int slot = 0;
    int pml_width = 2;
    BatchView view{};
    void *d_meta = nullptr, *d_meta_b = nullptr, *d_words = nullptr, *d_bytes = nullptr, *d_pml = nullptr, *d_cid = nullptr;
    unsigned long long *d_cursors = nullptr;
    uint64_t n_bases = 0, n_slot_tasks = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    DevicePlan plan;
};

static int check_width(int pml_width, uint32_t max_len)
{
    if (pml_width != COLBWT_PML_U8 && pml_width != COLBWT_PML_U16 && pml_width != COLBWT_PML_U32) {
        set_error("pml_width must be 1, 2 or 4");
        return COLBWT_ERR_ARG;
    }
    if (pml_width == COLBWT_PML_U8 && max_len > 255) {
        set_error("a read of %u bases does not fit 8-bit PML values; use COLBWT_PML_U16", max_len);
        return COLBWT_ERR_ARG;
    }
    if (pml_width == COLBWT_PML_U16 && max_len > 65535) {
        set_error("a read of %u bases does not fit 16-bit PML values; use COLBWT_PML_U32", max_len);
        return COLBWT_ERR_ARG;
    }
    return COLBWT_OK;
}

extern "C" void colbwt_batch_free(colbwt_batch *b)
{
    if (!b) return;
    if (b->idx && b->slot < (int)b->idx->dev.size()) cudaSetDevice(b->idx->dev[b->slot].device);
    cudaFree(b->d_meta);
    cudaFree(b->d_meta_b);
    cudaFree(b->d_words);
    cudaFree(b->d_bytes);
    cudaFree(b->d_pml);
    cudaFree(b->d_cid);
    cudaFree(b->d_cursors);
    b->plan.release();
    if (b->e0) cudaEventDestroy(b->e0);
    if (b->e1) cudaEventDestroy(b->e1);
    if (b->stream) cudaStreamDestroy(b->stream);
    delete b;
}

extern "C" int colbwt_batch_upload(colbwt_index *idx, int device_slot, const uint8_t *seqs, const uint64_t *off,
                                   uint64_t n_reads, int pml_width, colbwt_batch **out)
{
    if (!idx || !off || !out || (!seqs && n_reads && off[n_reads] > off[0]) || device_slot < 0 || device_slot >= (int)idx->dev.size()) {
        set_error("colbwt_batch_upload: bad argument");
        return COLBWT_ERR_ARG;
    }
    if (n_reads >= (1ull << 32)) {
        set_error("colbwt_batch_upload: at most 2^32-1 reads per batch");
        return COLBWT_ERR_ARG;
    }
    const DeviceTable &dt = idx->dev[device_slot];
    CB_CUDA(cudaSetDevice(dt.device));
    const uint64_t n_bases = off[n_reads] - off[0];
    uint64_t n_words = 0;
    for (uint64_t i = 0; i < n_reads; ++i) {
        if (off[i + 1] < off[i]) {
            set_error("colbwt_batch_upload: offsets must be non-decreasing (read %llu)", (unsigned long long)i);
            return COLBWT_ERR_ARG;
        }
        if (off[i + 1] - off[i] >= 0xFFFFFFFFull) {
            set_error("colbwt_batch_upload: a read of 2^32-1 or more bases is not supported");
            return COLBWT_ERR_ARG;
        }
        n_words += words_of(off[i + 1] - off[i]);
    }
    if (n_words >= (1ull << 32)) {
        set_error("colbwt_batch_upload: batch too large (packed words do not fit 32-bit offsets); split it");
        return COLBWT_ERR_ARG;
    }
    std::unique_ptr<colbwt_batch, void (*)(colbwt_batch *)> b(new colbwt_batch, colbwt_batch_free);
    b->idx = idx;
    b->slot = device_slot;
    b->pml_width = pml_width;
    b->n_bases = n_bases;

    std::vector<ReadMeta> meta(n_reads ? n_reads : 1), meta_b(n_reads ? n_reads : 1);
    std::vector<uint32_t> words(n_words + 2);
    std::vector<uint8_t> bytes(n_bases + 1);
    Staging st;
    st.meta = meta.data();
    st.meta_b = meta_b.data();
    st.words = words.data();
    st.bytes = bytes.data();
    prepare_reads(seqs, off, 0, n_reads, off[n_reads], (uint64_t)dt.sm_count * 1024, st);
    if (int rc = check_width(pml_width, st.max_len)) return rc;
    if (st.n_byte_bases >= (1ull << 32)) {
        set_error("colbwt_batch_upload: more than 4 Gi bases of irregular reads in one batch; split it");
        return COLBWT_ERR_ARG;
    }

    CB_CUDA(cudaStreamCreate(&b->stream));
    CB_CUDA(cudaEventCreate(&b->e0));
    CB_CUDA(cudaEventCreate(&b->e1));
    CB_CUDA(cudaMalloc(&b->d_meta, std::max<uint64_t>(16, n_reads * sizeof(ReadMeta))));
    CB_CUDA(cudaMalloc(&b->d_meta_b, std::max<uint64_t>(16, st.n_irregular * sizeof(ReadMeta))));
    CB_CUDA(cudaMalloc(&b->d_words, (n_words + 2) * 4));
    CB_CUDA(cudaMalloc(&b->d_bytes, std::max<uint64_t>(16, st.n_byte_bases)));
    CB_CUDA(cudaMalloc(&b->d_pml, (n_bases + 8) * (uint64_t)pml_width));
    CB_CUDA(cudaMalloc(&b->d_cid, n_bases + 8));
    CB_CUDA(cudaMalloc(&b->d_cursors, 4 * sizeof(unsigned long long)));
    CB_CUDA(cudaMemcpy(b->d_meta, meta.data(), n_reads * sizeof(ReadMeta), cudaMemcpyHostToDevice));
    CB_CUDA(cudaMemcpy(b->d_meta_b, meta_b.data(), st.n_irregular * sizeof(ReadMeta), cudaMemcpyHostToDevice));
    CB_CUDA(cudaMemcpy(b->d_words, words.data(), n_words * 4, cudaMemcpyHostToDevice));
    CB_CUDA(cudaMemcpy(b->d_bytes, bytes.data(), st.n_byte_bases, cudaMemcpyHostToDevice));
    b->view.meta = (const ReadMeta *)b->d_meta;
    b->view.meta_b = (const ReadMeta *)b->d_meta_b;
    b->view.words = (const uint32_t *)b->d_words;
    b->view.bytes = (const uint8_t *)b->d_bytes;
    b->view.pml = b->d_pml;
    b->view.cid = (uint8_t *)b->d_cid;
    b->view.n_packed = (uint32_t)n_reads;
    b->view.n_bytes = (uint32_t)st.n_irregular;
    if (int rc = upload_plan(st.plan, b->plan, b->view, b->stream)) return rc;
    b->n_slot_tasks = st.plan.by_slot.size();
    CB_CUDA(cudaMemset(b->d_cursors, 0, 4 * sizeof(unsigned long long)));
    CB_CUDA(cudaStreamSynchronize(b->stream));
    *out = b.release();
    return COLBWT_OK;
}

extern "C" int colbwt_batch_launches(const colbwt_batch *b)
{
    if (!b) return 0;
    return ((b->view.n_packed || b->view.n_tasks) ? 1 : 0) + ((b->view.n_bytes || b->view.n_tasks_b) ? 1 : 0) + (b->view.n_chains ? 1 : 0);
}

extern "C" int colbwt_batch_counters(colbwt_batch *b, uint64_t out[4])
{
    if (!b || !out) {
        set_error("colbwt_batch_counters: null argument");
        return COLBWT_ERR_ARG;
    }
    CB_CUDA(cudaSetDevice(b->idx->dev[b->slot].device));
    unsigned long long redone = 0;
    CB_CUDA(cudaMemcpy(&redone, b->d_cursors + 2, sizeof(redone), cudaMemcpyDeviceToHost));
    out[0] = (uint64_t)b->view.n_tasks + b->view.n_tasks_b;
    out[1] = b->view.n_chains;
    out[2] = redone;
    out[3] = b->n_slot_tasks;
    return COLBWT_OK;
}

extern "C" int colbwt_batch_run(colbwt_batch *b, int iters, float *ms_per_iter)
{
    if (!b || iters < 1) {
        set_error("colbwt_batch_run: bad argument");
        return COLBWT_ERR_ARG;
    }
    const DeviceTable &dt = b->idx->dev[b->slot];
    CB_CUDA(cudaSetDevice(dt.device));
    CB_CUDA(cudaEventRecord(b->e0, b->stream));
    for (int i = 0; i < iters; ++i)
        if (int rc = launch_traverse(dt, b->view, b->pml_width, b->d_cursors, b->stream)) return rc;
    CB_CUDA(cudaEventRecord(b->e1, b->stream));
    CB_CUDA(cudaEventSynchronize(b->e1));
    CB_CUDA(cudaGetLastError());
    if (ms_per_iter) {
        float ms = 0;
        CB_CUDA(cudaEventElapsedTime(&ms, b->e0, b->e1));
        *ms_per_iter = ms / (float)iters;
    }
    return COLBWT_OK;
}

extern "C" int colbwt_batch_download(colbwt_batch *b, void *pml, uint8_t *cid)
{
    if (!b) {
        set_error("colbwt_batch_download: null batch");
        return COLBWT_ERR_ARG;
    }
    CB_CUDA(cudaSetDevice(b->idx->dev[b->slot].device));
    if (pml) CB_CUDA(cudaMemcpy(pml, b->d_pml, b->n_bases * (uint64_t)b->pml_width, cudaMemcpyDeviceToHost));
    if (cid) CB_CUDA(cudaMemcpy(cid, b->d_cid, b->n_bases, cudaMemcpyDeviceToHost));
    return COLBWT_OK;
}

extern "C" int colbwt_batch_device_ptrs(colbwt_batch *b, void **pml_dev, void **cid_dev, uint64_t *n_bases)
{
    if (!b) return COLBWT_ERR_ARG;
    if (pml_dev) *pml_dev = b->d_pml;
    if (cid_dev) *cid_dev = b->d_cid;
    if (n_bases) *n_bases = b->n_bases;
    return COLBWT_OK;
}

// ---------------------------------------------------------------------------------------------------------
// Streaming query over host buffers
// ---------------------------------------------------------------------------------------------------------
namespace colbwt {

// Chunks in flight per device.  One chunk's H2D -> kernel -> D2H chain takes ~3x its D2H time, so fewer than ~6 slots
// leaves the D2H copy engine (the end-to-end bottleneck: 2-3 bytes out per base over PCIe) idle between chunks.
static const int SLOTS_PER_DEVICE = getenv("COLBWT_SLOTS") ? std::max(1, std::min(16, atoi(getenv("COLBWT_SLOTS")))) : 6;

// What a call hands back, and how the results cross the link.
enum OutKind {
    OUT_DENSE = 0,              // dense PML + CID arrays, copied as they are (2-5 bytes per base over PCIe)
    OUT_DENSE_VIA_COMPACT = 1,  // dense arrays for the caller, but the link carries the compact form and the host threads expand it
    OUT_COMPACT = 2,            // colbwt_query_compact: the compact form is the result
    OUT_DENSE_CID_COMPACT = 3,  // dense arrays for the caller: PML copied as it is, chain ids (sparse) cross the link in compact form
    OUT_DENSE_MIXED = 4         // dense arrays for the caller: chunks alternate between OUT_DENSE and OUT_DENSE_VIA_COMPACT, so that the
                                // copy engine and the host threads both write results (each chunk runs as one of those two kinds)
};

struct Slot {
    // pinned host staging
    ReadMeta *h_meta = nullptr, *h_meta_b = nullptr;
    uint32_t *h_words = nullptr;
    uint8_t *h_bytes = nullptr, *h_out = nullptr;   // h_out: dense results (pageable destination) or the compact form
    // device
    ReadMeta *d_meta = nullptr, *d_meta_b = nullptr;
    uint32_t *d_words = nullptr;
    uint8_t *d_bytes = nullptr, *d_pml = nullptr, *d_cid = nullptr;
    uint8_t *d_compact = nullptr, *d_values = nullptr;   // compact form of the chunk: fixed part, non-zero chain ids
    uint32_t *d_group_count = nullptr;
    void *d_scan_temp = nullptr;
    unsigned long long *d_cursors = nullptr;
    DevicePlan plan;
    cudaStream_t stream = nullptr;
    cudaEvent_t done = nullptr, fixed_done = nullptr;
    cudaEvent_t tev[4] = {nullptr, nullptr, nullptr, nullptr};   // COLBWT_TRACE=2: H2D start / kernel start / D2H start / end
    // the chunk in flight
    bool pending = false;
    OutKind kind = OUT_DENSE;   // how THIS chunk's results travel (the call's kind, or one of the two kinds a mixed call alternates)
    int phase = 0;              // compact forms: 1 = fixed part on its way, values not yet requested; 2 = everything enqueued
    size_t chunk = 0;           // index into the job's chunk list
    uint64_t out_base = 0, out_bases = 0, n_values = 0, values_off = 0;
};

struct Pipeline {
    colbwt_index *idx = nullptr;
    uint64_t cap_reads = 0, cap_bases = 0, h_out_bytes = 0, compact_fixed_cap = 0;
    size_t scan_temp_bytes = 0;
    int pml_width = 0;
    bool staged_out = false, compact = false;
    std::vector<Slot> slots;   // n_devices * SLOTS_PER_DEVICE
    ~Pipeline()
    {
        for (size_t s = 0; s < slots.size(); ++s) {
            cudaSetDevice(idx->dev[s / SLOTS_PER_DEVICE].device);
            Slot &k = slots[s];
            cudaFreeHost(k.h_meta);
            cudaFreeHost(k.h_meta_b);
            cudaFreeHost(k.h_words);
            cudaFreeHost(k.h_bytes);
            cudaFreeHost(k.h_out);
            cudaFree(k.d_meta);
            cudaFree(k.d_meta_b);
            cudaFree(k.d_words);
            cudaFree(k.d_bytes);
            cudaFree(k.d_pml);
            cudaFree(k.d_cid);
            cudaFree(k.d_compact);
            cudaFree(k.d_values);
            cudaFree(k.d_group_count);
            cudaFree(k.d_scan_temp);
            cudaFree(k.d_cursors);
            k.plan.release();
            if (k.done) cudaEventDestroy(k.done);
            if (k.fixed_done) cudaEventDestroy(k.fixed_done);
            for (auto e : k.tev) if (e) cudaEventDestroy(e);
            if (k.stream) cudaStreamDestroy(k.stream);
        }
    }
};

void destroy_pipeline(Pipeline *p) { delete p; }

static bool is_pinned(const void *p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost;
}

// Allocate (or reuse) the staging pipeline of an index.  A pipeline that is replaced hands its features on (staged dense
// output, compact buffers), so that alternating kinds of calls do not reallocate every time.
static int get_pipeline(colbwt_index *idx, uint64_t chunk_reads, uint64_t chunk_bases, int pml_width, bool staged_out, bool compact, Pipeline **out)
{
    Pipeline *p = idx->pipeline;
    if (p && p->cap_reads >= chunk_reads && p->cap_bases >= chunk_bases && p->pml_width >= pml_width && (p->staged_out || !staged_out) &&
        (p->compact || !compact)) {
        *out = p;
        return COLBWT_OK;
    }
    if (p) {
        chunk_reads = std::max(chunk_reads, p->cap_reads);
        chunk_bases = std::max(chunk_bases, p->cap_bases);
        pml_width = std::max(pml_width, p->pml_width);
        staged_out |= p->staged_out;
        compact |= p->compact;
    }
    delete p;
    idx->pipeline = nullptr;
    std::unique_ptr<Pipeline> pl(new Pipeline);
    pl->idx = idx;
    pl->cap_reads = chunk_reads;
    pl->cap_bases = chunk_bases;
    pl->pml_width = pml_width;
    pl->staged_out = staged_out;
    pl->compact = compact;
    pl->slots.resize(idx->dev.size() * SLOTS_PER_DEVICE);
    const uint64_t cap_words = chunk_bases / 16 + chunk_reads + 2;
    const CompactLayout lay(chunk_bases);
    pl->compact_fixed_cap = lay.fixed_bytes;
    pl->scan_temp_bytes = compact ? compact_scan_temp_bytes(lay.n_groups + 1) : 0;
    pl->h_out_bytes = std::max<uint64_t>(staged_out ? chunk_bases * (uint64_t)(pml_width + 1) + 64 : 0,
                                         compact ? lay.fixed_bytes + chunk_bases + 64 : 0);
    for (size_t s = 0; s < pl->slots.size(); ++s) {
        CB_CUDA(cudaSetDevice(idx->dev[s / SLOTS_PER_DEVICE].device));
        Slot &k = pl->slots[s];
        CB_CUDA(cudaMallocHost(&k.h_meta, chunk_reads * sizeof(ReadMeta)));
        CB_CUDA(cudaMallocHost(&k.h_meta_b, chunk_reads * sizeof(ReadMeta)));
        CB_CUDA(cudaMallocHost(&k.h_words, cap_words * 4));
        CB_CUDA(cudaMallocHost(&k.h_bytes, chunk_bases + 16));
        if (pl->h_out_bytes) CB_CUDA(cudaMallocHost(&k.h_out, pl->h_out_bytes));
        CB_CUDA(cudaMalloc(&k.d_meta, chunk_reads * sizeof(ReadMeta)));
        CB_CUDA(cudaMalloc(&k.d_meta_b, chunk_reads * sizeof(ReadMeta)));
        CB_CUDA(cudaMalloc(&k.d_words, cap_words * 4));
        CB_CUDA(cudaMalloc(&k.d_bytes, chunk_bases + 16));
        CB_CUDA(cudaMalloc(&k.d_pml, (chunk_bases + 8) * (uint64_t)pml_width));
        CB_CUDA(cudaMalloc(&k.d_cid, chunk_bases + 8));
        if (compact) {
            CB_CUDA(cudaMalloc(&k.d_compact, lay.fixed_bytes));
            CB_CUDA(cudaMalloc(&k.d_values, chunk_bases + 16));
            CB_CUDA(cudaMalloc(&k.d_group_count, (lay.n_groups + 2) * sizeof(uint32_t)));
            CB_CUDA(cudaMalloc(&k.d_scan_temp, std::max<size_t>(pl->scan_temp_bytes, 16)));
        }
        CB_CUDA(cudaMalloc(&k.d_cursors, 4 * sizeof(unsigned long long)));
        CB_CUDA(cudaStreamCreateWithFlags(&k.stream, cudaStreamNonBlocking));
        CB_CUDA(cudaEventCreateWithFlags(&k.done, cudaEventDisableTiming));
        CB_CUDA(cudaEventCreateWithFlags(&k.fixed_done, cudaEventDisableTiming));
        for (auto &e : k.tev) CB_CUDA(cudaEventCreate(&e));
    }
    idx->pipeline = pl.release();
    *out = idx->pipeline;
    return COLBWT_OK;
}

// ---- chunk planning: tasks.h (Chunk, Geometry, chunk_geometry, plan_chunks) --------------------------------------------
// One pass over the offsets (longest read, sanity), shared among the packing threads: 10 M reads are 80 MB.
static int scan_offsets(const uint64_t *off, uint64_t n_reads, uint32_t *max_len_out)
{
    const int T = n_reads >= (1u << 16) ? Pool::get().size() : 1;
    std::vector<uint64_t> part_max((size_t)T, 0), part_bad((size_t)T, UINT64_MAX);
    Pool::get().parallel_for(T, [&](int t) {
        const uint64_t a = n_reads * (uint64_t)t / (uint64_t)T, b = n_reads * (uint64_t)(t + 1) / (uint64_t)T;
        uint64_t mx = 0, bad = UINT64_MAX;
        for (uint64_t i = a; i < b; ++i) {
            if (off[i + 1] < off[i]) bad = std::min(bad, i);
            else mx = std::max(mx, off[i + 1] - off[i]);
        }
        part_max[(size_t)t] = mx;
        part_bad[(size_t)t] = bad;
    });
    const uint64_t bad = *std::min_element(part_bad.begin(), part_bad.end());
    if (bad != UINT64_MAX) {
        set_error("colbwt_query: offsets must be non-decreasing (read %llu)", (unsigned long long)bad);
        return COLBWT_ERR_ARG;
    }
    const uint64_t mx = *std::max_element(part_max.begin(), part_max.end());
    if (mx >= 0xFFFFFFFFull) {
        set_error("a read of 2^32-1 or more bases is not supported");
        return COLBWT_ERR_ARG;
    }
    *max_len_out = (uint32_t)mx;
    return COLBWT_OK;
}

// ---- one call ---------------------------------------------------------------------------------------------------
struct QueryJob {
    colbwt_index *idx = nullptr;
    Pipeline *pl = nullptr;
    const uint8_t *seqs = nullptr;
    const uint64_t *off = nullptr;
    uint64_t n_reads = 0;
    int pml_width = 0;
    OutKind kind = OUT_DENSE;
    bool device_pack = false, staged_out = false;
    uint8_t *pml = nullptr, *cid = nullptr;                // dense destination
    uint8_t *cbuf = nullptr;                               // compact destination (colbwt_query_compact)
    size_t ccap = 0;
    bool cbuf_pinned = false;
    colbwt_compact_segment *segments = nullptr;            // directory inside cbuf
    std::atomic<uint64_t> values_cursor{0};                // bump allocator for the value regions inside cbuf
    std::vector<Chunk> chunks;
    std::atomic<uint64_t> h2d_bytes{0}, d2h_bytes{0};      // what the copy engines were asked to move (colbwt_index_last_bytes)
    std::atomic<size_t> next_chunk{0};
    std::atomic<int> rc{COLBWT_OK};
    std::mutex err_m;
    std::string err;
    int trace = 0;
    cudaEvent_t origin = nullptr;                          // COLBWT_TRACE=2, single device: timeline origin
    struct ChunkTimes { float h2d0, k0, d2h0, end; };
    std::vector<ChunkTimes> timeline;
    struct DevTimes { double pack = 0, wait = 0, enqueue = 0, post = 0, loop = 0; uint64_t chunks = 0; };
    std::vector<DevTimes> times;

    void fail(int code)
    {
        std::lock_guard<std::mutex> g(err_m);
        if (rc.load() == COLBWT_OK) {
            err = colbwt_last_error();
            rc.store(code);
        }
    }
};

static inline double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

// Split reads [r0, r1) into T slices of about equal bases; slice t = [cut[t], cut[t+1]).
static std::vector<uint64_t> slice_reads(const uint64_t *off, uint64_t r0, uint64_t r1, int T)
{
    std::vector<uint64_t> cut((size_t)T + 1);
    const uint64_t b0 = off[r0], nb = off[r1] - b0;
    for (int t = 0; t <= T; ++t) cut[(size_t)t] = (t == T) ? r1 : (uint64_t)(std::lower_bound(off + r0, off + r1, b0 + nb * (uint64_t)t / (uint64_t)T) - off);
    cut[0] = r0;
    return cut;
}

// Compact forms, second phase: the fixed part (bit words + prefix) has arrived, so the number of non-zero chain ids is
// known: request exactly those bytes.
static int finish_values(QueryJob &J, Slot &k)
{
    if (k.phase != 1) return COLBWT_OK;
    CB_CUDA(cudaEventSynchronize(k.fixed_done));
    const CompactLayout lay(k.out_bases);
    const bool direct = J.kind == OUT_COMPACT && J.cbuf_pinned;
    const uint8_t *fixed = direct ? J.cbuf + J.segments[k.chunk].match_off : k.h_out;
    k.n_values = reinterpret_cast<const uint32_t *>(fixed + lay.prefix_off)[lay.n_groups];
    uint8_t *dst = k.h_out + J.pl->compact_fixed_cap;
    if (J.kind == OUT_COMPACT) {
        k.values_off = J.values_cursor.fetch_add((k.n_values + 15) & ~15ull);
        if (k.values_off + k.n_values > J.ccap) {
            set_error("colbwt_query_compact: result buffer of %zu bytes is too small (more than %llu needed; colbwt_compact_bound gives the worst case)",
                      J.ccap, (unsigned long long)(k.values_off + k.n_values));
            return COLBWT_ERR_NOMEM;
        }
        if (direct) dst = J.cbuf + k.values_off;
    }
    if (k.n_values) CB_CUDA(cudaMemcpyAsync(dst, k.d_values, k.n_values, cudaMemcpyDeviceToHost, k.stream));
    J.d2h_bytes += k.n_values;
    if (J.trace >= 2) CB_CUDA(cudaEventRecord(k.tev[3], k.stream));
    CB_CUDA(cudaEventRecord(k.done, k.stream));
    k.phase = 2;
    return COLBWT_OK;
}

// Wait for the slot's chunk and finish it on the host: copy out of / expand from the staging area, fill the directory.
static int drain(QueryJob &J, Slot &k, QueryJob::DevTimes &tm)
{
    if (!k.pending) return COLBWT_OK;
    double t0 = now_s();
    if (int rc = finish_values(J, k)) return rc;
    CB_CUDA(cudaEventSynchronize(k.done));
    tm.wait += now_s() - t0;
    if (J.trace >= 2 && J.origin) {
        QueryJob::ChunkTimes ct{};
        cudaEventElapsedTime(&ct.h2d0, J.origin, k.tev[0]);
        cudaEventElapsedTime(&ct.k0, J.origin, k.tev[1]);
        cudaEventElapsedTime(&ct.d2h0, J.origin, k.tev[2]);
        cudaEventElapsedTime(&ct.end, J.origin, k.tev[3]);
        J.timeline.push_back(ct);
    }
    t0 = now_s();
    const Chunk &c = J.chunks[k.chunk];
    Pool &pool = Pool::get();
    if (k.kind == OUT_DENSE) {
        if (J.staged_out) {   // pageable destination: copy out with all packing threads (first-touch page faults included)
            uint8_t *dst[2] = {J.pml + k.out_base * (uint64_t)J.pml_width, J.cid + k.out_base};
            const uint8_t *src[2] = {k.h_out, k.h_out + J.pl->cap_bases * (uint64_t)J.pml_width + 32};
            const uint64_t bytes[2] = {k.out_bases * (uint64_t)J.pml_width, k.out_bases};
            const int T = (int)std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)pool.size(), (bytes[0] + bytes[1]) >> 20));
            pool.parallel_for(2 * T, [&](int i) {
                const int a = i / T, t = i % T;
                const uint64_t lo = bytes[a] * (uint64_t)t / (uint64_t)T, hi = bytes[a] * (uint64_t)(t + 1) / (uint64_t)T;
                memcpy(dst[a] + lo, src[a] + lo, hi - lo);
            });
        }
    } else if (k.out_bases) {
        const CompactLayout lay(k.out_bases);
        if (k.kind == OUT_DENSE_CID_COMPACT) {
            // PML has landed in the caller's array by DMA; the chain ids are rebuilt group by group (2048 bases each)
            const uint8_t *fx = k.h_out, *vals = k.h_out + J.pl->compact_fixed_cap;
            const int T = (int)std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)pool.size() * 4, lay.n_groups / 16));
            pool.parallel_for(T, [&](int t) {
                expand_cid_groups(reinterpret_cast<const uint32_t *>(fx + lay.cid_off), reinterpret_cast<const uint32_t *>(fx + lay.prefix_off), vals, k.out_bases,
                                  lay.n_groups * (uint64_t)t / (uint64_t)T, lay.n_groups * (uint64_t)(t + 1) / (uint64_t)T, J.cid + k.out_base);
            });
        } else if (k.kind == OUT_DENSE_VIA_COMPACT) {
            const uint8_t *fx = k.h_out, *vals = k.h_out + J.pl->compact_fixed_cap;
            const int T = (int)std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)pool.size() * 4, k.out_bases >> 16));
            const std::vector<uint64_t> cut = slice_reads(J.off, c.r0, c.r1, T);
            pool.parallel_for(T, [&](int t) {
                expand_reads(reinterpret_cast<const uint32_t *>(fx + lay.match_off), reinterpret_cast<const uint32_t *>(fx + lay.cid_off),
                             reinterpret_cast<const uint32_t *>(fx + lay.prefix_off), vals, J.off, c.r0, cut[(size_t)t], cut[(size_t)t + 1],
                             J.pml + k.out_base * (uint64_t)J.pml_width, J.pml_width, J.cid + k.out_base);
            });
        } else {
            colbwt_compact_segment &sg = J.segments[k.chunk];
            sg.values_off = k.values_off;
            sg.n_values = k.n_values;
            if (!J.cbuf_pinned) {
                memcpy(J.cbuf + sg.match_off, k.h_out, lay.fixed_bytes);
                memcpy(J.cbuf + sg.values_off, k.h_out + J.pl->compact_fixed_cap, k.n_values);
            }
        }
    }
    k.pending = false;
    k.phase = 0;
    tm.post += now_s() - t0;
    return COLBWT_OK;
}

// Pack chunk c into the slot's staging and enqueue its H2D -> traversal -> D2H chain on the slot's stream.
static int enqueue_chunk(QueryJob &J, int d, Slot &k, size_t ci, QueryJob::DevTimes &tm)
{
    Pipeline &pl = *J.pl;
    const DeviceTable &dt = J.idx->dev[(size_t)d];
    const Chunk &c = J.chunks[ci];
    const uint64_t *off = J.off;
    double t0 = now_s();
    Staging st;
    st.meta = k.h_meta;
    st.meta_b = k.h_meta_b;
    st.words = k.h_words;
    st.bytes = k.h_bytes;
    if (J.device_pack) prepare_offsets_only(off, c.r0, c.r1, st);
    else prepare_reads(J.seqs, off, c.r0, c.r1, off[J.n_reads], (uint64_t)dt.sm_count * 1024 / (uint64_t)std::min<int>(SLOTS_PER_DEVICE, 4), st);
    tm.pack += now_s() - t0;
    t0 = now_s();

    if (J.trace >= 2) CB_CUDA(cudaEventRecord(k.tev[0], k.stream));
    CB_CUDA(cudaMemcpyAsync(k.d_meta, k.h_meta, st.n_reads * sizeof(ReadMeta), cudaMemcpyHostToDevice, k.stream));
    J.h2d_bytes += st.n_reads * sizeof(ReadMeta) + (J.device_pack ? st.n_bases : st.n_words * 4 + st.n_irregular * sizeof(ReadMeta) + st.n_byte_bases) +
                   st.plan.tasks.size() * sizeof(ChunkTask) + st.plan.by_slot.size() * sizeof(ChunkTask) + st.plan.chains.size() * sizeof(ChainDesc);
    if (J.device_pack) {
        CB_CUDA(cudaMemcpyAsync(k.d_bytes, J.seqs + off[c.r0], st.n_bases, cudaMemcpyHostToDevice, k.stream));
        if (int rc = launch_pack(dt, k.d_bytes, k.d_meta, (uint32_t)st.n_reads, k.d_words, k.d_meta_b, (uint32_t *)(k.d_cursors + 3), k.stream)) return rc;
    } else {
        CB_CUDA(cudaMemcpyAsync(k.d_words, k.h_words, st.n_words * 4, cudaMemcpyHostToDevice, k.stream));
        if (st.n_irregular) {
            CB_CUDA(cudaMemcpyAsync(k.d_meta_b, k.h_meta_b, st.n_irregular * sizeof(ReadMeta), cudaMemcpyHostToDevice, k.stream));
            CB_CUDA(cudaMemcpyAsync(k.d_bytes, k.h_bytes, st.n_byte_bases, cudaMemcpyHostToDevice, k.stream));
        }
    }
    BatchView bv{};
    bv.meta = k.d_meta;
    bv.meta_b = k.d_meta_b;
    bv.words = k.d_words;
    bv.bytes = k.d_bytes;
    bv.pml = k.d_pml;
    bv.cid = k.d_cid;
    bv.n_packed = (uint32_t)st.n_reads;
    bv.n_bytes = (uint32_t)st.n_irregular;
    bv.n_bytes_dev = J.device_pack ? (const uint32_t *)(k.d_cursors + 3) : nullptr;
    if (int rc = upload_plan(st.plan, k.plan, bv, k.stream)) return rc;
    if (J.trace >= 2) CB_CUDA(cudaEventRecord(k.tev[1], k.stream));
    if (int rc = launch_traverse(dt, bv, J.pml_width, k.d_cursors, k.stream)) return rc;
    const uint64_t ob = off[c.r0] - off[0];
    k.chunk = ci;
    k.out_base = ob;
    k.out_bases = st.n_bases;
    k.n_values = 0;
    k.values_off = 0;
    k.phase = 2;
    // a mixed call sends every other chunk as it is (the copy engine writes it) and the others in compact form (the host
    // threads write them): both ways of filling the caller's arrays work at the same time
    const OutKind kind = J.kind == OUT_DENSE_MIXED ? ((ci & 1) ? OUT_DENSE : OUT_DENSE_VIA_COMPACT) : J.kind;
    k.kind = kind;
    if (kind == OUT_DENSE) {
        if (J.trace >= 2) CB_CUDA(cudaEventRecord(k.tev[2], k.stream));
        if (J.staged_out) {
            CB_CUDA(cudaMemcpyAsync(k.h_out, k.d_pml, st.n_bases * (uint64_t)J.pml_width, cudaMemcpyDeviceToHost, k.stream));
            CB_CUDA(cudaMemcpyAsync(k.h_out + pl.cap_bases * (uint64_t)J.pml_width + 32, k.d_cid, st.n_bases, cudaMemcpyDeviceToHost, k.stream));
        } else {
            CB_CUDA(cudaMemcpyAsync(J.pml + ob * (uint64_t)J.pml_width, k.d_pml, st.n_bases * (uint64_t)J.pml_width, cudaMemcpyDeviceToHost, k.stream));
            CB_CUDA(cudaMemcpyAsync(J.cid + ob, k.d_cid, st.n_bases, cudaMemcpyDeviceToHost, k.stream));
        }
        J.d2h_bytes += st.n_bases * (uint64_t)(J.pml_width + 1);
        if (J.trace >= 2) CB_CUDA(cudaEventRecord(k.tev[3], k.stream));
        CB_CUDA(cudaEventRecord(k.done, k.stream));
    } else if (st.n_bases && kind == OUT_DENSE_CID_COMPACT) {
        // PML goes out dense (the copy engine writes it where the caller wants it); of the chain ids only the non-zero ones
        // cross the link (bit words + per-group prefix now, the values once their number is known: finish_values)
        const CompactLayout lay(st.n_bases);
        if (int rc = launch_compact(k.d_pml, J.pml_width, k.d_cid, st.n_bases, k.d_compact, k.d_group_count, k.d_scan_temp, pl.scan_temp_bytes, k.d_values, k.stream)) return rc;
        if (J.trace >= 2) CB_CUDA(cudaEventRecord(k.tev[2], k.stream));
        CB_CUDA(cudaMemcpyAsync(k.h_out + lay.cid_off, k.d_compact + lay.cid_off, lay.fixed_bytes - lay.cid_off, cudaMemcpyDeviceToHost, k.stream));
        CB_CUDA(cudaEventRecord(k.fixed_done, k.stream));
        CB_CUDA(cudaMemcpyAsync(J.pml + ob * (uint64_t)J.pml_width, k.d_pml, st.n_bases * (uint64_t)J.pml_width, cudaMemcpyDeviceToHost, k.stream));
        J.d2h_bytes += lay.fixed_bytes - lay.cid_off + st.n_bases * (uint64_t)J.pml_width;
        k.phase = 1;
    } else if (st.n_bases) {
        // dense results stay in HBM; the link carries the compact form (compact.cu) in two steps: the fixed-size part
        // now, the non-zero chain ids once their number is known (finish_values)
        const CompactLayout lay(st.n_bases);
        if (int rc = launch_compact(k.d_pml, J.pml_width, k.d_cid, st.n_bases, k.d_compact, k.d_group_count, k.d_scan_temp, pl.scan_temp_bytes, k.d_values, k.stream)) return rc;
        if (J.trace >= 2) CB_CUDA(cudaEventRecord(k.tev[2], k.stream));
        uint8_t *dst = (J.kind == OUT_COMPACT && J.cbuf_pinned) ? J.cbuf + J.segments[ci].match_off : k.h_out;
        CB_CUDA(cudaMemcpyAsync(dst, k.d_compact, lay.fixed_bytes, cudaMemcpyDeviceToHost, k.stream));
        J.d2h_bytes += lay.fixed_bytes;
        CB_CUDA(cudaEventRecord(k.fixed_done, k.stream));
        k.phase = 1;
    } else {
        if (J.trace >= 2) {
            CB_CUDA(cudaEventRecord(k.tev[2], k.stream));
            CB_CUDA(cudaEventRecord(k.tev[3], k.stream));
        }
        CB_CUDA(cudaEventRecord(k.done, k.stream));
    }
    k.pending = true;
    tm.enqueue += now_s() - t0;
    return COLBWT_OK;
}

// The feeder of one device: takes the next unassigned chunk, waits for a free slot, packs, enqueues.  With several
// replicas each device has its own feeder thread (reads are independent: src/pml_query.cpp:74-86), so a device never
// waits for another one's drain; the packing threads are shared (Pool).
static void feed_device(QueryJob &J, int d)
{
    QueryJob::DevTimes &tm = J.times[(size_t)d];
    const double t_begin = now_s();
    if (cudaSetDevice(J.idx->dev[(size_t)d].device) != cudaSuccess) {
        set_error("cudaSetDevice(%d) failed", J.idx->dev[(size_t)d].device);
        J.fail(COLBWT_ERR_CUDA);
        return;
    }
    Slot *slots = &J.pl->slots[(size_t)d * SLOTS_PER_DEVICE];
    // Compact forms: the values of a chunk are requested VALUES_LAG chunks after it was enqueued, so that waiting for its
    // fixed part (= for its kernel) never leaves the GPU without queued work; measured on C2 with a lag of 1: the feeder and
    // the GPU took turns (52 ms per 1.5 Gbases against 30 ms of kernels).
    const uint64_t VALUES_LAG = (uint64_t)std::max(1, std::min(3, SLOTS_PER_DEVICE - 2));
    int rc = COLBWT_OK;
    for (uint64_t local = 0; rc == COLBWT_OK && J.rc.load() == COLBWT_OK; ++local) {
        const size_t ci = J.next_chunk.fetch_add(1);
        if (ci >= J.chunks.size()) break;
        Slot &k = slots[local % (uint64_t)SLOTS_PER_DEVICE];
        if ((rc = drain(J, k, tm)) != COLBWT_OK) break;
        if ((rc = enqueue_chunk(J, d, k, ci, tm)) != COLBWT_OK) break;
        if (local >= VALUES_LAG) {
            const double t0 = now_s();
            rc = finish_values(J, slots[(local - VALUES_LAG) % (uint64_t)SLOTS_PER_DEVICE]);
            tm.wait += now_s() - t0;
        }
        ++tm.chunks;
    }
    tm.loop = now_s() - t_begin;
    for (int s = 0; s < SLOTS_PER_DEVICE && rc == COLBWT_OK; ++s) rc = drain(J, slots[s], tm);
    if (rc != COLBWT_OK) J.fail(rc);
}

struct CallSpec {
    OutKind kind = OUT_DENSE;              // OUT_DENSE: the library may still pick OUT_DENSE_VIA_COMPACT
    void *pml = nullptr;
    int pml_width = 0;
    uint8_t *cid = nullptr;
    void *cbuf = nullptr;
    size_t ccap = 0;
    size_t *cused = nullptr;
};

static int query_impl(colbwt_index *idx, const uint8_t *seqs, const uint64_t *off, uint64_t n_reads, const CallSpec &cs)
{
    const bool compact_api = cs.kind == OUT_COMPACT;
    const uint64_t total_bases = off[n_reads] - off[0];
    uint32_t max_len = 0;
    if (int rc = scan_offsets(off, n_reads, &max_len)) return rc;
    if (!seqs && total_bases) {
        set_error("colbwt_query: null sequence buffer");
        return COLBWT_ERR_ARG;
    }
    // the compact form carries no lengths, so the narrowest PML type that holds the longest read serves on the device
    const int pml_width = compact_api ? (max_len < 256 ? COLBWT_PML_U8 : max_len < 65536 ? COLBWT_PML_U16 : COLBWT_PML_U32) : cs.pml_width;
    if (int rc = check_width(pml_width, max_len)) return rc;
    // pageable result buffers are reached through one pinned staging area per slot
    const bool pageable_out = !compact_api && !(is_pinned(cs.pml) && is_pinned(cs.cid));
    Geometry geo = chunk_geometry(total_bases, n_reads, max_len, pageable_out ? (uint64_t)(pml_width + 1) : 0);
    std::vector<Chunk> planned;
    plan_chunks(off, n_reads, geo, planned);   // also leaves the staging capacities (largest chunk) in geo

    std::lock_guard<std::mutex> guard(idx->query_mutex);
    // ---- where to pack and how to cross the link: measured, not guessed -------------------------------------------
    // Mode bit 0: reads packed on the device (raw bytes DMA-ed as they are; needs a pinned input and no read that will be
    // split into chunk tasks).  Mode bit 1 (dense results only): compact transport.  One B200 fed by 16 cores: host packing
    // 19.5-22.7 Gbases/s end to end, device packing 17.0 (the extra 1 B/base of H2D slows the D2H stream on the shared
    // link); with few threads per process (one rank per GPU on a 32-core box) the host packer falls behind the link.  The
    // first large call follows a rule, every other allowed mode is then tried once, later calls take the fastest (5 %
    // hysteresis in favour of the rule).  COLBWT_DEVICE_PACK / COLBWT_COMPACT_D2H pin a bit.
    const SplitParams sp_query = SplitParams::from_env();
    const bool can_device_pack = is_pinned(seqs) && max_len < sp_query.min_len;
    const bool large_call = total_bases >= (64ull << 20);
    // mode = packing bit + 2 x transport; transport 0 = dense copies, 1 = compact form expanded by the host threads,
    // 2 = PML dense + chain ids compact (needs pinned result arrays: the copy engine writes the PML in place)
    // 3 = chunks alternate between transports 0 and 1 (copy engine and host threads fill the arrays together; pinned arrays)
    constexpr int N_MODES = 8;
    uint32_t allowed = 0;
    for (int m = 0; m < N_MODES; ++m) {
        if ((m & 1) && !can_device_pack) continue;
        if ((m >> 1) && compact_api) continue;
        if ((m >> 1) >= 2 && pageable_out) continue;
        allowed |= 1u << m;
    }
    auto pin_field = [&](const char *name, bool packing) {
        const char *e = getenv(name);
        if (!e) return;
        uint32_t keep = 0;
        for (int m = 0; m < N_MODES; ++m)
            if ((packing ? (m & 1) : (m >> 1)) == atoi(e)) keep |= 1u << m;
        if (allowed & keep) allowed &= keep;
    };
    pin_field("COLBWT_DEVICE_PACK", true);
    pin_field("COLBWT_COMPACT_D2H", false);
    double *rates = idx->mode_rate[compact_api ? 1 : 0];
    // Starting guess.  Plenty of host threads (one process, one GPU): pack on the host, copy the dense arrays.  Few threads per
    // process (one rank per GPU on a shared host: 4 each on the 8-GPU box) means a link shared with the other ranks as well:
    // fewest bytes and least host work first -- raw reads packed on the device, compact transport (8 ranks on C2: 12.3
    // Gbases/s per rank that way against 7.5-10.3 for the other modes, profiles/r2/r2_bench_c2_n8_trace.log).  With all ranks
    // trying modes at the same time a single sample is noisy, so the rule keeps its place unless clearly beaten.
    int rule = (Pool::get().size() < 8) ? 3 : 0;
    const int mode = choose_mode(rule, allowed, rates, N_MODES, large_call);   // tasks.h
    const bool device_pack = (mode & 1) != 0;
    const OutKind kind = compact_api ? OUT_COMPACT
                         : ((mode >> 1) == 1 ? OUT_DENSE_VIA_COMPACT : (mode >> 1) == 2 ? OUT_DENSE_CID_COMPACT : (mode >> 1) == 3 ? OUT_DENSE_MIXED : OUT_DENSE);
    idx->last_packing = device_pack ? 1 : 0;
    idx->last_transport = compact_api ? 1 : (mode >> 1);

    Pipeline *plp = nullptr;
    const bool staged_out = pageable_out && kind == OUT_DENSE;
    const Pipeline *before = idx->pipeline;
    // compact buffers are allocated as soon as a call could use them, so that trying that mode later does not reallocate
    const bool want_compact = kind != OUT_DENSE || (allowed & ~0x3u) != 0;
    if (int rc = get_pipeline(idx, geo.chunk_reads, geo.chunk_bases, pml_width, staged_out, want_compact, &plp)) return rc;
    const bool fresh_pipeline = plp != before;   // this call pays for the staging allocations: not a timing sample

    QueryJob J;
    J.idx = idx;
    J.pl = plp;
    J.seqs = seqs;
    J.off = off;
    J.n_reads = n_reads;
    J.pml_width = pml_width;
    J.kind = kind;
    J.device_pack = device_pack;
    J.staged_out = staged_out;
    J.pml = (uint8_t *)cs.pml;
    J.cid = cs.cid;
    J.trace = getenv("COLBWT_TRACE") ? atoi(getenv("COLBWT_TRACE")) : 0;
    J.chunks = std::move(planned);
    const int n_dev = (int)idx->dev.size();
    J.times.resize((size_t)n_dev);
    if (compact_api) {
        // result buffer: header | directory | fixed parts | value regions (bump-allocated as the chunks complete)
        J.cbuf = (uint8_t *)cs.cbuf;
        J.ccap = cs.ccap;
        J.cbuf_pinned = is_pinned(cs.cbuf);
        uint64_t at = sizeof(colbwt_compact_header) + J.chunks.size() * sizeof(colbwt_compact_segment);
        at = (at + 15) & ~15ull;
        const uint64_t dir_end = at;
        if (dir_end > J.ccap) {
            set_error("colbwt_query_compact: result buffer of %zu bytes cannot hold the directory of %zu segments", J.ccap, J.chunks.size());
            return COLBWT_ERR_NOMEM;
        }
        J.segments = reinterpret_cast<colbwt_compact_segment *>(J.cbuf + sizeof(colbwt_compact_header));
        for (size_t i = 0; i < J.chunks.size(); ++i) {
            const Chunk &c = J.chunks[i];
            const CompactLayout lay(off[c.r1] - off[c.r0]);
            colbwt_compact_segment sg{};
            sg.first_read = c.r0;
            sg.n_reads = c.r1 - c.r0;
            sg.first_base = off[c.r0] - off[0];
            sg.n_bases = off[c.r1] - off[c.r0];
            sg.match_off = at + lay.match_off;
            sg.cid_off = at + lay.cid_off;
            sg.prefix_off = at + lay.prefix_off;
            J.segments[i] = sg;
            at += sg.n_bases ? lay.fixed_bytes : 0;
        }
        if (at > J.ccap) {
            set_error("colbwt_query_compact: result buffer of %zu bytes is too small (%llu needed before any chain id)", J.ccap, (unsigned long long)at);
            return COLBWT_ERR_NOMEM;
        }
        J.values_cursor.store(at);
    }
    struct OriginEvent {   // COLBWT_TRACE=2 (single device): timeline origin
        cudaEvent_t e = nullptr;
        ~OriginEvent() { if (e) cudaEventDestroy(e); }
    } origin;
    if (J.trace >= 2 && n_dev == 1) {
        CB_CUDA(cudaSetDevice(idx->dev[0].device));
        CB_CUDA(cudaEventCreate(&origin.e));
        CB_CUDA(cudaEventRecord(origin.e, plp->slots[0].stream));
        J.origin = origin.e;
    } else if (J.trace >= 2) {
        J.trace = 1;   // events of different devices cannot be compared: per-device summaries only
    }

    const double t_begin = now_s();
    if (n_dev == 1) {
        feed_device(J, 0);
    } else {
        std::vector<std::thread> feeders;
        for (int d = 0; d < n_dev; ++d) feeders.emplace_back([&J, d] { feed_device(J, d); });
        for (auto &t : feeders) t.join();
    }
    const double t_total = now_s() - t_begin;
    if (J.rc.load() != COLBWT_OK) {
        set_error("%s", J.err.c_str());
        return J.rc.load();
    }
    if (compact_api) {
        colbwt_compact_header h{};
        h.magic = COLBWT_COMPACT_MAGIC;
        h.n_segments = J.chunks.size();
        h.n_reads = n_reads;
        h.n_bases = total_bases;
        h.bytes_used = J.values_cursor.load();
        memcpy(J.cbuf, &h, sizeof(h));
        if (cs.cused) *cs.cused = (size_t)h.bytes_used;
    }
    if (J.trace >= 2 && !J.timeline.empty()) {
        float h2d = 0, ker = 0, d2h = 0;
        for (auto &c : J.timeline) { h2d += c.k0 - c.h2d0; ker += c.d2h0 - c.k0; d2h += c.end - c.d2h0; }
        fprintf(stderr, "[colbwt_query] stream time per stage, summed over chunks: H2D %.1f ms, kernel %.1f ms, D2H %.1f ms; last chunk ends at %.1f ms\n",
                h2d, ker, d2h, J.timeline.back().end);
        for (size_t i = 0; i < J.timeline.size(); i += std::max<size_t>(1, J.timeline.size() / 8))
            fprintf(stderr, "    chunk %3zu: H2D %.2f..%.2f kernel ..%.2f D2H ..%.2f ms\n", i, J.timeline[i].h2d0, J.timeline[i].k0, J.timeline[i].d2h0, J.timeline[i].end);
    }
    if (J.trace) {
        for (int d = 0; d < n_dev; ++d) {
            const QueryJob::DevTimes &tm = J.times[(size_t)d];
            fprintf(stderr, "[colbwt_query] device %d: %llu chunks: pack %.1f ms, wait-for-slot %.1f ms, enqueue %.1f ms, host post-processing %.1f ms, loop %.1f ms\n",
                    idx->dev[(size_t)d].device, (unsigned long long)tm.chunks, tm.pack * 1e3, tm.wait * 1e3, tm.enqueue * 1e3, tm.post * 1e3, tm.loop * 1e3);
        }
        fprintf(stderr, "[colbwt_query] %zu chunks, %.1f Mbases, total %.1f ms = %.2f Gbases/s (%s, packing on the %s, %d host threads)\n", J.chunks.size(),
                total_bases / 1e6, t_total * 1e3, total_bases / t_total / 1e9,
                kind == OUT_DENSE ? (staged_out ? "dense via staging" : "dense into pinned buffers") : kind == OUT_COMPACT ? "compact result"
                : kind == OUT_DENSE_CID_COMPACT ? "dense, PML copied + chain ids compact"
                : kind == OUT_DENSE_MIXED ? "dense, chunks alternate between copies and compact transport" : "dense, compact transport",
                device_pack ? "device" : "host", Pool::get().size());
    }
    if (large_call && !fresh_pipeline) rates[mode] = (double)total_bases / std::max(1e-9, t_total);
    idx->last_h2d_bytes = J.h2d_bytes.load();
    idx->last_d2h_bytes = J.d2h_bytes.load();
    return COLBWT_OK;
}

static int run_query(colbwt_index *idx, const uint8_t *seqs, const uint64_t *off, uint64_t n_reads, const CallSpec &cs)
{
    const int rc = query_impl(idx, seqs, off, n_reads, cs);
    if (rc != COLBWT_OK && rc != COLBWT_ERR_ARG && idx->pipeline) {
        // a failed copy or launch may leave chunks in flight: wait, then drop the staging pipeline so that the next
        // call starts from a clean one
        std::string msg = colbwt_last_error();
        std::lock_guard<std::mutex> guard(idx->query_mutex);
        for (auto &d : idx->dev) {
            cudaSetDevice(d.device);
            cudaDeviceSynchronize();
        }
        cudaGetLastError();
        destroy_pipeline(idx->pipeline);
        idx->pipeline = nullptr;
        set_error("%s", msg.c_str());
    }
    return rc;
}

} // namespace colbwt

extern "C" int colbwt_query(colbwt_index *idx, const uint8_t *seqs, const uint64_t *off, uint64_t n_reads,
                            void *pml, int pml_width, uint8_t *cid)
{
    if (!idx || !off || !pml || !cid || idx->dev.empty()) {
        set_error("colbwt_query: bad argument");
        return COLBWT_ERR_ARG;
    }
    if (n_reads == 0) return COLBWT_OK;
    CallSpec cs;
    cs.kind = OUT_DENSE;
    cs.pml = pml;
    cs.pml_width = pml_width;
    cs.cid = cid;
    return run_query(idx, seqs, off, n_reads, cs);
}

extern "C" int colbwt_query_compact(colbwt_index *idx, const uint8_t *seqs, const uint64_t *off, uint64_t n_reads,
                                    void *result, size_t capacity, size_t *bytes_used)
{
    if (!idx || !off || !result || capacity < sizeof(colbwt_compact_header) || idx->dev.empty()) {
        set_error("colbwt_query_compact: bad argument");
        return COLBWT_ERR_ARG;
    }
    if (n_reads == 0) {
        colbwt_compact_header h{};
        h.magic = COLBWT_COMPACT_MAGIC;
        h.bytes_used = sizeof(h);
        memcpy(result, &h, sizeof(h));
        if (bytes_used) *bytes_used = sizeof(h);
        return COLBWT_OK;
    }
    CallSpec cs;
    cs.kind = OUT_COMPACT;
    cs.cbuf = result;
    cs.ccap = capacity;
    cs.cused = bytes_used;
    return run_query(idx, seqs, off, n_reads, cs);
}

extern "C" size_t colbwt_compact_bound(const uint64_t *off, uint64_t n_reads)
{
    if (!off || n_reads == 0) return sizeof(colbwt_compact_header);
    uint32_t max_len = 0;
    if (scan_offsets(off, n_reads, &max_len) != COLBWT_OK) return 0;
    const uint64_t total = off[n_reads] - off[0];
    Geometry geo = chunk_geometry(total, n_reads, max_len, 0);
    std::vector<Chunk> chunks;
    plan_chunks(off, n_reads, geo, chunks);
    uint64_t at = (sizeof(colbwt_compact_header) + chunks.size() * sizeof(colbwt_compact_segment) + 15) & ~15ull;
    for (const Chunk &c : chunks) {
        const uint64_t nb = off[c.r1] - off[c.r0];
        at += (nb ? CompactLayout(nb).fixed_bytes : 0) + ((nb + 15) & ~15ull);   // every chain id non-zero
    }
    return (size_t)at;
}

extern "C" int colbwt_compact_expand(const void *result, const uint64_t *off, uint64_t n_reads, void *pml, int pml_width, uint8_t *cid)
{
    if (!result || !off || !cid) {
        set_error("colbwt_compact_expand: null argument");
        return COLBWT_ERR_ARG;
    }
    colbwt_compact_header h;
    memcpy(&h, result, sizeof(h));
    if (h.magic != COLBWT_COMPACT_MAGIC || h.n_reads != n_reads || (n_reads && h.n_bases != off[n_reads] - off[0])) {
        set_error("colbwt_compact_expand: not a compact result of these %llu reads", (unsigned long long)n_reads);
        return COLBWT_ERR_ARG;
    }
    if (pml && pml_width != COLBWT_PML_U8 && pml_width != COLBWT_PML_U16 && pml_width != COLBWT_PML_U32) {
        set_error("pml_width must be 1, 2 or 4");
        return COLBWT_ERR_ARG;
    }
    const uint8_t *buf = (const uint8_t *)result;
    const colbwt_compact_segment *segs = reinterpret_cast<const colbwt_compact_segment *>(buf + sizeof(colbwt_compact_header));
    Pool &pool = Pool::get();
    for (uint64_t s = 0; s < h.n_segments; ++s) {
        const colbwt_compact_segment &sg = segs[s];
        if (sg.first_read + sg.n_reads > n_reads || off[sg.first_read] - off[0] != sg.first_base || off[sg.first_read + sg.n_reads] - off[sg.first_read] != sg.n_bases) {
            set_error("colbwt_compact_expand: segment %llu does not match the offsets", (unsigned long long)s);
            return COLBWT_ERR_ARG;
        }
        if (!sg.n_bases) continue;
        if (!pml) {   // chain ids only: group by group, no read boundary matters
            const CompactLayout lay(sg.n_bases);
            const int T = (int)std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)pool.size() * 4, lay.n_groups / 16));
            pool.parallel_for(T, [&](int t) {
                expand_cid_groups(reinterpret_cast<const uint32_t *>(buf + sg.cid_off), reinterpret_cast<const uint32_t *>(buf + sg.prefix_off), buf + sg.values_off,
                                  sg.n_bases, lay.n_groups * (uint64_t)t / (uint64_t)T, lay.n_groups * (uint64_t)(t + 1) / (uint64_t)T, cid + sg.first_base);
            });
            continue;
        }
        const int T = (int)std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)pool.size() * 4, sg.n_bases >> 16));
        const std::vector<uint64_t> cut = slice_reads(off, sg.first_read, sg.first_read + sg.n_reads, T);
        std::atomic<bool> too_long{false};
        pool.parallel_for(T, [&](int t) {
            if (pml_width < 4)
                for (uint64_t i = cut[(size_t)t]; i < cut[(size_t)t + 1]; ++i)
                    if (off[i + 1] - off[i] > (pml_width == 1 ? 255u : 65535u)) too_long = true;
            if (too_long) return;
            expand_reads(reinterpret_cast<const uint32_t *>(buf + sg.match_off), reinterpret_cast<const uint32_t *>(buf + sg.cid_off),
                         reinterpret_cast<const uint32_t *>(buf + sg.prefix_off), buf + sg.values_off, off, sg.first_read, cut[(size_t)t], cut[(size_t)t + 1],
                         (uint8_t *)pml + sg.first_base * (uint64_t)pml_width, pml_width, cid + sg.first_base);
        });
        if (too_long) {
            set_error("colbwt_compact_expand: a read does not fit %d-byte PML values", pml_width);
            return COLBWT_ERR_ARG;
        }
    }
    return COLBWT_OK;
}

extern "C" int colbwt_index_last_packing(const colbwt_index *idx) { return idx ? idx->last_packing : -1; }
extern "C" int colbwt_index_last_transport(const colbwt_index *idx) { return idx ? idx->last_transport : -1; }
extern "C" int colbwt_index_last_bytes(const colbwt_index *idx, uint64_t *h2d_bytes, uint64_t *d2h_bytes)
{
    if (!idx) return COLBWT_ERR_ARG;
    if (h2d_bytes) *h2d_bytes = idx->last_h2d_bytes;
    if (d2h_bytes) *d2h_bytes = idx->last_d2h_bytes;
    return COLBWT_OK;
}

extern "C" void *colbwt_host_alloc(size_t bytes)
{
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable) != cudaSuccess) {
        set_error("cudaHostAlloc(%zu) failed: %s", bytes, cudaGetErrorString(cudaGetLastError()));
        return nullptr;
    }
    return p;
}

extern "C" void colbwt_host_free(void *p)
{
    if (p) cudaFreeHost(p);
}
