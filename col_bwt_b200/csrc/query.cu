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
    void parallel_for(int n_tasks, const std::function<void(int)> &fn)
    {
        if (n_tasks <= 1 || workers_.empty()) {
            for (int i = 0; i < n_tasks; ++i) fn(i);
            return;
        }
        {
            std::lock_guard<std::mutex> g(m_);
            fn_ = &fn;
            n_tasks_ = n_tasks;
            next_.store(0);
            pending_ = (int)workers_.size();
            ++epoch_;
        }
        cv_.notify_all();
        run_tasks();
        std::unique_lock<std::mutex> g(m_);
        done_cv_.wait(g, [&] { return pending_ == 0; });
        fn_ = nullptr;
    }

private:
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
            ++epoch_;
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
            {
                std::unique_lock<std::mutex> g(m_);
                cv_.wait(g, [&] { return epoch_ != seen; });
                seen = epoch_;
                if (stop_) return;
            }
            run_tasks();
            {
                std::lock_guard<std::mutex> g(m_);
                if (--pending_ == 0) done_cv_.notify_one();
            }
        }
    }
    std::vector<std::thread> workers_;
    std::mutex m_;
    std::condition_variable cv_, done_cv_;
    const std::function<void(int)> *fn_ = nullptr;
    std::atomic<int> next_{0};
    int n_tasks_ = 0, pending_ = 0;
    uint64_t epoch_ = 0;
    bool stop_ = false;
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
    const int T = (int)std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)pool.size(), n_bases / 65536 + 1));
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
    const SplitParams sp = SplitParams::from_env();   // read per call: tests switch it with the environment
    if (sp.wanted(st.max_len, n_bases, lanes)) {
        for (uint64_t i = 0; i < n_reads; ++i)
            if (st.meta[i].len >= sp.min_len) {
                st.plan.add_read(st.meta[i].out_off, st.meta[i].in_off, st.meta[i].len, true, sp);
                st.meta[i].len = 0;
            }
        for (uint64_t i = 0; i < st.n_irregular; ++i)
            if (st.meta_b[i].len >= sp.min_len) {
                st.plan.add_read(st.meta_b[i].out_off, st.meta_b[i].in_off, st.meta_b[i].len, false, sp);
                st.meta_b[i].len = 0;
            }
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
    if (plan.by_slot.empty()) {
        bv.tasks = bv.by_slot = nullptr;
        bv.start_state = bv.end_state = nullptr;
        bv.chains = nullptr;
        return COLBWT_OK;
    }
    const size_t nt = plan.by_slot.size(), nc = plan.chains.size();
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
    CB_CUDA(cudaMemcpyAsync(dp.by_slot, plan.by_slot.data(), nt * sizeof(ChunkTask), cudaMemcpyHostToDevice, stream));
    CB_CUDA(cudaMemcpyAsync(dp.chains, plan.chains.data(), nc * sizeof(ChainDesc), cudaMemcpyHostToDevice, stream));
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
    uint64_t n_bases = 0;
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
    CB_CUDA(cudaStreamSynchronize(b->stream));
    *out = b.release();
    return COLBWT_OK;
}

extern "C" int colbwt_batch_launches(const colbwt_batch *b)
{
    if (!b) return 0;
    return ((b->view.n_packed || b->view.n_tasks) ? 1 : 0) + ((b->view.n_bytes || b->view.n_tasks_b) ? 1 : 0) + (b->view.n_chains ? 1 : 0);
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

struct Slot {
    // pinned host staging
    ReadMeta *h_meta = nullptr, *h_meta_b = nullptr;
    uint32_t *h_words = nullptr;
    uint8_t *h_bytes = nullptr, *h_out = nullptr;
    // device
    ReadMeta *d_meta = nullptr, *d_meta_b = nullptr;
    uint32_t *d_words = nullptr;
    uint8_t *d_bytes = nullptr, *d_pml = nullptr, *d_cid = nullptr;
    unsigned long long *d_cursors = nullptr;
    DevicePlan plan;
    cudaStream_t stream = nullptr;
    cudaEvent_t done = nullptr;
    cudaEvent_t tev[4] = {nullptr, nullptr, nullptr, nullptr};   // COLBWT_TRACE=2: H2D start / kernel start / D2H start / end
    // pending copy-out when the caller's buffers are not pinned
    bool pending = false;
    uint64_t out_base = 0, out_bases = 0;
};

struct Pipeline {
    colbwt_index *idx = nullptr;
    uint64_t cap_reads = 0, cap_bases = 0;
    int pml_width = 0;
    bool staged_out = false;
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
            cudaFree(k.d_cursors);
            k.plan.release();
            if (k.done) cudaEventDestroy(k.done);
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

// Allocate (or reuse) the staging pipeline of an index.
static int get_pipeline(colbwt_index *idx, uint64_t chunk_reads, uint64_t chunk_bases, int pml_width, bool staged_out, Pipeline **out)
{
    Pipeline *p = idx->pipeline;
    if (p && p->cap_reads >= chunk_reads && p->cap_bases >= chunk_bases && p->pml_width >= pml_width && (p->staged_out || !staged_out)) {
        *out = p;
        return COLBWT_OK;
    }
    delete p;
    idx->pipeline = nullptr;
    std::unique_ptr<Pipeline> pl(new Pipeline);
    pl->idx = idx;
    pl->cap_reads = chunk_reads;
    pl->cap_bases = chunk_bases;
    pl->pml_width = pml_width;
    pl->staged_out = staged_out;
    pl->slots.resize(idx->dev.size() * SLOTS_PER_DEVICE);
    const uint64_t cap_words = chunk_bases / 16 + chunk_reads + 2;
    const uint64_t out_bytes = chunk_bases * (uint64_t)(pml_width + 1) + 64;
    for (size_t s = 0; s < pl->slots.size(); ++s) {
        CB_CUDA(cudaSetDevice(idx->dev[s / SLOTS_PER_DEVICE].device));
        Slot &k = pl->slots[s];
        CB_CUDA(cudaMallocHost(&k.h_meta, chunk_reads * sizeof(ReadMeta)));
        CB_CUDA(cudaMallocHost(&k.h_meta_b, chunk_reads * sizeof(ReadMeta)));
        CB_CUDA(cudaMallocHost(&k.h_words, cap_words * 4));
        CB_CUDA(cudaMallocHost(&k.h_bytes, chunk_bases + 16));
        if (staged_out) CB_CUDA(cudaMallocHost(&k.h_out, out_bytes));
        CB_CUDA(cudaMalloc(&k.d_meta, chunk_reads * sizeof(ReadMeta)));
        CB_CUDA(cudaMalloc(&k.d_meta_b, chunk_reads * sizeof(ReadMeta)));
        CB_CUDA(cudaMalloc(&k.d_words, cap_words * 4));
        CB_CUDA(cudaMalloc(&k.d_bytes, chunk_bases + 16));
        CB_CUDA(cudaMalloc(&k.d_pml, (chunk_bases + 8) * (uint64_t)pml_width));
        CB_CUDA(cudaMalloc(&k.d_cid, chunk_bases + 8));
        CB_CUDA(cudaMalloc(&k.d_cursors, 4 * sizeof(unsigned long long)));
        CB_CUDA(cudaStreamCreateWithFlags(&k.stream, cudaStreamNonBlocking));
        CB_CUDA(cudaEventCreateWithFlags(&k.done, cudaEventDisableTiming));
        for (auto &e : k.tev) CB_CUDA(cudaEventCreate(&e));
    }
    idx->pipeline = pl.release();
    *out = idx->pipeline;
    return COLBWT_OK;
}

} // namespace colbwt

static int query_impl(colbwt_index *idx, const uint8_t *seqs, const uint64_t *off, uint64_t n_reads, void *pml, int pml_width, uint8_t *cid);

extern "C" int colbwt_query(colbwt_index *idx, const uint8_t *seqs, const uint64_t *off, uint64_t n_reads,
                            void *pml, int pml_width, uint8_t *cid)
{
    const int rc = query_impl(idx, seqs, off, n_reads, pml, pml_width, cid);
    if (rc != COLBWT_OK && rc != COLBWT_ERR_ARG && idx && idx->pipeline) {
        // a failed copy or launch may leave chunks in flight: wait, then drop the staging pipeline so that the next
        // call starts from a clean one
        std::lock_guard<std::mutex> guard(idx->query_mutex);
        for (auto &d : idx->dev) {
            cudaSetDevice(d.device);
            cudaDeviceSynchronize();
        }
        cudaGetLastError();
        destroy_pipeline(idx->pipeline);
        idx->pipeline = nullptr;
    }
    return rc;
}

static int query_impl(colbwt_index *idx, const uint8_t *seqs, const uint64_t *off, uint64_t n_reads, void *pml, int pml_width, uint8_t *cid)
{
    if (!idx || !off || !pml || !cid || idx->dev.empty()) {
        set_error("colbwt_query: bad argument");
        return COLBWT_ERR_ARG;
    }
    if (n_reads == 0) return COLBWT_OK;
    const uint64_t total_bases = off[n_reads] - off[0];
    // chunk geometry
    uint64_t chunk_bases = 96ull << 20;   // measured on C2 (profiles/r1/e2e_chunk_sweep.log): 16/24/32/48/96/128/192 M -> 83.6/85.0/81.1/76.3/71.8-73.2/72.0/73.1 ms
    if (const char *e = getenv("COLBWT_CHUNK_BASES")) chunk_bases = std::max<uint64_t>(1024, strtoull(e, nullptr, 10));
    // one pass over the offsets (longest read, sanity), shared among the packing threads: 10 M reads are 80 MB
    uint32_t max_len = 0;
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
        max_len = (uint32_t)std::min<uint64_t>(*std::max_element(part_max.begin(), part_max.end()), 0xFFFFFFFFull);
    }
    if (!seqs && total_bases) {
        set_error("colbwt_query: null sequence buffer");
        return COLBWT_ERR_ARG;
    }
    if (max_len == 0xFFFFFFFFu) {
        set_error("a read of 2^32-1 or more bases is not supported");
        return COLBWT_ERR_ARG;
    }
    if (int rc = check_width(pml_width, max_len)) return rc;
    // Long reads: a chunk must still hold enough reads to occupy the lanes (a lane works on one read at a time and
    // six chunks are in flight), so the chunk grows with the mean read length: 32 Ki reads per chunk, up to 512 Mbases.
    if (!getenv("COLBWT_CHUNK_BASES"))
        chunk_bases = std::max<uint64_t>(chunk_bases, std::min<uint64_t>(512ull << 20, (total_bases / n_reads) * 32768));
    // pageable result buffers are reached through one pinned staging area per slot: keep each under 256 MB
    const bool pageable_out = !(is_pinned(pml) && is_pinned(cid));
    if (pageable_out && !getenv("COLBWT_CHUNK_BASES")) chunk_bases = std::min<uint64_t>(chunk_bases, (256ull << 20) / (uint64_t)(pml_width + 1));
    chunk_bases = std::max<uint64_t>(chunk_bases, max_len);
    chunk_bases = std::min<uint64_t>(chunk_bases, std::max<uint64_t>(total_bases, 16));
    // a chunk also ends after chunk_bases/32 reads, which bounds the meta staging (16 B per read) for very short reads
    const uint64_t chunk_reads = std::min<uint64_t>(std::max<uint64_t>(chunk_bases / 32, 1024), n_reads);

    const int n_dev = (int)idx->dev.size();
    std::lock_guard<std::mutex> guard(idx->query_mutex);
    Pipeline *plp = nullptr;
    const bool staged_out = pageable_out;
    const Pipeline *before = idx->pipeline;
    if (int rc = get_pipeline(idx, chunk_reads, chunk_bases, pml_width, staged_out, &plp)) return rc;
    const bool fresh_pipeline = plp != before;   // this call pays for the staging allocations: not a timing sample
    Pipeline &pl = *plp;
    // Pack on the device when the input can be DMA-ed as it is (pinned) and no read will be split into chunk tasks:
    // H2D has headroom (the link is busy in the other direction), host cores often do not (one process per GPU).
    const SplitParams sp_query = SplitParams::from_env();
    const char *dp_env = getenv("COLBWT_DEVICE_PACK");
    // Where to pack is measured, not guessed.  One B200 fed by 16 cores: host 19.5-22.7 Gbases/s end to end, device 17.0 (the
    // extra 1 B/base of H2D slows the D2H stream on the shared link); two ranks with 8 threads each: 37.7 against 34.8;
    // with fewer threads per process the host packer (2.6 Gbases/s per thread) falls behind the link.  So: the first large
    // timed call follows a rule (device iff fewer than 4 packing threads), the next tries the other way once, later calls
    // take the faster of the two (5 % hysteresis) and re-try the other one every 64th large call.  COLBWT_DEVICE_PACK
    // pins the choice.
    const bool can_device_pack = is_pinned(seqs) && max_len < sp_query.min_len;
    const bool large_call = total_bases >= (64ull << 20);
    bool device_pack = false;
    if (dp_env) {
        device_pack = atoi(dp_env) != 0 && can_device_pack;
    } else if (can_device_pack) {
        device_pack = choose_device_pack(Pool::get().size() < 4 ? 1 : 0, idx->pack_rate, large_call, idx->large_calls);   // tasks.h
    }
    idx->last_packing = device_pack ? 1 : 0;

    static const int trace = getenv("COLBWT_TRACE") ? atoi(getenv("COLBWT_TRACE")) : 0;
    struct ChunkTimes { float h2d0, k0, d2h0, end; };
    std::vector<ChunkTimes> timeline;
    cudaEvent_t ev_origin = nullptr;
    if (trace >= 2) {
        CB_CUDA(cudaSetDevice(idx->dev[0].device));
        CB_CUDA(cudaEventCreate(&ev_origin));
        CB_CUDA(cudaEventRecord(ev_origin, pl.slots[0].stream));
    }
    uint8_t *pml_out = (uint8_t *)pml;
    auto drain = [&](Slot &k) -> int {   // wait for the slot's previous chunk; copy out if staged
        if (!k.pending) return COLBWT_OK;
        CB_CUDA(cudaEventSynchronize(k.done));
        if (trace >= 2) {
            ChunkTimes ct{};
            cudaEventElapsedTime(&ct.h2d0, ev_origin, k.tev[0]);
            cudaEventElapsedTime(&ct.k0, ev_origin, k.tev[1]);
            cudaEventElapsedTime(&ct.d2h0, ev_origin, k.tev[2]);
            cudaEventElapsedTime(&ct.end, ev_origin, k.tev[3]);
            timeline.push_back(ct);
        }
        if (staged_out) {   // pageable destination: copy out with all packing threads (first-touch page faults included)
            uint8_t *dst[2] = {pml_out + k.out_base * (uint64_t)pml_width, cid + k.out_base};
            const uint8_t *src[2] = {k.h_out, k.h_out + pl.cap_bases * (uint64_t)pml_width + 32};
            const uint64_t bytes[2] = {k.out_bases * (uint64_t)pml_width, k.out_bases};
            const int T = (int)std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)Pool::get().size(), (bytes[0] + bytes[1]) >> 20));
            Pool::get().parallel_for(2 * T, [&](int i) {
                const int a = i / T, t = i % T;
                const uint64_t lo = bytes[a] * (uint64_t)t / (uint64_t)T, hi = bytes[a] * (uint64_t)(t + 1) / (uint64_t)T;
                memcpy(dst[a] + lo, src[a] + lo, hi - lo);
            });
        }
        k.pending = false;
        return COLBWT_OK;
    };

    auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    double t_pack = 0, t_drain = 0, t_enqueue = 0;
    const double t_begin = now();
    uint64_t r0 = 0, chunk_no = 0;
    while (r0 < n_reads) {
        // next chunk [r0, r1): at most chunk_bases bases and chunk_reads reads, at least one read
        uint64_t r1 = (uint64_t)(std::upper_bound(off + r0, off + n_reads + 1, off[r0] + chunk_bases) - off) - 1;
        r1 = std::min(r1, r0 + chunk_reads);
        if (r1 <= r0) r1 = r0 + 1;
        const int d = (int)(chunk_no % (uint64_t)n_dev);
        Slot &k = pl.slots[(size_t)d * SLOTS_PER_DEVICE + (size_t)((chunk_no / (uint64_t)n_dev) % SLOTS_PER_DEVICE)];
        const DeviceTable &dt = idx->dev[d];
        CB_CUDA(cudaSetDevice(dt.device));
        double t0 = now();
        if (int rc = drain(k)) return rc;
        t_drain += now() - t0;
        t0 = now();

        Staging st;
        st.meta = k.h_meta;
        st.meta_b = k.h_meta_b;
        st.words = k.h_words;
        st.bytes = k.h_bytes;
        if (device_pack) prepare_offsets_only(off, r0, r1, st);
        else prepare_reads(seqs, off, r0, r1, off[n_reads], (uint64_t)dt.sm_count * 1024 / (uint64_t)std::min<int>(SLOTS_PER_DEVICE, 4), st);
        t_pack += now() - t0;
        t0 = now();

        if (trace >= 2) CB_CUDA(cudaEventRecord(k.tev[0], k.stream));
        CB_CUDA(cudaMemcpyAsync(k.d_meta, k.h_meta, st.n_reads * sizeof(ReadMeta), cudaMemcpyHostToDevice, k.stream));
        if (device_pack) {
            CB_CUDA(cudaMemcpyAsync(k.d_bytes, seqs + off[r0], st.n_bases, cudaMemcpyHostToDevice, k.stream));
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
        bv.n_bytes_dev = device_pack ? (const uint32_t *)(k.d_cursors + 3) : nullptr;
        if (int rc = upload_plan(st.plan, k.plan, bv, k.stream)) return rc;
        if (trace >= 2) CB_CUDA(cudaEventRecord(k.tev[1], k.stream));
        if (int rc = launch_traverse(dt, bv, pml_width, k.d_cursors, k.stream)) return rc;
        if (trace >= 2) CB_CUDA(cudaEventRecord(k.tev[2], k.stream));
        const uint64_t ob = off[r0] - off[0];
        if (staged_out) {
            CB_CUDA(cudaMemcpyAsync(k.h_out, k.d_pml, st.n_bases * (uint64_t)pml_width, cudaMemcpyDeviceToHost, k.stream));
            CB_CUDA(cudaMemcpyAsync(k.h_out + pl.cap_bases * (uint64_t)pml_width + 32, k.d_cid, st.n_bases, cudaMemcpyDeviceToHost, k.stream));
        } else {
            CB_CUDA(cudaMemcpyAsync(pml_out + ob * (uint64_t)pml_width, k.d_pml, st.n_bases * (uint64_t)pml_width, cudaMemcpyDeviceToHost, k.stream));
            CB_CUDA(cudaMemcpyAsync(cid + ob, k.d_cid, st.n_bases, cudaMemcpyDeviceToHost, k.stream));
        }
        if (trace >= 2) CB_CUDA(cudaEventRecord(k.tev[3], k.stream));
        CB_CUDA(cudaEventRecord(k.done, k.stream));
        k.pending = true;
        k.out_base = ob;
        k.out_bases = st.n_bases;
        r0 = r1;
        ++chunk_no;
        t_enqueue += now() - t0;
    }
    const double t_loop = now() - t_begin;
    for (size_t s = 0; s < pl.slots.size(); ++s) {
        CB_CUDA(cudaSetDevice(idx->dev[s / SLOTS_PER_DEVICE].device));
        if (int rc = drain(pl.slots[s])) return rc;
    }
    if (trace >= 2 && n_dev == 1) {
        float h2d = 0, ker = 0, d2h = 0;
        for (auto &c : timeline) { h2d += c.k0 - c.h2d0; ker += c.d2h0 - c.k0; d2h += c.end - c.d2h0; }
        fprintf(stderr, "[colbwt_query] stream time per stage, summed over chunks: H2D %.1f ms, kernel %.1f ms, D2H %.1f ms; last chunk ends at %.1f ms\n",
                h2d, ker, d2h, timeline.empty() ? 0.f : timeline.back().end);
        for (size_t i = 0; i < timeline.size(); i += std::max<size_t>(1, timeline.size() / 8))
            fprintf(stderr, "    chunk %3zu: H2D %.2f..%.2f kernel ..%.2f D2H ..%.2f ms\n", i, timeline[i].h2d0, timeline[i].k0, timeline[i].d2h0, timeline[i].end);
    }
    if (ev_origin) cudaEventDestroy(ev_origin);
    if (trace)
        fprintf(stderr, "[colbwt_query] %llu chunks, %.1f Mbases: pack %.1f ms, wait-for-slot %.1f ms, enqueue %.1f ms, loop %.1f ms, total %.1f ms (%s outputs, packing on the %s)\n",
                (unsigned long long)chunk_no, total_bases / 1e6, t_pack * 1e3, t_drain * 1e3, t_enqueue * 1e3, t_loop * 1e3,
                (now() - t_begin) * 1e3, staged_out ? "staged" : "pinned", device_pack ? "device" : "host");
    if (large_call && !dp_env && can_device_pack && !fresh_pipeline) {
        idx->pack_rate[device_pack ? 1 : 0] = (double)total_bases / std::max(1e-9, now() - t_begin);
        ++idx->large_calls;
    }
    return COLBWT_OK;
}

extern "C" int colbwt_index_last_packing(const colbwt_index *idx) { return idx ? idx->last_packing : -1; }

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
