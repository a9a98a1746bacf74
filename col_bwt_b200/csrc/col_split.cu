// col_split.cu -- multi-MUM sub-run marking on the GPU.
//
// Replaces src/col_split.cpp + include/col_split.hpp + the FL_table it walks (include/ds/FL_table.hpp):
//   FL_table(heads, lengths) / compute_table      FL_table.hpp:86-131, 343-379   -> build_fl_table (stable radix sort by
//                                                                                    character + scans + binary search)
//   col_split::split, FL_loop, FL_range           col_split.hpp:54-136, 226-247  -> frontier kernels below
//   col_split::find_col_runs                      col_split.hpp:258-338          -> resolve_marks (host, sequential sweep)
//   col_split::save / serialize_col_runs          col_split.hpp:138-157, 374-390 -> write_outputs
//
// The reference walks every multi-MUM forward through the text with FL steps and, every `split_rate` columns, marks
// the BWT range that the MUM's N suffixes occupy.  All MUMs are independent, so the walk is done level by level on
// the GPU: a *frontier* holds the current ranges of every MUM still alive; one kernel launch = one FL step of all of
// them (a range that straddles F-runs falls into pieces; in tunnel mode that ends the MUM, col_split.hpp:81,101).
// The marks are sorted into the reference's visiting order (MUM, column) and resolved on the host exactly as
// find_col_runs does (a priority-queue sweep over at most a few million marks).
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <cstring>
#include <queue>
#include <string>

#include "internal.h"

namespace colbwt {

struct FlTable {            // F-runs (BWT runs in stable character order) as a move structure, device pointers
    uint64_t *idx = nullptr;    // start of the F-run in [0, n)
    uint64_t *len = nullptr;
    uint32_t *dest = nullptr;   // F-run that holds the image of the run's first position
    uint64_t *off = nullptr;    // offset of that image inside dest
    uint64_t runs = 0, n = 0;
};

struct Range {              // one piece of a MUM's current BWT range
    uint32_t mum, interval;
    uint64_t offset;
    uint32_t height, pad;
};

struct Mark {
    uint64_t start;
    uint32_t mum, col;      // visiting order of the reference: MUM number, then column j
    uint32_t height, pad;
};

__global__ void k_fl_gather(const uint32_t *__restrict__ order, const uint64_t *__restrict__ lens, uint64_t runs, uint64_t *f_len)
{
    const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < runs) f_len[k] = lens[order[k]];
}

__device__ __forceinline__ uint32_t run_of(const uint64_t *idx, uint64_t runs, uint64_t pos)
{
    uint64_t lo = 0, hi = runs;                 // last run whose start is <= pos
    while (hi - lo > 1) {
        const uint64_t mid = (lo + hi) >> 1;
        if (idx[mid] <= pos) lo = mid; else hi = mid;
    }
    return (uint32_t)lo;
}

__global__ void k_fl_columns(const uint32_t *__restrict__ order, const uint64_t *__restrict__ l_start, FlTable t)
{
    const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= t.runs) return;
    const uint64_t p = l_start[order[k]];       // where this run's characters sit in L (FL_table.hpp:362-376)
    const uint32_t d = run_of(t.idx, t.runs, p);
    t.dest[k] = d;
    t.off[k] = p - t.idx[d];
}

__global__ void k_map_char(const uint8_t *__restrict__ heads, uint64_t runs, uint8_t *mapped)
{
    const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < runs) mapped[k] = (heads[k] <= 1 || heads[k] >= 128) ? 1 : heads[k];   // `char c; if (c <= TERMINATOR)` (FL_table.hpp:99)
}

// FL (FL_table.hpp:227-238): image of (run, offset) with fast-forward.
__device__ __forceinline__ void fl_step(const FlTable &t, uint32_t run, uint64_t offset, uint32_t *o_run, uint64_t *o_off)
{
    uint32_t d = t.dest[run];
    uint64_t o = t.off[run] + offset;
    while (o >= t.len[d]) o -= t.len[d++];
    *o_run = d;
    *o_off = o;
}

// FL_range (col_split.hpp:226-247) of one range: appends its pieces to `out`; returns the piece count through `n_out`
// so that tunnel mode can drop a MUM whose range fell apart.
__device__ __forceinline__ uint32_t expand(const FlTable &t, const Range &r, Range *out, unsigned long long *n_out, bool tunnels)
{
    // count pieces first (tunnel mode must not emit anything for a range that splits)
    uint32_t pieces = 0;
    {
        uint32_t run = r.interval, h = r.height;
        uint64_t o = r.offset;
        while (h > 0) {
            const uint64_t room = t.len[run] - o;
            const uint32_t covered = room < h ? (uint32_t)room : h;
            ++pieces;
            h -= covered;
            o = 0;
            ++run;
        }
    }
    if (tunnels && pieces > 1) return pieces;
    const unsigned long long base = atomicAdd(n_out, (unsigned long long)pieces);
    uint32_t run = r.interval, h = r.height, k = 0;
    uint64_t o = r.offset;
    while (h > 0) {
        const uint64_t room = t.len[run] - o;
        const uint32_t covered = room < h ? (uint32_t)room : h;
        Range q;
        q.mum = r.mum;
        fl_step(t, run, o, &q.interval, &q.offset);
        q.height = covered;
        q.pad = k;
        out[base + k++] = q;
        h -= covered;
        o = 0;
        ++run;
    }
    return pieces;
}

__global__ void k_init_frontier(FlTable t, const uint64_t *__restrict__ mum_pos, uint32_t n_mums, uint32_t num_docs, Range *out,
                                unsigned long long *n_out, int tunnels)
{
    const uint32_t m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= n_mums) return;
    Range r;
    r.mum = m;
    r.interval = run_of(t.idx, t.runs, mum_pos[m]);     // the F-run that holds the MUM's first row (col_split.hpp:72-75)
    r.offset = mum_pos[m] - t.idx[r.interval];
    r.height = num_docs;
    r.pad = 0;
    expand(t, r, out, n_out, tunnels != 0);
}

// One level of FL_loop (col_split.hpp:77-101): column j of every live range.
__global__ void k_frontier_step(FlTable t, const Range *__restrict__ cur, unsigned long long n_cur, const uint64_t *__restrict__ mum_len,
                                uint32_t j, uint32_t split_rate, Range *next, unsigned long long *n_next, Mark *marks,
                                unsigned long long *n_marks, int tunnels)
{
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_cur) return;
    const Range r = cur[i];
    if ((uint64_t)j >= mum_len[r.mum]) return;          // this MUM has been walked to its end
    if (j % split_rate == 0) {
        const unsigned long long k = atomicAdd(n_marks, 1ull);
        marks[k] = Mark{t.idx[r.interval] + r.offset, r.mum, j, r.height, 0};
    }
    if ((uint64_t)j + 1 < mum_len[r.mum]) expand(t, r, next, n_next, tunnels != 0);
}

namespace {
struct DevFree {
    std::vector<void *> p;
    ~DevFree() { for (void *q : p) cudaFree(q); }
    template <typename T> cudaError_t alloc(T **out, size_t n)
    {
        cudaError_t e = cudaMalloc((void **)out, std::max<size_t>(16, n * sizeof(T)));
        if (e == cudaSuccess) p.push_back(*out);
        return e;
    }
};

inline uint32_t bin_id(uint64_t id) { return id >= 256 ? (uint32_t)(id % 255) + 1 : (uint32_t)id; }   // col_split.hpp:222-224
} // namespace

// find_col_runs (col_split.hpp:258-338) on the marks in visiting order.  Returns set bits of col_runs + id per bit.
static void resolve_marks(uint64_t n, const std::vector<uint64_t> &run_starts, std::vector<Mark> &marks, bool mode_all,
                          std::vector<uint64_t> &out_pos, std::vector<uint8_t> &out_id)
{
    // second pass of split(): one (id, height) per distinct start (collect_ids, col_split.hpp:114-127)
    std::stable_sort(marks.begin(), marks.end(), [](const Mark &a, const Mark &b) { return a.mum != b.mum ? a.mum < b.mum : a.col < b.col; });
    struct Slot { uint64_t start; uint32_t id, height; uint64_t seq; };
    std::vector<Slot> slots(marks.size());
    for (size_t i = 0; i < marks.size(); ++i) slots[i] = Slot{marks[i].start, bin_id((uint64_t)marks[i].mum + 1), marks[i].height, i};
    std::stable_sort(slots.begin(), slots.end(), [](const Slot &a, const Slot &b) { return a.start != b.start ? a.start < b.start : a.seq < b.seq; });
    std::vector<Slot> uniq;
    for (size_t i = 0; i < slots.size();) {
        size_t e = i;
        uint32_t id = 0, h = 0;
        for (; e < slots.size() && slots[e].start == slots[i].start; ++e) {
            if (mode_all) {                     // keep the taller one; on ties the one visited first
                if (!(h >= slots[e].height)) id = slots[e].id;
                h = std::max(h, slots[e].height);
            } else {                            // tunnel mode: the last visit wins
                id = slots[e].id;
                h = slots[e].height;
            }
        }
        uniq.push_back(Slot{slots[i].start, id, h, 0});
        i = e;
    }
    // sweep
    struct Open { uint64_t end, start; uint32_t id; };
    auto later = [](const Open &a, const Open &b) { return a.end != b.end ? a.end > b.end : a.start > b.start; };
    std::priority_queue<Open, std::vector<Open>, decltype(later)> open(later);
    size_t run_cursor = 0;                      // next BWT run head not yet emitted
    uint32_t last_id = 0;
    auto update_bwt_pos = [&](uint64_t idx, uint32_t id) {
        while (run_cursor < run_starts.size() && run_starts[run_cursor] < idx) {
            out_pos.push_back(run_starts[run_cursor]);
            out_id.push_back((uint8_t)last_id);
            ++run_cursor;
        }
        if (run_cursor < run_starts.size() && run_starts[run_cursor] == idx) ++run_cursor;
        last_id = id;
    };
    auto update_col_ranges = [&](uint64_t idx) {
        while (!open.empty() && open.top().end <= idx) {
            const Open e = open.top();
            open.pop();
            if (open.size() == 1 && open.top().end > e.end) {
                update_bwt_pos(e.end, open.top().id);
                out_pos.push_back(e.end);
                out_id.push_back((uint8_t)open.top().id);
            } else if (open.empty() && e.end < idx) {
                update_bwt_pos(e.end, 0);
                out_pos.push_back(e.end);
                out_id.push_back(0);
            }
        }
    };
    for (const Slot &s : uniq) {
        update_col_ranges(s.start);
        open.push(Open{s.start + s.height, s.start, s.id});
        if (open.size() == 1 && s.id > 0) {
            update_bwt_pos(s.start, s.id);
            out_pos.push_back(s.start);
            out_id.push_back((uint8_t)s.id);
        }
    }
    update_col_ranges(n);
    update_bwt_pos(n, 0);
}

static bool read_all(const std::string &path, std::vector<uint8_t> &out)
{
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) return false;
    fseek(f, 0, SEEK_END);
    const long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    out.resize(sz > 0 ? (size_t)sz : 0);
    const bool ok = out.empty() || fread(out.data(), 1, out.size(), f) == out.size();
    fclose(f);
    return ok;
}

static void unpack_u40(const uint8_t *raw, size_t count, std::vector<uint64_t> &out)
{
    out.resize(count);
    for (size_t i = 0; i < count; ++i) {
        uint64_t v = 0;
        memcpy(&v, raw + 5 * i, 5);
        out[i] = v;
    }
}

} // namespace colbwt

using namespace colbwt;

extern "C" int colbwt_col_split(const char *prefix, int mode_all, int split_rate, int device, uint64_t *n_set_bits, uint64_t *n_marked)
{
    if (!prefix || split_rate < 1) {
        set_error("colbwt_col_split: bad argument");
        return COLBWT_ERR_ARG;
    }
    const std::string p(prefix);
    std::vector<uint8_t> heads, raw;
    std::vector<uint64_t> lens, mums;
    if (!read_all(p + ".bwt.heads", heads) || heads.empty()) {
        set_error("cannot read %s.bwt.heads", prefix);
        return COLBWT_ERR_IO;
    }
    if (!read_all(p + ".bwt.len", raw) || raw.size() != heads.size() * 5) {
        set_error("cannot read %s.bwt.len (or it disagrees with .bwt.heads)", prefix);
        return COLBWT_ERR_IO;
    }
    unpack_u40(raw.data(), heads.size(), lens);
    if (!read_all(p + ".col_mums", raw) || raw.size() < 5) {      // col_split.cpp:90-106: num_docs, then (len, pos) pairs
        set_error("cannot read %s.col_mums", prefix);
        return COLBWT_ERR_IO;
    }
    unpack_u40(raw.data(), raw.size() / 5, mums);
    const uint32_t num_docs = (uint32_t)mums[0];
    const uint32_t n_mums = (uint32_t)((mums.size() - 1) / 2);
    std::vector<uint64_t> mum_len(n_mums), mum_pos(n_mums);
    uint64_t max_len = 0;
    for (uint32_t i = 0; i < n_mums; ++i) {
        mum_len[i] = mums[1 + 2 * i];
        mum_pos[i] = mums[2 + 2 * i];
        max_len = std::max(max_len, mum_len[i]);
    }
    const uint64_t runs = heads.size();
    std::vector<uint64_t> run_starts(runs);
    uint64_t n = 0;
    for (uint64_t i = 0; i < runs; ++i) {
        run_starts[i] = n;
        n += lens[i];
    }
    int have = 0;
    if (cudaGetDeviceCount(&have) != cudaSuccess || have == 0) {
        cudaGetLastError();
        set_error("no CUDA device available; this library has no CPU path");
        return COLBWT_ERR_CUDA;
    }
    CB_CUDA(cudaSetDevice(device));

    // ---- FL table on the device ------------------------------------------------------------------------------
    DevFree mem;
    uint8_t *d_heads, *d_mapped, *d_keys_out;
    uint64_t *d_lens, *d_lstart, *d_mum_len, *d_mum_pos;
    uint32_t *d_iota, *d_order;
    FlTable t;
    t.runs = runs;
    t.n = n;
    CB_CUDA(mem.alloc(&d_heads, runs));
    CB_CUDA(mem.alloc(&d_mapped, runs));
    CB_CUDA(mem.alloc(&d_keys_out, runs));
    CB_CUDA(mem.alloc(&d_lens, runs));
    CB_CUDA(mem.alloc(&d_lstart, runs));
    CB_CUDA(mem.alloc(&d_iota, runs));
    CB_CUDA(mem.alloc(&d_order, runs));
    CB_CUDA(mem.alloc(&t.idx, runs));
    CB_CUDA(mem.alloc(&t.len, runs));
    CB_CUDA(mem.alloc(&t.dest, runs));
    CB_CUDA(mem.alloc(&t.off, runs));
    CB_CUDA(mem.alloc(&d_mum_len, n_mums));
    CB_CUDA(mem.alloc(&d_mum_pos, n_mums));
    CB_CUDA(cudaMemcpy(d_heads, heads.data(), runs, cudaMemcpyHostToDevice));
    CB_CUDA(cudaMemcpy(d_lens, lens.data(), runs * 8, cudaMemcpyHostToDevice));
    CB_CUDA(cudaMemcpy(d_lstart, run_starts.data(), runs * 8, cudaMemcpyHostToDevice));
    CB_CUDA(cudaMemcpy(d_mum_len, mum_len.data(), (size_t)n_mums * 8, cudaMemcpyHostToDevice));
    CB_CUDA(cudaMemcpy(d_mum_pos, mum_pos.data(), (size_t)n_mums * 8, cudaMemcpyHostToDevice));
    const unsigned gr = (unsigned)((runs + 255) / 256);
    k_map_char<<<gr, 256>>>(d_heads, runs, d_mapped);
    std::vector<uint32_t> iota(runs);
    for (uint64_t i = 0; i < runs; ++i) iota[i] = (uint32_t)i;
    CB_CUDA(cudaMemcpy(d_iota, iota.data(), runs * 4, cudaMemcpyHostToDevice));
    size_t tb = 0;
    CB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, (const uint8_t *)d_mapped, d_keys_out, (const uint32_t *)d_iota, d_order, (int64_t)runs, 0, 8));
    size_t tb2 = 0;
    CB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb2, t.len, t.idx, (int64_t)runs));
    uint8_t *d_temp;
    CB_CUDA(mem.alloc(&d_temp, std::max(tb, tb2) + 256));
    tb = std::max(tb, tb2) + 256;
    CB_CUDA(cub::DeviceRadixSort::SortPairs(d_temp, tb, (const uint8_t *)d_mapped, d_keys_out, (const uint32_t *)d_iota, d_order, (int64_t)runs, 0, 8));
    k_fl_gather<<<gr, 256>>>(d_order, d_lens, runs, t.len);
    tb = std::max(tb, tb2) + 256;
    CB_CUDA(cub::DeviceScan::ExclusiveSum(d_temp, tb, t.len, t.idx, (int64_t)runs));
    k_fl_columns<<<gr, 256>>>(d_order, d_lstart, t);
    CB_CUDA(cudaGetLastError());

    // ---- frontier walk ---------------------------------------------------------------------------------------------
    const uint64_t cap_ranges = mode_all ? (uint64_t)n_mums * num_docs + 1024 : (uint64_t)n_mums + 1024;
    uint64_t total_cols = 0;
    for (uint32_t i = 0; i < n_mums; ++i) total_cols += (mum_len[i] + split_rate - 1) / split_rate;
    const uint64_t cap_marks = total_cols * (mode_all ? num_docs : 1) + 1024;
    Range *d_cur, *d_next;
    Mark *d_marks;
    unsigned long long *d_counts;   // [0] cur, [1] next, [2] marks
    CB_CUDA(mem.alloc(&d_cur, cap_ranges));
    CB_CUDA(mem.alloc(&d_next, cap_ranges));
    CB_CUDA(mem.alloc(&d_marks, cap_marks));
    CB_CUDA(mem.alloc(&d_counts, 4));
    CB_CUDA(cudaMemset(d_counts, 0, 32));
    if (n_mums) k_init_frontier<<<(n_mums + 255) / 256, 256>>>(t, d_mum_pos, n_mums, num_docs, d_cur, d_counts + 0, !mode_all);
    unsigned long long h_counts[4] = {0, 0, 0, 0};
    for (uint32_t j = 0; j < max_len; ++j) {
        CB_CUDA(cudaMemcpy(h_counts, d_counts, 32, cudaMemcpyDeviceToHost));
        if (h_counts[0] == 0) break;
        if (h_counts[0] > cap_ranges || h_counts[2] > cap_marks) {
            set_error("internal: frontier overflow (%llu ranges, %llu marks)", h_counts[0], h_counts[2]);
            return COLBWT_ERR_NOMEM;
        }
        CB_CUDA(cudaMemset(d_counts + 1, 0, 8));
        k_frontier_step<<<(unsigned)((h_counts[0] + 255) / 256), 256>>>(t, d_cur, h_counts[0], d_mum_len, j, (uint32_t)split_rate, d_next, d_counts + 1,
                                                                       d_marks, d_counts + 2, !mode_all);
        CB_CUDA(cudaGetLastError());
        CB_CUDA(cudaMemcpy(d_counts + 0, d_counts + 1, 8, cudaMemcpyDeviceToDevice));
        std::swap(d_cur, d_next);
    }
    CB_CUDA(cudaMemcpy(h_counts, d_counts, 32, cudaMemcpyDeviceToHost));
    if (h_counts[2] > cap_marks) {
        set_error("internal: mark buffer overflow");
        return COLBWT_ERR_NOMEM;
    }
    std::vector<Mark> marks(h_counts[2]);
    CB_CUDA(cudaMemcpy(marks.data(), d_marks, marks.size() * sizeof(Mark), cudaMemcpyDeviceToHost));

    // ---- overlaps + run heads, then the two output files -------------------------------------------------------------
    std::vector<uint64_t> pos;
    std::vector<uint8_t> ids;
    if (!marks.empty()) resolve_marks(n, run_starts, marks, mode_all != 0, pos, ids);   // find_col_runs returns early without marks (col_split.hpp:259)
    std::vector<uint64_t> words((n + 63) / 64, 0);
    for (uint64_t q : pos) words[q >> 6] |= 1ull << (q & 63);
    FILE *f = fopen((p + ".col_runs").c_str(), "wb");
    if (!f) {
        set_error("cannot write %s.col_runs", prefix);
        return COLBWT_ERR_IO;
    }
    fwrite(&n, 8, 1, f);                        // sdsl bit_vector: length in bits, then the words (col_split.hpp:384-386)
    fwrite(words.data(), 8, words.size(), f);
    fclose(f);
    f = fopen((p + ".col_ids").c_str(), "wb");
    if (!f) {
        set_error("cannot write %s.col_ids", prefix);
        return COLBWT_ERR_IO;
    }
    fwrite(ids.data(), 1, ids.size(), f);       // one ID_BYTES = 1 byte per set bit (col_split.hpp:147-155)
    fclose(f);
    if (n_set_bits) *n_set_bits = pos.size();
    if (n_marked) {
        uint64_t m = 0;
        for (uint8_t v : ids) m += v != 0;
        *n_marked = m;
    }
    return COLBWT_OK;
}
