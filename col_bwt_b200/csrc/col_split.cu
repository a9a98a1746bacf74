// col_split.cu -- multi-MUM sub-run marking on the GPU.
//
// Replaces src/col_split.cpp + include/col_split.hpp + the FL_table it walks (include/ds/FL_table.hpp):
//   FL_table(heads, lengths) / compute_table      FL_table.hpp:86-131, 343-379   -> build_fl_table (stable radix sort by
//                                                                                    character + scans + binary search)
//   col_split::split, FL_loop, FL_range           col_split.hpp:54-136, 226-247  -> frontier kernels below
//   col_split::find_col_runs                      col_split.hpp:258-338          -> resolve_marks_device (sorts + prefix sums)
//   col_split::save / serialize_col_runs          col_split.hpp:138-157, 374-390 -> write_outputs
//
// The reference walks every multi-MUM forward through the text with FL steps and, every `split_rate` columns, marks
// the BWT range that the MUM's N suffixes occupy.  All MUMs are independent, so the walk is done level by level on
// the GPU: a *frontier* holds the current ranges of every MUM still alive; one level = one FL step of all of them (a
// range that straddles F-runs falls into pieces; in tunnel mode that ends the MUM, col_split.hpp:81,101), and all levels
// run inside ONE cooperative launch with a grid-wide barrier between them.  The marks are then reduced to one per start
// and resolved against each other and against the BWT run heads by sorting and prefix sums on the device -- the open-range
// counts of a sweep decide everything the reference's priority queue decides (see resolve_marks_device).
#include <cooperative_groups.h>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <chrono>
#include <cstring>
#include <string>

#include "internal.h"

namespace cg = cooperative_groups;

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

static_assert(sizeof(Range) == 24, "k_frontier_walk reads a Range as three 64-bit words");

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
} // namespace

// ---------------------------------------------------------------------------------------------------------------------
// The whole FL walk as ONE cooperative launch: every level of FL_loop (col_split.hpp:77-101) is a grid-stride pass over the
// live ranges followed by a grid-wide barrier, so the host neither launches per column nor reads a counter back (multi-MUMs
// are 1e5-1e6 columns long on a real pangenome).  counts[0..2] rotate as (current, next, to be zeroed); counts[3] = marks.
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_frontier_walk(FlTable t, Range *buf_a, Range *buf_b, const uint64_t *__restrict__ mum_len, uint32_t max_len,
                                                        uint32_t split_rate, Mark *marks, unsigned long long *counts, unsigned long long cap_ranges,
                                                        unsigned long long cap_marks, int tunnels, unsigned int *status)
{
    cg::grid_group grid = cg::this_grid();
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x, tid = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    Range *cur = buf_a, *next = buf_b;
    for (uint32_t j = 0; j < max_len; ++j) {
        // counters and ranges are written by other SMs one level earlier and were read here two or three levels ago: read them
        // past the L1 (ld.global.cg), which a grid-wide barrier does not invalidate
        const unsigned long long n_cur = __ldcg(counts + j % 3);
        if (n_cur == 0) break;                                   // same value for every thread: read after the barrier below
        if (n_cur > cap_ranges || __ldcg(counts + 3) > cap_marks) {       // cannot happen (the capacities are true upper bounds); never write past them
            if (tid == 0) *status = 1;
            break;
        }
        if (tid == 0) counts[(j + 2) % 3] = 0;                   // last read one level ago, written again one level from now
        unsigned long long *n_next = counts + (j + 1) % 3;
        for (unsigned long long i = tid; i < n_cur; i += stride) {
            const unsigned long long *rp = reinterpret_cast<const unsigned long long *>(cur + i);
            const unsigned long long w0 = __ldcg(rp), w1 = __ldcg(rp + 1), w2 = __ldcg(rp + 2);
            Range r;
            r.mum = (uint32_t)w0;
            r.interval = (uint32_t)(w0 >> 32);
            r.offset = w1;
            r.height = (uint32_t)w2;
            r.pad = (uint32_t)(w2 >> 32);
            if ((uint64_t)j >= mum_len[r.mum]) continue;         // this MUM has been walked to its end
            if (j % split_rate == 0) {
                const unsigned long long k = atomicAdd(counts + 3, 1ull);
                if (k < cap_marks) marks[k] = Mark{t.idx[r.interval] + r.offset, r.mum, j, r.height, 0};
            }
            if ((uint64_t)j + 1 < mum_len[r.mum]) expand(t, r, next, n_next, tunnels != 0);
        }
        grid.sync();
        Range *tmp = cur;
        cur = next;
        next = tmp;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Overlap resolution (col_split.hpp:105-132 second pass + find_col_runs :258-338) as sorts and scans.
//
// What the reference's priority-queue sweep emits depends only on how many marked ranges are open around each event of a
// sweep over the range starts and ends (ends before starts at equal positions, ends ordered by (end, start)):
//   * a range that starts while none is open begins a sub-run with its id;
//   * a range that ends leaving exactly ONE open (which ends later) hands the position to that one: sub-run with ITS id;
//   * a range that ends leaving none open begins an unmarked sub-run (id 0), unless another range starts right there or
//     the position is n;
//   * BWT run heads that are none of these positions become sub-runs carrying the id in force (that of the last such
//     boundary before them, 0 if none).
// The number of open ranges is a prefix sum of +1/-1 over the sorted events; WHICH range is the one left open is the
// prefix sum of +/-(index+1), read where the count is 1.  No queue, no sequential state.
// ---------------------------------------------------------------------------------------------------------------------
struct UniqueMark { uint64_t start, end; uint32_t id, pad; };

__device__ __forceinline__ uint32_t bin_id_dev(uint64_t id) { return id >= 256 ? (uint32_t)(id % 255) + 1 : (uint32_t)id; }   // col_split.hpp:222-224

__global__ void k_mark_keys(const Mark *__restrict__ marks, uint64_t n_marks, uint64_t *visit_key, uint32_t *iota)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_marks) return;
    visit_key[i] = ((uint64_t)marks[i].mum << 32) | marks[i].col;   // the reference's visiting order: MUM, then column
    iota[i] = (uint32_t)i;
}

__global__ void k_start_keys(const Mark *__restrict__ marks, const uint32_t *__restrict__ by_visit, uint64_t n_marks, uint64_t *start_key)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_marks) start_key[i] = marks[by_visit[i]].start;
}

__global__ void k_segment_heads(const uint64_t *__restrict__ start_sorted, uint64_t n_marks, uint32_t *is_head)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_marks) is_head[i] = (i == 0 || start_sorted[i] != start_sorted[i - 1]) ? 1u : 0u;
}

// One thread per distinct start: marks with that start are consecutive and in visiting order (collect_ids, col_split.hpp:114-127).
__global__ void k_unique_marks(const Mark *__restrict__ marks, const uint32_t *__restrict__ order, const uint64_t *__restrict__ start_sorted,
                               const uint32_t *__restrict__ is_head, const uint32_t *__restrict__ head_rank, uint64_t n_marks, int mode_all, UniqueMark *out)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_marks || !is_head[i]) return;
    uint32_t id = 0, h = 0;
    for (uint64_t e = i; e < n_marks && start_sorted[e] == start_sorted[i]; ++e) {
        const Mark m = marks[order[e]];
        const uint32_t mid = bin_id_dev((uint64_t)m.mum + 1);
        if (mode_all) {                     // the taller range wins; on ties the one visited first
            if (!(h >= m.height)) id = mid;
            h = max(h, m.height);
        } else {                            // tunnel mode: the last visit wins
            id = mid;
            h = m.height;
        }
    }
    out[head_rank[i]] = UniqueMark{start_sorted[i], start_sorted[i] + h, id, 0};
}

__global__ void k_event_keys(const UniqueMark *__restrict__ u, uint64_t m, uint64_t *key, uint32_t *payload)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    key[i] = u[i].end << 1;                 // ends first at equal positions; ends keep their start order (stable sort)
    payload[i] = (uint32_t)i;
    key[m + i] = (u[i].start << 1) | 1;
    payload[m + i] = (uint32_t)i;
}

__global__ void k_event_deltas(const uint64_t *__restrict__ key, const uint32_t *__restrict__ payload, uint64_t n_events, long long *d_count, long long *d_sum)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_events) return;
    const long long sign = (key[i] & 1) ? 1 : -1;
    d_count[i] = sign;
    d_sum[i] = sign * ((long long)payload[i] + 1);
}

__global__ void k_event_flags(const uint64_t *__restrict__ key, const uint32_t *__restrict__ payload, const long long *__restrict__ count,
                              const long long *__restrict__ sum, const UniqueMark *__restrict__ u, uint64_t n_events, uint64_t n, uint32_t *flag, uint32_t *id_out)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_events) return;
    const uint64_t pos = key[i] >> 1;
    uint32_t f = 0, id = 0;
    if (key[i] & 1) {                                           // a range starts
        if (count[i] == 1) { f = 1; id = u[payload[i]].id; }    // nothing else open
    } else if (count[i] == 1) {                                 // a range ends and exactly one stays open
        const UniqueMark &left = u[sum[i] - 1];
        if (left.end > pos) { f = 1; id = left.id; }
    } else if (count[i] == 0) {                                 // a range ends and nothing stays open
        const bool start_here = i + 1 < n_events && key[i + 1] == ((pos << 1) | 1);
        if (!start_here && pos < n) f = 1;
    }
    flag[i] = f;
    id_out[i] = id;
}

__global__ void k_compact_boundaries(const uint64_t *__restrict__ key, const uint32_t *__restrict__ flag, const uint32_t *__restrict__ rank,
                                     const uint32_t *__restrict__ id_in, uint64_t n_events, uint64_t *b_pos, uint8_t *b_id, unsigned long long *bits)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_events || !flag[i]) return;
    const uint64_t pos = key[i] >> 1;
    b_pos[rank[i]] = pos;
    b_id[rank[i]] = (uint8_t)id_in[i];
    atomicOr(bits + (pos >> 6), 1ull << (pos & 63));
}

__global__ void k_set_head_bits(const uint64_t *__restrict__ run_start, uint64_t runs, unsigned long long *bits)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < runs) atomicOr(bits + (run_start[i] >> 6), 1ull << (run_start[i] & 63));
}

__global__ void k_popc_words(const unsigned long long *__restrict__ bits, uint64_t nw, uint64_t *cnt)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nw) cnt[i] = (uint64_t)__popcll(bits[i]);
}

__device__ __forceinline__ uint64_t bit_rank(const unsigned long long *bits, const uint64_t *prefix, uint64_t pos)
{
    return prefix[pos >> 6] + (uint64_t)__popcll(bits[pos >> 6] & ((1ull << (pos & 63)) - 1ull));
}

// Run heads: id in force = id of the last boundary before the head; a head that is itself a boundary is written by k_boundary_ids.
__global__ void k_head_ids(const uint64_t *__restrict__ run_start, uint64_t runs, const uint64_t *__restrict__ b_pos, const uint8_t *__restrict__ b_id,
                           uint64_t n_b, const unsigned long long *__restrict__ bits, const uint64_t *__restrict__ prefix, uint8_t *ids)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= runs) return;
    const uint64_t p = run_start[i];
    uint64_t lo = 0, hi = n_b;                  // first boundary at or after p
    while (lo < hi) {
        const uint64_t mid = (lo + hi) >> 1;
        if (b_pos[mid] < p) lo = mid + 1; else hi = mid;
    }
    if (lo < n_b && b_pos[lo] == p) return;
    ids[bit_rank(bits, prefix, p)] = lo ? b_id[lo - 1] : 0;
}

__global__ void k_boundary_ids(const uint64_t *__restrict__ b_pos, const uint8_t *__restrict__ b_id, uint64_t n_b, const unsigned long long *__restrict__ bits,
                               const uint64_t *__restrict__ prefix, uint8_t *ids)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_b) ids[bit_rank(bits, prefix, b_pos[i])] = b_id[i];
}

// marks (device, any order) -> words of the col_runs bit vector + one id per set bit, both on the host.
static int resolve_marks_device(DevFree &mem, const Mark *d_marks, uint64_t n_marks, const uint64_t *d_run_start, uint64_t runs, uint64_t n, bool mode_all,
                                std::vector<uint64_t> &words, std::vector<uint8_t> &ids)
{
    auto grid = [](uint64_t k) { return (unsigned)((k + 255) / 256); };
    // ---- one (id, height) per distinct start: sort by visit, then stably by start -----------------------------------------
    uint64_t *d_k1, *d_k2;
    uint32_t *d_v1, *d_v2, *d_head, *d_head_rank;
    CB_CUDA(mem.alloc(&d_k1, 2 * n_marks));
    CB_CUDA(mem.alloc(&d_k2, 2 * n_marks));
    CB_CUDA(mem.alloc(&d_v1, 2 * n_marks));
    CB_CUDA(mem.alloc(&d_v2, 2 * n_marks));
    CB_CUDA(mem.alloc(&d_head, 2 * n_marks));
    CB_CUDA(mem.alloc(&d_head_rank, 2 * n_marks + 1));
    size_t tb = 0, tb2 = 0, tb3 = 0;
    CB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, (const uint64_t *)d_k1, d_k2, (const uint32_t *)d_v1, d_v2, (int64_t)(2 * n_marks), 0, 64));
    CB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb2, (const uint32_t *)d_head, d_head_rank, (int64_t)(2 * n_marks + 1)));
    CB_CUDA(cub::DeviceScan::InclusiveSum(nullptr, tb3, (const long long *)nullptr, (long long *)nullptr, (int64_t)(2 * n_marks)));
    const uint64_t nw = (n + 63) / 64;
    size_t tb4 = 0;
    CB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb4, (const uint64_t *)nullptr, (uint64_t *)nullptr, (int64_t)nw));
    const size_t temp_bytes = std::max(std::max(tb, tb2), std::max(tb3, tb4)) + 256;
    uint8_t *d_temp;
    CB_CUDA(mem.alloc(&d_temp, temp_bytes));
    k_mark_keys<<<grid(n_marks), 256>>>(d_marks, n_marks, d_k1, d_v1);
    tb = temp_bytes;
    CB_CUDA(cub::DeviceRadixSort::SortPairs(d_temp, tb, (const uint64_t *)d_k1, d_k2, (const uint32_t *)d_v1, d_v2, (int64_t)n_marks, 0, 64));
    k_start_keys<<<grid(n_marks), 256>>>(d_marks, d_v2, n_marks, d_k1);                 // d_v2 = marks in visiting order
    tb = temp_bytes;
    CB_CUDA(cub::DeviceRadixSort::SortPairs(d_temp, tb, (const uint64_t *)d_k1, d_k2, (const uint32_t *)d_v2, d_v1, (int64_t)n_marks, 0, 40));
    // now d_k2 = starts ascending, d_v1 = mark index; equal starts in visiting order (the sort is stable)
    k_segment_heads<<<grid(n_marks), 256>>>(d_k2, n_marks, d_head);
    tb = temp_bytes;
    CB_CUDA(cub::DeviceScan::ExclusiveSum(d_temp, tb, (const uint32_t *)d_head, d_head_rank, (int64_t)(n_marks + 1)));
    uint32_t m32 = 0;
    CB_CUDA(cudaMemcpy(&m32, d_head_rank + n_marks, 4, cudaMemcpyDeviceToHost));
    const uint64_t m = m32;
    UniqueMark *d_u;
    CB_CUDA(mem.alloc(&d_u, m));
    k_unique_marks<<<grid(n_marks), 256>>>(d_marks, d_v1, d_k2, d_head, d_head_rank, n_marks, mode_all ? 1 : 0, d_u);
    // ---- the sweep as a sort of 2m events and two prefix sums -----------------------------------------------------------
    const uint64_t n_events = 2 * m;
    long long *d_dc, *d_ds;
    uint32_t *d_flag = d_head, *d_rank = d_head_rank, *d_eid;
    CB_CUDA(mem.alloc(&d_dc, n_events));
    CB_CUDA(mem.alloc(&d_ds, n_events));
    CB_CUDA(mem.alloc(&d_eid, n_events));
    k_event_keys<<<grid(m), 256>>>(d_u, m, d_k1, d_v1);
    tb = temp_bytes;
    CB_CUDA(cub::DeviceRadixSort::SortPairs(d_temp, tb, (const uint64_t *)d_k1, d_k2, (const uint32_t *)d_v1, d_v2, (int64_t)n_events, 0, 42));
    k_event_deltas<<<grid(n_events), 256>>>(d_k2, d_v2, n_events, d_dc, d_ds);
    tb = temp_bytes;
    CB_CUDA(cub::DeviceScan::InclusiveSum(d_temp, tb, (const long long *)d_dc, d_dc, (int64_t)n_events));
    tb = temp_bytes;
    CB_CUDA(cub::DeviceScan::InclusiveSum(d_temp, tb, (const long long *)d_ds, d_ds, (int64_t)n_events));
    k_event_flags<<<grid(n_events), 256>>>(d_k2, d_v2, d_dc, d_ds, d_u, n_events, n, d_flag, d_eid);
    tb = temp_bytes;
    CB_CUDA(cub::DeviceScan::ExclusiveSum(d_temp, tb, (const uint32_t *)d_flag, d_rank, (int64_t)(n_events + 1)));
    uint32_t nb32 = 0;
    CB_CUDA(cudaMemcpy(&nb32, d_rank + n_events, 4, cudaMemcpyDeviceToHost));
    const uint64_t n_b = nb32;
    // ---- boundaries + run heads -> bit vector, ranks, ids ------------------------------------------------------------------
    uint64_t *d_bpos, *d_cnt, *d_prefix;
    uint8_t *d_bid, *d_ids;
    unsigned long long *d_bits;
    CB_CUDA(mem.alloc(&d_bpos, n_b));
    CB_CUDA(mem.alloc(&d_bid, n_b));
    CB_CUDA(mem.alloc(&d_bits, nw));
    CB_CUDA(mem.alloc(&d_cnt, nw));
    CB_CUDA(mem.alloc(&d_prefix, nw));
    CB_CUDA(cudaMemset(d_bits, 0, nw * 8));
    k_compact_boundaries<<<grid(n_events), 256>>>(d_k2, d_flag, d_rank, d_eid, n_events, d_bpos, d_bid, d_bits);
    k_set_head_bits<<<grid(runs), 256>>>(d_run_start, runs, d_bits);
    k_popc_words<<<grid(nw), 256>>>(d_bits, nw, d_cnt);
    tb = temp_bytes;
    CB_CUDA(cub::DeviceScan::ExclusiveSum(d_temp, tb, (const uint64_t *)d_cnt, d_prefix, (int64_t)nw));
    uint64_t last_pre = 0, last_cnt = 0;
    CB_CUDA(cudaMemcpy(&last_pre, d_prefix + (nw - 1), 8, cudaMemcpyDeviceToHost));
    CB_CUDA(cudaMemcpy(&last_cnt, d_cnt + (nw - 1), 8, cudaMemcpyDeviceToHost));
    const uint64_t set_bits = last_pre + last_cnt;
    CB_CUDA(mem.alloc(&d_ids, set_bits));
    k_head_ids<<<grid(runs), 256>>>(d_run_start, runs, d_bpos, d_bid, n_b, d_bits, d_prefix, d_ids);
    k_boundary_ids<<<grid(n_b), 256>>>(d_bpos, d_bid, n_b, d_bits, d_prefix, d_ids);
    CB_CUDA(cudaGetLastError());
    words.resize(nw);
    ids.resize(set_bits);
    CB_CUDA(cudaMemcpy(words.data(), d_bits, nw * 8, cudaMemcpyDeviceToHost));
    CB_CUDA(cudaMemcpy(ids.data(), d_ids, set_bits, cudaMemcpyDeviceToHost));
    return COLBWT_OK;
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
    const bool trace = getenv("COLBWT_TRACE") && atoi(getenv("COLBWT_TRACE")) > 0;
    auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t_begin = now();
    double t_read = 0, t_table = 0, t_walk = 0, t_resolve = 0;
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
    // The reference consumes the MUM positions with ONE forward cursor over the BWT runs (FL_loop, col_split.hpp:70-100): a
    // position smaller than its predecessor stalls that cursor and every later MUM is silently dropped.  mumemto writes them
    // ascending; anything else is refused here rather than reproduced.
    for (uint32_t i = 1; i < n_mums; ++i)
        if (mum_pos[i] < mum_pos[i - 1]) {
            set_error("%s.col_mums: positions must be ascending (entry %u: %llu after %llu)", prefix, i, (unsigned long long)mum_pos[i],
                      (unsigned long long)mum_pos[i - 1]);
            return COLBWT_ERR_FORMAT;
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
    t_read = now() - t_begin;

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

    if (trace) CB_CUDA(cudaDeviceSynchronize());
    t_table = now() - t_begin - t_read;

    // ---- frontier walk ---------------------------------------------------------------------------------------------
    const uint64_t cap_ranges = mode_all ? (uint64_t)n_mums * num_docs + 1024 : (uint64_t)n_mums + 1024;
    uint64_t total_cols = 0;
    for (uint32_t i = 0; i < n_mums; ++i) total_cols += (mum_len[i] + split_rate - 1) / split_rate;
    const uint64_t cap_marks = total_cols * (mode_all ? num_docs : 1) + 1024;
    Range *d_cur, *d_next;
    Mark *d_marks;
    unsigned long long *d_counts;   // [0..2] live ranges of three consecutive levels (rotating), [3] marks
    unsigned int *d_status;
    CB_CUDA(mem.alloc(&d_cur, cap_ranges));
    CB_CUDA(mem.alloc(&d_next, cap_ranges));
    CB_CUDA(mem.alloc(&d_marks, cap_marks));
    CB_CUDA(mem.alloc(&d_counts, 4));
    CB_CUDA(mem.alloc(&d_status, 1));
    CB_CUDA(cudaMemset(d_counts, 0, 32));
    CB_CUDA(cudaMemset(d_status, 0, 4));
    unsigned long long h_counts[4] = {0, 0, 0, 0};
    if (n_mums) {
        k_init_frontier<<<(n_mums + 255) / 256, 256>>>(t, d_mum_pos, n_mums, num_docs, d_cur, d_counts + 0, !mode_all);
        CB_CUDA(cudaGetLastError());
        // every level of the walk inside one cooperative launch (grid-wide barrier between levels): no per-column launch,
        // no counter read back by the host
        int per_sm = 0, sms = 0, coop = 0;
        CB_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device));
        CB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
        CB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_frontier_walk, 256, 0));
        if (!coop || per_sm < 1) {
            set_error("colbwt_col_split: the device cannot run a cooperative launch");
            return COLBWT_ERR_CUDA;
        }
        uint32_t max_len32 = (uint32_t)std::min<uint64_t>(max_len, 0xFFFFFFFFull), rate32 = (uint32_t)split_rate;
        unsigned long long cap_r = cap_ranges, cap_m = cap_marks;
        int tunnels = !mode_all;
        void *args[] = {&t, &d_cur, &d_next, &d_mum_len, &max_len32, &rate32, &d_marks, &d_counts, &cap_r, &cap_m, &tunnels, &d_status};
        CB_CUDA(cudaLaunchCooperativeKernel((const void *)k_frontier_walk, dim3((unsigned)(sms * std::min(per_sm, 4))), dim3(256), args, 0, nullptr));
        CB_CUDA(cudaDeviceSynchronize());
    }
    unsigned int h_status = 0;
    CB_CUDA(cudaMemcpy(&h_status, d_status, 4, cudaMemcpyDeviceToHost));
    CB_CUDA(cudaMemcpy(h_counts, d_counts, 32, cudaMemcpyDeviceToHost));
    if (h_status || h_counts[3] > cap_marks) {
        set_error("internal: frontier overflow (%llu marks of %llu)", h_counts[3], (unsigned long long)cap_marks);
        return COLBWT_ERR_NOMEM;
    }
    const uint64_t n_marks = h_counts[3];
    t_walk = now() - t_begin - t_read - t_table;

    // ---- overlaps + run heads (on the device), then the two output files -----------------------------------------------------
    std::vector<uint64_t> words;
    std::vector<uint8_t> ids;
    uint64_t n_bits = 0;                         // find_col_runs returns early without marks (col_split.hpp:259): col_runs stays empty
    if (n_marks) {
        if (int rc = resolve_marks_device(mem, d_marks, n_marks, d_lstart, runs, n, mode_all != 0, words, ids)) return rc;
        n_bits = n;
    }
    t_resolve = now() - t_begin - t_read - t_table - t_walk;
    FILE *f = fopen((p + ".col_runs").c_str(), "wb");
    if (!f) {
        set_error("cannot write %s.col_runs", prefix);
        return COLBWT_ERR_IO;
    }
    fwrite(&n_bits, 8, 1, f);                   // sdsl bit_vector: length in bits, then the words (col_split.hpp:384-386)
    fwrite(words.data(), 8, words.size(), f);
    fclose(f);
    f = fopen((p + ".col_ids").c_str(), "wb");
    if (!f) {
        set_error("cannot write %s.col_ids", prefix);
        return COLBWT_ERR_IO;
    }
    fwrite(ids.data(), 1, ids.size(), f);       // one ID_BYTES = 1 byte per set bit (col_split.hpp:147-155)
    fclose(f);
    if (trace)
        fprintf(stderr, "[colbwt_col_split] %u multi-MUMs (longest %llu), %llu marks: read inputs %.1f ms, FL table %.1f ms, walk %.1f ms, resolve %.1f ms, write %.1f ms\n",
                n_mums, (unsigned long long)max_len, (unsigned long long)n_marks, t_read * 1e3, t_table * 1e3, t_walk * 1e3, t_resolve * 1e3,
                (now() - t_begin - t_read - t_table - t_walk - t_resolve) * 1e3);
    if (n_set_bits) *n_set_bits = ids.size();
    if (n_marked) {
        uint64_t m = 0;
        for (uint8_t v : ids) m += v != 0;
        *n_marked = m;
    }
    return COLBWT_OK;
}
