// index.cu -- upload of the reference's `.col_pml` rows and device-side construction of the packed move table.
//
// Input layout (include/ds/LF_table.hpp:33-84 + include/col_bwt.hpp:40-115, GCC packed, 18 bytes per row):
//   byte 0 character | 1-5 idx u40 | 6-9 interval u32 | 10-11 offset u16 | 12 col_id | 13-17 threshold u40
// The rows are streamed to the GPU as they are (pinned double buffer), split into columns by k_unpack_rows and
// turned into 16-byte packed rows by k_build_rows (colbwt_core.cuh: build_row).  The per-character row lists
// used by the exact reposition search come from one stable radix sort of the row numbers by row byte.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <cstring>

#include "internal.h"

namespace colbwt {

constexpr int REF_ROW_BYTES = 18;
constexpr uint64_t CHUNK_ROWS = 4u << 20;   // 72 MiB of raw rows per staging buffer

__global__ void k_unpack_rows(const uint8_t *__restrict__ raw, uint64_t first, uint32_t count, uint8_t *ch8, uint64_t *idx,
                              uint64_t *thr, uint32_t *dest, uint16_t *doff, uint8_t *cid)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const uint8_t *p = raw + (uint64_t)i * REF_ROW_BYTES;
    uint8_t b[REF_ROW_BYTES];
#pragma unroll
    for (int k = 0; k < REF_ROW_BYTES; ++k) b[k] = p[k];
    const uint64_t row = first + i;
    ch8[row] = b[0];
    idx[row] = (uint64_t)b[1] | ((uint64_t)b[2] << 8) | ((uint64_t)b[3] << 16) | ((uint64_t)b[4] << 24) | ((uint64_t)b[5] << 32);
    dest[row] = (uint32_t)b[6] | ((uint32_t)b[7] << 8) | ((uint32_t)b[8] << 16) | ((uint32_t)b[9] << 24);
    doff[row] = (uint16_t)(b[10] | (b[11] << 8));
    cid[row] = b[12];
    thr[row] = (uint64_t)b[13] | ((uint64_t)b[14] << 8) | ((uint64_t)b[15] << 16) | ((uint64_t)b[16] << 24) | ((uint64_t)b[17] << 32);
}

// counters: [0] OR of build flags, [1] marked rows, [2] slow rows, [3] max row length
__global__ void k_build_rows(BuildView b, Row *rows, unsigned long long *counters)
{
    const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t flags = 0, len = 0, marked = 0;
    if (k < b.r) {
        Row r = build_row(b, (uint32_t)k, &flags);
        *reinterpret_cast<uint4 *>(rows + k) = make_uint4(r.dest, r.offlen, r.m0, r.m1);
        marked = row_cid(r) != 0;
        len = row_len(r);
    }
    // one atomic per warp and counter
    const uint32_t all_flags = __reduce_or_sync(0xffffffffu, flags);
    const uint32_t n_marked = __popc(__ballot_sync(0xffffffffu, marked));
    const uint32_t n_slow = __popc(__ballot_sync(0xffffffffu, (flags & BUILD_FLAG_SLOW) != 0));
    const uint32_t max_len = __reduce_max_sync(0xffffffffu, len);
    if ((threadIdx.x & 31) == 0) {
        if (all_flags & (BUILD_FLAG_BAD_LEN | BUILD_FLAG_BAD_LF)) atomicOr(counters + 0, (unsigned long long)all_flags);
        if (n_marked) atomicAdd(counters + 1, (unsigned long long)n_marked);
        if (n_slow) atomicAdd(counters + 2, (unsigned long long)n_slow);
        atomicMax(counters + 3, (unsigned long long)max_len);
    }
}

__global__ void k_narrow_rows(const Row *__restrict__ rows, uint64_t r, uint64_t *hot, uint64_t *cold)
{
    const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= r) return;
    const Row row = rows[k];
    hot[k] = hot_from_row(row);
    cold[k] = cold_from_row(row);
}

__global__ void k_iota(uint32_t *v, uint64_t n)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] = (uint32_t)i;
}

__global__ void k_byte_histogram(const uint8_t *__restrict__ ch8, uint64_t n, unsigned long long *hist)
{
    __shared__ unsigned int h[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) h[i] = 0;
    __syncthreads();
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        atomicAdd(&h[ch8[i]], 1u);
    __syncthreads();
    for (int i = threadIdx.x; i < 256; i += blockDim.x)
        if (h[i]) atomicAdd(hist + i, (unsigned long long)h[i]);
}

void free_device_table(DeviceTable &dt)
{
    if (dt.device < 0) return;
    cudaSetDevice(dt.device);
    cudaFree(dt.d_rows);
    cudaFree(dt.d_hot);
    cudaFree(dt.d_cold);
    cudaFree(dt.d_ch8);
    cudaFree(dt.d_idx);
    cudaFree(dt.d_thr);
    cudaFree(dt.d_char_rows);
    cudaFree(dt.d_char_start);
    cudaFree(dt.d_code_lut);
    dt = DeviceTable{};
}

namespace {
struct Scratch {   // freed on every exit path
    std::vector<void *> dev;
    std::vector<void *> pinned;
    cudaStream_t streams[2] = {nullptr, nullptr};
    ~Scratch()
    {
        for (void *p : dev) cudaFree(p);
        for (void *p : pinned) cudaFreeHost(p);
        for (auto s : streams)
            if (s) cudaStreamDestroy(s);
    }
};
} // namespace

// Back-end shared by both front-ends: columns (dt.d_ch8/d_idx/d_thr + dest/doff/cid) -> packed rows, per-byte row
// lists, code table, stats.
static int finish_table(DeviceTable &dt, const uint32_t *d_dest, const uint16_t *d_doff, const uint8_t *d_cid, uint64_t n, uint64_t r,
                        colbwt_stats *stats, uint8_t *code_lut_out)
{
    Scratch sc;
    unsigned long long *d_counters = nullptr;
    CB_CUDA(cudaMalloc(&d_counters, 8 * (256 + 8)));
    sc.dev.push_back(d_counters);
    CB_CUDA(cudaMemset(d_counters, 0, 8 * (256 + 8)));
    // ---- packed rows ---------------------------------------------------------------------------------------
    CB_CUDA(cudaMalloc(&dt.d_rows, r * sizeof(Row)));
    BuildView b{(const uint8_t *)dt.d_ch8, (const uint64_t *)dt.d_idx, (const uint64_t *)dt.d_thr, d_dest, d_doff, d_cid, n, (uint32_t)r};
    k_build_rows<<<(unsigned)((r + 255) / 256), 256>>>(b, (Row *)dt.d_rows, d_counters);
    CB_CUDA(cudaGetLastError());

    // ---- per-character row lists -----------------------------------------------------------------------------
    unsigned long long *d_hist = d_counters + 8;
    k_byte_histogram<<<std::max(1, dt.sm_count * 4), 256>>>((const uint8_t *)dt.d_ch8, r, d_hist);
    CB_CUDA(cudaGetLastError());
    uint32_t *d_iota = nullptr;
    uint8_t *d_keys_out = nullptr;
    CB_CUDA(cudaMalloc(&d_iota, r * 4));
    sc.dev.push_back(d_iota);
    CB_CUDA(cudaMalloc(&d_keys_out, r));
    sc.dev.push_back(d_keys_out);
    CB_CUDA(cudaMalloc(&dt.d_char_rows, r * 4));
    k_iota<<<(unsigned)((r + 255) / 256), 256>>>(d_iota, r);
    CB_CUDA(cudaGetLastError());
    size_t temp_bytes = 0;
    CB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes, (const uint8_t *)dt.d_ch8, d_keys_out, (const uint32_t *)d_iota,
                                            (uint32_t *)dt.d_char_rows, (int64_t)r, 0, 8));
    void *d_temp = nullptr;
    CB_CUDA(cudaMalloc(&d_temp, std::max<size_t>(temp_bytes, 16)));
    sc.dev.push_back(d_temp);
    CB_CUDA(cub::DeviceRadixSort::SortPairs(d_temp, temp_bytes, (const uint8_t *)dt.d_ch8, d_keys_out, (const uint32_t *)d_iota,
                                            (uint32_t *)dt.d_char_rows, (int64_t)r, 0, 8));

    unsigned long long h_counters[256 + 8];
    CB_CUDA(cudaMemcpy(h_counters, d_counters, sizeof(h_counters), cudaMemcpyDeviceToHost));
    if (h_counters[0] & BUILD_FLAG_BAD_LEN) {
        set_error("a row has length 0 or >= 65536: the reference's 16-bit offset field (LF_table.hpp:39) cannot address it");
        return COLBWT_ERR_ROW_TOO_LONG;
    }
    if (h_counters[0] & BUILD_FLAG_BAD_LF) {
        set_error("a row's LF destination is outside the table");
        return COLBWT_ERR_FORMAT;
    }
    uint32_t h_start[257];
    uint64_t acc = 0;
    for (int c = 0; c < 256; ++c) {
        h_start[c] = (uint32_t)acc;
        acc += h_counters[8 + c];
        // byte -> traversal code: ACGT keep their 2-bit code even if absent (the row builder then routes them
        // to the exact search, which finds nothing)
        const int pc = primary_code((uint8_t)c);
        code_lut_out[c] = pc >= 0 ? (uint8_t)pc : (h_counters[8 + c] ? CODE_OTHER : CODE_ABSENT);
    }
    h_start[256] = (uint32_t)acc;   // == r (r < 2^32)
    CB_CUDA(cudaMalloc(&dt.d_char_start, sizeof(h_start)));
    CB_CUDA(cudaMemcpy(dt.d_char_start, h_start, sizeof(h_start), cudaMemcpyHostToDevice));
    CB_CUDA(cudaMalloc(&dt.d_code_lut, 256));
    CB_CUDA(cudaMemcpy(dt.d_code_lut, code_lut_out, 256, cudaMemcpyHostToDevice));

    uint64_t last_idx = 0;
    CB_CUDA(cudaMemcpy(&last_idx, (const uint64_t *)dt.d_idx + (r - 1), 8, cudaMemcpyDeviceToHost));
    CB_CUDA(cudaDeviceSynchronize());

    // Narrow layout (half-size gather array), opt-in with COLBWT_NARROW=1 when every row is short enough.  Measured on
    // C2 it gains 4 % (DRAM lines per base 0.91 -> 0.87: the L2 keeps few random lines either way) and it costs one more
    // dependent gather per mismatch, which lengthens the serial chain of long noisy reads; see DESIGN.md section 4.
    if (h_counters[3] <= NARROW_MAX_LEN && getenv("COLBWT_NARROW") && atoi(getenv("COLBWT_NARROW")) == 1) {
        CB_CUDA(cudaMalloc(&dt.d_hot, r * 8));
        CB_CUDA(cudaMalloc(&dt.d_cold, r * 8));
        k_narrow_rows<<<(unsigned)((r + 255) / 256), 256>>>((const Row *)dt.d_rows, r, (uint64_t *)dt.d_hot, (uint64_t *)dt.d_cold);
        CB_CUDA(cudaGetLastError());
        CB_CUDA(cudaDeviceSynchronize());
    }
    dt.view.hot = (const uint64_t *)dt.d_hot;
    dt.view.cold = (const uint64_t *)dt.d_cold;
    dt.view.rows = (const Row *)dt.d_rows;
    dt.view.ch8 = (const uint8_t *)dt.d_ch8;
    dt.view.idx = (const uint64_t *)dt.d_idx;
    dt.view.thr = (const uint64_t *)dt.d_thr;
    dt.view.char_rows = (const uint32_t *)dt.d_char_rows;
    dt.view.char_start = (const uint32_t *)dt.d_char_start;
    dt.view.n = n;
    dt.view.r = (uint32_t)r;
    dt.view.last_len = (uint32_t)(n - last_idx);
    dt.bytes = r * (sizeof(Row) + (dt.d_hot ? 16 : 0) + 1 + 8 + 8 + 4) + sizeof(h_start) + 256;
    if (stats) {
        stats->marked_rows = h_counters[1];
        stats->slow_rows = h_counters[2];
        stats->max_row_len = (uint32_t)h_counters[3];
        stats->device_bytes = dt.bytes;
    }
    return COLBWT_OK;
}


int build_device_table(DeviceTable &dt, int device, const void *rows_host, FILE *fp, uint64_t n, uint64_t r,
                       colbwt_stats *stats, uint8_t *code_lut_out)
{
    CB_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    CB_CUDA(cudaGetDeviceProperties(&prop, device));
    dt.device = device;
    dt.sm_count = prop.multiProcessorCount;

    Scratch sc;
    uint32_t *d_dest = nullptr;
    uint16_t *d_doff = nullptr;
    uint8_t *d_cid = nullptr;
    CB_CUDA(cudaMalloc(&dt.d_ch8, r));
    CB_CUDA(cudaMalloc(&dt.d_idx, r * 8));
    CB_CUDA(cudaMalloc(&dt.d_thr, r * 8));
    CB_CUDA(cudaMalloc(&d_dest, r * 4));
    sc.dev.push_back(d_dest);
    CB_CUDA(cudaMalloc(&d_doff, r * 2));
    sc.dev.push_back(d_doff);
    CB_CUDA(cudaMalloc(&d_cid, r));
    sc.dev.push_back(d_cid);

    // ---- stream the raw rows through two pinned staging buffers --------------------------------------------
    const uint64_t chunk_rows = std::min<uint64_t>(CHUNK_ROWS, r);
    uint8_t *h_stage[2] = {nullptr, nullptr}, *d_stage[2] = {nullptr, nullptr};
    struct StageEvents {   // destroyed on every exit path (CB_CUDA returns early)
        cudaEvent_t e[2] = {nullptr, nullptr};
        ~StageEvents() { for (auto x : e) if (x) cudaEventDestroy(x); }
        cudaEvent_t &operator[](int i) { return e[i]; }
    } done;
    for (int s = 0; s < 2; ++s) {
        CB_CUDA(cudaMallocHost(&h_stage[s], chunk_rows * REF_ROW_BYTES));
        sc.pinned.push_back(h_stage[s]);
        CB_CUDA(cudaMalloc(&d_stage[s], chunk_rows * REF_ROW_BYTES));
        sc.dev.push_back(d_stage[s]);
        CB_CUDA(cudaStreamCreate(&sc.streams[s]));
        CB_CUDA(cudaEventCreateWithFlags(&done[s], cudaEventDisableTiming));
    }
    int rc = COLBWT_OK;
    // rows already in device memory (a table produced on the GPU, e.g. by a build tool): no host hop -- unpack in place
    // when they sit on this device, device-to-device copy into the staging buffer otherwise
    bool rows_on_device = false, rows_on_this_device = false;
    if (rows_host) {
        cudaPointerAttributes pa;
        if (cudaPointerGetAttributes(&pa, rows_host) == cudaSuccess && pa.type == cudaMemoryTypeDevice) {
            rows_on_device = true;
            rows_on_this_device = pa.device == device;
        }
        cudaGetLastError();
    }
    for (uint64_t first = 0, c = 0; first < r; first += chunk_rows, ++c) {
        const int s = (int)(c & 1);
        const uint64_t count = std::min<uint64_t>(chunk_rows, r - first);
        CB_CUDA(cudaEventSynchronize(done[s]));   // staging buffer s free again
        if (rows_on_device) {
            const uint8_t *src = (const uint8_t *)rows_host + first * REF_ROW_BYTES;
            if (!rows_on_this_device) {
                CB_CUDA(cudaMemcpyAsync(d_stage[s], src, count * REF_ROW_BYTES, cudaMemcpyDefault, sc.streams[s]));
                src = d_stage[s];
            }
            k_unpack_rows<<<(unsigned)((count + 255) / 256), 256, 0, sc.streams[s]>>>(
                src, first, (uint32_t)count, (uint8_t *)dt.d_ch8, (uint64_t *)dt.d_idx, (uint64_t *)dt.d_thr, d_dest, d_doff, d_cid);
            CB_CUDA(cudaGetLastError());
            CB_CUDA(cudaEventRecord(done[s], sc.streams[s]));
            continue;
        }
        if (rows_host) {
            memcpy(h_stage[s], (const uint8_t *)rows_host + first * REF_ROW_BYTES, count * REF_ROW_BYTES);
        } else if (fread(h_stage[s], REF_ROW_BYTES, count, fp) != count) {
            set_error("index file is shorter than its header says (row %llu of %llu)", (unsigned long long)first, (unsigned long long)r);
            rc = COLBWT_ERR_IO;
            break;
        }
        CB_CUDA(cudaMemcpyAsync(d_stage[s], h_stage[s], count * REF_ROW_BYTES, cudaMemcpyHostToDevice, sc.streams[s]));
        k_unpack_rows<<<(unsigned)((count + 255) / 256), 256, 0, sc.streams[s]>>>(
            d_stage[s], first, (uint32_t)count, (uint8_t *)dt.d_ch8, (uint64_t *)dt.d_idx, (uint64_t *)dt.d_thr, d_dest, d_doff, d_cid);
        CB_CUDA(cudaGetLastError());
        CB_CUDA(cudaEventRecord(done[s], sc.streams[s]));
    }
    CB_CUDA(cudaDeviceSynchronize());
    if (rc != COLBWT_OK) return rc;

    return finish_table(dt, d_dest, d_doff, d_cid, n, r, stats, code_lut_out);
}


// =========================================================================================================
// Front-end 2: build the table from the primaries `col-bwt build --keep` leaves on disk, entirely on the GPU.
// Replaces src/build_col_bwt.cpp and the constructors it calls:
//   col_bwt(heads, lengths, col_ids, splits)   include/col_bwt.hpp:124-230  rows = runs cut at every set bit of splits;
//                                                                           id of a row = id of the last set bit <= its start
//   LF_table::compute_table                    include/ds/LF_table.hpp:365-387  rows in stable character order tile F
//   col_pml::read_thresholds                   include/col_bwt.hpp:440-457  one threshold per BWT run, copied to its rows
// =========================================================================================================
__global__ void k_mark_heads(const uint64_t *__restrict__ run_start, uint64_t runs, unsigned long long *union_bits, unsigned long long *head_bits)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= runs) return;
    const uint64_t p = run_start[i];
    atomicOr(union_bits + (p >> 6), 1ull << (p & 63));
    atomicOr(head_bits + (p >> 6), 1ull << (p & 63));
}

__global__ void k_popc(const unsigned long long *__restrict__ bits, uint64_t nw, uint64_t *counts)
{
    const uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w < nw) counts[w] = (uint64_t)__popcll(bits[w]);
}

__global__ void k_emit_row_starts(const unsigned long long *__restrict__ bits, const uint64_t *__restrict__ prefix, uint64_t nw, uint64_t *idx)
{
    const uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= nw) return;
    unsigned long long x = bits[w];
    uint64_t k = prefix[w];
    while (x) {
        idx[k++] = w * 64 + (uint64_t)(__ffsll((long long)x) - 1);
        x &= x - 1;
    }
}

// number of set bits at positions <= p
__device__ __forceinline__ uint64_t rank_incl(const unsigned long long *bits, const uint64_t *prefix, uint64_t p)
{
    const uint64_t w = p >> 6;
    const unsigned b = (unsigned)(p & 63);
    const unsigned long long mask = (b == 63) ? ~0ull : ((1ull << (b + 1)) - 1);
    return prefix[w] + (uint64_t)__popcll(bits[w] & mask);
}

__global__ void k_fill_rows(const uint64_t *__restrict__ idx, uint64_t r, const unsigned long long *head_bits, const uint64_t *head_prefix,
                            const unsigned long long *orig_bits, const uint64_t *orig_prefix, const uint8_t *__restrict__ heads,
                            const uint64_t *__restrict__ thr_run, const uint8_t *__restrict__ ids, uint64_t n_ids, uint8_t *ch8,
                            uint64_t *thr, uint8_t *cid, unsigned long long *err)
{
    const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= r) return;
    const uint64_t p = idx[k];
    const uint64_t run = rank_incl(head_bits, head_prefix, p) - 1;     // position 0 is always a run head
    const uint8_t c = heads[run];
    ch8[k] = (c <= 1 || c >= 128) ? 1 : c;                              // `char c; if (c <= TERMINATOR) c = TERMINATOR` (col_bwt.hpp:165-171)
    thr[k] = thr_run[run];
    const uint64_t s = rank_incl(orig_bits, orig_prefix, p);            // set bits of .col_runs at positions <= p
    if (s > n_ids) atomicOr(err, 1ull);
    cid[k] = (s == 0 || s > n_ids) ? 0 : ids[s - 1];
}

__global__ void k_gather_len(const uint32_t *__restrict__ order, const uint64_t *__restrict__ idx, uint64_t n, uint64_t r, uint64_t *len_sorted)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= r) return;
    const uint64_t k = order[i];
    len_sorted[i] = ((k + 1 < r) ? idx[k + 1] : n) - idx[k];
}

__global__ void k_lf_columns(const uint32_t *__restrict__ order, const uint64_t *__restrict__ fpos_sorted, const uint64_t *__restrict__ idx,
                             uint64_t r, uint32_t *dest, uint16_t *doff, unsigned long long *err)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= r) return;
    const uint64_t f = fpos_sorted[i];
    uint64_t lo = 0, hi = r;                    // last row whose start is <= f
    while (hi - lo > 1) {
        const uint64_t mid = (lo + hi) >> 1;
        if (idx[mid] <= f) lo = mid; else hi = mid;
    }
    const uint64_t o = f - idx[lo];
    if (o > 0xFFFF) atomicOr(err, 2ull);        // LF_row::offset is 16 bit (LF_table.hpp:39)
    dest[order[i]] = (uint32_t)lo;
    doff[order[i]] = (uint16_t)o;
}

int build_device_table_from_primaries(DeviceTable &dt, int device, const Primaries &pr, colbwt_stats *stats, uint8_t *code_lut_out,
                                      uint64_t *n_out, uint64_t *r_out)
{
    CB_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    CB_CUDA(cudaGetDeviceProperties(&prop, device));
    dt.device = device;
    dt.sm_count = prop.multiProcessorCount;
    Scratch sc;
    const uint64_t runs = pr.heads.size(), nw = (pr.n_bits + 63) / 64;
    auto dev_alloc = [&](void **p, size_t bytes) -> cudaError_t {
        cudaError_t e = cudaMalloc(p, std::max<size_t>(bytes, 16));
        if (e == cudaSuccess) sc.dev.push_back(*p);
        return e;
    };
    uint8_t *d_heads, *d_ids;
    uint64_t *d_lens, *d_run_start, *d_thr_run, *d_cnt, *d_pre_union, *d_pre_head, *d_pre_orig;
    unsigned long long *d_union, *d_head, *d_orig, *d_err;
    CB_CUDA(dev_alloc((void **)&d_heads, runs));
    CB_CUDA(dev_alloc((void **)&d_ids, pr.ids.size()));
    CB_CUDA(dev_alloc((void **)&d_lens, runs * 8));
    CB_CUDA(dev_alloc((void **)&d_run_start, runs * 8));
    CB_CUDA(dev_alloc((void **)&d_thr_run, runs * 8));
    CB_CUDA(dev_alloc((void **)&d_union, nw * 8));
    CB_CUDA(dev_alloc((void **)&d_head, nw * 8));
    CB_CUDA(dev_alloc((void **)&d_orig, nw * 8));
    CB_CUDA(dev_alloc((void **)&d_cnt, nw * 8));
    CB_CUDA(dev_alloc((void **)&d_pre_union, nw * 8));
    CB_CUDA(dev_alloc((void **)&d_pre_head, nw * 8));
    CB_CUDA(dev_alloc((void **)&d_pre_orig, nw * 8));
    CB_CUDA(dev_alloc((void **)&d_err, 8));
    CB_CUDA(cudaMemset(d_err, 0, 8));
    CB_CUDA(cudaMemcpy(d_heads, pr.heads.data(), runs, cudaMemcpyHostToDevice));
    CB_CUDA(cudaMemcpy(d_ids, pr.ids.data(), pr.ids.size(), cudaMemcpyHostToDevice));
    CB_CUDA(cudaMemcpy(d_lens, pr.lens.data(), runs * 8, cudaMemcpyHostToDevice));
    CB_CUDA(cudaMemcpy(d_thr_run, pr.thr.data(), runs * 8, cudaMemcpyHostToDevice));
    CB_CUDA(cudaMemcpy(d_orig, pr.bits.data(), nw * 8, cudaMemcpyHostToDevice));
    CB_CUDA(cudaMemcpy(d_union, d_orig, nw * 8, cudaMemcpyDeviceToDevice));
    CB_CUDA(cudaMemset(d_head, 0, nw * 8));

    size_t temp_bytes = 0, tb = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, temp_bytes, d_lens, d_run_start, (int64_t)std::max(runs, nw));
    void *d_temp = nullptr;
    CB_CUDA(dev_alloc(&d_temp, temp_bytes + 256));
    tb = temp_bytes + 256;
    CB_CUDA(cub::DeviceScan::ExclusiveSum(d_temp, tb, d_lens, d_run_start, (int64_t)runs));
    const unsigned gw = (unsigned)((nw + 255) / 256), gr = (unsigned)((runs + 255) / 256);
    k_mark_heads<<<gr, 256>>>(d_run_start, runs, d_union, d_head);
    const unsigned long long *bitsets[3] = {d_union, d_head, d_orig};
    uint64_t *prefixes[3] = {d_pre_union, d_pre_head, d_pre_orig};
    uint64_t totals[3] = {0, 0, 0};
    for (int b = 0; b < 3; ++b) {
        k_popc<<<gw, 256>>>(bitsets[b], nw, d_cnt);
        tb = temp_bytes + 256;
        CB_CUDA(cub::DeviceScan::ExclusiveSum(d_temp, tb, d_cnt, prefixes[b], (int64_t)nw));
        uint64_t last_pre = 0, last_cnt = 0;
        CB_CUDA(cudaMemcpy(&last_pre, prefixes[b] + (nw - 1), 8, cudaMemcpyDeviceToHost));
        CB_CUDA(cudaMemcpy(&last_cnt, d_cnt + (nw - 1), 8, cudaMemcpyDeviceToHost));
        totals[b] = last_pre + last_cnt;
    }
    const uint64_t r = totals[0], n = pr.n_bits;
    if (totals[1] != runs || totals[2] != pr.ids.size()) {
        set_error("primaries disagree: %llu run heads marked for %llu runs, %llu set bits in .col_runs for %zu ids in .col_ids",
                  (unsigned long long)totals[1], (unsigned long long)runs, (unsigned long long)totals[2], pr.ids.size());
        return COLBWT_ERR_FORMAT;
    }
    if (r == 0 || r >= (1ull << 32)) {
        set_error("table would have %llu rows (need 0 < r < 2^32)", (unsigned long long)r);
        return COLBWT_ERR_FORMAT;
    }
    uint32_t *d_dest = nullptr, *d_order = nullptr, *d_iota = nullptr;
    uint16_t *d_doff = nullptr;
    uint8_t *d_cid = nullptr, *d_keys_out = nullptr;
    uint64_t *d_len_sorted = nullptr, *d_fpos = nullptr;
    CB_CUDA(cudaMalloc(&dt.d_ch8, r));
    CB_CUDA(cudaMalloc(&dt.d_idx, r * 8));
    CB_CUDA(cudaMalloc(&dt.d_thr, r * 8));
    CB_CUDA(dev_alloc((void **)&d_dest, r * 4));
    CB_CUDA(dev_alloc((void **)&d_doff, r * 2));
    CB_CUDA(dev_alloc((void **)&d_cid, r));
    CB_CUDA(dev_alloc((void **)&d_order, r * 4));
    CB_CUDA(dev_alloc((void **)&d_iota, r * 4));
    CB_CUDA(dev_alloc((void **)&d_keys_out, r));
    CB_CUDA(dev_alloc((void **)&d_len_sorted, r * 8));
    CB_CUDA(dev_alloc((void **)&d_fpos, r * 8));
    const unsigned grr = (unsigned)((r + 255) / 256);
    k_emit_row_starts<<<gw, 256>>>(d_union, d_pre_union, nw, (uint64_t *)dt.d_idx);
    k_fill_rows<<<grr, 256>>>((const uint64_t *)dt.d_idx, r, d_head, d_pre_head, d_orig, d_pre_orig, d_heads, d_thr_run, d_ids, pr.ids.size(),
                              (uint8_t *)dt.d_ch8, (uint64_t *)dt.d_thr, d_cid, d_err);
    // compute_table: stable sort of the rows by character, F positions = running sum of their lengths in that order
    k_iota<<<grr, 256>>>(d_iota, r);
    size_t sort_bytes = 0;
    CB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, (const uint8_t *)dt.d_ch8, d_keys_out, (const uint32_t *)d_iota, d_order, (int64_t)r, 0, 8));
    void *d_sort_temp = nullptr;
    CB_CUDA(dev_alloc(&d_sort_temp, sort_bytes));
    CB_CUDA(cub::DeviceRadixSort::SortPairs(d_sort_temp, sort_bytes, (const uint8_t *)dt.d_ch8, d_keys_out, (const uint32_t *)d_iota, d_order, (int64_t)r, 0, 8));
    k_gather_len<<<grr, 256>>>(d_order, (const uint64_t *)dt.d_idx, n, r, d_len_sorted);
    size_t scan_bytes = 0;
    CB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, d_len_sorted, d_fpos, (int64_t)r));
    void *d_scan_temp = nullptr;
    CB_CUDA(dev_alloc(&d_scan_temp, scan_bytes));
    CB_CUDA(cub::DeviceScan::ExclusiveSum(d_scan_temp, scan_bytes, d_len_sorted, d_fpos, (int64_t)r));
    k_lf_columns<<<grr, 256>>>(d_order, d_fpos, (const uint64_t *)dt.d_idx, r, d_dest, d_doff, d_err);
    CB_CUDA(cudaGetLastError());
    unsigned long long h_err = 0;
    CB_CUDA(cudaMemcpy(&h_err, d_err, 8, cudaMemcpyDeviceToHost));
    if (h_err & 1) {
        set_error(".col_ids holds fewer ids than .col_runs has set bits");
        return COLBWT_ERR_FORMAT;
    }
    if (h_err & 2) {
        set_error("an LF offset does not fit the reference's 16-bit field (a row of >= 65536 symbols)");
        return COLBWT_ERR_ROW_TOO_LONG;
    }
    *n_out = n;
    *r_out = r;
    return finish_table(dt, d_dest, d_doff, d_cid, n, r, stats, code_lut_out);
}

// Inverse of k_unpack_rows: the 18-byte rows of `.col_pml` from the device columns (col_bwt::serialize, col_bwt.hpp:360-370).
__global__ void k_export_rows(const Row *__restrict__ rows, const uint8_t *__restrict__ ch8, const uint64_t *__restrict__ idx,
                              const uint64_t *__restrict__ thr, uint64_t first, uint32_t count, uint8_t *out)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const uint64_t k = first + i;
    const Row r = rows[k];
    uint8_t *p = out + (uint64_t)i * REF_ROW_BYTES;
    const uint64_t ix = idx[k], t = thr[k];
    p[0] = ch8[k];
    for (int b = 0; b < 5; ++b) p[1 + b] = (uint8_t)(ix >> (8 * b));
    for (int b = 0; b < 4; ++b) p[6 + b] = (uint8_t)(r.dest >> (8 * b));
    p[10] = (uint8_t)row_doff(r);
    p[11] = (uint8_t)(row_doff(r) >> 8);
    p[12] = (uint8_t)row_cid(r);
    for (int b = 0; b < 5; ++b) p[13 + b] = (uint8_t)(t >> (8 * b));
}

int export_rows(const DeviceTable &dt, uint64_t first, uint64_t count, void *host_out)
{
    CB_CUDA(cudaSetDevice(dt.device));
    uint8_t *d_out = nullptr;
    CB_CUDA(cudaMalloc(&d_out, std::max<uint64_t>(16, count * REF_ROW_BYTES)));
    k_export_rows<<<(unsigned)((count + 255) / 256), 256>>>(dt.view.rows, dt.view.ch8, dt.view.idx, dt.view.thr, first, (uint32_t)count, d_out);
    cudaError_t e = cudaMemcpy(host_out, d_out, count * REF_ROW_BYTES, cudaMemcpyDeviceToHost);
    cudaFree(d_out);
    CB_CUDA(e);
    return COLBWT_OK;
}

} // namespace colbwt
