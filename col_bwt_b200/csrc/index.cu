// index.cu -- upload of the reference's `.col_pml` rows and device-side construction of the packed move table.
//
// Input layout (include/ds/LF_table.hpp:33-84 + include/col_bwt.hpp:40-115, GCC packed, 18 bytes per row):
//   byte 0 character | 1-5 idx u40 | 6-9 interval u32 | 10-11 offset u16 | 12 col_id | 13-17 threshold u40
// The rows are streamed to the GPU as they are (pinned double buffer), split into columns by k_unpack_rows and
// turned into 16-byte packed rows by k_build_rows (colbwt_core.cuh: build_row).  The per-character row lists
// used by the exact reposition search come from one stable radix sort of the row numbers by row byte.
#include <cub/device/device_radix_sort.cuh>

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
    unsigned long long *d_counters = nullptr;
    CB_CUDA(cudaMalloc(&dt.d_ch8, r));
    CB_CUDA(cudaMalloc(&dt.d_idx, r * 8));
    CB_CUDA(cudaMalloc(&dt.d_thr, r * 8));
    CB_CUDA(cudaMalloc(&d_dest, r * 4));
    sc.dev.push_back(d_dest);
    CB_CUDA(cudaMalloc(&d_doff, r * 2));
    sc.dev.push_back(d_doff);
    CB_CUDA(cudaMalloc(&d_cid, r));
    sc.dev.push_back(d_cid);
    CB_CUDA(cudaMalloc(&d_counters, 8 * (256 + 8)));
    sc.dev.push_back(d_counters);
    CB_CUDA(cudaMemset(d_counters, 0, 8 * (256 + 8)));

    // ---- stream the raw rows through two pinned staging buffers --------------------------------------------
    const uint64_t chunk_rows = std::min<uint64_t>(CHUNK_ROWS, r);
    uint8_t *h_stage[2] = {nullptr, nullptr}, *d_stage[2] = {nullptr, nullptr};
    cudaEvent_t done[2];
    for (int s = 0; s < 2; ++s) {
        CB_CUDA(cudaMallocHost(&h_stage[s], chunk_rows * REF_ROW_BYTES));
        sc.pinned.push_back(h_stage[s]);
        CB_CUDA(cudaMalloc(&d_stage[s], chunk_rows * REF_ROW_BYTES));
        sc.dev.push_back(d_stage[s]);
        CB_CUDA(cudaStreamCreate(&sc.streams[s]));
        CB_CUDA(cudaEventCreateWithFlags(&done[s], cudaEventDisableTiming));
    }
    int rc = COLBWT_OK;
    for (uint64_t first = 0, c = 0; first < r; first += chunk_rows, ++c) {
        const int s = (int)(c & 1);
        const uint64_t count = std::min<uint64_t>(chunk_rows, r - first);
        CB_CUDA(cudaEventSynchronize(done[s]));   // staging buffer s free again
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
    for (int s = 0; s < 2; ++s) cudaEventDestroy(done[s]);
    if (rc != COLBWT_OK) return rc;

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

} // namespace colbwt
