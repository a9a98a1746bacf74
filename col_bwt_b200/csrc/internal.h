// internal.h -- host-side structures shared by index.cu / query.cu / capi.cpp (not part of the C-ABI).
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/colbwt_b200.h"
#include "colbwt_core.cuh"

namespace colbwt {

void set_error(const char *fmt, ...);

#define CB_CUDA(call)                                                                                   \
    do {                                                                                                \
        cudaError_t e__ = (call);                                                                       \
        if (e__ != cudaSuccess) {                                                                       \
            colbwt::set_error("%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__));      \
            return COLBWT_ERR_CUDA;                                                                     \
        }                                                                                               \
    } while (0)

// The table in the HBM of one GPU.
struct DeviceTable {
    int device = -1;
    TableView view{};
    void *d_rows = nullptr, *d_hot = nullptr, *d_cold = nullptr, *d_ch8 = nullptr, *d_idx = nullptr, *d_thr = nullptr, *d_char_rows = nullptr,
         *d_char_start = nullptr, *d_code_lut = nullptr;
    uint64_t bytes = 0;
    int sm_count = 0;
};

} // namespace colbwt

namespace colbwt { struct Pipeline; void destroy_pipeline(Pipeline *); }

struct colbwt_index {
    colbwt_stats stats{};
    std::vector<colbwt::DeviceTable> dev;
    uint8_t code_lut[256];
    std::mutex query_mutex;                 // one colbwt_query at a time per index
    colbwt::Pipeline *pipeline = nullptr;   // staging buffers + streams, kept between colbwt_query calls
    // Where colbwt_query packs the reads (query.cu): bases/s of the last large call packed on the host [0] / on the device [1],
    // the number of large calls so far, and what the last call did.
    double pack_rate[2] = {0.0, 0.0};
    uint32_t large_calls = 0;
    int last_packing = 0;
};

namespace colbwt {

// index.cu: upload + device-side build of one replica from the raw 18-byte rows (host memory, or streamed
// from `fp` when rows == nullptr).
int build_device_table(DeviceTable &dt, int device, const void *rows, FILE *fp, uint64_t n, uint64_t r,
                       colbwt_stats *stats, uint8_t *code_lut_out);
void free_device_table(DeviceTable &dt);

// The files `col-bwt build --keep` leaves on disk, read into host memory (capi.cpp: load_primaries).
struct Primaries {
    std::vector<uint8_t> heads;     // .bwt.heads   1 byte per BWT run
    std::vector<uint64_t> lens;     // .bwt.len     5-byte LE per run
    std::vector<uint64_t> thr;      // .thr_pos     5-byte LE per run
    std::vector<uint64_t> bits;     // .col_runs    sdsl bit_vector words (LSB first), n_bits of them used
    std::vector<uint8_t> ids;       // .col_ids     1 byte per set bit
    uint64_t n_bits = 0;
};
int build_device_table_from_primaries(DeviceTable &dt, int device, const Primaries &pr, colbwt_stats *stats, uint8_t *code_lut_out,
                                      uint64_t *n_out, uint64_t *r_out);
// 18-byte reference rows [first, first+count) reconstructed from the device columns.
int export_rows(const DeviceTable &dt, uint64_t first, uint64_t count, void *host_out);

// pack.cpp: host-side 2-bit packer.
// One read: returns false if it contains a byte outside ACGT (its words are then garbage, the caller ships bytes).
bool pack_read_2bit(const uint8_t *seq, uint64_t len, uint32_t *words);
// Packs reads [r0, r1) of a batch (one thread's slice): metas at meta[i - r_base], words from word index w on;
// seq_end = number of readable bytes in seqs (bounds the 32-byte over-reads); irregular read numbers go to irr.
void pack_slice(const uint8_t *seqs, const uint64_t *off, uint64_t r0, uint64_t r1, uint64_t r_base, uint64_t base0,
                uint64_t seq_end, uint32_t *words, uint64_t w, ReadMeta *meta, std::vector<uint64_t> &irr);

// Kernels launched through host wrappers (traverse.cu).
// Device-side packing of a chunk's raw bytes (traverse.cu: k_pack_reads).
int launch_pack(const DeviceTable &dt, const uint8_t *d_bytes, ReadMeta *d_meta, uint32_t n_reads, uint32_t *d_words, ReadMeta *d_meta_b,
                uint32_t *d_n_irregular, cudaStream_t stream);
// One traversal of a batch: packed pass, byte pass, chain fix-up (those that have work).  d_counters: 4 x u64 scratch
// ([0],[1] work cursors, [2] chunks re-traversed by k_fixup, [3] irregular-read count of the device packer).
int launch_traverse(const DeviceTable &dt, const BatchView &bv, int pml_width, unsigned long long *d_counters,
                    cudaStream_t stream);

} // namespace colbwt
