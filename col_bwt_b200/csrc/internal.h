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

// Compact result form of one chunk (compact.cu / expand.cpp): match words | chain-id words | prefix, then the values.
constexpr uint32_t COMPACT_GROUP_WORDS = 64;   // one prefix entry per 64 words = 2048 bases
struct CompactLayout {
    uint64_t n_words, n_groups, match_off, cid_off, prefix_off, fixed_bytes;
    explicit CompactLayout(uint64_t n_bases)
    {
        n_words = (n_bases + 31) / 32;
        n_groups = (n_words + COMPACT_GROUP_WORDS - 1) / COMPACT_GROUP_WORDS;
        const uint64_t wb = (n_words * 4 + 15) & ~15ull;
        match_off = 0;
        cid_off = wb;
        prefix_off = 2 * wb;
        fixed_bytes = prefix_off + (((n_groups + 1) * 4 + 15) & ~15ull);
    }
};

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
    // How colbwt_query [0] / colbwt_query_compact [1] ran (query.cu): bases/s of the last large call in each mode
    // (0 = not measured yet), and what the last call did.
    // mode = packing bit + 2 x transport (0 dense copies, 1 compact form expanded on the host, 2 PML dense + chain ids compact)
    double mode_rate[2][8] = {{0.0}, {0.0}};
    int last_packing = 0, last_transport = 0;
    uint64_t last_h2d_bytes = 0, last_d2h_bytes = 0;   // bytes the last call asked the copy engines to move
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

// compact.cu: dense PML/CID of one chunk (still in HBM) -> compact form.  d_out holds CompactLayout(n_bases).fixed_bytes,
// d_group_count n_groups + 1 words, d_values up to n_bases bytes.
size_t compact_scan_temp_bytes(uint64_t max_groups);
int launch_compact(const void *d_pml, int pml_width, const uint8_t *d_cid, uint64_t n_bases, uint8_t *d_out, uint32_t *d_group_count,
                   void *d_scan_temp, size_t scan_temp_bytes, uint8_t *d_values, cudaStream_t stream);
// expand.cpp: reads [ra, rb) of the segment starting at read r_first, back to dense arrays (pml / cid point at the segment's base 0).
void expand_reads(const uint32_t *match, const uint32_t *cid_words, const uint32_t *prefix, const uint8_t *values, const uint64_t *off,
                  uint64_t r_first, uint64_t ra, uint64_t rb, void *pml, int pml_width, uint8_t *cid);
// Chain ids alone, prefix groups [g0, g1) of a chunk of n_bases bases (cid points at the chunk's base 0).
void expand_cid_groups(const uint32_t *cid_words, const uint32_t *prefix, const uint8_t *values, uint64_t n_bases, uint64_t g0, uint64_t g1, uint8_t *cid);

// Kernels launched through host wrappers (traverse.cu).
// Device-side packing of a chunk's raw bytes (traverse.cu: k_pack_reads).
int launch_pack(const DeviceTable &dt, const uint8_t *d_bytes, ReadMeta *d_meta, uint32_t n_reads, uint32_t *d_words, ReadMeta *d_meta_b,
                uint32_t *d_n_irregular, cudaStream_t stream);
// One traversal of a batch: packed pass, byte pass, chain fix-up (those that have work).  d_counters: 4 x u64 scratch
// ([0],[1] work cursors, [2] chunks re-traversed by k_fixup, [3] irregular-read count of the device packer).
int launch_traverse(const DeviceTable &dt, const BatchView &bv, int pml_width, unsigned long long *d_counters,
                    cudaStream_t stream);

} // namespace colbwt
