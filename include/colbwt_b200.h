/* colbwt_b200.h -- C-ABI of the B200-native col-bwt query path (PML + chain ids per base).
 *
 * The reference (drnatebrown/col-bwt) has no FFI; its boundary for this path is the C++ class
 * `col_pml` plus the `pml_query` executable.  Each entry point below names the reference interface it
 * replaces (paths relative to the reference tree):
 *
 *   colbwt_index_load        col_pml::load / col_bwt::load           include/col_bwt.hpp:375-380
 *                            + LF_table::load                        include/ds/LF_table.hpp:347-357
 *                            (called from src/pml_query.cpp:110-113)
 *   colbwt_index_from_rows   the same, from memory (rows as read_vec would fill them,
 *                            include/common/common.hpp:318-323)
 *   colbwt_index_from_primaries / colbwt_index_save
 *                            src/build_col_bwt.cpp:8-64 (constructors + serialize), on the GPU
 *   colbwt_col_split         src/col_split.cpp + include/col_split.hpp (+ FL_table.hpp), on the GPU
 *   colbwt_rlbwt_to_bwt      src/rlbwt_to_bwt.cpp:8-34, on the GPU
 *   colbwt_index_stats       col_bwt::bwt_stats / runs / size        include/col_bwt.hpp:331-344
 *   colbwt_query             col_pml::query_pml(const char*, size_t) include/col_bwt.hpp:409-412, for a whole
 *                            batch of reads (the per-read loop of src/pml_query.cpp:74-86)
 *   colbwt_query_compact     the same result in compact form (one match bit per base + the non-zero chain ids) for
 *                            consumers that do not need dense arrays; colbwt_compact_expand rebuilds the dense
 *                            arrays of col_pml::query_pml from it bit-exactly
 *   colbwt_batch_*           the same split into upload / traverse / download so the traversal can be
 *                            timed with its inputs resident in HBM
 *   colbwt_format_stats      the text writer of pml_to_vec           src/pml_query.cpp:79-85
 *
 * Plain pointers and sizes only.  All functions return COLBWT_OK (0) or a negative colbwt_status;
 * colbwt_last_error() gives a message for the calling thread.  There is no CPU fallback: without a CUDA
 * device every entry point that computes returns COLBWT_ERR_CUDA.
 *
 * Semantics (bit-exact with the reference, SURVEY.md section 8a checklist): reads are raw bytes, compared
 * byte-exactly with the row characters; outputs are indexed by read position, pml[off[i]+j] / cid[off[i]+j]
 * for base j of read i; a zero-length read produces no values.
 */
#ifndef COLBWT_B200_H
#define COLBWT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum colbwt_status {
    COLBWT_OK = 0,
    COLBWT_ERR_IO = -1,          /* file missing or short                                   */
    COLBWT_ERR_FORMAT = -2,      /* header inconsistent (size != r, r == 0, r >= 2^32)      */
    COLBWT_ERR_ROW_TOO_LONG = -3,/* a row of >= 65536 symbols (16-bit offset field wraps in the reference, LF_table.hpp:39) */
    COLBWT_ERR_CUDA = -4,        /* no device / CUDA runtime failure                        */
    COLBWT_ERR_ARG = -5,         /* bad argument (null pointer, PML width too small, ...)   */
    COLBWT_ERR_NOMEM = -6        /* allocation failed, or a caller buffer is too small      */
} colbwt_status;

typedef struct colbwt_index colbwt_index;   /* move table replicated in the HBM of one or more GPUs */
typedef struct colbwt_batch colbwt_batch;   /* one batch of reads resident on one GPU               */

typedef struct colbwt_stats {
    uint64_t n;            /* BWT length                        (LF_table.hpp:359) */
    uint64_t r;            /* rows = sub-runs                   (LF_table.hpp:360) */
    uint64_t bwt_r;        /* BWT runs before sub-run splitting (col_bwt.hpp:382)  */
    uint64_t marked_rows;  /* rows with col_id != 0                                 */
    uint64_t slow_rows;    /* rows whose reposition data did not fit the packed row (resolved by exact search) */
    uint64_t device_bytes; /* HBM bytes held per device                              */
    uint32_t max_row_len;
    int32_t  n_devices;
} colbwt_stats;

/* Width in bytes of one PML value in caller buffers: 1 (every read shorter than 256 bases), 2 (shorter than
 * 65536) or 4.  A PML never exceeds the read length, so the narrow widths lose nothing; they cut the D2H bytes. */
#define COLBWT_PML_U8 1
#define COLBWT_PML_U16 2
#define COLBWT_PML_U32 4

/* ---- index ------------------------------------------------------------------------------------------- */

/* Load PATH (the reference's `<prefix>.col_pml`; if PATH does not exist, PATH + ".col_pml" is tried, which
 * is how pml_query.cpp:110-111 forms the name) and upload it to `n_devices` GPUs (`devices[i]` = CUDA
 * ordinal; devices == NULL means 0..n_devices-1). */
int colbwt_index_load(const char *path, const int *devices, int n_devices, colbwt_index **out);

/* Same from memory: `rows` = r packed 18-byte col_thr rows exactly as they sit in the file.  `rows` may be host memory or
 * device memory (a table produced on a GPU): device rows are unpacked without a host hop. */
int colbwt_index_from_rows(const void *rows, uint64_t bwt_r, uint64_t n, uint64_t r,
                           const int *devices, int n_devices, colbwt_index **out);

/* Build the table on the GPU from the primaries that `col-bwt build --keep` leaves on disk, without the `.col_pml`
 * intermediate: PREFIX.bwt.heads, PREFIX.bwt.len, PREFIX.thr_pos, PREFIX.col_runs (sdsl bit_vector as written by
 * col_split.hpp:384-386) and PREFIX.col_ids.  Replaces src/build_col_bwt.cpp:8-64, i.e. the constructor
 * col_bwt(heads, lengths, col_ids, splits) (include/col_bwt.hpp:124-230), LF_table::compute_table
 * (include/ds/LF_table.hpp:365-387) and col_pml::read_thresholds (include/col_bwt.hpp:440-457). */
int colbwt_index_from_primaries(const char *prefix, const int *devices, int n_devices, colbwt_index **out);

/* Write the table as `.col_pml` (col_bwt::serialize, include/col_bwt.hpp:360-370 + LF_table.hpp:325-342): byte-identical
 * to what the reference's build_col_bwt writes for the same primaries. */
int colbwt_index_save(const colbwt_index *idx, const char *path);

/* Multi-MUM sub-run marking on the GPU: reads PREFIX.bwt.heads, PREFIX.bwt.len and PREFIX.col_mums, writes
 * PREFIX.col_runs (sdsl bit_vector) and PREFIX.col_ids exactly as the reference's `col_split PREFIX -m MODE -s RATE`
 * (src/col_split.cpp:62-141, include/col_split.hpp:54-157, walking include/ds/FL_table.hpp) does -- without the
 * intermediate PREFIX.FL_table.  mode_all = 0: `-m tunnels`, 1: `-m all`; overlap handling is the reference's default
 * (`append`).  Optional outputs: number of set bits of col_runs, number of them with a non-zero chain id. */
int colbwt_col_split(const char *prefix, int mode_all, int split_rate, int device, uint64_t *n_set_bits, uint64_t *n_marked);

/* Run-length BWT -> plain BWT: reads PREFIX.bwt.heads and PREFIX.bwt.len, writes PREFIX.bwt exactly as the reference's
 * `rlbwt_to_bwt PREFIX` (src/rlbwt_to_bwt.cpp:8-34) does -- each head byte, unmodified, repeated `len` times.
 * Optional output: the number of bytes written. */
int colbwt_rlbwt_to_bwt(const char *prefix, int device, uint64_t *n_out);

int colbwt_index_stats(const colbwt_index *idx, colbwt_stats *out);
void colbwt_index_free(colbwt_index *idx);

/* ---- query, host buffers (the drop-in call) ------------------------------------------------------------ */

/* n_reads reads: bytes seqs[off[i] .. off[i+1]).  pml has off[n_reads] elements of `pml_width` bytes, cid has
 * off[n_reads] bytes.  Reads are packed on the host (2 bit/base; reads with a byte outside ACGT travel as
 * bytes), streamed through pinned multi-buffered copies (6 chunks in flight per GPU), traversed on the GPU(s) and the results copied back
 * in input order.  With several devices, chunks of reads are dealt round-robin; no inter-GPU communication. */
int colbwt_query(colbwt_index *idx, const uint8_t *seqs, const uint64_t *off, uint64_t n_reads,
                 void *pml, int pml_width, uint8_t *cid);

/* ---- query, compact result ------------------------------------------------------------------------------
 * The dense result of col_pml::query_pml (include/col_bwt.hpp:409-412) is highly redundant: the length is reset to 0 on a
 * mismatch and incremented on a match (col_bwt.hpp:516-523), so a read's PML array is a function of one match bit per
 * base (PML[j] = distance from j to the first mismatch at or right of j, or to the end of the read), and chain ids are 0
 * off the marked sub-runs.  The compact result holds exactly that -- about 0.3 bytes per base on BASELINE configs[1]
 * instead of 2-5 -- so PCIe stops bounding the call.  Layout of the caller's buffer (all offsets in bytes from its start):
 *   colbwt_compact_header | colbwt_compact_segment[n_segments] | per segment: match words, chain-id words, prefix | values
 * A segment covers consecutive reads; bit b of word w (u32, little endian) of its match / chain-id words refers to base
 * 32w+b counted from the segment's first base; prefix[g] (u32) = number of non-zero chain ids before base 2048g of the
 * segment, prefix[ceil(n_bases/2048)] = n_values; values = the non-zero chain ids (u8) in base order. */
#define COLBWT_COMPACT_MAGIC 0x31504d4354574243ull   /* "CBWTCMP1" */
typedef struct colbwt_compact_header {
    uint64_t magic, n_segments, n_reads, n_bases, bytes_used, reserved[3];
} colbwt_compact_header;
typedef struct colbwt_compact_segment {
    uint64_t first_read, n_reads;     /* reads [first_read, first_read + n_reads) of the call                    */
    uint64_t first_base, n_bases;     /* first_base = off[first_read] - off[0]                                   */
    uint64_t match_off, cid_off;      /* ceil(n_bases/32) u32 words each                                         */
    uint64_t prefix_off;              /* ceil(n_bases/2048) + 1 u32 entries                                      */
    uint64_t values_off, n_values;
} colbwt_compact_segment;

/* Bytes that always suffice for colbwt_query_compact on these reads (every chain id non-zero); 0 on bad offsets. */
size_t colbwt_compact_bound(const uint64_t *off, uint64_t n_reads);
/* Same reads, same traversal as colbwt_query; `result` (capacity bytes, ideally pinned: then the GPU writes into it
 * directly) receives the compact form.  COLBWT_ERR_NOMEM if it does not fit.  *bytes_used (may be NULL) = bytes written. */
int colbwt_query_compact(colbwt_index *idx, const uint8_t *seqs, const uint64_t *off, uint64_t n_reads,
                         void *result, size_t capacity, size_t *bytes_used);
/* Host-side: rebuild the dense arrays colbwt_query would have returned (pml_width-byte PML, u8 chain ids) for the
 * same reads.  Runs on the library's host threads; needs no GPU.  pml may be NULL: only the chain ids are rebuilt. */
int colbwt_compact_expand(const void *result, const uint64_t *off, uint64_t n_reads, void *pml, int pml_width, uint8_t *cid);

/* Where the last colbwt_query on this index packed the reads into 2 bits per base: 0 on the host, 1 on the device (raw
 * bytes copied as they are; only when `seqs` is pinned and the reads are short).  The library measures both ways on large
 * calls and keeps the faster one; COLBWT_DEVICE_PACK=0|1 pins the choice.  No reference counterpart (diagnostic). */
int colbwt_index_last_packing(const colbwt_index *idx);
/* How the dense results of the last colbwt_query crossed the link: 0 as they are, 1 in the compact form above, expanded
 * into the caller's arrays by the library's host threads, 2 PML copied as it is and only the chain ids (sparse) in compact
 * form, 3 chunks alternating between 0 and 1 so that the copy engine and the host threads fill the arrays together (2 and 3
 * need pinned result arrays).  Measured the same way as the packing; COLBWT_COMPACT_D2H=0|1|2|3 pins it. */
int colbwt_index_last_transport(const colbwt_index *idx);
/* Bytes the last colbwt_query / colbwt_query_compact on this index asked the copy engines to move, host to device and
 * device to host (diagnostic; what bench.py reports as h2d/d2h bytes per step). */
int colbwt_index_last_bytes(const colbwt_index *idx, uint64_t *h2d_bytes, uint64_t *d2h_bytes);

/* Pinned host memory for seqs / pml / cid buffers (lets colbwt_query copy without a staging hop). */
void *colbwt_host_alloc(size_t bytes);
void colbwt_host_free(void *p);

/* ---- query, device-resident batch (for kernel-only timing and for callers that keep results on device) -- */

int colbwt_batch_upload(colbwt_index *idx, int device_slot, const uint8_t *seqs, const uint64_t *off,
                        uint64_t n_reads, int pml_width, colbwt_batch **out);
/* Run the traversal `iters` times back to back; *ms_per_iter (may be NULL) = CUDA-event time / iters measured
 * on the launching stream.  Outputs stay on the device. */
int colbwt_batch_run(colbwt_batch *b, int iters, float *ms_per_iter);
int colbwt_batch_download(colbwt_batch *b, void *pml, uint8_t *cid);
/* Device pointers of the results (pml: pml_width-byte elements, cid: bytes), valid until colbwt_batch_free. */
int colbwt_batch_device_ptrs(colbwt_batch *b, void **pml_dev, void **cid_dev, uint64_t *n_bases);
/* Kernel launches issued by one traversal of this batch (for bench.py's gpu_launches). */
int colbwt_batch_launches(const colbwt_batch *b);
/* Long-read bookkeeping of this batch (diagnostic): out[0] = work items scheduled longest-first (chunk tasks of the reads
 * that were cut + whole reads ordered among them), out[1] = reads that were cut, out[2] = chunk tasks whose speculative
 * start did not converge and were re-traversed in the last colbwt_batch_run traversal, out[3] = chunk tasks of cut reads. */
int colbwt_batch_counters(colbwt_batch *b, uint64_t out[4]);
void colbwt_batch_free(colbwt_batch *b);

/* ---- text output of pml_query (src/pml_query.cpp:79-85) -------------------------------------------------- */

/* Formats one read: ">" id " \n", every value followed by " ", then "\n".  Returns bytes written, or the
 * bytes needed if cap is too small (nothing is written then).  width = element bytes of `values` (1, 2, 4). */
size_t colbwt_format_stats(char *buf, size_t cap, const char *id, size_t id_len,
                           const void *values, int width, uint64_t m);

/* ---- measurement helpers ------------------------------------------------------------------------------- */

/* Random-gather roofline microbenchmark on `device`: `loads` independent 16-byte loads, each from a random
 * 32-byte sector of a `bytes`-sized buffer.  bit 0 of `dependent` chains each thread's next address on the loaded
 * value (latency-bound variant with the same occupancy); bits 8.. select the load flavour under test (0 = ld.global.nc).  *sectors_per_s receives the measured rate. */
int colbwt_gather_bench(int device, uint64_t bytes, uint64_t loads, int dependent, double *sectors_per_s);

const char *colbwt_last_error(void);
const char *colbwt_version(void);

#ifdef __cplusplus
}
#endif
#endif /* COLBWT_B200_H */
