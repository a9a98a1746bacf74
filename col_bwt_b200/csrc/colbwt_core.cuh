// colbwt_core.cuh -- packed move-table row, row builder and the per-lane traversal state machine.
//
// Everything here is `__host__ __device__` so that tests/native/emulate.cpp can run the *same* logic the
// kernels run, one lane at a time on the CPU, against the oracle (test infrastructure only: the product
// never executes this on the host).
//
// Reference semantics being reproduced (paths relative to the reference tree):
//   col_pml::_query_pml        include/col_bwt.hpp:498-529   per-base loop, CID sampled before reposition
//   col_pml::threshold_step    include/col_bwt.hpp:531-574   successor / threshold / predecessor choice
//   LF_table::LF               include/ds/LF_table.hpp:251-262 LF step + fast-forward
//   LF_table::pred/succ_char   include/ds/LF_table.hpp:271-298 (replaced by precomputed per-row distances,
//                              with an exact search over per-character row lists when they do not fit)
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define CB_HD __host__ __device__ __forceinline__
#else
#define CB_HD inline
#endif

namespace colbwt {

// ---------------------------------------------------------------------------------------------------------
// Packed row: 16 bytes, one 128-bit load.
//   x            dest   LF destination row                                    (LF_row::interval, 32 bit)
//   y[ 0:16)     doff   offset inside the destination row                     (LF_row::offset,   16 bit)
//   y[16:32)     len    row length, 1..65535                                  (idx[k+1]-idx[k])
//   meta = z | w<<32:
//     [ 0: 8)    cid    chain id                                              (col_row::col_id)
//     [ 8:11)    chc    0..3 = A,C,T,G ((byte>>1)&3), 4 = any other byte (exact byte in ch8[])
//     [11:17)    mode   3 x 2 bit, slot s = (c - chc - 1) & 3 for read code c != chc:
//                         0 jump to successor row (offset 0)          1 jump to predecessor row (offset len-1)
//                         2 predecessor if offset < fx else successor  3 resolve by exact search (slow path)
//     [17:47)    dist   3 x (dp:5 | ds:5<<5): distance to the predecessor / successor row of that character
//     [47:63)    fx     the one in-row flip offset = clamp(thr[succ] - idx[k], 0, len)
// The reference's test `pos < thr[succ]` (col_bwt.hpp:560) with pos = idx[k]+offset is offset < thr[succ]-idx[k];
// for almost every (row, character) that is constant over the row, so two bits replace two 40-bit compares.
// ---------------------------------------------------------------------------------------------------------
struct Row {
    uint32_t dest, offlen, m0, m1;
};

constexpr int CHC_OTHER = 4;
constexpr int CODE_OTHER = 4;   // read byte present in the table but not A/C/G/T
constexpr int CODE_ABSENT = 5;  // read byte that no row carries
constexpr uint32_t MAX_DIST = 31;
constexpr uint32_t MAX_ROW_LEN = 65535;

CB_HD int primary_code(uint8_t c) { return (c == 'A') ? 0 : (c == 'C') ? 1 : (c == 'T') ? 2 : (c == 'G') ? 3 : -1; }
CB_HD uint8_t primary_byte(int code) { return code == 0 ? 'A' : code == 1 ? 'C' : code == 2 ? 'T' : 'G'; }

CB_HD uint32_t row_len(const Row &r) { return r.offlen >> 16; }
CB_HD uint32_t row_doff(const Row &r) { return r.offlen & 0xFFFFu; }
CB_HD uint32_t row_cid(const Row &r) { return r.m0 & 0xFFu; }
CB_HD uint32_t row_chc(const Row &r) { return (r.m0 >> 8) & 7u; }
CB_HD uint64_t row_meta(const Row &r) { return (uint64_t)r.m0 | ((uint64_t)r.m1 << 32); }

// ---------------------------------------------------------------------------------------------------------
// Narrow layout: the same row split into two 8-byte words so that the array every step gathers from is half as
// large (twice the L2 reach; an L2 miss on this machine always costs a full 128-byte line).
//   hot  = dest:32 | doff:10 | len:10 | chc:3 | cid:8      (every step)
//   cold = mode:3x2 | dist:3x10 | fx:16                    (= wide meta >> 11; only after a mismatch)
// Usable when every row is shorter than 1024 symbols.
// ---------------------------------------------------------------------------------------------------------
constexpr uint32_t NARROW_MAX_LEN = 1023;
CB_HD uint64_t hot_from_row(const Row &r)
{
    return (uint64_t)r.dest | ((uint64_t)(row_doff(r) & 1023u) << 32) | ((uint64_t)(row_len(r) & 1023u) << 42) |
           ((uint64_t)row_chc(r) << 52) | ((uint64_t)row_cid(r) << 55);
}
CB_HD uint64_t cold_from_row(const Row &r) { return row_meta(r) >> 11; }

// Device-resident table (all pointers in HBM).
struct TableView {
    const Row *rows;            // r packed rows (wide layout)
    const uint64_t *hot;        // narrow layout: dest/doff/len/chc/cid per row (nullptr when not built)
    const uint64_t *cold;       // narrow layout: reposition modes/distances/flip offset per row
    const uint8_t *ch8;         // r exact row bytes                      (slow path, "other" characters)
    const uint64_t *idx;        // r BWT start positions                  (slow path: pos = idx[k]+offset)
    const uint64_t *thr;        // r thresholds                           (slow path)
    const uint32_t *char_rows;  // row numbers sorted by (byte, row)      (slow path: exact pred/succ search)
    const uint32_t *char_start; // 257 offsets into char_rows
    uint64_t n;
    uint32_t r;
    uint32_t last_len;          // len(r-1): initial offset is last_len-1 (col_bwt.hpp:503-505)
};

struct BuildView {              // inputs of build_row: unpacked reference columns
    const uint8_t *ch8;
    const uint64_t *idx;
    const uint64_t *thr;
    const uint32_t *dest;
    const uint16_t *doff;
    const uint8_t *cid;
    uint64_t n;
    uint32_t r;
};

constexpr uint32_t BUILD_FLAG_BAD_LEN = 1u;   // row of length 0 or >= 65536
constexpr uint32_t BUILD_FLAG_SLOW = 2u;      // some (row, character) needs the exact search
constexpr uint32_t BUILD_FLAG_BAD_LF = 4u;    // dest >= r

#if defined(__CUDACC__) && defined(__CUDA_ARCH__)
// Flavour of the row gather (development knob; 0 is what ships unless a measurement says otherwise):
//   0 ld.global.nc (LDG.E.128.CONSTANT, allocates in the L1)   1 ld.global.cg (L2 only)
//   2 ld.global.nc.L1::no_allocate                               3 ld.global.L1::no_allocate
#ifndef COLBWT_ROW_LOAD
#define COLBWT_ROW_LOAD 0
#endif
CB_HD Row ld_row(const Row *p)
{
    uint4 v;
#if COLBWT_ROW_LOAD == 1
    v = __ldcg(reinterpret_cast<const uint4 *>(p));
#elif COLBWT_ROW_LOAD == 2
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
#elif COLBWT_ROW_LOAD == 3
    asm volatile("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
#else
    v = __ldg(reinterpret_cast<const uint4 *>(p));   // LDG.E.128.CONSTANT: one 128-bit read-only gather
#endif
    return Row{v.x, v.y, v.z, v.w};
}
CB_HD uint64_t ld_row64(const uint64_t *p) { return __ldg(reinterpret_cast<const unsigned long long *>(p)); }
template <typename T> CB_HD T ld_ro(const T *p) { return __ldg(p); }
CB_HD void st_v4(void *p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) { *reinterpret_cast<uint4 *>(p) = make_uint4(a, b, c, d); }
CB_HD void st_v2(void *p, uint32_t a, uint32_t b) { *reinterpret_cast<uint2 *>(p) = make_uint2(a, b); }
#else
CB_HD Row ld_row(const Row *p) { return *p; }
CB_HD uint64_t ld_row64(const uint64_t *p) { return *p; }
template <typename T> CB_HD T ld_ro(const T *p) { return *p; }
#endif
// (L2 eviction-priority hints -- evict_last on the table, evict_first on the streams -- were measured and dropped:
//  no gain on this access pattern, profiles/r1/variant_sweep1.log.)

// Build the packed row k from the reference columns.  Pure function of the columns.
CB_HD Row build_row(const BuildView &b, uint32_t k, uint32_t *flags)
{
    const uint64_t i0 = b.idx[k];
    const uint64_t i1 = (k + 1 < b.r) ? b.idx[k + 1] : b.n;
    const uint64_t len = i1 - i0;
    if (i1 <= i0 || len > MAX_ROW_LEN) *flags |= BUILD_FLAG_BAD_LEN;
    if (b.dest[k] >= b.r) *flags |= BUILD_FLAG_BAD_LF;
    const int rc = primary_code(b.ch8[k]);
    const uint32_t chc = rc < 0 ? CHC_OTHER : (uint32_t)rc;
    uint64_t meta = (uint64_t)b.cid[k] | ((uint64_t)chc << 8);
    if (rc < 0) {
        // every mismatch on this row goes through the exact search: all mode fields (and the two bits a
        // slot-3 decode would read) are 3
        meta |= (uint64_t)0x3F << 11;
        meta |= (uint64_t)0x3FFFFFFF << 17;
        *flags |= BUILD_FLAG_SLOW;
    } else {
        bool fx_set = false;
        uint64_t fx = 0;
        for (int slot = 0; slot < 3; ++slot) {
            const uint8_t cb = primary_byte((rc + 1 + slot) & 3);
            uint32_t dp = 0, ds = 0;
            for (uint32_t d = 1; d <= MAX_DIST && d <= k; ++d)
                if (b.ch8[k - d] == cb) { dp = d; break; }
            for (uint32_t d = 1; d <= MAX_DIST && (uint64_t)k + d < b.r; ++d)
                if (b.ch8[k + d] == cb) { ds = d; break; }
            const bool pred_none = (dp == 0) && (k <= MAX_DIST);                        // scanned to row 0
            const bool succ_none = (ds == 0) && ((uint64_t)k + MAX_DIST >= b.r - 1);    // scanned to row r-1
            uint32_t mode = 3;
            if (ds) {
                const uint64_t t = b.thr[k + ds];
                const uint64_t flip = (t <= i0) ? 0 : ((t - i0 >= len) ? len : t - i0);
                if (flip == 0 || pred_none) mode = 0;            // successor whatever the offset
                else if (dp) {
                    if (flip >= len) mode = 1;                   // predecessor whatever the offset
                    else if (!fx_set || fx == flip) { mode = 2; fx = flip; fx_set = true; }
                }
            } else if (succ_none && dp) {
                mode = 1;                                        // no successor: thr = n, predecessor wins
            }
            if (mode == 3) *flags |= BUILD_FLAG_SLOW;
            meta |= (uint64_t)mode << (11 + 2 * slot);
            meta |= (uint64_t)(dp | (ds << 5)) << (17 + 10 * slot);
        }
        meta |= (fx & 0xFFFF) << 47;
    }
    Row r;
    r.dest = b.dest[k];
    r.offlen = (uint32_t)b.doff[k] | ((uint32_t)(len & 0xFFFF) << 16);
    r.m0 = (uint32_t)meta;
    r.m1 = (uint32_t)(meta >> 32);
    return r;
}

// Exact restatement of threshold_step's search (col_bwt.hpp:531-574) over the sorted per-character row lists.
// Returns false when neither a successor nor a predecessor exists (state unchanged).
CB_HD bool slow_reposition(const TableView &t, uint32_t cur, uint32_t off, uint8_t c, uint32_t *tgt, bool *use_pred)
{
    uint32_t lo = ld_ro(t.char_start + c), hi = ld_ro(t.char_start + c + 1);
    if (lo == hi) return false;
    const uint32_t base = lo, end = hi;
    while (lo < hi) {                                   // first list entry > cur (cur itself has another byte)
        uint32_t mid = lo + ((hi - lo) >> 1);
        if (ld_ro(t.char_rows + mid) > cur) hi = mid; else lo = mid + 1;
    }
    const bool has_s = lo < end, has_p = lo > base;
    uint64_t thr = t.n;
    bool found = false;
    if (has_s) {
        const uint32_t s = ld_ro(t.char_rows + lo);
        thr = ld_ro(t.thr + s);
        *tgt = s;
        *use_pred = false;
        found = true;
    }
    const uint64_t pos = ld_ro(t.idx + cur) + off;
    if (pos < thr && has_p) {
        *tgt = ld_ro(t.char_rows + lo - 1);
        *use_pred = true;
        found = true;
    }
    return found;
}

// ---------------------------------------------------------------------------------------------------------
// One batch of reads on the device.
// ---------------------------------------------------------------------------------------------------------
struct ReadMeta {           // 16 bytes per read
    uint64_t out_off;       // index of base 0 of the read in pml[] / cid[]
    uint32_t len;
    uint32_t in_off;        // packed reads: 32-bit word index into words[]; byte reads: byte index into bytes[]
};

// Long reads are cut into chunk tasks so that one read's serial chain does not bound the whole batch.  A read is
// traversed right to left, so chunk 0 is the top chunk [hi0=len .. lo0) and starts from the true initial state; every
// other chunk starts SPECULATIVELY from the initial state `warm` bases above its range (top = hi + warm), runs the
// warm-up without emitting, and records the state in which it arrives at base hi-1.  After the main pass, the chain
// of a read is verified top-down (fixup_chain): a chunk whose recorded arrival state equals the true state left by
// the chunk above it is exact as it stands (states merge at reposition jumps, so a few hundred noisy bases are
// enough); any other chunk is re-traversed from the true state by one lane.  The result is bit-identical to the
// serial traversal whatever the warm-up length; the warm-up only decides how often the re-traversal is needed.
constexpr uint32_t NO_SLOT = 0xFFFFFFFFu;
struct ChunkTask {          // 32 bytes
    uint64_t out_off;       // index of base 0 of the READ in pml[] / cid[]
    uint32_t in_off;        // the read's first word (packed) or byte
    uint32_t lo, hi;        // bases [lo, hi) are emitted by this task
    uint32_t top;           // bases [hi, top) are the warm-up (top == hi for the first chunk of a read)
    uint32_t slot;          // index of this chunk's ChainState records
    uint32_t len;           // top - lo: work in this task (sort key)
};
struct ChainState {         // traversal state just before a base is consumed
    uint32_t row, off, plen, pad;
};
struct ChainDesc {          // one split read: its chunk tasks occupy slots [first_slot, first_slot + n_chunks), top chunk first
    uint32_t first_slot, n_chunks, packed, pad;
};

struct BatchView {
    const ReadMeta *meta;   // n_packed entries, input order; reads shipped as bytes (or split into tasks) have len = 0 here
    const ReadMeta *meta_b; // n_bytes entries: the reads that contain something outside ACGT
    const ChunkTask *tasks;     // n_tasks + n_tasks_b chunk tasks in scheduling order (packed ones first)
    const ChunkTask *by_slot;   // the same tasks indexed by slot (for fixup_chain)
    ChainState *start_state;    // [slots] state recorded on arrival at base hi-1
    ChainState *end_state;      // [slots] state after base lo was consumed
    const ChainDesc *chains;    // n_chains split reads
    uint32_t n_tasks, n_tasks_b, n_chains, pad;
    const uint32_t *words;  // 2 bit per base, 16 bases per word, base j of a read at bits 2*(j&15) of word j>>4
    const uint8_t *bytes;   // raw bytes of the reads that contain something outside ACGT
    void *pml;              // PmlT[total bases]
    uint8_t *cid;           // [total bases]
    uint32_t n_packed, n_bytes;
    const uint32_t *n_bytes_dev;   // when set (reads packed on the device), the byte-read count lives here instead of n_bytes
};

enum LaneState : uint32_t { LANE_IDLE = 0, LANE_LF = 1, LANE_REPOS_SUCC = 2, LANE_REPOS_PRED = 3,
                            LANE_COLD = 4 /* narrow: gather cold[addr]; read code in bits 8-9, row chc in bits 10-11 */,
                            LANE_RELOAD = 5 /* narrow: nothing found, gather hot[addr] again and LF through it */ };

// Per-lane registers.  Every iteration of the traversal loop performs exactly one row gather per lane
// (`rows[addr]`) whatever the lane needs next -- LF destination, fast-forward neighbour or reposition target --
// so no lane ever waits for another lane's extra gather, and lanes that finish a read take the next one.
template <typename PmlT> struct Lane {
    uint32_t state = LANE_IDLE;
    uint32_t addr = 0;      // row to gather this iteration
    uint32_t off = 0;       // LANE_LF: offset carried into rows[addr] (may exceed its length: fast-forward)
    uint32_t plen = 0;      // current pseudo matching length
    uint32_t j = 0;         // bases still to process; the next base is j-1 (right to left, col_bwt.hpp:512)
    uint32_t in_off = 0;
    uint32_t j_stop = 0;    // stop after base j_stop (0 for whole reads, lo for chunk tasks)
    uint32_t emit_top = 0;  // bases >= emit_top are warm-up: traversed, not emitted
    uint32_t slot = NO_SLOT;
    uint32_t rw = 0;        // current packed word
    uint64_t out_base = 0;
    uint32_t hi_slot = 64;  // highest valid slot of the output block being staged (>= block size: nothing staged yet)
    CB_HD static PmlT plen_type() { return PmlT(); }   // lets generic test code name PmlT
    uint32_t flush = 0;     // deferred flush request (warp-cooperative flush in k_traverse): 0 = none, else highest slot + 1
};

// ---------------------------------------------------------------------------------------------------------
// Output staging.  Scattered 8/16-byte stores from ~150 k lanes cost 20 % of the random-gather rate (each becomes its
// own partial DRAM write: profiles/r1/l2stream*.log), 64-byte bursts per lane do not.  Every lane therefore stages one
// aligned block of 64 output positions -- 64 CID bytes and 64 PML values -- in shared memory and writes it out as
// consecutive 16-byte vectors when the block is complete (or the read ends).  Words are interleaved across the
// threads of the CTA (word w of thread t at [w * stride + t]), so a warp's accesses never conflict on a bank.
// ---------------------------------------------------------------------------------------------------------
struct Stage {
    uint32_t *base;     // this thread's word 0
    uint32_t stride;    // words between consecutive words of one thread
    uint32_t swz;       // word w lives at index w ^ swz (< 16): k_traverse keeps each lane's block contiguous and rotates
                        // it by the lane number, so that 32 lanes touching the same word number hit different banks
};
// Positions staged per lane.  The stage competes with the L1 for the SM's 256 KB, and the L1 is what holds the row
// gathers in flight and the line a fast-forward step re-reads: with 4 x 256 lanes per SM, 32 KB of stage per CTA leaves
// a 124 KB L1 (49.8 Gbases/s on C2), 34 KB pushes the carve-out to the next step and leaves 92 KB (42.0), a full
// carve-out 28 KB (24.2) -- profiles/r1/l1_capacity_probe.log, stage_sweep.log.  So 64 positions with 8-bit PML
// (64 + 64 bytes per lane), 32 with 16-bit PML (32 + 64 bytes), 16 with 32-bit PML (16 + 64 bytes).
#ifndef COLBWT_STAGE_BLOCK_U8
#define COLBWT_STAGE_BLOCK_U8 64
#endif
#ifndef COLBWT_STAGE_BLOCK_U16
#define COLBWT_STAGE_BLOCK_U16 32
#endif
#ifndef COLBWT_STAGE_BLOCK_U32
#define COLBWT_STAGE_BLOCK_U32 16
#endif
template <typename PmlT> CB_HD constexpr uint32_t stage_block()
{
    return sizeof(PmlT) == 1 ? COLBWT_STAGE_BLOCK_U8 : (sizeof(PmlT) == 2 ? COLBWT_STAGE_BLOCK_U16 : COLBWT_STAGE_BLOCK_U32);
}
template <typename PmlT> CB_HD constexpr uint32_t stage_cid_words() { return stage_block<PmlT>() / 4; }
template <typename PmlT> CB_HD constexpr uint32_t stage_words() { return stage_cid_words<PmlT>() + stage_block<PmlT>() * (uint32_t)sizeof(PmlT) / 4; }

// Bank rotation mask of a lane's block (see Stage::swz): the rotation must stay inside the block.
template <typename PmlT> CB_HD constexpr uint32_t stage_swz_mask() { return ((stage_words<PmlT>() & (0u - stage_words<PmlT>())) - 1u) & 15u; }

CB_HD uint32_t &stage_word(const Stage &sg, uint32_t w) { return sg.base[(w ^ sg.swz) * sg.stride]; }

// Write slots [lo, hi] of one staged array (element size ES bytes, first word w0) to dst (= address of slot 0).
template <int ES, uint32_t BLOCK> CB_HD void stage_write(const Stage &sg, uint32_t w0, uint8_t *dst, uint32_t lo, uint32_t hi)
{
    constexpr uint32_t PER_VEC = 16 / ES;                       // slots per 16-byte vector
#if defined(__CUDACC__) && defined(__CUDA_ARCH__)
    if (lo == 0 && hi == BLOCK - 1) {                     // complete block: one straight burst
#pragma unroll
        for (uint32_t q = 0; q < BLOCK / PER_VEC; ++q)
            st_v4(dst + q * 16, stage_word(sg, w0 + 4 * q), stage_word(sg, w0 + 4 * q + 1), stage_word(sg, w0 + 4 * q + 2), stage_word(sg, w0 + 4 * q + 3));
        return;
    }
#endif
#pragma unroll
    for (uint32_t q = 0; q < BLOCK / PER_VEC; ++q) {
        const uint32_t s0 = q * PER_VEC, s1 = s0 + PER_VEC - 1;
        if (s1 < lo || s0 > hi) continue;
        const uint32_t w = w0 + q * 4;
        if (s0 >= lo && s1 <= hi) {
#if defined(__CUDACC__) && defined(__CUDA_ARCH__)
            st_v4(dst + s0 * ES, stage_word(sg, w), stage_word(sg, w + 1), stage_word(sg, w + 2), stage_word(sg, w + 3));
#else
            for (uint32_t k = 0; k < 4; ++k) {
                const uint32_t v = stage_word(sg, w + k);
                for (uint32_t b = 0; b < 4; ++b) dst[s0 * ES + 4 * k + b] = (uint8_t)(v >> (8 * b));
            }
#endif
        } else {                                                // ragged edge of a read: element by element
            const uint32_t a = s0 > lo ? s0 : lo, e = s1 < hi ? s1 : hi;
            for (uint32_t sl = a; sl <= e; ++sl) {
                const uint32_t v = stage_word(sg, w0 + sl * ES / 4) >> (8 * ((sl * ES) & 3));
                if (ES == 1) dst[sl] = (uint8_t)v;
                else if (ES == 2) reinterpret_cast<uint16_t *>(dst)[sl] = (uint16_t)v;
                else reinterpret_cast<uint32_t *>(dst)[sl] = v;
            }
        }
    }
}

template <typename PmlT> CB_HD void lane_flush(Lane<PmlT> &L, const Stage &sg, const BatchView &bv, uint64_t g)
{
    constexpr uint32_t BLOCK = stage_block<PmlT>(), CIDW = stage_cid_words<PmlT>();
    const uint32_t lo = (uint32_t)(g & (BLOCK - 1));
    const uint64_t blk = g - lo;
    stage_write<1, BLOCK>(sg, 0, bv.cid + blk, lo, L.hi_slot);
    if (sizeof(PmlT) == 1) stage_write<1, BLOCK>(sg, CIDW, reinterpret_cast<uint8_t *>(bv.pml) + blk, lo, L.hi_slot);
    else if (sizeof(PmlT) == 2) stage_write<2, BLOCK>(sg, CIDW, reinterpret_cast<uint8_t *>(bv.pml) + 2 * blk, lo, L.hi_slot);
    else stage_write<4, BLOCK>(sg, CIDW, reinterpret_cast<uint8_t *>(bv.pml) + 4 * blk, lo, L.hi_slot);
    L.hi_slot = BLOCK - 1;                                // the next block below starts full-width
}

// Warp-cooperative flush (k_traverse): a lane that completes a block only posts a request (Lane::flush) and the whole
// warp then moves the block, one 32-bit word per thread -- word w of the flushing lane's stage goes to its place in
// CID (w < 16) or PML -- instead of one thread walking through ~230 instructions while 31 wait.  `block` = word 0 of
// the flushing lane's stage (contiguous words, rotated by swz as in Stage), [lo, hi] = valid slots, blk = global position of slot 0.  Words cut by
// the edge of a read are written element by element (the neighbouring elements belong to another read).
template <typename PmlT> CB_HD void flush_word(const uint32_t *block, uint32_t swz, uint32_t w, const BatchView &bv, uint64_t blk, uint32_t lo, uint32_t hi)
{
    if (w >= stage_words<PmlT>()) return;
    const bool is_cid = w < stage_cid_words<PmlT>();
    const uint32_t es = is_cid ? 1u : (uint32_t)sizeof(PmlT), per = 4u / es;
    const uint32_t s0 = (is_cid ? w : w - stage_cid_words<PmlT>()) * per, s1 = s0 + per - 1;
    if (s1 < lo || s0 > hi) return;
    uint8_t *dst = is_cid ? bv.cid + blk + s0 : reinterpret_cast<uint8_t *>(bv.pml) + (blk + s0) * es;
    const uint32_t v = block[w ^ swz];
    if (s0 >= lo && s1 <= hi) {
#if defined(__CUDACC__) && defined(__CUDA_ARCH__)
        *reinterpret_cast<uint32_t *>(dst) = v;
#else
        for (uint32_t b = 0; b < 4; ++b) dst[b] = (uint8_t)(v >> (8 * b));
#endif
        return;
    }
    for (uint32_t e = 0; e < per; ++e) {
        if (s0 + e < lo || s0 + e > hi) continue;
        if (es == 1) dst[e] = (uint8_t)(v >> (8 * e));
        else if (es == 2) reinterpret_cast<uint16_t *>(dst)[e] = (uint16_t)(v >> (16 * e));
        else *reinterpret_cast<uint32_t *>(dst) = v;
    }
}

template <bool DEFER = false, typename PmlT>
CB_HD void lane_emit(Lane<PmlT> &L, const Stage &sg, const BatchView &bv, uint32_t jj, uint32_t plen, uint32_t cid)   // jj < emit_top
{
    constexpr uint32_t BLOCK = stage_block<PmlT>(), CIDW = stage_cid_words<PmlT>();
    const uint64_t g = L.out_base + jj;
    const uint32_t slot = (uint32_t)(g & (BLOCK - 1));
    if (L.hi_slot >= BLOCK) L.hi_slot = slot;             // first output of this read / chunk
    reinterpret_cast<uint8_t *>(&stage_word(sg, slot >> 2))[slot & 3] = (uint8_t)cid;
    if (sizeof(PmlT) == 1) reinterpret_cast<uint8_t *>(&stage_word(sg, CIDW + (slot >> 2)))[slot & 3] = (uint8_t)plen;
    else if (sizeof(PmlT) == 2) reinterpret_cast<uint16_t *>(&stage_word(sg, CIDW + (slot >> 1)))[slot & 1] = (uint16_t)plen;
    else stage_word(sg, CIDW + slot) = plen;
    if (slot == 0 || jj == L.j_stop) {
        if (DEFER) {
            L.flush = L.hi_slot + 1;                            // the warp writes [slot, hi_slot] of this block after the step
            L.hi_slot = BLOCK - 1;
        } else {
            lane_flush(L, sg, bv, g);
        }
    }
}

// Start read `m` on this lane (zero-length reads are skipped by the caller).
template <bool PACKED, typename PmlT> CB_HD void lane_begin(Lane<PmlT> &L, const TableView &t, const BatchView &bv, const ReadMeta &m)
{
    L.state = LANE_LF;
    L.addr = t.r - 1;               // col_bwt.hpp:504
    L.off = t.last_len - 1;         // col_bwt.hpp:505
    L.plen = 0;
    L.j = m.len;
    L.j_stop = 0;
    L.emit_top = m.len;
    L.slot = NO_SLOT;
    L.in_off = m.in_off;
    L.out_base = m.out_off;
    L.hi_slot = stage_block<PmlT>();
    if (PACKED) L.rw = ld_ro(bv.words + m.in_off + ((m.len - 1) >> 4));
}

// Start a chunk task: from the initial state at base top-1 (speculative unless top is the read length) ...
template <bool PACKED, typename PmlT> CB_HD void lane_begin_task(Lane<PmlT> &L, const TableView &t, const BatchView &bv, const ChunkTask &k)
{
    L.state = LANE_LF;
    L.addr = t.r - 1;
    L.off = t.last_len - 1;
    L.plen = 0;
    L.j = k.top;
    L.j_stop = k.lo;
    L.emit_top = k.hi;
    L.slot = k.slot;
    L.in_off = k.in_off;
    L.out_base = k.out_off;
    L.hi_slot = stage_block<PmlT>();
    if (PACKED) L.rw = ld_ro(bv.words + k.in_off + ((k.top - 1) >> 4));
}

// ... or from a known state at base hi-1 (re-traversal of a chunk whose speculation failed).
template <bool PACKED, typename PmlT>
CB_HD void lane_begin_from_state(Lane<PmlT> &L, const BatchView &bv, const ChunkTask &k, const ChainState &s)
{
    L.state = LANE_LF;
    L.addr = s.row;
    L.off = s.off;
    L.plen = s.plen;
    L.j = k.hi;
    L.j_stop = k.lo;
    L.emit_top = k.hi;
    L.slot = k.slot;
    L.in_off = k.in_off;
    L.out_base = k.out_off;
    L.hi_slot = stage_block<PmlT>();
    if (PACKED) L.rw = ld_ro(bv.words + k.in_off + ((k.hi - 1) >> 4));
}

// In-trip resolution (template parameter INTRIP): a fast-forward neighbour or a reposition target that lies in the SAME
// 128-byte line as the row just gathered (8 rows of 16 bytes per line; the row array is line-aligned) is read at once instead
// of costing the lane another trip round the warp loop: fewer trips per base (C2 1.38 -> ~1.1), but every trip now waits
// for those dependent loads too, and by the time a warp has all its 32 rows the earliest lines have often left the L1
// (1024 lanes per SM share 992 lines), so the second load is an L2 access.  Measured (profiles/r2/r2_intrip_*.log): worse
// wherever the table is L2-resident or DRAM bandwidth is the limit (c2small 110.7 -> 71.8, C2 51.3 -> 46.1, c3small
// 43.9 -> 36.7 Gbases/s), better only where every gather is a DRAM + TLB miss and a trip is long anyway (c5mid, 4 GB table:
// 23.1 -> 25.8).  The launcher therefore turns it on for tables of 2 GiB and more only (traverse.cu; COLBWT_INTRIP=0|1 in the
// environment pins it, -DCOLBWT_INTRIP=1 makes it the default of the host emulation).
#ifndef COLBWT_INTRIP
#define COLBWT_INTRIP 0
#endif
CB_HD bool same_line(uint32_t a, uint32_t b) { return (a >> 3) == (b >> 3); }

// Advance the lane by one gathered row.  code_lut: 256-entry byte -> {0..3, CODE_OTHER, CODE_ABSENT} (byte reads only).
template <bool PACKED, bool DEFER = false, bool INTRIP = (COLBWT_INTRIP != 0), typename PmlT>
CB_HD void lane_step(Lane<PmlT> &L, const Stage &sg, const TableView &t, const BatchView &bv, Row row, const uint8_t *code_lut)
{
    uint32_t len = row_len(row);
    if (L.state != LANE_LF) {
        // rows[addr] is the reposition target (LF_table.hpp:282,297): predecessor at len-1, successor at 0
        const uint32_t o = (L.state == LANE_REPOS_PRED) ? len - 1 : 0;
        L.off = row_doff(row) + o;
        L.addr = row.dest;
        L.state = LANE_LF;
        return;
    }
    if (INTRIP) {
        while (L.off >= len && L.addr + 1 < t.r) {   // fast-forward (LF_table.hpp:256-259)
            L.off -= len;
            ++L.addr;
            if (!same_line(L.addr, L.addr - 1)) return;   // the neighbour starts another line: that is the next trip's gather
            row = ld_row(t.rows + L.addr);                // same 128-byte line as the row just gathered: no trip of its own
            len = row_len(row);
        }
    } else if (L.off >= len && L.addr + 1 < t.r) {   // fast-forward (LF_table.hpp:256-259)
        L.off -= len;
        ++L.addr;
        return;
    }
    // settled on (addr, off)
    if (L.slot != NO_SLOT) {
        if (L.j == L.emit_top) bv.start_state[L.slot] = ChainState{L.addr, L.off, L.plen, 0};   // arrival at base hi-1
        if (L.j == L.j_stop) {                                                                   // base lo consumed and stepped
            bv.end_state[L.slot] = ChainState{L.addr, L.off, L.plen, 0};
            L.state = LANE_IDLE;
            return;
        }
    }
    // process the next base
    const uint32_t jj = --L.j;
    uint32_t code;
    uint8_t cbyte = 0;
    if (PACKED) {
        code = (L.rw >> (2 * (jj & 15))) & 3u;
        if ((jj & 15) == 0 && jj != 0) L.rw = ld_ro(bv.words + L.in_off + ((jj - 1) >> 4));
    } else {
        cbyte = ld_ro(bv.bytes + L.in_off + jj);
        code = code_lut[cbyte];
    }
    const uint32_t chc = row_chc(row);
    const uint32_t cid = row_cid(row);          // sampled before any reposition (col_bwt.hpp:513)
    bool match = (code == chc);
    if (!PACKED && code >= CODE_OTHER)
        match = (code == CODE_OTHER) && (chc == CHC_OTHER) && (ld_ro(t.ch8 + L.addr) == cbyte);
    if (match) {
        ++L.plen;                                // col_bwt.hpp:516-518
    } else {
        L.plen = 0;                              // col_bwt.hpp:520-523
    }
    if (jj < L.emit_top) lane_emit<DEFER>(L, sg, bv, jj, L.plen, cid);
    if (jj == L.j_stop && L.slot == NO_SLOT) {   // whole read: the reference's last LF step has no observable effect
        L.state = LANE_IDLE;
        return;
    }
    if (!match) {
        bool found = false, use_pred = false;
        uint32_t tgt = 0;
        if (PACKED || code < CODE_OTHER) {
            const uint64_t meta = row_meta(row);
            const uint32_t slot = (code - chc - 1u) & 3u;
            const uint32_t mode = (uint32_t)(meta >> (11 + 2 * slot)) & 3u;
            const uint32_t d = (uint32_t)(meta >> (17 + 10 * slot)) & 1023u;
            if (mode == 3) {
                if (chc != CHC_OTHER && (d & 31u) && (d >> 5)) {
                    // both neighbours are known, only the in-row flip offset did not fit (a second threshold inside this
                    // row): evaluate `pos < thr[succ]` (col_bwt.hpp:560) directly -- two loads instead of a search
                    use_pred = ld_ro(t.idx + L.addr) + L.off < ld_ro(t.thr + L.addr + (d >> 5));
                    tgt = use_pred ? L.addr - (d & 31u) : L.addr + (d >> 5);
                    found = true;
                } else {
                    found = slow_reposition(t, L.addr, L.off, primary_byte((int)code), &tgt, &use_pred);
                }
            } else {
                const uint32_t fx = (uint32_t)(meta >> 47) & 0xFFFFu;
                use_pred = (mode == 1) || (mode == 2 && L.off < fx);
                tgt = use_pred ? L.addr - (d & 31u) : L.addr + (d >> 5);
                found = true;
            }
        } else if (code == CODE_OTHER) {
            found = slow_reposition(t, L.addr, L.off, cbyte, &tgt, &use_pred);
        }
        if (found) {
            if (INTRIP && same_line(tgt, L.addr)) {   // the target shares the gathered line: jump and LF through it in this trip
                const Row tr = ld_row(t.rows + tgt);
                L.off = row_doff(tr) + (use_pred ? row_len(tr) - 1 : 0);
                L.addr = tr.dest;
                return;
            }
            L.addr = tgt;
            L.state = use_pred ? LANE_REPOS_PRED : LANE_REPOS_SUCC;
            return;
        }
        // nothing found: state kept, LF proceeds through the current row (col_bwt.hpp:572-573)
    }
    L.off = row_doff(row) + L.off;               // LF_table.hpp:253-254
    L.addr = row.dest;
}


// Narrow-layout twin of lane_step.  `w` = hot[addr], or cold[addr] when the lane is in LANE_COLD.
template <bool PACKED, bool DEFER = false, typename PmlT>
CB_HD void lane_step_narrow(Lane<PmlT> &L, const Stage &sg, const TableView &t, const BatchView &bv, const uint64_t w, const uint8_t *code_lut)
{
    const uint32_t st = L.state & 7u;
    if (st == LANE_COLD) {
        const uint32_t code = (L.state >> 8) & 3u, rchc = (L.state >> 10) & 3u;
        const uint32_t slot = (code - rchc - 1u) & 3u;
        const uint32_t mode = (uint32_t)(w >> (2 * slot)) & 3u;
        bool found, use_pred = false;
        uint32_t tgt = 0;
        const uint32_t d = (uint32_t)(w >> (6 + 10 * slot)) & 1023u;
        if (mode == 3) {
            if ((d & 31u) && (d >> 5)) {   // second threshold inside this row: exact compare, see lane_step
                use_pred = ld_ro(t.idx + L.addr) + L.off < ld_ro(t.thr + L.addr + (d >> 5));
                tgt = use_pred ? L.addr - (d & 31u) : L.addr + (d >> 5);
                found = true;
            } else {
                found = slow_reposition(t, L.addr, L.off, primary_byte((int)code), &tgt, &use_pred);
            }
        } else {
            const uint32_t fx = (uint32_t)(w >> 36) & 0xFFFFu;
            use_pred = (mode == 1) || (mode == 2 && L.off < fx);
            tgt = use_pred ? L.addr - (d & 31u) : L.addr + (d >> 5);
            found = true;
        }
        if (found) {
            L.addr = tgt;
            L.state = use_pred ? LANE_REPOS_PRED : LANE_REPOS_SUCC;
        } else {
            L.state = LANE_RELOAD;
        }
        return;
    }
    const uint32_t dest = (uint32_t)w;
    const uint32_t doff = (uint32_t)(w >> 32) & 1023u;
    const uint32_t len = (uint32_t)(w >> 42) & 1023u;
    const uint32_t chc = (uint32_t)(w >> 52) & 7u;
    const uint32_t cid = (uint32_t)(w >> 55) & 255u;
    if (st != LANE_LF) {
        uint32_t o = L.off;                                   // LANE_RELOAD: keep the offset
        if (st == LANE_REPOS_PRED) o = len - 1;
        else if (st == LANE_REPOS_SUCC) o = 0;
        L.off = doff + o;
        L.addr = dest;
        L.state = LANE_LF;
        return;
    }
    if (L.off >= len && L.addr + 1 < t.r) {
        L.off -= len;
        ++L.addr;
        return;
    }
    if (L.slot != NO_SLOT) {
        if (L.j == L.emit_top) bv.start_state[L.slot] = ChainState{L.addr, L.off, L.plen, 0};
        if (L.j == L.j_stop) {
            bv.end_state[L.slot] = ChainState{L.addr, L.off, L.plen, 0};
            L.state = LANE_IDLE;
            return;
        }
    }
    const uint32_t jj = --L.j;
    uint32_t code;
    uint8_t cbyte = 0;
    if (PACKED) {
        code = (L.rw >> (2 * (jj & 15))) & 3u;
        if ((jj & 15) == 0 && jj != 0) L.rw = ld_ro(bv.words + L.in_off + ((jj - 1) >> 4));
    } else {
        cbyte = ld_ro(bv.bytes + L.in_off + jj);
        code = code_lut[cbyte];
    }
    bool match = (code == chc);
    if (!PACKED && code >= CODE_OTHER)
        match = (code == CODE_OTHER) && (chc == CHC_OTHER) && (ld_ro(t.ch8 + L.addr) == cbyte);
    L.plen = match ? L.plen + 1 : 0;
    if (jj < L.emit_top) lane_emit<DEFER>(L, sg, bv, jj, L.plen, cid);
    if (jj == L.j_stop && L.slot == NO_SLOT) {
        L.state = LANE_IDLE;
        return;
    }
    if (!match) {
        if ((PACKED || code < CODE_OTHER) && chc != CHC_OTHER) {
            L.state = LANE_COLD | (code << 8) | (chc << 10);   // next trip gathers cold[addr]
            return;
        }
        bool found = false, use_pred = false;
        uint32_t tgt = 0;
        if (PACKED || code < CODE_OTHER) found = slow_reposition(t, L.addr, L.off, primary_byte((int)code), &tgt, &use_pred);
        else if (code == CODE_OTHER) found = slow_reposition(t, L.addr, L.off, cbyte, &tgt, &use_pred);
        if (found) {
            L.addr = tgt;
            L.state = use_pred ? LANE_REPOS_PRED : LANE_REPOS_SUCC;
            return;
        }
    }
    L.off = doff + L.off;
    L.addr = dest;
}


// Run a lane to completion on its own (fixup re-traversal on the device, every lane in the host emulation).
template <bool PACKED, bool NARROW, typename PmlT>
CB_HD void lane_run(Lane<PmlT> &L, const Stage &sg, const TableView &t, const BatchView &bv, const uint8_t *code_lut)
{
    while (L.state != LANE_IDLE) {
        if (NARROW) {
            const uint64_t *base = ((L.state & 7u) == LANE_COLD) ? t.cold : t.hot;
            lane_step_narrow<PACKED>(L, sg, t, bv, ld_row64(base + L.addr), code_lut);
        } else {
            lane_step<PACKED>(L, sg, t, bv, ld_row(t.rows + L.addr), code_lut);
        }
    }
}

// Verify / repair the chunk chain of one split read, top chunk first (see ChunkTask).  Returns the number of chunks
// that had to be re-traversed.
template <bool PACKED, bool NARROW, typename PmlT>
CB_HD uint32_t fixup_chain(const Stage &sg, const TableView &t, const BatchView &bv, const ChainDesc &c, const uint8_t *code_lut)
{
    uint32_t redone = 0;
    PmlT *pml = reinterpret_cast<PmlT *>(bv.pml);
    for (uint32_t i = 1; i < c.n_chunks; ++i) {
        const uint32_t slot = c.first_slot + i;
        const ChainState truth = bv.end_state[slot - 1];          // exact by induction (chunk 0 starts from the true state)
        const ChainState spec = bv.start_state[slot];
        const ChunkTask k = bv.by_slot[slot];
        if (spec.row == truth.row && spec.off == truth.off) {
            // same position: everything below is identical except that the running match length may have started
            // before the warm-up window; shift it until the first mismatch of the chunk
            const uint32_t delta = truth.plen - spec.plen;
            if (delta != 0) {
                uint32_t j = k.hi;
                while (j > k.lo && pml[k.out_off + j - 1] != 0) {
                    pml[k.out_off + j - 1] = (PmlT)(pml[k.out_off + j - 1] + delta);
                    --j;
                }
                if (j == k.lo) bv.end_state[slot].plen += delta;  // no mismatch in the whole chunk: carry on
            }
        } else {
            Lane<PmlT> L;
            lane_begin_from_state<PACKED>(L, bv, k, truth);
            lane_run<PACKED, NARROW>(L, sg, t, bv, code_lut);
            ++redone;
        }
    }
    return redone;
}

} // namespace colbwt
