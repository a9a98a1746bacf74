// expand.cpp -- host side of the compact result form (compact.cu): rebuild the dense per-base arrays the reference
// returns (col_pml::query_pml, include/col_bwt.hpp:409-412) from one match bit per base + the sparse chain ids.
//
// Exactness: the reference's loop (col_bwt.hpp:510-528) sets length = 0 on a mismatch and ++length on a match, walking
// the read right to left, and stores the length AFTER that update.  So with nm(j) = the first mismatching base at or
// right of j (or the read length if there is none), PML[j] = nm(j) - j: a descending ramp ending in 0 at every
// mismatch, and ending in 1 at the last base of a read whose tail matches.  Chain ids are copied (non-zero) or 0.
#include <cstdint>
#include <cstring>

#include "internal.h"

namespace colbwt {

// First position p in [pos, e) whose match bit is 0; e if there is none.
static inline uint64_t next_mismatch(const uint32_t *match, uint64_t pos, uint64_t e)
{
    while (pos < e) {
        const uint32_t inv = ~match[pos >> 5] >> (pos & 31);
        if (inv) {
            const uint64_t p = pos + (uint64_t)__builtin_ctz(inv);
            return p < e ? p : e;
        }
        pos = (pos | 31) + 1;
    }
    return e;
}

template <typename T> static inline void ramp_down(T *dst, uint64_t count, uint64_t first)
{
    for (uint64_t t = 0; t < count; ++t) dst[t] = (T)(first - t);   // vectorised by the compiler
}

template <typename T>
static void expand_reads_t(const uint32_t *match, const uint32_t *cid_words, const uint32_t *prefix, const uint8_t *values, const uint64_t *off,
                           uint64_t r_first, uint64_t ra, uint64_t rb, T *pml, uint8_t *cid)
{
    if (ra >= rb) return;
    const uint64_t base0 = off[r_first];
    // chain-id cursor: values before the first base of read ra
    uint64_t s0 = off[ra] - base0;
    uint64_t w = s0 >> 5;
    const uint64_t g = w / COMPACT_GROUP_WORDS;
    uint64_t k = prefix[g];
    for (uint64_t x = g * COMPACT_GROUP_WORDS; x < w; ++x) k += (uint64_t)__builtin_popcount(cid_words[x]);
    if (s0 & 31) k += (uint64_t)__builtin_popcount(cid_words[w] & ((1u << (s0 & 31)) - 1u));
    for (uint64_t i = ra; i < rb; ++i) {
        const uint64_t s = off[i] - base0, e = off[i + 1] - base0;
        if (s == e) continue;
        // ---- PML ------------------------------------------------------------------------------------------
        for (uint64_t pos = s; pos < e;) {
            const uint64_t p = next_mismatch(match, pos, e);
            if (p < e) {
                ramp_down(pml + pos, p - pos + 1, p - pos);   // ... 2 1 0: the mismatch itself is 0
                pos = p + 1;
            } else {
                ramp_down(pml + pos, e - pos, e - pos);       // matching tail of the read: ... 2 1
                pos = e;
            }
        }
        // ---- chain ids ------------------------------------------------------------------------------------------
        memset(cid + s, 0, e - s);
        for (uint64_t pos = s; pos < e;) {
            uint32_t bits = cid_words[pos >> 5] >> (pos & 31);
            const uint64_t word_end = (pos | 31) + 1;
            if (word_end > e) bits &= (1u << (e - pos)) - 1u;   // e - pos < 32 here
            while (bits) {
                const uint32_t b = (uint32_t)__builtin_ctz(bits);
                bits &= bits - 1;
                cid[pos + b] = values[k++];
            }
            pos = word_end;
        }
    }
}

// Reads [ra, rb) of the segment that starts at read r_first; pml / cid point at the segment's base 0.
void expand_reads(const uint32_t *match, const uint32_t *cid_words, const uint32_t *prefix, const uint8_t *values, const uint64_t *off,
                  uint64_t r_first, uint64_t ra, uint64_t rb, void *pml, int pml_width, uint8_t *cid)
{
    if (pml_width == 1) expand_reads_t(match, cid_words, prefix, values, off, r_first, ra, rb, (uint8_t *)pml, cid);
    else if (pml_width == 2) expand_reads_t(match, cid_words, prefix, values, off, r_first, ra, rb, (uint16_t *)pml, cid);
    else expand_reads_t(match, cid_words, prefix, values, off, r_first, ra, rb, (uint32_t *)pml, cid);
}

} // namespace colbwt
