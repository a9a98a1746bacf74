// expand.cpp -- host side of the compact result form (compact.cu): rebuild the dense per-base arrays the reference
// returns (col_pml::query_pml, include/col_bwt.hpp:409-412) from one match bit per base + the sparse chain ids.
//
// Exactness: the reference's loop (col_bwt.hpp:510-528) sets length = 0 on a mismatch and ++length on a match, walking
// the read right to left, and stores the length AFTER that update.  So with nm(j) = the first mismatching base at or
// right of j (or the read length if there is none), PML[j] = nm(j) - j: a descending ramp ending in 0 at every
// mismatch, and ending in 1 at the last base of a read whose tail matches.  Chain ids are copied (non-zero) or 0.
#include <algorithm>
#include <cstdint>
#include <cstring>

#if defined(__x86_64__)
#include <emmintrin.h>
#endif

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

// 16 bits of a bit array starting at bit `pos` (bits past the array's last word read as 0).
static inline uint32_t bits16(const uint32_t *words, uint64_t n_words, uint64_t pos)
{
    const uint64_t w = pos >> 5;
    const uint32_t sh = (uint32_t)(pos & 31);
    uint64_t v = words[w];
    if (sh > 16 && w + 1 < n_words) v |= (uint64_t)words[w + 1] << 32;
    return (uint32_t)(v >> sh) & 0xFFFFu;
}

#if defined(__x86_64__)
// PML of 16 consecutive bases [pos, pos + cnt) of a read, cnt <= 16, given the PML `carry` of base pos + cnt (0 at the end
// of the read): data-independent (no branch per mismatch), which is what matters for noisy long reads where a mismatch
// comes every third base.  SSE2 only.  Lane i holds base pos + i:
//   nm(i) = index of the first mismatch at or right of i inside the block (suffix minimum, 4 shift+min steps), or none;
//   PML   = nm(i) - i            if there is one,
//           cnt - i + carry      otherwise (the match runs on into the bases already done).
template <typename T> static inline uint64_t pml_block16(T *dst, uint32_t match16, uint32_t cnt, uint64_t carry)
{
    const __m128i idx = _mm_setr_epi8(0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15);
    const __m128i bit = _mm_setr_epi8(1, 2, 4, 8, 16, 32, 64, (char)128, 1, 2, 4, 8, 16, 32, 64, (char)128);
    if (cnt < 16) match16 = (match16 | (0xFFFFu << cnt)) & 0xFFFFu;           // lanes past the block never stop a run
    const __m128i bytes = _mm_set_epi64x((long long)(0x0101010101010101ull * (match16 >> 8)), (long long)(0x0101010101010101ull * (match16 & 0xFF)));
    const __m128i is_match = _mm_cmpeq_epi8(_mm_and_si128(bytes, bit), bit);  // 0xFF where the base matched
    __m128i s = _mm_or_si128(is_match, idx);                                   // mismatch lanes: own index; match lanes: 0xFF
    s = _mm_min_epu8(s, _mm_or_si128(_mm_srli_si128(s, 1), _mm_set_epi8(-1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0)));
    s = _mm_min_epu8(s, _mm_or_si128(_mm_srli_si128(s, 2), _mm_set_epi8(-1, -1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0)));
    s = _mm_min_epu8(s, _mm_or_si128(_mm_srli_si128(s, 4), _mm_set_epi8(-1, -1, -1, -1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0)));
    s = _mm_min_epu8(s, _mm_or_si128(_mm_srli_si128(s, 8), _mm_set_epi8(-1, -1, -1, -1, -1, -1, -1, -1, 0, 0, 0, 0, 0, 0, 0, 0)));
    const __m128i none = _mm_cmpeq_epi8(s, _mm_set1_epi8(-1));                // no mismatch at or right of the lane
    const __m128i in_block = _mm_andnot_si128(none, _mm_sub_epi8(s, idx));    // nm - i where there is one (< 16), else 0
    const uint64_t tail = (uint64_t)cnt + carry;                              // PML of lane 0 if the whole block matches
    const int lane0_runs_on = _mm_cvtsi128_si32(none) & 1;
    const uint64_t next_carry = lane0_runs_on ? tail : (uint64_t)(_mm_cvtsi128_si32(in_block) & 0xFF);
    alignas(16) T out[16];
    if (sizeof(T) == 1) {
        // reads shorter than 256 bases: everything fits a byte (tail - i <= read length)
        const __m128i run = _mm_and_si128(none, _mm_sub_epi8(_mm_set1_epi8((char)tail), idx));
        _mm_store_si128(reinterpret_cast<__m128i *>(out), _mm_or_si128(in_block, run));
    } else if (sizeof(T) == 2) {
        const __m128i z = _mm_setzero_si128();
        const __m128i t16 = _mm_set1_epi16((short)tail);
        const __m128i lo_i = _mm_unpacklo_epi8(idx, z), hi_i = _mm_unpackhi_epi8(idx, z);
        const __m128i lo_n = _mm_unpacklo_epi8(none, none), hi_n = _mm_unpackhi_epi8(none, none);
        _mm_store_si128(reinterpret_cast<__m128i *>(out), _mm_or_si128(_mm_unpacklo_epi8(in_block, z), _mm_and_si128(lo_n, _mm_sub_epi16(t16, lo_i))));
        _mm_store_si128(reinterpret_cast<__m128i *>(out) + 1, _mm_or_si128(_mm_unpackhi_epi8(in_block, z), _mm_and_si128(hi_n, _mm_sub_epi16(t16, hi_i))));
    } else {
        const __m128i z = _mm_setzero_si128();
        const __m128i t32 = _mm_set1_epi32((int)tail);
        const __m128i i16[2] = {_mm_unpacklo_epi8(idx, z), _mm_unpackhi_epi8(idx, z)};
        const __m128i n16[2] = {_mm_unpacklo_epi8(none, none), _mm_unpackhi_epi8(none, none)};
        const __m128i b16[2] = {_mm_unpacklo_epi8(in_block, z), _mm_unpackhi_epi8(in_block, z)};
        for (int h = 0; h < 2; ++h) {
            const __m128i i_lo = _mm_unpacklo_epi16(i16[h], z), i_hi = _mm_unpackhi_epi16(i16[h], z);
            const __m128i n_lo = _mm_unpacklo_epi16(n16[h], n16[h]), n_hi = _mm_unpackhi_epi16(n16[h], n16[h]);
            _mm_store_si128(reinterpret_cast<__m128i *>(out) + 2 * h, _mm_or_si128(_mm_unpacklo_epi16(b16[h], z), _mm_and_si128(n_lo, _mm_sub_epi32(t32, i_lo))));
            _mm_store_si128(reinterpret_cast<__m128i *>(out) + 2 * h + 1, _mm_or_si128(_mm_unpackhi_epi16(b16[h], z), _mm_and_si128(n_hi, _mm_sub_epi32(t32, i_hi))));
        }
    }
    if (cnt == 16) memcpy(dst, out, 16 * sizeof(T));
    else memcpy(dst, out, cnt * sizeof(T));
    return next_carry;
}
#endif

template <typename T>
static void expand_reads_t(const uint32_t *match, const uint32_t *cid_words, const uint32_t *prefix, const uint8_t *values, const uint64_t *off,
                           uint64_t r_first, uint64_t ra, uint64_t rb, uint64_t seg_words, T *pml, uint8_t *cid)
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
        // ---- PML: right to left in blocks of 16 bases ------------------------------------------------------------
#if defined(__x86_64__)
        uint64_t carry = 0;
        for (uint64_t hi = e; hi > s;) {
            const uint32_t cnt = (uint32_t)(hi - s < 16 ? hi - s : 16);
            const uint64_t pos = hi - cnt;
            carry = pml_block16(pml + pos, bits16(match, seg_words, pos), cnt, carry);
            hi = pos;
        }
#else
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
#endif
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

// Copy a block that was just built in a small local buffer to its place in a large output array without reading the
// destination first: non-temporal stores for the 16-byte aligned body (measured on the GPU box's host, 16 cores:
// 201 GB/s against 90 GB/s for ordinary stores, which fetch every line before overwriting it -- profiles/r2/r2_host_membw.log).
static inline void stream_out(uint8_t *dst, const uint8_t *src, size_t n)
{
#if defined(__x86_64__)
    size_t head = (size_t)((16 - ((uintptr_t)dst & 15)) & 15);
    if (head > n) head = n;
    memcpy(dst, src, head);
    dst += head;
    src += head;
    n -= head;
    const size_t body = n & ~(size_t)15;
    for (size_t i = 0; i < body; i += 16)
        _mm_stream_si128(reinterpret_cast<__m128i *>(dst + i), _mm_loadu_si128(reinterpret_cast<const __m128i *>(src + i)));
    memcpy(dst + body, src + body, n - body);
#else
    memcpy(dst, src, n);
#endif
}

// Chain ids only (the transport that copies PML as it is): every group of 2048 bases is built in a local buffer -- zeros,
// then its non-zero ids dropped in from values[prefix[g] ...] -- and streamed out; no read boundary matters.
void expand_cid_groups(const uint32_t *cid_words, const uint32_t *prefix, const uint8_t *values, uint64_t n_bases, uint64_t g0, uint64_t g1, uint8_t *cid)
{
    const uint64_t n_words = (n_bases + 31) / 32;
    alignas(64) uint8_t buf[COMPACT_GROUP_WORDS * 32];
    for (uint64_t g = g0; g < g1; ++g) {
        const uint64_t w0 = g * COMPACT_GROUP_WORDS, w1 = std::min<uint64_t>(n_words, w0 + COMPACT_GROUP_WORDS);
        const uint64_t b0 = w0 * 32, b1 = std::min<uint64_t>(n_bases, w1 * 32);
        memset(buf, 0, sizeof(buf));
        uint64_t k = prefix[g];
        for (uint64_t w = w0; w < w1; ++w) {
            uint32_t bits = cid_words[w];
            while (bits) {
                const uint32_t b = (uint32_t)__builtin_ctz(bits);
                bits &= bits - 1;
                buf[(w - w0) * 32 + b] = values[k++];
            }
        }
        stream_out(cid + b0, buf, (size_t)(b1 - b0));
    }
#if defined(__x86_64__)
    _mm_sfence();   // the streamed lines are visible before the caller is told the chunk is complete
#endif
}

// Reads [ra, rb) of the segment that starts at read r_first; pml / cid point at the segment's base 0.
void expand_reads(const uint32_t *match, const uint32_t *cid_words, const uint32_t *prefix, const uint8_t *values, const uint64_t *off,
                  uint64_t r_first, uint64_t ra, uint64_t rb, void *pml, int pml_width, uint8_t *cid)
{
    const uint64_t seg_words = (off[rb] - off[r_first] + 31) / 32 + (off[rb] < off[r_first] ? 0 : 0);
    // words of the whole segment are at least those up to read rb; bits16 only needs a bound that keeps it inside the array
    if (pml_width == 1) expand_reads_t(match, cid_words, prefix, values, off, r_first, ra, rb, seg_words, (uint8_t *)pml, cid);
    else if (pml_width == 2) expand_reads_t(match, cid_words, prefix, values, off, r_first, ra, rb, seg_words, (uint16_t *)pml, cid);
    else expand_reads_t(match, cid_words, prefix, values, off, r_first, ra, rb, seg_words, (uint32_t *)pml, cid);
}

} // namespace colbwt
