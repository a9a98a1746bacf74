// expand.cpp -- host side of the compact result form (compact.cu): rebuild the dense per-base arrays the reference
// returns (col_pml::query_pml, include/col_bwt.hpp:409-412) from one match bit per base + the sparse chain ids.
//
// Exactness: the reference's loop (col_bwt.hpp:510-528) sets length = 0 on a mismatch and ++length on a match, walking
// the read right to left, and stores the length AFTER that update.  So with nm(j) = the first mismatching base at or
// right of j (or the read length if there is none), PML[j] = nm(j) - j: a descending ramp ending in 0 at every
// mismatch, and ending in 1 at the last base of a read whose tail matches.  Chain ids are copied (non-zero) or 0.
//
// Speed: the arrays are 2-5 bytes per base that the host cores must WRITE, so the expander is built like a copy engine:
// results are produced in a cache-resident window and leave it through non-temporal stores (no read-for-ownership of
// the destination: 201 GB/s against 90 GB/s for ordinary stores on the 16 cores of the GPU box,
// profiles/r2/r2_host_membw.log), and the PML ramp of 8 bases comes from one 8-byte table entry indexed by their match
// bits -- no branch per mismatch, which matters for noisy long reads where every third base is one.
#include <algorithm>
#include <cstdint>
#include <cstring>

#include <cstdlib>

#if defined(__x86_64__)
#include <immintrin.h>
#endif

#include "internal.h"

namespace colbwt {

// 8 bits of a bit array starting at bit `pos` (bits past the array's last word read as 0).
static inline uint32_t bits8(const uint32_t *words, uint64_t n_words, uint64_t pos)
{
    const uint64_t w = pos >> 5;
    const uint32_t sh = (uint32_t)(pos & 31);
    uint64_t v = words[w];
    if (sh > 24 && w + 1 < n_words) v |= (uint64_t)words[w + 1] << 32;
    return (uint32_t)(v >> sh) & 0xFFu;
}

// RAMP.v[m][i], m = match bits of 8 consecutive bases (bit i = base i matched), i = lane:
//   d          (0..7)  a mismatch exists at or right of lane i inside the group, d = its distance: the PML of the lane
//   0x80 | 8-i         none: the run continues into the bases right of the group, PML = (8 - i) + PML of the base after the group
struct RampTable {
    alignas(64) uint8_t v[256][8];
    RampTable()
    {
        for (int m = 0; m < 256; ++m) {
            int next = -1;
            for (int i = 7; i >= 0; --i) {
                if (!((m >> i) & 1)) next = i;
                v[m][i] = (uint8_t)(next >= 0 ? next - i : (0x80 | (8 - i)));
            }
        }
    }
};
static const RampTable RAMP;

typedef uint64_t __attribute__((aligned(1), may_alias)) unaligned_u64;

// PML of the 8 bases of one full group into out[0..8); `carry` = PML of the base right after the group (0 past the end of
// the read).  Returns the PML of lane 0 = the carry of the group to the left.
template <typename T> static inline uint64_t ramp8(T *out, uint32_t m, uint64_t carry)
{
    const uint8_t *tv = RAMP.v[m];
    if (sizeof(T) == 1) {
        // SWAR: no byte overflows because a read with 8-bit PML is shorter than 256 bases (carry + 8 <= 255)
        const uint64_t x = *reinterpret_cast<const uint64_t *>(tv);   // table rows are 8-byte aligned
        const uint64_t flag = ((x >> 7) & 0x0101010101010101ull) * 0xFFull;
        const uint64_t r = (x & 0x7F7F7F7F7F7F7F7Full) + (flag & (carry * 0x0101010101010101ull));
        *reinterpret_cast<unaligned_u64 *>(out) = r;   // (a memcpy into the local window compiles to a checked call per group)
    } else {
#if defined(__x86_64__)
        const __m128i z = _mm_setzero_si128();
        const __m128i v16 = _mm_unpacklo_epi8(_mm_loadl_epi64(reinterpret_cast<const __m128i *>(tv)), z);
        const __m128i c80 = _mm_set1_epi16(0x80);
        const __m128i flag = _mm_cmpeq_epi16(_mm_and_si128(v16, c80), c80);
        const __m128i val = _mm_and_si128(v16, _mm_set1_epi16(0x7F));
        if (sizeof(T) == 2) {
            _mm_storeu_si128(reinterpret_cast<__m128i *>(out), _mm_add_epi16(val, _mm_and_si128(flag, _mm_set1_epi16((short)carry))));
        } else {
            const __m128i c32 = _mm_set1_epi32((int)carry);
            _mm_storeu_si128(reinterpret_cast<__m128i *>(out), _mm_add_epi32(_mm_unpacklo_epi16(val, z), _mm_and_si128(_mm_unpacklo_epi16(flag, flag), c32)));
            _mm_storeu_si128(reinterpret_cast<__m128i *>(out) + 1, _mm_add_epi32(_mm_unpackhi_epi16(val, z), _mm_and_si128(_mm_unpackhi_epi16(flag, flag), c32)));
        }
#else
        for (int i = 0; i < 8; ++i) out[i] = (T)((tv[i] & 0x80) ? (tv[i] & 0x7F) + carry : tv[i]);
#endif
    }
    return (tv[0] & 0x80) ? (uint64_t)(tv[0] & 0x7F) + carry : (uint64_t)tv[0];
}

// Copy a block that was just built in a small local buffer to its place in a large output array without reading the
// destination first: non-temporal stores for the 16-byte aligned body.
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

static inline void stores_visible()
{
#if defined(__x86_64__)
    _mm_sfence();   // the streamed lines are visible before the caller is told the work is complete
#endif
}

// Largest p <= x such that dst + p starts a 64-byte line (elements are naturally aligned).
template <typename T> static inline uint64_t line_start(const T *dst, uint64_t x)
{
    return x - (uint64_t)(((uintptr_t)(dst + x) & 63) / sizeof(T));
}

// PML of reads [ra, rb): one descending sweep over their (contiguous) bases.  The sweep fills a window of WINDOW
// positions from the top down -- groups of 8 bases through ramp8, the ragged left end of a read and of the window bit by
// bit -- and streams each finished window out; the running length resets at every read end.
template <typename T>
static void expand_pml(const uint32_t *match, uint64_t n_words, const uint64_t *off, uint64_t base0, uint64_t ra, uint64_t rb, T *pml)
{
    constexpr uint64_t WINDOW = 8192;
    alignas(64) T buf[WINDOW + 8];
    const uint64_t S = off[ra] - base0, E = off[rb] - base0;
    if (S >= E) return;
    uint64_t i = rb - 1;                                   // current read, [s, e) in segment positions
    uint64_t s = off[i] - base0, p = E, carry = 0;        // bases [p, e) of it are done
    for (uint64_t hi = E; hi > S;) {
        uint64_t lo = std::max<uint64_t>(S, hi > WINDOW ? hi - WINDOW : 0);
        if (lo > S) {                                      // windows meet on 64-byte line boundaries of the destination
            const uint64_t aligned = line_start(pml, lo + 64 / sizeof(T) - 1);   // first line start at or above lo
            if (aligned < hi) lo = aligned;
        }
        while (p > lo) {
            if (p == s) {                                  // read done: the one before it ends here (empty reads fall through)
                --i;
                s = off[i] - base0;
                carry = 0;
                continue;
            }
            const uint64_t room = std::min<uint64_t>(p - s, p - lo);
            if (room >= 8) {
                const uint64_t q = p - 8;
                carry = ramp8(buf + (q - lo), bits8(match, n_words, q), carry);
                p = q;
            } else {                                       // ragged edge: bit by bit (col_bwt.hpp:516-523 as it stands)
                const uint64_t q = p - room;
                const uint32_t m = bits8(match, n_words, q);
                for (uint64_t t = room; t-- > 0;) {
                    carry = ((m >> t) & 1u) ? carry + 1 : 0;
                    buf[q - lo + t] = (T)carry;
                }
                p = q;
            }
        }
        stream_out(reinterpret_cast<uint8_t *>(pml + lo), reinterpret_cast<const uint8_t *>(buf), (size_t)(hi - lo) * sizeof(T));
        hi = lo;
    }
}

static inline uint64_t values_before(const uint32_t *cid_words, const uint32_t *prefix, uint64_t b);

// Chain ids of segment positions [b0, b1): windows of up to 2048 bytes are zeroed in a local buffer, the non-zero ids are
// dropped in from values[...] (cursor from the per-group prefix + popcounts), and the window is streamed out.
static void expand_cid_range(const uint32_t *cid_words, const uint32_t *prefix, const uint8_t *values, uint64_t b0, uint64_t b1, uint8_t *cid)
{
    if (b0 >= b1) return;
    constexpr uint64_t WINDOW = 2048;
    alignas(64) uint8_t buf[WINDOW];
    uint64_t k = values_before(cid_words, prefix, b0);
    for (uint64_t pos = b0; pos < b1;) {
        uint64_t end = std::min<uint64_t>(b1, pos + WINDOW);
        if (end < b1 && end - pos > 128) end = line_start(cid, end);   // windows meet on line boundaries
        memset(buf, 0, (size_t)(end - pos));
        for (uint64_t x = pos; x < end;) {
            uint32_t bits = cid_words[x >> 5] >> (x & 31);
            const uint64_t word_end = (x | 31) + 1;
            if (word_end > end) bits &= (1u << (end - x)) - 1u;         // end - x < 32 here
            while (bits) {
                const uint32_t b = (uint32_t)__builtin_ctz(bits);
                bits &= bits - 1;
                buf[x - pos + b] = values[k++];
            }
            x = word_end;
        }
        stream_out(cid + pos, buf, (size_t)(end - pos));
        pos = end;
    }
}

// Number of non-zero chain ids before segment position b (b <= n_bases): per-group prefix + popcounts.
static inline uint64_t values_before(const uint32_t *cid_words, const uint32_t *prefix, uint64_t b)
{
    const uint64_t w = b >> 5, g = w / COMPACT_GROUP_WORDS;
    uint64_t k = prefix[g];
    for (uint64_t x = g * COMPACT_GROUP_WORDS; x < w; ++x) k += (uint64_t)__builtin_popcount(cid_words[x]);
    if (b & 31) k += (uint64_t)__builtin_popcount(cid_words[w] & ((1u << (b & 31)) - 1u));
    return k;
}

#if defined(__x86_64__)
// ---------------------------------------------------------------------------------------------------------------------
// AVX-512 path (VBMI + VBMI2: Ice Lake / Sapphire Rapids / Zen 4 and later; picked at run time): 64 bases per step.
//   PML   V[i] = i where base i mismatched, i + 1 where it is the (matching) last base of a read, 0xFF otherwise; a
//         suffix minimum over the 64 byte lanes (six permute + min steps) gives every lane its nearest stop, PML = stop -
//         lane; lanes without a stop continue into the block on the right: (64 - lane) + PML of that block's lane 0.
//   CID   vpexpandb drops the next popcount(mask) values into the lanes whose chain-id bit is set, zeros elsewhere.
// Blocks are aligned to multiples of 64 segment positions (one 64-bit load per bit array); lanes outside the caller's
// range are computed and not stored.  Same windows and streamed stores as the portable path.
// ---------------------------------------------------------------------------------------------------------------------
#define CB_AVX512 __attribute__((target("avx512f,avx512bw,avx512vl,avx512dq,avx512vbmi,avx512vbmi2,bmi,bmi2,popcnt,lzcnt")))

static bool have_avx512()
{
    static const bool cpu = __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512vl") &&
                            __builtin_cpu_supports("avx512dq") && __builtin_cpu_supports("avx512vbmi") && __builtin_cpu_supports("avx512vbmi2");
    const char *e = getenv("COLBWT_NO_AVX512");   // read per call: tests run both paths
    return cpu && !(e && atoi(e) != 0);
}

static inline uint64_t bits64(const uint32_t *words, uint64_t n_words, uint64_t block_base)
{
    const uint64_t w = block_base >> 5;
    if (w >= n_words) return 0;
    return (uint64_t)words[w] | (w + 1 < n_words ? (uint64_t)words[w + 1] << 32 : 0);
}

CB_AVX512 static inline void stream_out512(uint8_t *dst, const uint8_t *src, size_t n)
{
    size_t head = (size_t)((64 - ((uintptr_t)dst & 63)) & 63);
    if (head > n) head = n;
    memcpy(dst, src, head);
    dst += head;
    src += head;
    n -= head;
    const size_t body = n & ~(size_t)63;
    for (size_t i = 0; i < body; i += 64) _mm512_stream_si512(reinterpret_cast<__m512i *>(dst + i), _mm512_loadu_si512(src + i));
    memcpy(dst + body, src + body, n - body);
}

struct Avx512Consts {
    alignas(64) uint8_t idx[64];
    alignas(64) uint8_t shifted[6][64];   // lane j of step t reads lane j + 2^t
    Avx512Consts()
    {
        for (int j = 0; j < 64; ++j) {
            idx[j] = (uint8_t)j;
            for (int t = 0; t < 6; ++t) shifted[t][j] = (uint8_t)((j + (1 << t)) & 63);
        }
    }
};
static const Avx512Consts K512;

// Positions [S, E) of the segment (S, E on read boundaries of reads [ra, rb)): PML into pml[S..E), chain ids into cid[S..E).
template <typename T>
CB_AVX512 static void expand_avx512(const uint32_t *match, const uint32_t *cid_words, const uint32_t *prefix, const uint8_t *values, uint64_t n_words,
                                    const uint64_t *off, uint64_t base0, uint64_t ra, uint64_t rb, T *pml, uint8_t *cid)
{
    constexpr uint64_t WINDOW = 4096;                     // positions per window, a multiple of 64
    alignas(64) T pbuf[WINDOW + 64];
    alignas(64) uint8_t cbuf[WINDOW + 64];
    uint64_t ends[WINDOW / 64 + 2];                       // bit = last base of a read, window-relative blocks
    const uint64_t S = off[ra] - base0, E = off[rb] - base0;
    if (S >= E) return;
    const __m512i idx = _mm512_load_si512(K512.idx), ff = _mm512_set1_epi8((char)0xFF);
    const __m512i idx1 = _mm512_add_epi8(idx, _mm512_set1_epi8(1));
    const __m512i rest = _mm512_sub_epi8(_mm512_set1_epi8(64), idx);               // 64 - lane
    __m512i perm[6];
    __mmask64 keep[6];
    for (int t = 0; t < 6; ++t) {
        perm[t] = _mm512_load_si512(K512.shifted[t]);
        keep[t] = ~0ull >> (1u << t);                                              // lanes whose partner j + 2^t exists
    }
    uint64_t k_end = values_before(cid_words, prefix, E);                          // values before the top of the next block
    uint64_t carry = 0;                                                            // PML of the base right above the current block
    uint64_t ie = rb;                                                              // reads >= ie have had their end bit handed out
    for (uint64_t hi = E; hi > S;) {
        // window [lo, hi): lo on a block boundary (or S)
        uint64_t lo = hi > WINDOW ? ((hi - 1) / 64 * 64 + 64 - WINDOW) : 0;
        if (lo < S) lo = S;
        const uint64_t b_top = (hi - 1) / 64 * 64, b_bot = lo / 64 * 64;
        const uint64_t n_blocks = (b_top - b_bot) / 64 + 1;
        for (uint64_t q = 0; q < n_blocks; ++q) ends[q] = 0;
        while (ie > ra) {                                  // read ends inside the window, descending
            const uint64_t e = off[ie] - base0;            // end (exclusive) of read ie - 1
            if (e == off[ie - 1] - base0) { --ie; continue; }   // empty read
            if (e - 1 < lo) break;
            ends[(e - 1 - b_bot) >> 6] |= 1ull << ((e - 1 - b_bot) & 63);
            --ie;
        }
        for (uint64_t B = b_top;; B -= 64) {
            const uint64_t a = B < lo ? lo - B : 0, z = std::min<uint64_t>(64, hi - B);   // lanes [a, z) are stored
            const __mmask64 valid = (z == 64 ? ~0ull : ((1ull << z) - 1)) & ~((1ull << a) - 1);
            // ---- PML ----------------------------------------------------------------------------------------------
            const __mmask64 mism = ~bits64(match, n_words, B), last = ends[(B - b_bot) >> 6];
            __m512i v = _mm512_mask_blend_epi8(last, ff, idx1);
            v = _mm512_mask_blend_epi8(mism, v, idx);
#pragma GCC unroll 6
            for (int t = 0; t < 6; ++t) v = _mm512_min_epu8(v, _mm512_mask_permutexvar_epi8(ff, keep[t], perm[t], v));
            const __mmask64 none = _mm512_cmpeq_epi8_mask(v, ff);
            const __m512i d = _mm512_sub_epi8(v, idx);                                 // stop - lane (garbage where none)
            const uint32_t lane0 = (uint32_t)_mm_cvtsi128_si32(_mm512_castsi512_si128(v)) & 0xFFu;
            T *pout = pbuf + (B + a - lo) - a;                                          // lane j of the block -> pout[j]
            if (sizeof(T) == 1) {
                const __m512i r = _mm512_mask_add_epi8(d, none, rest, _mm512_set1_epi8((char)carry));
                _mm512_mask_storeu_epi8(pout, valid, r);
            } else if (sizeof(T) == 2) {
                const __m512i c16 = _mm512_set1_epi16((short)carry);
#pragma GCC unroll 2
                for (int h = 0; h < 2; ++h) {
                    const __m512i d16 = _mm512_cvtepu8_epi16(_mm512_extracti64x4_epi64(d, h));
                    const __m512i r16 = _mm512_cvtepu8_epi16(_mm512_extracti64x4_epi64(rest, h));
                    const __m512i r = _mm512_mask_add_epi16(d16, (__mmask32)(none >> (32 * h)), r16, c16);
                    _mm512_mask_storeu_epi16(reinterpret_cast<uint16_t *>(pout) + 32 * h, (__mmask32)(valid >> (32 * h)), r);
                }
            } else {
                const __m512i c32 = _mm512_set1_epi32((int)carry);
#pragma GCC unroll 4
                for (int h = 0; h < 4; ++h) {
                    const __m512i d32 = _mm512_cvtepu8_epi32(_mm512_extracti32x4_epi32(d, h));
                    const __m512i r32 = _mm512_cvtepu8_epi32(_mm512_extracti32x4_epi32(rest, h));
                    const __m512i r = _mm512_mask_add_epi32(d32, (__mmask16)(none >> (16 * h)), r32, c32);
                    _mm512_mask_storeu_epi32(reinterpret_cast<uint32_t *>(pout) + 16 * h, (__mmask16)(valid >> (16 * h)), r);
                }
            }
            carry = lane0 == 0xFFu ? 64 + carry : lane0;
            // ---- chain ids ----------------------------------------------------------------------------------------
            const __mmask64 cm = bits64(cid_words, n_words, B) & valid;
            const uint64_t cnt = (uint64_t)__builtin_popcountll(cm);
            k_end -= cnt;
            const __m512i vals = _mm512_maskz_loadu_epi8(cnt == 64 ? ~0ull : ((1ull << cnt) - 1), values + k_end);
            _mm512_mask_storeu_epi8(cbuf + (B + a - lo) - a, valid, _mm512_maskz_expand_epi8(cm, vals));
            if (B == b_bot) break;
        }
        stream_out512(reinterpret_cast<uint8_t *>(pml + lo), reinterpret_cast<const uint8_t *>(pbuf), (size_t)(hi - lo) * sizeof(T));
        stream_out512(cid + lo, cbuf, (size_t)(hi - lo));
        hi = lo;
    }
}

// 64 bits of a bit array starting at bit `pos` (any alignment, pos may be negative: bits outside the array read as 0).
static inline uint64_t bits64_at(const uint32_t *words, uint64_t n_words, int64_t pos)
{
    if (pos < 0) return pos <= -64 ? 0 : bits64_at(words, n_words, 0) << (uint64_t)(-pos);
    const uint64_t w = (uint64_t)pos >> 5, sh = (uint64_t)pos & 31;
    if (w + 3 <= n_words) {                                // interior: three words, no bounds to check
        uint64_t lo;
        memcpy(&lo, words + w, 8);
        return sh ? (lo >> sh) | ((uint64_t)words[w + 2] << (64 - sh)) : lo;
    }
    auto word = [&](uint64_t i) -> uint64_t { return i < n_words ? words[i] : 0; };
    const uint64_t lo = word(w) | (word(w + 1) << 32);
    return sh ? (lo >> sh) | (word(w + 2) << (64 - sh)) : lo;
}

// Same computation with the blocks aligned to the DESTINATION (64 chain-id bytes = one line) instead of to the bit arrays:
// every full block leaves through one non-temporal 64-byte store per line, no window, no second copy.  Needs the PML
// array to be line-aligned at the same positions (true for page-aligned caller arrays, pinned ones in particular).
template <typename T>
CB_AVX512 static void expand_avx512_direct(const uint32_t *match, const uint32_t *cid_words, const uint32_t *prefix, const uint8_t *values, uint64_t n_words,
                                           const uint64_t *off, uint64_t base0, uint64_t ra, uint64_t rb, T *pml, uint8_t *cid)
{
    const int64_t S = (int64_t)(off[ra] - base0), E = (int64_t)(off[rb] - base0);
    if (S >= E) return;
    const int64_t c0 = (int64_t)((64 - ((uintptr_t)cid & 63)) & 63);             // positions p = c0 (mod 64) start a line of cid
    const __m512i idx = _mm512_load_si512(K512.idx), ff = _mm512_set1_epi8((char)0xFF);
    const __m512i idx1 = _mm512_add_epi8(idx, _mm512_set1_epi8(1));
    const __m512i rest = _mm512_sub_epi8(_mm512_set1_epi8(64), idx);
    __m512i perm[6];
    __mmask64 keep[6];
    for (int t = 0; t < 6; ++t) {
        perm[t] = _mm512_load_si512(K512.shifted[t]);
        keep[t] = ~0ull >> (1u << t);
    }
    __m512i rest_w[4];                                                             // (64 - lane) widened to T, per part of the block
    if (sizeof(T) == 2)
        for (int h = 0; h < 2; ++h) rest_w[h] = _mm512_cvtepu8_epi16(_mm512_extracti64x4_epi64(rest, h));
    else if (sizeof(T) == 4)
        for (int h = 0; h < 4; ++h) rest_w[h] = _mm512_cvtepu8_epi32(_mm512_extracti32x4_epi32(rest, h));
    uint64_t k_end = values_before(cid_words, prefix, (uint64_t)E);
    uint64_t carry = 0;
    uint64_t ie = rb;
    // blocks [B, B + 64) with B = c0 (mod 64), from the one holding E - 1 down to the one holding S (B may be negative)
    auto floor_block = [&](int64_t x) { const int64_t r = ((x - c0) % 64 + 64) % 64; return x - r; };
    const int64_t b_top = floor_block(E - 1), b_bot = floor_block(S);
    // Read-end bits are laid out for a span of 64 blocks at a time (one short loop over the reads that end in the span: a
    // per-block search for them costs a fifth of the whole expansion in mispredicted branches); then two blocks per trip:
    // their suffix-minimum scans are independent dependency chains of ~25 cycles each.
    constexpr int64_t SPAN_BLOCKS = 64;
    uint64_t ends[SPAN_BLOCKS];
    int64_t span_bot = b_top + 64;                           // blocks >= span_bot have their end bits consumed
    for (int64_t Bp = b_top; Bp >= b_bot; Bp -= 128) {
        const int nb = Bp - 64 >= b_bot ? 2 : 1;
        if (Bp < span_bot) {                                 // next span: blocks [span_bot, span_bot + 64 * SPAN_BLOCKS) below the old one
            span_bot = std::max<int64_t>(b_bot, Bp - 64 * (SPAN_BLOCKS - 1));
            for (int q = 0; q < SPAN_BLOCKS; ++q) ends[q] = 0;
            while (ie > ra) {                                // reads ending inside the span, descending; empty reads add nothing
                const int64_t e1 = (int64_t)(off[ie] - base0) - 1;
                if (e1 < span_bot) break;
                if (off[ie] != off[ie - 1]) ends[(e1 - span_bot) >> 6] |= 1ull << (uint64_t)((e1 - span_bot) & 63);
                --ie;
            }
        }
        __m512i v[2] = {ff, ff};
#pragma GCC unroll 2
        for (int u = 0; u < 2; ++u) {
            if (u >= nb) break;
            const int64_t B = Bp - 64 * u;
            const uint64_t last = ends[(B - span_bot) >> 6];
            const __mmask64 mism = ~bits64_at(match, n_words, B);
            v[u] = _mm512_mask_blend_epi8(mism, _mm512_mask_blend_epi8(last, ff, idx1), idx);
        }
#pragma GCC unroll 6
        for (int t = 0; t < 6; ++t) {
            v[0] = _mm512_min_epu8(v[0], _mm512_mask_permutexvar_epi8(ff, keep[t], perm[t], v[0]));
            v[1] = _mm512_min_epu8(v[1], _mm512_mask_permutexvar_epi8(ff, keep[t], perm[t], v[1]));
        }
#pragma GCC unroll 2
        for (int u = 0; u < 2; ++u) {
            if (u >= nb) break;
            const int64_t B = Bp - 64 * u;
            const uint64_t a = B < S ? (uint64_t)(S - B) : 0, z = (uint64_t)std::min<int64_t>(64, E - B);   // lanes [a, z) are stored
            const __mmask64 valid = (z == 64 ? ~0ull : ((1ull << z) - 1)) & ~((1ull << a) - 1);
            const __mmask64 none = _mm512_cmpeq_epi8_mask(v[u], ff);
            const __m512i d = _mm512_sub_epi8(v[u], idx);
            const uint32_t lane0 = (uint32_t)_mm_cvtsi128_si32(_mm512_castsi512_si128(v[u])) & 0xFFu;
            T *pout = pml + B;                                                      // lane j -> pout[j]; never dereferenced outside [a, z)
            const bool full = valid == ~0ull;
            if (sizeof(T) == 1) {
                const __m512i r = _mm512_mask_add_epi8(d, none, rest, _mm512_set1_epi8((char)carry));
                if (full) _mm512_stream_si512(reinterpret_cast<__m512i *>(pout), r);
                else _mm512_mask_storeu_epi8(pout, valid, r);
            } else if (sizeof(T) == 2) {
                const __m512i c16 = _mm512_set1_epi16((short)carry);
#pragma GCC unroll 2
                for (int h = 0; h < 2; ++h) {
                    const __m512i r = _mm512_mask_add_epi16(_mm512_cvtepu8_epi16(_mm512_extracti64x4_epi64(d, h)), (__mmask32)(none >> (32 * h)), rest_w[h], c16);
                    if (full) _mm512_stream_si512(reinterpret_cast<__m512i *>(reinterpret_cast<uint16_t *>(pout) + 32 * h), r);
                    else _mm512_mask_storeu_epi16(reinterpret_cast<uint16_t *>(pout) + 32 * h, (__mmask32)(valid >> (32 * h)), r);
                }
            } else {
                const __m512i c32 = _mm512_set1_epi32((int)carry);
#pragma GCC unroll 4
                for (int h = 0; h < 4; ++h) {
                    const __m512i r = _mm512_mask_add_epi32(_mm512_cvtepu8_epi32(_mm512_extracti32x4_epi32(d, h)), (__mmask16)(none >> (16 * h)), rest_w[h], c32);
                    if (full) _mm512_stream_si512(reinterpret_cast<__m512i *>(reinterpret_cast<uint32_t *>(pout) + 16 * h), r);
                    else _mm512_mask_storeu_epi32(reinterpret_cast<uint32_t *>(pout) + 16 * h, (__mmask16)(valid >> (16 * h)), r);
                }
            }
            carry = lane0 == 0xFFu ? 64 + carry : lane0;
            const __mmask64 cm = bits64_at(cid_words, n_words, B) & valid;
            const uint64_t cnt = (uint64_t)__builtin_popcountll(cm);
            k_end -= cnt;
            const __m512i vals = _mm512_maskz_loadu_epi8(cnt == 64 ? ~0ull : ((1ull << cnt) - 1), values + k_end);
            const __m512i cr = _mm512_maskz_expand_epi8(cm, vals);
            if (full) _mm512_stream_si512(reinterpret_cast<__m512i *>(cid + B), cr);
            else _mm512_mask_storeu_epi8(cid + B, valid, cr);
        }
    }
}

// Chain ids of positions [b0, b1) alone.
CB_AVX512 static void expand_cid_avx512(const uint32_t *cid_words, const uint32_t *prefix, const uint8_t *values, uint64_t n_words, uint64_t b0, uint64_t b1,
                                        uint8_t *cid)
{
    constexpr uint64_t WINDOW = 8192;
    alignas(64) uint8_t cbuf[WINDOW + 64];
    if (b0 >= b1) return;
    uint64_t k = values_before(cid_words, prefix, b0);
    for (uint64_t lo = b0; lo < b1;) {
        const uint64_t hi = std::min<uint64_t>(b1, lo / 64 * 64 + WINDOW);
        for (uint64_t B = lo / 64 * 64; B < hi; B += 64) {
            const uint64_t a = B < lo ? lo - B : 0, z = std::min<uint64_t>(64, hi - B);
            const __mmask64 valid = (z == 64 ? ~0ull : ((1ull << z) - 1)) & ~((1ull << a) - 1);
            const __mmask64 cm = bits64(cid_words, n_words, B) & valid;
            const uint64_t cnt = (uint64_t)__builtin_popcountll(cm);
            const __m512i vals = _mm512_maskz_loadu_epi8(cnt == 64 ? ~0ull : ((1ull << cnt) - 1), values + k);
            k += cnt;
            _mm512_mask_storeu_epi8(cbuf + (B + a - lo) - a, valid, _mm512_maskz_expand_epi8(cm, vals));
        }
        stream_out512(cid + lo, cbuf, (size_t)(hi - lo));
        lo = hi;
    }
}
#endif

// Chain ids only (the transport that copies PML as it is), prefix groups [g0, g1) of a chunk of n_bases bases.
void expand_cid_groups(const uint32_t *cid_words, const uint32_t *prefix, const uint8_t *values, uint64_t n_bases, uint64_t g0, uint64_t g1, uint8_t *cid)
{
    const uint64_t group_bases = (uint64_t)COMPACT_GROUP_WORDS * 32;
    const uint64_t b0 = std::min(n_bases, g0 * group_bases), b1 = std::min(n_bases, g1 * group_bases);
#if defined(__x86_64__)
    if (have_avx512()) {
        expand_cid_avx512(cid_words, prefix, values, (n_bases + 31) / 32, b0, b1, cid);
        stores_visible();
        return;
    }
#endif
    expand_cid_range(cid_words, prefix, values, b0, b1, cid);
    stores_visible();
}

// Reads [ra, rb) of the segment that starts at read r_first; pml / cid point at the segment's base 0.
void expand_reads(const uint32_t *match, const uint32_t *cid_words, const uint32_t *prefix, const uint8_t *values, const uint64_t *off,
                  uint64_t r_first, uint64_t ra, uint64_t rb, void *pml, int pml_width, uint8_t *cid)
{
    if (ra >= rb) return;
    const uint64_t base0 = off[r_first];
    // bits8 only needs a bound that keeps it inside the match array: the words up to read rb
    const uint64_t n_words = (off[rb] - base0 + 31) / 32;
#if defined(__x86_64__)
    if (have_avx512()) {
        // lines of the chain-id array and of the PML array start at the same positions: blocks aligned to them, one
        // streamed store per line; otherwise the windowed variant
        const uint64_t c0 = (64 - ((uintptr_t)cid & 63)) & 63;
        const bool direct = (((uintptr_t)pml + c0 * (uint64_t)pml_width) & 63) == 0 && !getenv("COLBWT_EXPAND_WINDOWED");
        if (direct) {
            if (pml_width == 1) expand_avx512_direct(match, cid_words, prefix, values, n_words, off, base0, ra, rb, (uint8_t *)pml, cid);
            else if (pml_width == 2) expand_avx512_direct(match, cid_words, prefix, values, n_words, off, base0, ra, rb, (uint16_t *)pml, cid);
            else expand_avx512_direct(match, cid_words, prefix, values, n_words, off, base0, ra, rb, (uint32_t *)pml, cid);
        } else {
            if (pml_width == 1) expand_avx512(match, cid_words, prefix, values, n_words, off, base0, ra, rb, (uint8_t *)pml, cid);
            else if (pml_width == 2) expand_avx512(match, cid_words, prefix, values, n_words, off, base0, ra, rb, (uint16_t *)pml, cid);
            else expand_avx512(match, cid_words, prefix, values, n_words, off, base0, ra, rb, (uint32_t *)pml, cid);
        }
        stores_visible();
        return;
    }
#endif
    if (pml_width == 1) expand_pml(match, n_words, off, base0, ra, rb, (uint8_t *)pml);
    else if (pml_width == 2) expand_pml(match, n_words, off, base0, ra, rb, (uint16_t *)pml);
    else expand_pml(match, n_words, off, base0, ra, rb, (uint32_t *)pml);
    expand_cid_range(cid_words, prefix, values, off[ra] - base0, off[rb] - base0, cid);
    stores_visible();
}

} // namespace colbwt
