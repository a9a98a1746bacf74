// pack.cpp -- host-side 2-bit packing of reads (the input format the traversal kernel streams from HBM).
//
// Replaces the byte-at-a-time access `pattern[m - i - 1]` of include/col_bwt.hpp:512: a read made only of
// A/C/G/T is stored 2 bit/base, code = (byte >> 1) & 3 (A=0 C=1 T=2 G=3), 16 bases per little-endian 32-bit
// word, base j at bits 2*(j & 15) of word j >> 4.  Any other byte (N, lower case, IUPAC, ...) makes the read
// "irregular": it is shipped as raw bytes and traversed by the byte kernel, so comparisons stay byte-exact.
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "colbwt_core.cuh"

#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace colbwt {

static inline bool is_acgt(uint8_t c) { return c == 'A' || c == 'C' || c == 'G' || c == 'T'; }

#if defined(__x86_64__)
// 32 bases -> 64 bits.  Returns false if any byte is outside ACGT.
__attribute__((target("avx2"))) static inline bool pack32_avx2(const uint8_t *p, uint64_t *out)
{
    const __m256i x = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(p));
    const __m256i code = _mm256_and_si256(_mm256_srli_epi16(x, 1), _mm256_set1_epi8(3));
    // expected byte for each code: 0->A 1->C 2->T 3->G
    const __m256i lut = _mm256_setr_epi8('A', 'C', 'T', 'G', 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0,
                                         'A', 'C', 'T', 'G', 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0);
    const __m256i expect = _mm256_shuffle_epi8(lut, code);
    const bool ok = _mm256_movemask_epi8(_mm256_cmpeq_epi8(expect, x)) == -1;
    // pairs of bytes -> 4 bit, pairs of those -> 8 bit
    const __m256i n4 = _mm256_maddubs_epi16(code, _mm256_set1_epi16(0x0401));          // b0 + 4*b1 in each 16-bit lane
    const __m256i n8 = _mm256_madd_epi16(n4, _mm256_set1_epi32(0x00100001));           // lo + 16*hi in each 32-bit lane
    // gather the low byte of the 8 dwords: within each 128-bit half bytes 0,4,8,12
    const __m256i sh = _mm256_setr_epi8(0, 4, 8, 12, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1,
                                        0, 4, 8, 12, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1);
    const __m256i g = _mm256_shuffle_epi8(n8, sh);
    const uint32_t lo = (uint32_t)_mm256_extract_epi32(g, 0), hi = (uint32_t)_mm256_extract_epi32(g, 4);
    *out = (uint64_t)lo | ((uint64_t)hi << 32);
    return ok;
}

__attribute__((target("avx2"))) static bool pack_avx2(const uint8_t *seq, uint64_t len, uint32_t *words)
{
    bool ok = true;
    uint64_t j = 0;
    for (; j + 32 <= len; j += 32) {
        uint64_t w;
        ok &= pack32_avx2(seq + j, &w);
        memcpy(words + (j >> 4), &w, 8);
    }
    if (j < len) {
        uint8_t tail[32];
        memset(tail, 'A', sizeof(tail));
        memcpy(tail, seq + j, len - j);
        uint64_t w;
        ok &= pack32_avx2(tail, &w);
        const uint64_t nwords = (len - j + 15) >> 4;
        memcpy(words + (j >> 4), &w, nwords * 4);
    }
    return ok;
}
#endif

static bool pack_scalar(const uint8_t *seq, uint64_t len, uint32_t *words)
{
    bool ok = true;
    for (uint64_t w = 0; w * 16 < len; ++w) {
        uint32_t v = 0;
        const uint64_t e = (len - w * 16 < 16) ? len - w * 16 : 16;
        for (uint64_t k = 0; k < e; ++k) {
            const uint8_t c = seq[w * 16 + k];
            ok &= is_acgt(c);
            v |= (uint32_t)((c >> 1) & 3) << (2 * k);
        }
        words[w] = v;
    }
    return ok;
}

#if defined(__x86_64__)
// Slice packer: reads [r0, r1), words written from word index `w` on, metas at meta[i - r_base].  Reads whose bytes
// are followed by >= 32 more readable bytes use whole 32-byte loads (the bytes past the read are masked out of the
// validity test and only the read's own words are stored).  Irregular reads are appended to `irr`.
__attribute__((target("avx2"))) static void pack_slice_avx2(const uint8_t *seqs, const uint64_t *off, uint64_t r0, uint64_t r1,
                                                          uint64_t r_base, uint64_t base0, uint64_t seq_end, uint32_t *words,
                                                          uint64_t w, ReadMeta *meta, std::vector<uint64_t> &irr)
{
    for (uint64_t i = r0; i < r1; ++i) {
        const uint64_t beg = off[i], len = off[i + 1] - beg;
        ReadMeta m{beg - base0, (uint32_t)len, (uint32_t)w};
        bool ok = true;
        if (len) {
            const uint8_t *p = seqs + beg;
            uint32_t *out = words + w;
            if (beg + ((len + 31) & ~31ull) <= seq_end) {
                uint64_t j = 0;
                for (; j + 32 <= len; j += 32) {
                    uint64_t v;
                    ok &= pack32_avx2(p + j, &v);
                    memcpy(out + (j >> 4), &v, 8);
                }
                if (j < len) {   // tail: over-read, judge only the read's own bytes
                    const __m256i x = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(p + j));
                    const __m256i code = _mm256_and_si256(_mm256_srli_epi16(x, 1), _mm256_set1_epi8(3));
                    const __m256i lut = _mm256_setr_epi8('A', 'C', 'T', 'G', 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0,
                                                         'A', 'C', 'T', 'G', 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0);
                    const uint32_t good = (uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(_mm256_shuffle_epi8(lut, code), x));
                    const uint32_t need = (uint32_t)((1ull << (len - j)) - 1);
                    ok &= (good & need) == need;
                    const __m256i n4 = _mm256_maddubs_epi16(code, _mm256_set1_epi16(0x0401));
                    const __m256i n8 = _mm256_madd_epi16(n4, _mm256_set1_epi32(0x00100001));
                    const __m256i sh = _mm256_setr_epi8(0, 4, 8, 12, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1,
                                                        0, 4, 8, 12, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1);
                    const __m256i g = _mm256_shuffle_epi8(n8, sh);
                    out[j >> 4] = (uint32_t)_mm256_extract_epi32(g, 0);
                    if (len - j > 16) out[(j >> 4) + 1] = (uint32_t)_mm256_extract_epi32(g, 4);
                }
            } else {
                ok = pack_avx2(p, len, out);
            }
        }
        if (!ok) {
            m.len = 0;   // skipped by the packed kernel, shipped as bytes
            irr.push_back(i);
        }
        meta[i - r_base] = m;
        w += (len + 15) >> 4;
    }
}
#endif

#if defined(__x86_64__)
// AVX-512 (F + BW + VL) slice packer: 64 bases per step; the ragged end of a read is a masked load (nothing past the read is
// touched, masked-off lanes pack as code 0) and a masked store of the words it fills.
__attribute__((target("avx512f,avx512bw,avx512vl"))) static void pack_slice_avx512(const uint8_t *seqs, const uint64_t *off, uint64_t r0, uint64_t r1,
                                                                                  uint64_t r_base, uint64_t base0, uint32_t *words, uint64_t w,
                                                                                  ReadMeta *meta, std::vector<uint64_t> &irr)
{
    const __m512i lut = _mm512_broadcast_i32x4(_mm_setr_epi8('A', 'C', 'T', 'G', 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0));
    const __m512i three = _mm512_set1_epi8(3), m4 = _mm512_set1_epi16(0x0401), m16 = _mm512_set1_epi32(0x00100001);
    for (uint64_t i = r0; i < r1; ++i) {
        const uint64_t beg = off[i], len = off[i + 1] - beg;
        const uint8_t *p = seqs + beg;
        uint32_t *out = words + w;
        __mmask64 bad = 0;
        uint64_t j = 0;
        for (; j + 64 <= len; j += 64) {
            const __m512i x = _mm512_loadu_si512(p + j);
            const __m512i code = _mm512_and_si512(_mm512_srli_epi16(x, 1), three);
            bad |= _mm512_cmpneq_epi8_mask(_mm512_shuffle_epi8(lut, code), x);
            const __m512i n8 = _mm512_madd_epi16(_mm512_maddubs_epi16(code, m4), m16);   // 4 codes -> the low byte of every dword
            _mm_storeu_si128(reinterpret_cast<__m128i *>(out + (j >> 4)), _mm512_cvtepi32_epi8(n8));
        }
        if (j < len) {
            const uint64_t rem = len - j;
            const __mmask64 k = (1ull << rem) - 1;
            const __m512i x = _mm512_maskz_loadu_epi8(k, p + j);
            const __m512i code = _mm512_and_si512(_mm512_srli_epi16(x, 1), three);
            bad |= _mm512_cmpneq_epi8_mask(_mm512_shuffle_epi8(lut, code), x) & k;
            const __m512i n8 = _mm512_madd_epi16(_mm512_maddubs_epi16(code, m4), m16);
            _mm_mask_storeu_epi32(out + (j >> 4), (__mmask8)((1u << ((rem + 15) >> 4)) - 1), _mm512_cvtepi32_epi8(n8));
        }
        ReadMeta m{beg - base0, (uint32_t)len, (uint32_t)w};
        if (bad) {
            m.len = 0;   // skipped by the packed kernel, shipped as bytes
            irr.push_back(i);
        }
        meta[i - r_base] = m;
        w += (len + 15) >> 4;
    }
}
#endif

void pack_slice(const uint8_t *seqs, const uint64_t *off, uint64_t r0, uint64_t r1, uint64_t r_base, uint64_t base0,
                uint64_t seq_end, uint32_t *words, uint64_t w, ReadMeta *meta, std::vector<uint64_t> &irr)
{
#if defined(__x86_64__)
    static const bool have_avx512 = __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512vl");
    if (have_avx512 && !getenv("COLBWT_NO_AVX512")) {
        pack_slice_avx512(seqs, off, r0, r1, r_base, base0, words, w, meta, irr);
        return;
    }
    static const bool have_avx2 = __builtin_cpu_supports("avx2");
    if (have_avx2) {
        pack_slice_avx2(seqs, off, r0, r1, r_base, base0, seq_end, words, w, meta, irr);
        return;
    }
#endif
    for (uint64_t i = r0; i < r1; ++i) {
        const uint64_t len = off[i + 1] - off[i];
        ReadMeta m{off[i] - base0, (uint32_t)len, (uint32_t)w};
        if (len && !pack_scalar(seqs + off[i], len, words + w)) {
            m.len = 0;
            irr.push_back(i);
        }
        meta[i - r_base] = m;
        w += (len + 15) >> 4;
    }
}

bool pack_read_2bit(const uint8_t *seq, uint64_t len, uint32_t *words)
{
#if defined(__x86_64__)
    static const bool have_avx2 = __builtin_cpu_supports("avx2");
    if (have_avx2) return pack_avx2(seq, len, words);
#endif
    return pack_scalar(seq, len, words);
}

} // namespace colbwt
