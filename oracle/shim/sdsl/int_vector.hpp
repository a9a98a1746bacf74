/* Stand-in for sdsl-lite's int_vector.hpp (un-vendored dependency of the reference).
 * TEST INFRASTRUCTURE ONLY: lets the reference's own sources compile from /root/reference
 * into oracle/_ref/. Provides just the surface the reference touches:
 *   sdsl::bit_vector (ctor(n,v), operator[], size, serialize, load, rank_1_type, select_1_type)
 * bit_vector::serialize follows sdsl's published layout: u64 bit count, then ceil(n/64) u64 words,
 * LSB first -- this is the `.col_runs` format written by col_split.hpp:384-386. */
#ifndef ORACLE_SHIM_SDSL_INT_VECTOR_HPP
#define ORACLE_SHIM_SDSL_INT_VECTOR_HPP
#include <algorithm>
#include <array>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <iostream>
#include <iterator>
#include <optional>
#include <queue>
#include <sstream>
#include <string>
#include <tuple>
#include <vector>

namespace sdsl {

class structure_tree_node {};

class nullstream : public std::ostream {
    struct nullbuf : public std::streambuf {
        int overflow(int c) override { return c; }
        std::streamsize xsputn(const char *, std::streamsize n) override { return n; }
    } m_buf;
public:
    nullstream() : std::ostream(&m_buf) {}
};

class bit_vector {
public:
    typedef uint64_t size_type;
    class reference {
        uint64_t *w; uint64_t mask;
    public:
        reference(uint64_t *w_, unsigned b) : w(w_), mask(uint64_t(1) << b) {}
        operator bool() const { return (*w & mask) != 0; }
        reference &operator=(bool v) { if (v) *w |= mask; else *w &= ~mask; return *this; }
        reference &operator=(const reference &o) { return *this = bool(o); }
    };

    bit_vector() : m_size(0) {}
    bit_vector(size_type n, bool v = false) : m_size(n), m_words((n + 63) / 64 + 1, v ? ~uint64_t(0) : 0) {
        if (v && (n & 63)) m_words[n / 64] &= (uint64_t(1) << (n & 63)) - 1;
        if (v) m_words.back() = 0;
    }
    size_type size() const { return m_size; }
    bool operator[](size_type i) const { return (m_words[i >> 6] >> (i & 63)) & 1; }
    reference operator[](size_type i) { return reference(&m_words[i >> 6], unsigned(i & 63)); }

    size_type serialize(std::ostream &out, structure_tree_node * = nullptr, std::string = "") const {
        uint64_t n = m_size;
        out.write((const char *)&n, 8);
        size_type nw = (m_size + 63) / 64;
        out.write((const char *)m_words.data(), nw * 8);
        return 8 + nw * 8;
    }
    void load(std::istream &in) {
        uint64_t n = 0;
        in.read((char *)&n, 8);
        m_size = n;
        m_words.assign((n + 63) / 64 + 1, 0);
        in.read((char *)m_words.data(), ((n + 63) / 64) * 8);
    }
    const std::vector<uint64_t> &words() const { return m_words; }

    /* rank(i) = number of ones in [0, i) */
    class rank_1_type {
        const bit_vector *m_bv = nullptr;
        std::vector<uint64_t> m_cum;
    public:
        rank_1_type() {}
        rank_1_type(const bit_vector *bv) : m_bv(bv) {
            size_t nw = bv->m_words.size();
            m_cum.resize(nw + 1, 0);
            for (size_t w = 0; w < nw; ++w) m_cum[w + 1] = m_cum[w] + __builtin_popcountll(bv->m_words[w]);
        }
        uint64_t operator()(uint64_t i) const {
            uint64_t w = i >> 6, b = i & 63;
            uint64_t r = m_cum[w];
            if (b) r += __builtin_popcountll(m_bv->m_words[w] & ((uint64_t(1) << b) - 1));
            return r;
        }
        uint64_t rank(uint64_t i) const { return (*this)(i); }
    };

    /* select(i) = position of the i-th one, i is 1-based */
    class select_1_type {
        std::vector<uint64_t> m_pos;
    public:
        select_1_type() {}
        select_1_type(const bit_vector *bv) {
            for (size_t w = 0; w < bv->m_words.size(); ++w) {
                uint64_t x = bv->m_words[w];
                while (x) { m_pos.push_back(w * 64 + __builtin_ctzll(x)); x &= x - 1; }
            }
        }
        uint64_t operator()(uint64_t i) const { return (i >= 1 && i <= m_pos.size()) ? m_pos[i - 1] : UINT64_MAX; }
        uint64_t select(uint64_t i) const { return (*this)(i); }
    };

private:
    size_type m_size;
    std::vector<uint64_t> m_words;
};

} // namespace sdsl
#endif
