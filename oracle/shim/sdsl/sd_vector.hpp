/* Stand-in for sdsl-lite's sd_vector.hpp (un-vendored dependency of the reference).
 * TEST INFRASTRUCTURE ONLY. sd_vector<> is kept as a sorted position list (no Elias-Fano);
 * its serialize/load format is private to oracle/_ref (u64 n, u64 m, m x u64 positions) and is
 * NOT sdsl's on-disk layout -- it only round-trips between tools built from this shim. */
#ifndef ORACLE_SHIM_SDSL_SD_VECTOR_HPP
#define ORACLE_SHIM_SDSL_SD_VECTOR_HPP
#include "int_vector.hpp"

namespace sdsl {

class sd_vector_builder {
public:
    sd_vector_builder() : m_n(0) {}
    sd_vector_builder(uint64_t n, uint64_t m) : m_n(n) { m_pos.reserve(m); }
    void set(uint64_t i) { m_pos.push_back(i); }
    uint64_t m_n;
    std::vector<uint64_t> m_pos;
};

template <class A = void, class B = void, class C = void>
class sd_vector {
public:
    typedef uint64_t size_type;
    sd_vector() : m_n(0) {}
    sd_vector(const bit_vector &bv) : m_n(bv.size()) {
        const std::vector<uint64_t> &w = bv.words();
        for (size_t k = 0; k < w.size(); ++k) {
            uint64_t x = w[k];
            while (x) { m_pos.push_back(k * 64 + __builtin_ctzll(x)); x &= x - 1; }
        }
    }
    sd_vector(sd_vector_builder &b) : m_n(b.m_n), m_pos(b.m_pos) {}
    size_type size() const { return m_n; }
    bool operator[](size_type i) const { return std::binary_search(m_pos.begin(), m_pos.end(), i); }

    size_type serialize(std::ostream &out, structure_tree_node * = nullptr, std::string = "") const {
        uint64_t n = m_n, m = m_pos.size();
        out.write((const char *)&n, 8);
        out.write((const char *)&m, 8);
        out.write((const char *)m_pos.data(), m * 8);
        return 16 + m * 8;
    }
    void load(std::istream &in) {
        uint64_t n = 0, m = 0;
        in.read((char *)&n, 8);
        in.read((char *)&m, 8);
        m_n = n;
        m_pos.assign(m, 0);
        in.read((char *)m_pos.data(), m * 8);
    }

    class rank_1_type {
        const sd_vector *m_v = nullptr;
    public:
        rank_1_type() {}
        rank_1_type(const sd_vector *v) : m_v(v) {}
        uint64_t operator()(uint64_t i) const {
            return std::lower_bound(m_v->m_pos.begin(), m_v->m_pos.end(), i) - m_v->m_pos.begin();
        }
        uint64_t rank(uint64_t i) const { return (*this)(i); }
    };
    class select_1_type {
        const sd_vector *m_v = nullptr;
    public:
        select_1_type() {}
        select_1_type(const sd_vector *v) : m_v(v) {}
        uint64_t operator()(uint64_t i) const {
            return (i >= 1 && i <= m_v->m_pos.size()) ? m_v->m_pos[i - 1] : UINT64_MAX;
        }
        uint64_t select(uint64_t i) const { return (*this)(i); }
    };

    uint64_t m_n;
    std::vector<uint64_t> m_pos;
};

} // namespace sdsl
#endif
