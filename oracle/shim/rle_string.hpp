/* Stand-in for r-index's rle_string.hpp (un-vendored). include/ds/r_index.hpp is included by
 * col_split.cpp:25 and build_FL.cpp:23 but never instantiated, so only names must resolve. */
#ifndef ORACLE_SHIM_RLE_STRING_HPP
#define ORACLE_SHIM_RLE_STRING_HPP
#include <sdsl/int_vector.hpp>
#include <sdsl/sd_vector.hpp>
#include <string>
#include <utility>
using namespace sdsl;
namespace ri {
typedef std::pair<unsigned long, unsigned long> range_t;
class rle_string_sd {
public:
    rle_string_sd() {}
    template <class... T> rle_string_sd(T &...) {}
    unsigned long rank(unsigned long, unsigned char) { return 0; }
    unsigned long select(unsigned long, unsigned char) { return 0; }
    unsigned long size() { return 0; }
    unsigned long number_of_runs() { return 0; }
    unsigned char operator[](unsigned long) { return 0; }
    unsigned long run_of_position(unsigned long) { return 0; }
    range_t run_range(unsigned long) { return range_t(0, 0); }
    unsigned long serialize(std::ostream &) { return 0; }
    void load(std::istream &) {}
    template <class... T> unsigned long run_at(T...) { return 0; }
};
}
#endif
