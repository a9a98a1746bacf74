// expand_bench.cpp -- one-core rate of the host expander (col_bwt_b200/csrc/expand.cpp) on synthetic match / chain-id bits
// (development probe).  From the repo root:
//   g++ -O3 -std=c++17 -I/usr/local/cuda/include -o /tmp/expand_bench tools/probe/expand_bench.cpp col_bwt_b200/csrc/expand.cpp
//   /tmp/expand_bench <read length> <PML bytes 1|2|4> <match probability>     (COLBWT_NO_AVX512=1 / COLBWT_EXPAND_WINDOWED=1: other paths)
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cstdint>
namespace colbwt {
void expand_cid_groups(const uint32_t *cid_words, const uint32_t *prefix, const uint8_t *values, uint64_t n_bases, uint64_t g0, uint64_t g1, uint8_t *cid);
void expand_reads(const uint32_t *match, const uint32_t *cid_words, const uint32_t *prefix, const uint8_t *values, const uint64_t *off,
                  uint64_t r_first, uint64_t ra, uint64_t rb, void *pml, int pml_width, uint8_t *cid);
void set_error(const char *, ...) {}
}
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main(int argc, char **argv)
{
    const uint64_t L = argc > 1 ? atoll(argv[1]) : 150; const int width = argc > 2 ? atoi(argv[2]) : 1; const double pm = argc > 3 ? atof(argv[3]) : 0.845;
    const uint64_t n = 64000000 / L * L, nr = n / L, nw = (n + 31) / 32, ng = (nw + 63) / 64;
    std::vector<uint32_t> match(nw + 4), cw(nw + 4), prefix(ng + 2);
    std::vector<uint8_t> values; std::vector<uint64_t> off(nr + 1);
    for (uint64_t i = 0; i <= nr; ++i) off[i] = i * L;
    srand(1);
    uint64_t k = 0;
    for (uint64_t b = 0; b < n; ++b) {
        if (b % 2048 == 0) prefix[b / 2048] = (uint32_t)k;
        if ((rand() % 1000) < pm * 1000) match[b >> 5] |= 1u << (b & 31);
        if ((rand() % 100) < 7) { cw[b >> 5] |= 1u << (b & 31); values.push_back((uint8_t)(1 + rand() % 255)); ++k; }
    }
    prefix[ng] = (uint32_t)k;
    std::vector<uint8_t> pml(n * width + 64), cid(n + 64);
    for (int rep = 0; rep < 3; ++rep) {
        double t0 = now();
        colbwt::expand_cid_groups(cw.data(), prefix.data(), values.data(), n, 0, ng, cid.data());
        double t1 = now();
        colbwt::expand_reads(match.data(), cw.data(), prefix.data(), values.data(), off.data(), 0, 0, nr, pml.data(), width, cid.data());
        double t2 = now();
        printf("L=%lu width=%d: cid only %.3f Gbases/s, full %.3f Gbases/s (1 thread)\n", L, width, n / (t1 - t0) / 1e9, n / (t2 - t1) / 1e9);
    }
    unsigned long s = 0; for (uint64_t i = 0; i < n; i += 997) s += pml[i * width] + cid[i];
    printf("checksum %lu\n", s);
}
