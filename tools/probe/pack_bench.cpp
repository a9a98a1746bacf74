// pack_bench.cpp -- one-core rate of the host 2-bit packer (col_bwt_b200/csrc/pack.cpp) (development probe).  From the repo root:
//   g++ -O3 -std=c++17 -I/usr/local/cuda/include -o /tmp/pack_bench tools/probe/pack_bench.cpp col_bwt_b200/csrc/pack.cpp
//   /tmp/pack_bench <read length>      (COLBWT_NO_AVX512=1: the AVX2 path)
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cstdint>
#include "../../col_bwt_b200/csrc/colbwt_core.cuh"
namespace colbwt {
void pack_slice(const uint8_t *seqs, const uint64_t *off, uint64_t r0, uint64_t r1, uint64_t r_base, uint64_t base0,
                uint64_t seq_end, uint32_t *words, uint64_t w, ReadMeta *meta, std::vector<uint64_t> &irr);
}
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main(int argc, char **argv)
{
    const uint64_t L = argc > 1 ? atoll(argv[1]) : 150;
    const uint64_t n = 128000000 / L * L, nr = n / L;
    std::vector<uint8_t> seq(n + 64);
    for (uint64_t i = 0; i < n; ++i) seq[i] = "ACGT"[rand() & 3];
    std::vector<uint64_t> off(nr + 1);
    for (uint64_t i = 0; i <= nr; ++i) off[i] = i * L;
    std::vector<uint32_t> words(n / 16 + nr + 16);
    std::vector<colbwt::ReadMeta> meta(nr);
    std::vector<uint64_t> irr;
    for (int rep = 0; rep < 3; ++rep) {
        double t0 = now();
        colbwt::pack_slice(seq.data(), off.data(), 0, nr, 0, 0, n, words.data(), 0, meta.data(), irr);
        double t1 = now();
        printf("L=%lu: pack %.3f Gbases/s (1 thread), irr %zu\n", L, n / (t1 - t0) / 1e9, irr.size());
    }
    unsigned long s = 0; for (uint64_t i = 0; i < words.size(); i += 97) s += words[i];
    printf("checksum %lu\n", s);
}
