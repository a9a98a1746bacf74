// host_membw.cpp -- what the host cores of the box can read and write per second (development probe).
// The end-to-end query path either lets the GPU's copy engine write the dense results into host memory or has the
// packing threads rebuild them from the compact form (expand.cpp); which one wins is a question of how fast T threads
// write memory (with and without read-for-ownership) next to how fast they read the raw reads they pack.
//   g++ -O2 -march=native -pthread -o host_membw host_membw.cpp && ./host_membw [MiB per thread = 256]
// Prints one JSON line per (threads, operation): GB/s summed over the threads.
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
#include <immintrin.h>

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

static uint64_t op_read(const uint8_t *p, size_t n)
{
    const __m128i *v = (const __m128i *)p;
    __m128i acc = _mm_setzero_si128();
    for (size_t i = 0; i < n / 16; ++i) acc = _mm_xor_si128(acc, _mm_load_si128(v + i));
    return (uint64_t)_mm_cvtsi128_si64(acc);
}
static uint64_t op_write(uint8_t *p, size_t n)
{
    __m128i x = _mm_set1_epi8(7);
    for (size_t i = 0; i < n / 16; ++i) _mm_store_si128((__m128i *)p + i, x);
    return 0;
}
static uint64_t op_write_nt(uint8_t *p, size_t n)
{
    __m128i x = _mm_set1_epi8(9);
    for (size_t i = 0; i < n / 16; ++i) _mm_stream_si128((__m128i *)p + i, x);
    _mm_sfence();
    return 0;
}

int main(int argc, char **argv)
{
    const size_t mib = argc > 1 ? (size_t)atoi(argv[1]) : 256;
    const size_t n = mib << 20;
    const int hw = (int)std::thread::hardware_concurrency();
    std::vector<int> counts;
    for (int t = 1; t < hw; t *= 2) counts.push_back(t);
    counts.push_back(hw);
    std::vector<uint8_t *> bufs((size_t)hw);
    for (int t = 0; t < hw; ++t) {
        bufs[(size_t)t] = (uint8_t *)aligned_alloc(4096, n);
        memset(bufs[(size_t)t], 1, n);   // first touch
    }
    const char *names[3] = {"read", "write", "write_nt"};
    for (int T : counts)
        for (int op = 0; op < 3; ++op) {
            double best = 1e30;
            for (int rep = 0; rep < 3; ++rep) {
                std::vector<std::thread> th;
                std::vector<uint64_t> sink((size_t)T);
                const double t0 = now();
                for (int t = 0; t < T; ++t)
                    th.emplace_back([&, t] {
                        sink[(size_t)t] = op == 0 ? op_read(bufs[(size_t)t], n) : op == 1 ? op_write(bufs[(size_t)t], n) : op_write_nt(bufs[(size_t)t], n);
                    });
                for (auto &x : th) x.join();
                const double dt = now() - t0;
                if (dt < best) best = dt;
                if (sink[0] == 0x123456789abcdefull) printf("!");
            }
            printf("{\"threads\": %d, \"op\": \"%s\", \"GBps\": %.1f, \"MiB_per_thread\": %zu}\n", T, names[op], (double)T * n / best / 1e9, mib);
            fflush(stdout);
        }
    return 0;
}
