/* TEST INFRASTRUCTURE (oracle/_ref/libref_harness.so). NOT part of the product path.
 * Exposes the reference's own header-only implementation col_pml::load (col_bwt.hpp:375-380)
 * and col_pml::query_pml(const char*, size_t) (col_bwt.hpp:409-412) through a tiny C ABI so
 * that tests/ and bench.py's cpu_baseline / --impl reference legs can call the UNMODIFIED
 * reference algorithm in-process. Compiled from /root/reference headers where they lie; the
 * only tweak is MULTI_THREAD off (same trick as pml_query_nomt.cpp), which SURVEY.md section 6
 * shows is output-identical. The query loop touches no globals, so one table can be shared by
 * several threads. */
#include <common.hpp>
#ifndef REF_HARNESS_KEEP_MULTI_THREAD
#undef MULTI_THREAD
#endif
#include <col_bwt.hpp>
#include <atomic>
#include <thread>

extern "C" {

void *ref_load(const char *col_pml_path)
{
    std::ifstream in(col_pml_path, std::ios::binary);
    if (!in) return nullptr;
    col_pml *t = new col_pml();
    t->load(in);
    return t;
}

void ref_free(void *h) { delete static_cast<col_pml *>(h); }

unsigned long ref_n(void *h) { return static_cast<col_pml *>(h)->size(); }
unsigned long ref_r(void *h) { return static_cast<col_pml *>(h)->runs(); }
unsigned long ref_bwt_r(void *h) { return static_cast<col_pml *>(h)->bwt_runs(); }

/* One read: out arrays hold m values each (u32 PML, u8 CID). */
void ref_query(void *h, const char *seq, unsigned long m, uint32_t *pml, uint8_t *cid)
{
    auto res = static_cast<col_pml *>(h)->query_pml(seq, m);
    for (unsigned long i = 0; i < m; ++i) { pml[i] = (uint32_t)res.first[i]; cid[i] = (uint8_t)res.second[i]; }
}

/* Batch of reads (concatenated bytes + n_reads+1 offsets), dealt to `threads` std::threads in
 * blocks of 64 reads via an atomic cursor. pml/cid may be null (timing only). Returns bases done. */
unsigned long ref_query_batch(void *h, const char *seqs, const uint64_t *off, uint64_t n_reads,
                              uint32_t *pml, uint8_t *cid, int threads)
{
    col_pml *t = static_cast<col_pml *>(h);
    std::atomic<uint64_t> cursor(0), bases(0);
    auto work = [&]() {
        uint64_t local = 0, sink = 0;
        for (;;) {
            uint64_t b = cursor.fetch_add(64);
            if (b >= n_reads) break;
            uint64_t e = std::min<uint64_t>(b + 64, n_reads);
            for (uint64_t i = b; i < e; ++i) {
                uint64_t m = off[i + 1] - off[i];
                auto res = t->query_pml(seqs + off[i], m);
                if (pml) for (uint64_t j = 0; j < m; ++j) pml[off[i] + j] = (uint32_t)res.first[j];
                if (cid) for (uint64_t j = 0; j < m; ++j) cid[off[i] + j] = (uint8_t)res.second[j];
                if (!pml && m) sink += res.first[0] + res.second[m - 1];
                local += m;
            }
        }
        bases += local + (sink & 0);
    };
    if (threads <= 1) work();
    else {
        std::vector<std::thread> pool;
        for (int i = 0; i < threads; ++i) pool.emplace_back(work);
        for (auto &th : pool) th.join();
    }
    return bases.load();
}

} // extern "C"
