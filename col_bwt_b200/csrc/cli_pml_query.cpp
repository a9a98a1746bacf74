// pml_query_b200 -- drop-in for the reference's `pml_query` executable (src/pml_query.cpp:92-143) running the
// query on B200 GPUs through the C-ABI.  Same command line (`<prefix> -p <reads> [-v] [-l]`, one shared getopt
// string "rvlN:p:m:s:o:", include/common/common.hpp:231), same input (`<prefix>.col_pml`), same outputs
// (`<reads>.pml`, `<reads>.cid`, text of pml_to_vec, src/pml_query.cpp:65-90).  Extra: -G <n> GPUs (default 1).
// Unlike the reference (whose main always returns 0 and whose -l path is broken for more than one read,
// SURVEY.md section 3.1), failures give a non-zero exit code and -l is accepted as a no-op.
#include <getopt.h>
#include <algorithm>
#include <unistd.h>

#include <chrono>
#include <condition_variable>
#include <deque>
#include <memory>
#include <mutex>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "colbwt_b200.h"
#include "fastx.h"

static double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int main(int argc, char *const argv[])
{
    std::string pattern;
    bool verbose = false;
    int gpus = 1, c;
    while ((c = getopt(argc, argv, "rvlN:p:m:s:o:G:")) != -1) {
        switch (c) {
        case 'v': verbose = true; break;
        case 'p': pattern = optarg; break;
        case 'G': gpus = atoi(optarg); break;
        case '?': printf("ERROR: Unknown option.\n"); break;
        default: break;   // r l N m s o: accepted for command-line compatibility, unused by the query
        }
    }
    if (argc != optind + 1) { fprintf(stderr, "[ERROR]: Invalid number of arguments\n"); return 2; }
    if (pattern.empty()) { fprintf(stderr, "[ERROR]: Pattern file not provided\n"); return 2; }
    const std::string prefix = argv[optind];

    printf("[INFO]: Loading BWT table supporting LF mapping: \n");
    double t0 = now_s();
    colbwt_index *idx = nullptr;
    if (colbwt_index_load((prefix + ".col_pml").c_str(), nullptr, gpus, &idx) != COLBWT_OK) {
        fprintf(stderr, "[ERROR]: %s\n", colbwt_last_error());
        return 1;
    }
    colbwt_stats st;
    colbwt_index_stats(idx, &st);
    if (verbose) {
        printf("\t[LOG]: Number of Col equal-letter runs: r = %llu\n", (unsigned long long)st.r);
        printf("\t[LOG]: Number of BWT equal-letter runs: bwt_r = %llu\n", (unsigned long long)st.bwt_r);
        printf("\t[LOG]: Length of complete BWT: n = %llu\n", (unsigned long long)st.n);
        printf("\t[LOG]: Marked rows: %llu, rows using exact search: %llu, HBM bytes per GPU: %llu, GPUs: %d\n",
               (unsigned long long)st.marked_rows, (unsigned long long)st.slow_rows, (unsigned long long)st.device_bytes, st.n_devices);
    }
    printf("\t[INFO]: Load Complete\n\t[INFO]: Elapsed time (s): %g\n", now_s() - t0);

    printf("[INFO]: Computing PML Queries: \n");
    double t1 = now_s();
    colbwt::FastxReader reader(pattern);
    if (!reader.ok()) { fprintf(stderr, "[ERROR]: cannot open %s\n", pattern.c_str()); colbwt_index_free(idx); return 1; }
    const std::string pml_name = pattern + ".pml", cid_name = pattern + ".cid";
    FILE *f_pml = fopen(pml_name.c_str(), "wb"), *f_cid = fopen(cid_name.c_str(), "wb");
    if (!f_pml || !f_cid) { fprintf(stderr, "[ERROR]: cannot write %s / %s\n", pml_name.c_str(), cid_name.c_str()); return 1; }

    // Reader thread: parses the next batch (gz + FASTA/FASTQ) while this thread queries, formats and writes the current one.
    struct Batch {
        std::vector<uint8_t> seqs;
        std::vector<uint64_t> off;
        std::vector<std::string> ids;
    };
    std::mutex mtx;
    std::condition_variable cv;
    std::deque<std::unique_ptr<Batch>> queue;
    bool reader_done = false;
    std::thread reader_thread([&] {
        for (;;) {
            std::unique_ptr<Batch> b(new Batch);
            const size_t n = reader.next_batch(b->seqs, b->off, b->ids, 256ull << 20, 8ull << 20);
            std::unique_lock<std::mutex> lk(mtx);
            cv.wait(lk, [&] { return queue.size() < 2; });
            if (n) queue.push_back(std::move(b));
            else reader_done = true;
            cv.notify_all();
            if (!n) return;
        }
    });
    // result buffers in pinned memory (grown on demand): colbwt_query then copies device -> host without a staging hop
    uint8_t *pml = nullptr, *cid = nullptr;
    size_t pml_cap = 0, cid_cap = 0;
    uint64_t total_bases = 0, total_reads = 0;
    double t_wait = 0, t_query = 0, t_format = 0, t_write = 0;
    int rc = 0;
    for (;;) {
        double t = now_s();
        std::unique_ptr<Batch> bp;
        {
            std::unique_lock<std::mutex> lk(mtx);
            cv.wait(lk, [&] { return !queue.empty() || reader_done; });
            if (queue.empty()) break;
            bp = std::move(queue.front());
            queue.pop_front();
            cv.notify_all();
        }
        t_wait += now_s() - t;
        t = now_s();
        const std::vector<uint8_t> &seqs = bp->seqs;
        const std::vector<uint64_t> &off = bp->off;
        const std::vector<std::string> &ids = bp->ids;
        uint64_t max_len = 0;
        for (size_t i = 0; i + 1 < off.size(); ++i) max_len = std::max(max_len, off[i + 1] - off[i]);
        const int width = max_len < 256 ? COLBWT_PML_U8 : max_len < 65536 ? COLBWT_PML_U16 : COLBWT_PML_U32;   // fewer bytes over PCIe
        if ((seqs.size() + 1) * (size_t)width > pml_cap) {
            colbwt_host_free(pml);
            pml_cap = (seqs.size() + 1) * (size_t)width;
            pml = (uint8_t *)colbwt_host_alloc(pml_cap);
        }
        if (seqs.size() + 1 > cid_cap) {
            colbwt_host_free(cid);
            cid_cap = seqs.size() + 1;
            cid = (uint8_t *)colbwt_host_alloc(cid_cap);
        }
        if (!pml || !cid) {
            fprintf(stderr, "[ERROR]: %s\n", colbwt_last_error());
            rc = 1;
        }
        if (rc == 0 && colbwt_query(idx, seqs.data(), off.data(), ids.size(), pml, width, cid) != COLBWT_OK) {
            fprintf(stderr, "[ERROR]: %s\n", colbwt_last_error());
            rc = 1;
        }
        if (rc) continue;   // keep draining the reader
        t_query += now_s() - t;
        t = now_s();
        // format in parallel slices, write in order
        const int T = (int)std::max<size_t>(1, std::min<size_t>(std::thread::hardware_concurrency(), ids.size() / 256 + 1));
        std::vector<std::string> out_p(T), out_c(T);
        std::vector<std::thread> th;
        for (int t2 = 0; t2 < T; ++t2)
            th.emplace_back([&, t2] {
                const size_t a = ids.size() * t2 / T, b = ids.size() * (t2 + 1) / T;
                std::string &sp = out_p[t2], &sc = out_c[t2];
                for (size_t i = a; i < b; ++i) {   // append with worst-case room, then trim: one pass per value
                    const uint64_t m = off[i + 1] - off[i];
                    size_t at = sp.size();
                    sp.resize(at + ids[i].size() + 4 + m * 11);
                    sp.resize(at + colbwt_format_stats(&sp[at], sp.size() - at, ids[i].data(), ids[i].size(), pml + off[i] * width, width, m));
                    at = sc.size();
                    sc.resize(at + ids[i].size() + 4 + m * 4);
                    sc.resize(at + colbwt_format_stats(&sc[at], sc.size() - at, ids[i].data(), ids[i].size(), cid + off[i], 1, m));
                }
            });
        for (auto &x : th) x.join();
        t_format += now_s() - t;
        t = now_s();
        for (int t2 = 0; t2 < T; ++t2) {
            fwrite(out_p[t2].data(), 1, out_p[t2].size(), f_pml);
            fwrite(out_c[t2].data(), 1, out_c[t2].size(), f_cid);
        }
        t_write += now_s() - t;
        total_bases += seqs.size();
        total_reads += ids.size();
    }
    reader_thread.join();
    colbwt_host_free(pml);
    colbwt_host_free(cid);
    fclose(f_pml);
    fclose(f_cid);
    colbwt_index_free(idx);
    if (rc) { unlink(pml_name.c_str()); unlink(cid_name.c_str()); return rc; }   // col-bwt.py:70-77 deletes outputs on failure
    const double dt = now_s() - t1;
    printf("\t[INFO]: Query Complete\n\t[INFO]: Elapsed time (s): %g\n", dt);
    if (verbose)
        printf("\t[LOG]: %llu reads, %llu bases, %.3f Mbases/s end to end (waiting for the parser %.2f s, query %.2f s, formatting %.2f s, writing %.2f s)\n",
               (unsigned long long)total_reads, (unsigned long long)total_bases, total_bases / dt / 1e6, t_wait, t_query, t_format, t_write);
    printf("\t[INFO]: PMLs written to: %s\n\t[INFO]: CIDs written to: %s\n[INFO]: Done\n", pml_name.c_str(), cid_name.c_str());
    return 0;
}
