// pml_query_b200 -- drop-in for the reference's `pml_query` executable (src/pml_query.cpp:92-143) running the
// query on B200 GPUs through the C-ABI.  Same command line (`<prefix> -p <reads> [-v] [-l]`, one shared getopt
// string "rvlN:p:m:s:o:", include/common/common.hpp:231), same input (`<prefix>.col_pml`), same outputs
// (`<reads>.pml`, `<reads>.cid`, text of pml_to_vec, src/pml_query.cpp:65-90).  Extra: -G <n> GPUs (default 1).
// Unlike the reference (whose main always returns 0 and whose -l path is broken for more than one read,
// SURVEY.md section 3.1), failures give a non-zero exit code and -l is accepted as a no-op.
#include <getopt.h>
#include <unistd.h>

#include <chrono>
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

    std::vector<uint8_t> seqs;
    std::vector<uint64_t> off;
    std::vector<std::string> ids;
    std::vector<uint32_t> pml;
    std::vector<uint8_t> cid;
    uint64_t total_bases = 0, total_reads = 0;
    int rc = 0;
    while (reader.next_batch(seqs, off, ids, 256ull << 20, 8ull << 20)) {
        pml.resize(seqs.size() + 1);
        cid.resize(seqs.size() + 1);
        if (colbwt_query(idx, seqs.data(), off.data(), ids.size(), pml.data(), COLBWT_PML_U32, cid.data()) != COLBWT_OK) {
            fprintf(stderr, "[ERROR]: %s\n", colbwt_last_error());
            rc = 1;
            break;
        }
        // format in parallel slices, write in order
        const int T = (int)std::max<size_t>(1, std::min<size_t>(std::thread::hardware_concurrency(), ids.size() / 256 + 1));
        std::vector<std::string> out_p(T), out_c(T);
        std::vector<std::thread> th;
        for (int t = 0; t < T; ++t)
            th.emplace_back([&, t] {
                const size_t a = ids.size() * t / T, b = ids.size() * (t + 1) / T;
                std::string &sp = out_p[t], &sc = out_c[t];
                for (size_t i = a; i < b; ++i) {
                    const uint64_t m = off[i + 1] - off[i];
                    size_t need = colbwt_format_stats(nullptr, 0, ids[i].data(), ids[i].size(), pml.data() + off[i], 4, m);
                    size_t at = sp.size();
                    sp.resize(at + need);
                    colbwt_format_stats(&sp[at], need, ids[i].data(), ids[i].size(), pml.data() + off[i], 4, m);
                    need = colbwt_format_stats(nullptr, 0, ids[i].data(), ids[i].size(), cid.data() + off[i], 1, m);
                    at = sc.size();
                    sc.resize(at + need);
                    colbwt_format_stats(&sc[at], need, ids[i].data(), ids[i].size(), cid.data() + off[i], 1, m);
                }
            });
        for (auto &x : th) x.join();
        for (int t = 0; t < T; ++t) {
            fwrite(out_p[t].data(), 1, out_p[t].size(), f_pml);
            fwrite(out_c[t].data(), 1, out_c[t].size(), f_cid);
        }
        total_bases += seqs.size();
        total_reads += ids.size();
    }
    fclose(f_pml);
    fclose(f_cid);
    colbwt_index_free(idx);
    if (rc) { unlink(pml_name.c_str()); unlink(cid_name.c_str()); return rc; }   // col-bwt.py:70-77 deletes outputs on failure
    const double dt = now_s() - t1;
    printf("\t[INFO]: Query Complete\n\t[INFO]: Elapsed time (s): %g\n", dt);
    if (verbose) printf("\t[LOG]: %llu reads, %llu bases, %.3f Mbases/s end to end\n", (unsigned long long)total_reads,
                        (unsigned long long)total_bases, total_bases / dt / 1e6);
    printf("\t[INFO]: PMLs written to: %s\n\t[INFO]: CIDs written to: %s\n[INFO]: Done\n", pml_name.c_str(), cid_name.c_str());
    return 0;
}
