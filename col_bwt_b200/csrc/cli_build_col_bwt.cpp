// build_col_bwt_b200 -- drop-in for the reference's `build_col_bwt` executable (src/build_col_bwt.cpp:8-64): reads
// <prefix>.bwt.heads / .bwt.len / .thr_pos / .col_runs / .col_ids and writes <prefix>.col_pml, with the table
// constructed on the GPU (colbwt_index_from_primaries).  Difference: `.col_runs` is read as the plain sdsl
// bit_vector that col_split actually writes (col_split.hpp:384-386); the reference binary expects an sd_vector
// there and so cannot consume col_split's own output.
#include <cstdio>
#include <string>

#include "colbwt_b200.h"

int main(int argc, char **argv)
{
    std::string prefix;
    for (int i = 1; i < argc; ++i)
        if (argv[i][0] != '-') prefix = argv[i];
    if (prefix.empty()) {
        fprintf(stderr, "[ERROR]: Invalid number of arguments\nusage: build_col_bwt_b200 <prefix> [-v]\n");
        return 2;
    }
    printf("[INFO]: Building Col BWT supporting Co-linearity Statistics: \n");
    colbwt_index *idx = nullptr;
    if (colbwt_index_from_primaries(prefix.c_str(), nullptr, 1, &idx) != COLBWT_OK) {
        fprintf(stderr, "[ERROR]: %s\n", colbwt_last_error());
        return 1;
    }
    colbwt_stats st;
    colbwt_index_stats(idx, &st);
    printf("Col-BWT runs: %llu\nBWT runs: %llu\nText length: %llu\n", (unsigned long long)st.r, (unsigned long long)st.bwt_r, (unsigned long long)st.n);
    printf("[INFO]: Serializing\n");
    const std::string out = prefix + ".col_pml";
    const int rc = colbwt_index_save(idx, out.c_str());
    if (rc != COLBWT_OK) fprintf(stderr, "[ERROR]: %s\n", colbwt_last_error());
    colbwt_index_free(idx);
    if (rc == COLBWT_OK) printf("[INFO]: Done\n");
    return rc == COLBWT_OK ? 0 : 1;
}
