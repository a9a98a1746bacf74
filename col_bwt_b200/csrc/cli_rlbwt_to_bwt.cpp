// rlbwt_to_bwt_b200 -- drop-in for the reference's `rlbwt_to_bwt` executable (src/rlbwt_to_bwt.cpp:8-34): same argument
// (`<prefix>`, plus the shared getopt string of include/common/common.hpp:231), same output (<prefix>.bwt).
#include <getopt.h>

#include <cstdio>

#include "colbwt_b200.h"

int main(int argc, char *const argv[])
{
    int c;
    while ((c = getopt(argc, argv, "rvlN:p:m:s:o:")) != -1) {
    }
    if (argc != optind + 1) {
        fprintf(stderr, "[ERROR]: Invalid number of arguments\n");
        return 2;
    }
    printf("[INFO]: Creating BWT from RLBWT\n");
    uint64_t n = 0;
    if (colbwt_rlbwt_to_bwt(argv[optind], 0, &n) != COLBWT_OK) {
        fprintf(stderr, "[ERROR]: %s\n", colbwt_last_error());
        return 1;
    }
    printf("\t[INFO]: %llu characters written to %s.bwt\n[INFO]: Done\n", (unsigned long long)n, argv[optind]);
    return 0;
}
