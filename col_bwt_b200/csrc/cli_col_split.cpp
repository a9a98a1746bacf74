// col_split_b200 -- drop-in for the reference's `col_split` executable (src/col_split.cpp:62-141): same arguments
// (`<prefix> -m tunnels|all -s <rate> [-v]`, the shared getopt string of include/common/common.hpp:231), same outputs
// (<prefix>.col_runs as a plain sdsl bit_vector, <prefix>.col_ids).  Reads <prefix>.bwt.heads/.bwt.len/.col_mums directly,
// so the build_FL step and its <prefix>.FL_table file are not needed.  `-o truncate` is not implemented (the reference's
// default `append` is; find_col_runs ignores the option as well).
#include <getopt.h>

#include <cstdio>
#include <cstdlib>
#include <string>

#include "colbwt_b200.h"

int main(int argc, char *const argv[])
{
    std::string mode = "default";
    int rate = 1, c;   // Args defaults (common.hpp:215,221)
    while ((c = getopt(argc, argv, "rvlN:p:m:s:o:")) != -1) {
        if (c == 'm') mode = optarg;
        else if (c == 's') rate = atoi(optarg);
    }
    if (argc != optind + 1) {
        fprintf(stderr, "[ERROR]: Invalid number of arguments\n");
        return 2;
    }
    if (mode != "all" && mode != "tunnels") {   // parse_mode (col_split.cpp:34-46) throws on anything else
        fprintf(stderr, "[ERROR]: Invalid split mode: %s. Must be one of: all, tunnels\n", mode.c_str());
        return 2;
    }
    printf("[INFO]: Splitting runs based on multi-MUM Positions using FL Table\n");
    uint64_t bits = 0, marked = 0;
    if (colbwt_col_split(argv[optind], mode == "all", rate, 0, &bits, &marked) != COLBWT_OK) {
        fprintf(stderr, "[ERROR]: %s\n", colbwt_last_error());
        return 1;
    }
    printf("Col runs: %llu\nTotal runs: %llu\n[INFO]: Done\n", (unsigned long long)marked, (unsigned long long)bits);
    return 0;
}
