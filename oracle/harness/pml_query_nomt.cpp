/* TEST INFRASTRUCTURE (oracle). Builds the reference's own pml_query main from
 * /root/reference/src/pml_query.cpp with the MULTI_THREAD macro (include/common/common.hpp:50)
 * switched off, without copying or editing any reference source: common.hpp is include-guarded,
 * so including it first and #undef-ing the macro makes col_bwt.hpp:537-550 take its
 * single-threaded branch. Output is byte-identical to the shipped build. */
#include <common.hpp>
#undef MULTI_THREAD
#include REF_PML_QUERY_CPP
