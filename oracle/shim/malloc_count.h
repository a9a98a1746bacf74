/* Stand-in for the un-vendored malloc_count dependency (test infrastructure only).
 * The reference only prints malloc_count_peak() (common.hpp:118-120). */
#ifndef ORACLE_SHIM_MALLOC_COUNT_H
#define ORACLE_SHIM_MALLOC_COUNT_H
#include <stddef.h>
static inline size_t malloc_count_peak(void) { return 0; }
#endif
