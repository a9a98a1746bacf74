/* oracle/colbwt_oracle.h -- TEST INFRASTRUCTURE ONLY (never linked or called by the product).
 *
 * Plain-C CPU restatement of col-bwt's query hot path (PML + chain ids per base), i.e. of
 *   col_pml::_query_pml / threshold_step   /root/reference/include/col_bwt.hpp:498-574
 *   LF_table::LF / LF_idx / pred_char / succ_char / get_length
 *                                           /root/reference/include/ds/LF_table.hpp:204-298
 *   col_bwt::load + LF_table::load          col_bwt.hpp:375-380, LF_table.hpp:347-357
 *   pml_to_vec text output                  /root/reference/src/pml_query.cpp:65-90
 *
 * Parity pin: checked against (a) the known-answer vector of SURVEY.md section 4.3 and (b) outputs of
 * the reference's own sources compiled verbatim into oracle/_ref (tests/test_oracle.py,
 * tests/golden/, generator tests/golden/make_golden.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library -- as the checker, never as the thing shipped.
 */
#ifndef COLBWT_ORACLE_H
#define COLBWT_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    uint64_t bwt_r;     /* BWT runs before sub-run splitting (col_bwt.hpp:382) */
    uint64_t n;         /* BWT length (LF_table.hpp:359)                       */
    uint64_t r;         /* number of rows (LF_table.hpp:360)                   */
    uint8_t  *ch;       /* row character, raw byte (LF_table.hpp:36)           */
    uint64_t *idx;      /* BWT start of row, 40 bit (LF_table.hpp:37)          */
    uint32_t *interval; /* LF destination row (LF_table.hpp:38)                */
    uint16_t *offset;   /* offset inside destination row (LF_table.hpp:39)     */
    uint8_t  *col_id;   /* chain id, 0 = unmarked (col_bwt.hpp:43)             */
    uint64_t *thr;      /* threshold, absolute BWT position (col_bwt.hpp:84)   */
} oracle_table;

/* Load PREFIX.col_pml given its full path. Returns NULL on a missing/short file. */
oracle_table *oracle_load(const char *col_pml_path);
/* Build a table from caller-owned columns (copied). */
oracle_table *oracle_from_columns(uint64_t bwt_r, uint64_t n, uint64_t r, const uint8_t *ch, const uint64_t *idx,
                                  const uint32_t *interval, const uint16_t *offset, const uint8_t *col_id,
                                  const uint64_t *thr);
void oracle_free(oracle_table *t);

/* One read of m raw bytes -> m PML values and m chain ids, indexed by read position. */
void oracle_query(const oracle_table *t, const uint8_t *pattern, uint64_t m, uint32_t *pml, uint8_t *cid);

/* n_reads reads (concatenated bytes, n_reads+1 offsets); outputs at the same offsets.
 * Sequential, one thread: the CPU "port" baseline. Returns the number of bases processed.
 * pml/cid may be NULL (timing only: a checksum keeps the work alive). */
uint64_t oracle_query_batch(const oracle_table *t, const uint8_t *seqs, const uint64_t *off, uint64_t n_reads,
                            uint32_t *pml, uint8_t *cid);

/* Text writer of pml_query.cpp:65-90 for one read: ">id \n" then "v " per base then "\n".
 * Returns bytes written into buf (cap must be >= oracle_text_bound(id_len, m)). */
size_t oracle_text_bound(size_t id_len, uint64_t m);
size_t oracle_format_u32(char *buf, const char *id, size_t id_len, const uint32_t *v, uint64_t m);
size_t oracle_format_u8(char *buf, const char *id, size_t id_len, const uint8_t *v, uint64_t m);

#ifdef __cplusplus
}
#endif
#endif
