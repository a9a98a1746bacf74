/* oracle/colbwt_oracle.c -- TEST INFRASTRUCTURE ONLY. See colbwt_oracle.h for scope and parity pin.
 * A plain restatement of the reference algorithm; deliberately simple (column arrays, scalar loops). */
#include "colbwt_oracle.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define ROW_BYTES 18 /* sizeof(col_thr) packed: 1+5+4+2+1+5 (LF_table.hpp:33-84, col_bwt.hpp:40-115) */

static uint64_t le_bytes(const uint8_t *p, int k)
{
    uint64_t v = 0;
    for (int i = 0; i < k; ++i) v |= (uint64_t)p[i] << (8 * i);
    return v;
}

static oracle_table *alloc_table(uint64_t r)
{
    oracle_table *t = (oracle_table *)calloc(1, sizeof(*t));
    if (!t) return NULL;
    size_t rr = r ? r : 1;
    t->r = r;
    t->ch = (uint8_t *)malloc(rr);
    t->idx = (uint64_t *)malloc(rr * 8);
    t->interval = (uint32_t *)malloc(rr * 4);
    t->offset = (uint16_t *)malloc(rr * 2);
    t->col_id = (uint8_t *)malloc(rr);
    t->thr = (uint64_t *)malloc(rr * 8);
    if (!t->ch || !t->idx || !t->interval || !t->offset || !t->col_id || !t->thr) { oracle_free(t); return NULL; }
    return t;
}

/* col_bwt::load (col_bwt.hpp:375-380): bwt_r; LF_table::load (LF_table.hpp:347-357): n, r, size, then
 * read_vec (common.hpp:318-323) = one raw read of size*sizeof(col_thr) bytes. */
oracle_table *oracle_load(const char *path)
{
    FILE *f = fopen(path, "rb");
    if (!f) return NULL;
    uint64_t hdr[4];
    if (fread(hdr, 8, 4, f) != 4) { fclose(f); return NULL; }
    uint64_t size = hdr[3];
    oracle_table *t = alloc_table(size);
    if (!t) { fclose(f); return NULL; }
    t->bwt_r = hdr[0];
    t->n = hdr[1];
    t->r = hdr[2];
    uint8_t row[ROW_BYTES];
    for (uint64_t i = 0; i < size; ++i) {
        if (fread(row, 1, ROW_BYTES, f) != ROW_BYTES) { fclose(f); oracle_free(t); return NULL; }
        t->ch[i] = row[0];
        t->idx[i] = le_bytes(row + 1, 5);
        t->interval[i] = (uint32_t)le_bytes(row + 6, 4);
        t->offset[i] = (uint16_t)le_bytes(row + 10, 2);
        t->col_id[i] = row[12];
        t->thr[i] = le_bytes(row + 13, 5);
    }
    fclose(f);
    if (t->r != size) { oracle_free(t); return NULL; }
    return t;
}

oracle_table *oracle_from_columns(uint64_t bwt_r, uint64_t n, uint64_t r, const uint8_t *ch, const uint64_t *idx,
                                  const uint32_t *interval, const uint16_t *offset, const uint8_t *col_id,
                                  const uint64_t *thr)
{
    oracle_table *t = alloc_table(r);
    if (!t) return NULL;
    t->bwt_r = bwt_r;
    t->n = n;
    memcpy(t->ch, ch, r);
    memcpy(t->idx, idx, r * 8);
    memcpy(t->interval, interval, r * 4);
    memcpy(t->offset, offset, r * 2);
    memcpy(t->col_id, col_id, r);
    memcpy(t->thr, thr, r * 8);
    return t;
}

void oracle_free(oracle_table *t)
{
    if (!t) return;
    free(t->ch); free(t->idx); free(t->interval); free(t->offset); free(t->col_id); free(t->thr);
    free(t);
}

/* LF_table::get_length (LF_table.hpp:204-207) */
static inline uint64_t row_len(const oracle_table *t, uint64_t i)
{
    return (i == t->r - 1) ? (t->n - t->idx[i]) : (t->idx[i + 1] - t->idx[i]);
}

/* LF_table::succ_char (LF_table.hpp:286-298): first row >= run with character c; 0 = none. */
static int succ_char(const oracle_table *t, uint64_t run, uint8_t c, uint64_t *row)
{
    while (t->ch[run] != c) {
        if (run == t->r - 1) return 0;
        ++run;
    }
    *row = run;
    return 1;
}

/* LF_table::pred_char (LF_table.hpp:271-283): last row <= run with character c; 0 = none. */
static int pred_char(const oracle_table *t, uint64_t run, uint8_t c, uint64_t *row)
{
    while (t->ch[run] != c) {
        if (run == 0) return 0;
        --run;
    }
    *row = run;
    return 1;
}

/* col_pml::threshold_step (col_bwt.hpp:531-574), single-threaded branch. pos is the BWT position
 * BEFORE the jump. Successor lands at offset 0, predecessor at len-1; nothing found => unchanged. */
static void threshold_step(const oracle_table *t, uint64_t *interval, uint64_t *offset, uint64_t pos, uint8_t c)
{
    uint64_t new_interval = *interval, new_offset = *offset, thr = t->n, s, p;
    if (succ_char(t, *interval, c, &s)) {
        thr = t->thr[s];
        new_interval = s;
        new_offset = 0;
    }
    if (pos < thr) {
        if (pred_char(t, *interval, c, &p)) {
            new_interval = p;
            new_offset = row_len(t, p) - 1;
        }
    }
    *interval = new_interval;
    *offset = new_offset;
}

/* col_pml::_query_pml core loop (col_bwt.hpp:498-529) + LF_idx (LF_table.hpp:251-268). */
void oracle_query(const oracle_table *t, const uint8_t *pattern, uint64_t m, uint32_t *pml, uint8_t *cid)
{
    if (t->r == 0) return;
    uint64_t pos = t->n - 1;                       /* col_bwt.hpp:503 */
    uint64_t interval = t->r - 1;                  /* :504 */
    uint64_t offset = row_len(t, interval) - 1;    /* :505 */
    uint64_t length = 0;
    for (uint64_t i = 0; i < m; ++i) {
        uint8_t c = pattern[m - i - 1];            /* :512 */
        uint8_t id = t->col_id[interval];          /* :513, sampled before any reposition */
        if (t->ch[interval] == c) {
            ++length;                              /* :516-518 */
        } else {
            length = 0;                            /* :520-523 */
            threshold_step(t, &interval, &offset, pos, c);
        }
        if (pml) pml[m - i - 1] = (uint32_t)length; /* :525 */
        if (cid) cid[m - i - 1] = id;
        /* LF (LF_table.hpp:251-262): follow the (possibly repositioned) row, then fast-forward */
        uint64_t next_interval = t->interval[interval];
        uint64_t next_offset = (uint64_t)t->offset[interval] + offset;
        while (next_offset >= row_len(t, next_interval)) next_offset -= row_len(t, next_interval++);
        interval = next_interval;
        offset = next_offset;
        pos = t->idx[interval] + offset;           /* to_idx, LF_table.hpp:214-217 */
    }
}

/* Experiment support: like oracle_query, and also records the state (row, offset) the traversal is in just before it
 * consumes base j (so state[m-1] is the initial state). */
void oracle_query_states(const oracle_table *t, const uint8_t *pattern, uint64_t m, uint32_t *pml, uint32_t *row_out, uint32_t *off_out)
{
    uint64_t pos = t->n - 1, interval = t->r - 1, offset = row_len(t, interval) - 1, length = 0;
    for (uint64_t i = 0; i < m; ++i) {
        const uint64_t j = m - i - 1;
        uint8_t c = pattern[j];
        row_out[j] = (uint32_t)interval;
        off_out[j] = (uint32_t)offset;
        if (t->ch[interval] == c) ++length;
        else { length = 0; threshold_step(t, &interval, &offset, pos, c); }
        pml[j] = (uint32_t)length;
        uint64_t ni = t->interval[interval], no = (uint64_t)t->offset[interval] + offset;
        while (no >= row_len(t, ni)) no -= row_len(t, ni++);
        interval = ni; offset = no; pos = t->idx[interval] + offset;
    }
}

uint64_t oracle_query_batch(const oracle_table *t, const uint8_t *seqs, const uint64_t *off, uint64_t n_reads,
                            uint32_t *pml, uint8_t *cid)
{
    uint64_t bases = 0, cap = 0;
    uint32_t *tmp_p = NULL;
    uint8_t *tmp_c = NULL;
    volatile uint64_t sink = 0;
    for (uint64_t i = 0; i < n_reads; ++i) {
        uint64_t m = off[i + 1] - off[i];
        if (pml && cid) {
            oracle_query(t, seqs + off[i], m, pml + off[i], cid + off[i]);
        } else {
            if (m > cap) {
                cap = m;
                tmp_p = (uint32_t *)realloc(tmp_p, cap * 4);
                tmp_c = (uint8_t *)realloc(tmp_c, cap);
            }
            oracle_query(t, seqs + off[i], m, tmp_p, tmp_c);
            if (m) sink += tmp_p[0] + tmp_c[m - 1];
        }
        bases += m;
    }
    free(tmp_p);
    free(tmp_c);
    (void)sink;
    return bases;
}

/* pml_query.cpp:79-85: fs << '>' << id << " \n"; each value followed by " "; then "\n". */
size_t oracle_text_bound(size_t id_len, uint64_t m) { return id_len + 4 + (size_t)m * 11 + 1; }

static size_t put_u(char *p, uint32_t v)
{
    char tmp[12];
    int k = 0;
    do { tmp[k++] = (char)('0' + v % 10); v /= 10; } while (v);
    for (int i = 0; i < k; ++i) p[i] = tmp[k - 1 - i];
    return (size_t)k;
}

size_t oracle_format_u32(char *buf, const char *id, size_t id_len, const uint32_t *v, uint64_t m)
{
    char *p = buf;
    *p++ = '>';
    memcpy(p, id, id_len); p += id_len;
    *p++ = ' '; *p++ = '\n';
    for (uint64_t i = 0; i < m; ++i) { p += put_u(p, v[i]); *p++ = ' '; }
    *p++ = '\n';
    return (size_t)(p - buf);
}

size_t oracle_format_u8(char *buf, const char *id, size_t id_len, const uint8_t *v, uint64_t m)
{
    char *p = buf;
    *p++ = '>';
    memcpy(p, id, id_len); p += id_len;
    *p++ = ' '; *p++ = '\n';
    for (uint64_t i = 0; i < m; ++i) { p += put_u(p, v[i]); *p++ = ' '; }
    *p++ = '\n';
    return (size_t)(p - buf);
}
