/* Stand-in for klib's kseq.h (un-vendored dependency of the reference; test infrastructure only).
 * Written from kseq's documented behaviour: FASTA/FASTQ records, name = header up to the first
 * whitespace, sequence = all non-newline bytes of the sequence lines concatenated verbatim.
 * Surface used by the reference: KSEQ_INIT, kseq_t{name,seq}, kseq_init, kseq_read (io.hpp:1-35). */
#ifndef ORACLE_SHIM_KSEQ_H
#define ORACLE_SHIM_KSEQ_H
#include <stdlib.h>
#include <string.h>
#include <ctype.h>

typedef struct { size_t l, m; char *s; } kstring_t;

#define KSEQ_BUFSZ 65536

#define KSEQ_INIT(type_t, __read)                                                              \
  typedef struct {                                                                             \
    kstring_t name, comment, seq, qual;                                                        \
    int last_char;                                                                             \
    type_t f;                                                                                  \
    unsigned char *buf; int beg, end, eof;                                                     \
  } kseq_t;                                                                                    \
  static inline kseq_t *kseq_init(type_t fd) {                                                 \
    kseq_t *s = (kseq_t *)calloc(1, sizeof(kseq_t));                                           \
    s->f = fd; s->buf = (unsigned char *)malloc(KSEQ_BUFSZ); return s;                         \
  }                                                                                            \
  static inline void kseq_destroy(kseq_t *s) {                                                 \
    if (!s) return; free(s->name.s); free(s->comment.s); free(s->seq.s); free(s->qual.s);      \
    free(s->buf); free(s);                                                                     \
  }                                                                                            \
  static inline int ks_getc_(kseq_t *s) {                                                      \
    if (s->beg >= s->end) {                                                                    \
      if (s->eof) return -1;                                                                   \
      s->beg = 0; s->end = __read(s->f, s->buf, KSEQ_BUFSZ);                                   \
      if (s->end <= 0) { s->eof = 1; s->end = 0; return -1; }                                  \
    }                                                                                          \
    return (int)s->buf[s->beg++];                                                              \
  }                                                                                            \
  static inline void ks_push_(kstring_t *k, int c) {                                           \
    if (k->l + 2 > k->m) { k->m = k->m ? k->m * 2 : 256; k->s = (char *)realloc(k->s, k->m); } \
    k->s[k->l++] = (char)c; k->s[k->l] = 0;                                                    \
  }                                                                                            \
  static inline int kseq_read(kseq_t *s) {                                                     \
    int c;                                                                                     \
    if (s->last_char == 0) {                                                                   \
      while ((c = ks_getc_(s)) != -1 && c != '>' && c != '@') {}                               \
      if (c == -1) return -1;                                                                  \
      s->last_char = c;                                                                        \
    }                                                                                          \
    s->name.l = s->comment.l = s->seq.l = s->qual.l = 0;                                       \
    ks_push_(&s->name, 0); s->name.l = 0; ks_push_(&s->seq, 0); s->seq.l = 0;                  \
    while ((c = ks_getc_(s)) != -1 && !isspace(c)) ks_push_(&s->name, c);                      \
    if (c == -1) return -1;                                                                    \
    if (c != '\n') while ((c = ks_getc_(s)) != -1 && c != '\n') ks_push_(&s->comment, c);      \
    while ((c = ks_getc_(s)) != -1 && c != '>' && c != '+' && c != '@') {                      \
      if (c == '\n') continue;                                                                 \
      ks_push_(&s->seq, c);                                                                    \
      while ((c = ks_getc_(s)) != -1 && c != '\n') ks_push_(&s->seq, c);                       \
      if (s->seq.l > 1 && s->seq.s[s->seq.l - 1] == '\r') s->seq.s[--s->seq.l] = 0;            \
    }                                                                                          \
    if (c == '>' || c == '@') s->last_char = c;                                                \
    if (c != '+') { if (c == -1) s->last_char = 0, s->eof = 1; return (int)s->seq.l; }         \
    while ((c = ks_getc_(s)) != -1 && c != '\n') {}                                            \
    if (c == -1) return -2;                                                                    \
    while ((c = ks_getc_(s)) != -1 && s->qual.l < s->seq.l) { if (c != '\n') ks_push_(&s->qual, c); } \
    s->last_char = 0;                                                                          \
    if (s->seq.l != s->qual.l) return -2;                                                      \
    return (int)s->seq.l;                                                                      \
  }
#endif
