/*
 * oracle/fasta_ref.h -- TEST INFRASTRUCTURE ONLY (CPU oracle). Never linked into the product.
 *
 * FASTA/FASTQ record reader shared by the two oracle programs.  It restates the record
 * framing of the parser both upstream tools use (kseq.h over zlib): SURVEY.md Appendix A.1.
 * The upstream sources are NOT under /root/reference (btb-phylo only shells out to the
 * binaries: btbphylo/phylogeny.py:133 and :149), so this is written from the behavioural
 * spec, not from upstream code.   PARITY UNPINNED against real binaries.
 *
 * Framing rules restated here:
 *   - input is read through zlib (plain or gzip);
 *   - bytes before the first '>' / '@' are skipped;
 *   - name = bytes after the marker up to the first whitespace; rest of header line dropped;
 *   - sequence = concatenation of following lines until a line whose FIRST byte is
 *     '>', '@' or '+'; '\n' dropped, one trailing '\r' per line dropped, empty lines skipped,
 *     every other byte (spaces included) kept;
 *   - a '+' line starts a FASTQ quality block, which is consumed (>= seq length bytes).
 */
#ifndef BTB_ORACLE_FASTA_REF_H
#define BTB_ORACLE_FASTA_REF_H

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

typedef struct {
    gzFile fp;
    unsigned char *buf;
    int beg, end, eof;
    int last;          /* marker byte already consumed for the next record, or 0 */
    char *name;  size_t name_l, name_m;
    char *seq;   size_t seq_l, seq_m;
} fa_reader;

#define FA_BUFSZ (1 << 16)

static int fa_getc(fa_reader *r) {
    if (r->beg >= r->end) {
        if (r->eof) return -1;
        r->beg = 0;
        r->end = gzread(r->fp, r->buf, FA_BUFSZ);
        if (r->end <= 0) { r->eof = 1; r->end = 0; return -1; }
    }
    return r->buf[r->beg++];
}

static void fa_push(char **s, size_t *l, size_t *m, int c) {
    if (*l + 2 > *m) {
        *m = *m ? *m * 2 : 256;
        *s = (char *)realloc(*s, *m);
        if (!*s) { fprintf(stderr, "oracle: out of memory\n"); exit(2); }
    }
    (*s)[(*l)++] = (char)c;
    (*s)[*l] = 0;
}

static fa_reader *fa_open(const char *path) {
    gzFile fp = gzopen(path, "r");
    if (!fp) return NULL;
    fa_reader *r = (fa_reader *)calloc(1, sizeof(fa_reader));
    r->fp = fp;
    r->buf = (unsigned char *)malloc(FA_BUFSZ);
    return r;
}

static void fa_close(fa_reader *r) {
    if (!r) return;
    gzclose(r->fp);
    free(r->buf); free(r->name); free(r->seq); free(r);
}

/* Read the rest of the current line into (s,l,m) when s != NULL, else discard it.
 * Returns the terminating byte ('\n') or -1 at EOF. */
static int fa_rest_of_line(fa_reader *r, char **s, size_t *l, size_t *m) {
    int c;
    while ((c = fa_getc(r)) >= 0 && c != '\n')
        if (s) fa_push(s, l, m, c);
    /* one trailing CR per line is dropped (whole-string test, so a line that is only "\r"
     * is dropped too unless it is the very first sequence byte) */
    if (s && *l > 1 && (*s)[*l - 1] == '\r') { (*l)--; (*s)[*l] = 0; }
    return c;
}

/* Returns sequence length (>= 0), or -1 at end of input. */
static long fa_read(fa_reader *r) {
    int c;
    if (r->last == 0) {
        while ((c = fa_getc(r)) >= 0 && c != '>' && c != '@') {}
        if (c < 0) return -1;
        r->last = c;
    }
    r->name_l = 0; r->seq_l = 0;
    if (r->name) r->name[0] = 0;
    if (r->seq) r->seq[0] = 0;
    /* header: name up to first whitespace */
    while ((c = fa_getc(r)) >= 0 && c != ' ' && c != '\t' && c != '\n' && c != '\r' &&
           c != '\v' && c != '\f')
        fa_push(&r->name, &r->name_l, &r->name_m, c);
    if (c < 0) { r->last = 0; if (!r->name) fa_push(&r->name, &r->name_l, &r->name_m, 0), r->name_l = 0; return 0; }
    if (c != '\n') fa_rest_of_line(r, NULL, NULL, NULL);
    if (!r->name) { fa_push(&r->name, &r->name_l, &r->name_m, 0); r->name_l = 0; }
    if (!r->seq) { fa_push(&r->seq, &r->seq_l, &r->seq_m, 0); r->seq_l = 0; }
    /* sequence lines */
    while ((c = fa_getc(r)) >= 0 && c != '>' && c != '+' && c != '@') {
        if (c == '\n') continue;
        fa_push(&r->seq, &r->seq_l, &r->seq_m, c);
        fa_rest_of_line(r, &r->seq, &r->seq_l, &r->seq_m);
    }
    if (c == '>' || c == '@') { r->last = c; return (long)r->seq_l; }
    r->last = 0;
    if (c == '+') {            /* FASTQ quality block: skip '+' line, then >= seq_l bytes */
        size_t q = 0;
        fa_rest_of_line(r, NULL, NULL, NULL);
        while (q < r->seq_l && (c = fa_getc(r)) >= 0)
            if (c != '\n' && c != '\r') q++;
        if (q) fa_rest_of_line(r, NULL, NULL, NULL);
    }
    return (long)r->seq_l;
}

#endif
