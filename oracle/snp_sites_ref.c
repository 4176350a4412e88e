/*
 * oracle/snp_sites_ref.c -- TEST INFRASTRUCTURE ONLY (CPU oracle). Never linked into the product.
 *
 * CPU restatement of the variable-column extraction that btb-phylo obtains by running
 *     snp-sites <multi_fasta> -c -o <snps.fas>          (btbphylo/phylogeny.py:128-141, cmd at :133)
 * The algorithm lives in the third-party program snp-sites (sanger-pathogens/snp-sites, the
 * Ubuntu 22.04 apt package = upstream 2.5.1; install/install-snp-sites.sh:8,10) whose source is
 * absent from /root/reference.  This file restates its published behaviour from SURVEY.md
 * section 8a row a1 and Appendix A.2 -- it is not a copy of upstream code.
 * PARITY UNPINNED: no real snp-sites binary or reference golden vector exists offline; the
 * oracle is pinned only by the upstream README example (KAT-2) and hand-derived vectors in
 * tests/golden/.
 *
 * Rules (Appendix A.2):
 *   pure mode (-c): column kept  <=>  every byte in "ACGTacgt"  and  >= 2 distinct toupper() values.
 *   default mode  : bytes N n - ? are "unknown" and ignored; every other byte takes part after
 *                   toupper(); column kept <=> >= 2 distinct non-unknown values.
 *   all records must have the length of the first, else exit 1 with the upstream message;
 *   zero kept columns -> warning on stderr, exit 1, no output file;
 *   output: for each record ">name\n" + toupper(bytes at kept columns) + "\n".
 *
 * Build:  make -C oracle        (gcc -O3 -std=c99, links zlib)
 * Usage:  snp-sites [-c] [-o out] <file>      (other upstream flags are rejected)
 * Library entry points (ctypes, used by tests/bench only): oracle_snp_sites_file(),
 * oracle_snp_sites_mask().
 */
#include <ctype.h>
#include <stdint.h>
#include "fasta_ref.h"

/* per-column state (16-bit so that every byte value can be a 'single value seen') */
typedef uint16_t colst;
#define ST_UNSET 0x100
#define ST_VAR 0x101
#define ST_EXCL 0x102

static int is_unknown(int c) { return c == 'N' || c == 'n' || c == '-' || c == '?'; }
static int is_pure(int c) {
    switch (c) { case 'A': case 'C': case 'G': case 'T': case 'a': case 'c': case 'g': case 't': return 1; }
    return 0;
}

static colst *colst_new(long G) {
    colst *st = (colst *)malloc(sizeof(colst) * (size_t)(G > 0 ? G : 1));
    for (long k = 0; k < G; k++) st[k] = ST_UNSET;
    return st;
}

static void column_update(colst *st, const char *s, size_t G, int pure) {
    for (size_t k = 0; k < G; k++) {
        int c = (unsigned char)s[k];
        if (st[k] == ST_EXCL) continue;
        if (pure) {
            if (!is_pure(c)) { st[k] = ST_EXCL; continue; }
        } else if (is_unknown(c)) continue;
        int u = toupper(c);
        if (st[k] == ST_UNSET) st[k] = (colst)u;
        else if (st[k] != ST_VAR && st[k] != u) st[k] = ST_VAR;
    }
}

/* In-memory form used by tests: seqs = n rows of G bytes (row-major).  keep[k] = 1 if column k is
 * kept.  Returns the number of kept columns. */
long oracle_snp_sites_mask(const char *seqs, long n, long G, int pure, unsigned char *keep) {
    colst *st = colst_new(G);
    for (long i = 0; i < n; i++) column_update(st, seqs + (size_t)i * G, (size_t)G, pure);
    long L = 0;
    for (long k = 0; k < G; k++) { keep[k] = (st[k] == ST_VAR); L += keep[k]; }
    free(st);
    return L;
}

/* File form: returns 0 on success, 1 on any upstream exit-1 condition.  n_out/G_out/L_out may be NULL. */
int oracle_snp_sites_file(const char *in_path, const char *out_path, int pure,
                          long *n_out, long *G_out, long *L_out) {
    fa_reader *r = fa_open(in_path);
    if (!r) { fprintf(stderr, "ERROR: cannot open %s\n", in_path); return 1; }
    long G = -1, n = 0, l;
    colst *st = NULL;
    while ((l = fa_read(r)) >= 0) {
        if (G < 0) { G = l; st = colst_new(G); }
        if (l != G) {
            fprintf(stderr, "Alignment %s contains sequences of unequal length. Expected length is %i but got %i in sequence %s\n\n",
                    in_path, (int)G, (int)l, r->name);
            fa_close(r); free(st); return 1;
        }
        column_update(st, r->seq, (size_t)G, pure);
        n++;
    }
    fa_close(r);
    long L = 0;
    for (long k = 0; k < G; k++) L += (st[k] == ST_VAR);
    if (n_out) *n_out = n;
    if (G_out) *G_out = G < 0 ? 0 : G;
    if (L_out) *L_out = L;
    if (n == 0 || L == 0) {
        fprintf(stderr, "Warning: No SNPs were detected so there is nothing to output.\n");
        free(st); return 1;
    }
    uint32_t *idx = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)L);
    long j = 0;
    for (long k = 0; k < G; k++) if (st[k] == ST_VAR) idx[j++] = (uint32_t)k;
    free(st);
    /* pass 2 */
    r = fa_open(in_path);
    FILE *out = out_path ? fopen(out_path, "w") : stdout;
    if (!r || !out) { fprintf(stderr, "ERROR: cannot open output %s\n", out_path ? out_path : "-"); free(idx); return 1; }
    char *line = (char *)malloc((size_t)L + 2);
    while ((l = fa_read(r)) >= 0) {
        for (long q = 0; q < L; q++) line[q] = (char)toupper((unsigned char)r->seq[idx[q]]);
        line[L] = '\n';
        fputc('>', out); fwrite(r->name, 1, r->name_l, out); fputc('\n', out);
        fwrite(line, 1, (size_t)L + 1, out);
    }
    free(line); free(idx); fa_close(r);
    if (out_path) fclose(out); else fflush(out);
    return 0;
}

#ifndef ORACLE_NO_MAIN
int main(int argc, char **argv) {
    const char *in = NULL, *out = NULL;
    int pure = 0;
    for (int i = 1; i < argc; i++) {
        if (!strcmp(argv[i], "-c")) pure = 1;
        else if (!strcmp(argv[i], "-o") && i + 1 < argc) out = argv[++i];
        else if (!strcmp(argv[i], "-m")) {}
        else if (!strcmp(argv[i], "-V")) { printf("snp-sites 2.5.1\n"); return 0; }
        else if (argv[i][0] == '-' && argv[i][1]) { fprintf(stderr, "oracle snp-sites: unsupported flag %s\n", argv[i]); return 1; }
        else in = argv[i];
    }
    if (!in) { fprintf(stderr, "Usage: snp-sites [-c] [-o output_filename] <file>\n"); return 1; }
    return oracle_snp_sites_file(in, out, pure, NULL, NULL, NULL);
}
#endif
