/*
 * oracle/snp_dists_ref.c -- TEST INFRASTRUCTURE ONLY (CPU oracle). Never linked into the product.
 *
 * CPU restatement of the all-vs-all SNP distance matrix that btb-phylo obtains by running
 *     snp-dists -c -j <threads> <snps.fas> > <snps.csv>   (btbphylo/phylogeny.py:144-150, cmd at :149)
 * The algorithm lives in the third-party program snp-dists (tseemann/snp-dists, built from
 * un-pinned git HEAD by install/install-snp-dists.sh:9-11; the version the reference's authors
 * observed and BASELINE.json pins is 0.8.2, Readme.md:64).  Its source is absent from
 * /root/reference; this file restates the published behaviour from SURVEY.md section 8a row a2
 * and Appendix A.3 -- it is not a copy of upstream code.
 * PARITY UNPINNED: no real snp-dists binary or reference golden vector exists offline; the
 * oracle is pinned by the upstream README example (KAT-1), the three stderr banner lines the
 * reference quotes (Readme.md:64-66) and hand-derived vectors in tests/golden/.
 *
 * Rules (Appendix A.3):
 *   load: toupper() every byte unless -k; unless -a, every byte not in {A,C,G,T} becomes '.';
 *   d(a,b) = #{k : a[k] != b[k]  and  a[k] != '.'  and  b[k] != '.'};
 *   full square is computed row by row, the inner (column) loop is the OpenMP parallel-for (-j);
 *   output: corner "snp-dists 0.8.2" (blank with -b), SEP name..., then rows name SEP d...;
 *   -m molten output; -c comma instead of TAB; -q no banners.
 *   errors (exit 1): unequal length, > 100000 sequences, no sequences, unopenable file.
 *
 * Build:  make -C oracle   (gcc -O3 -std=c99 -fopenmp, the upstream Makefile's flags; links zlib)
 * Library entry points (ctypes; tests and bench.py's cpu_baseline only):
 *   oracle_snp_dists_rows()  -- rows [row0,row1) of the matrix for in-memory sequences;
 *   oracle_snp_dists_prepare() / _rows_prepared() / _release() -- the same split into snp-dists' load step
 *   (once) and its loop nest (what bench.py's CPU arm times).
 */
#include <ctype.h>
#include <stdint.h>
#include <unistd.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "fasta_ref.h"

#define REF_VERSION "0.8.2"
#define REF_EXENAME "snp-dists"
#define REF_MAX_SEQ 100000
#define REF_IGNORE '.'

static size_t ref_distance(const char *restrict a, const char *restrict b, size_t L) {
    size_t diff = 0;
    for (size_t i = 0; i < L; i++)
        if (a[i] != b[i] && a[i] != REF_IGNORE && b[i] != REF_IGNORE) diff++;
    return diff;
}

static void ref_prepare(char *s, size_t L, int allchars, int keepcase) {
    if (!keepcase) for (size_t i = 0; i < L; i++) s[i] = (char)toupper((unsigned char)s[i]);
    if (!allchars)
        for (size_t i = 0; i < L; i++)
            if (s[i] != 'A' && s[i] != 'C' && s[i] != 'G' && s[i] != 'T') s[i] = REF_IGNORE;
}

/* In-memory forms.  snp-dists loads and prepares every sequence ONCE (main() below does the same), then
 * computes the square row by row: oracle_snp_dists_prepare() is that load step, and
 * oracle_snp_dists_rows_prepared() the loop nest -- outer row j serial, inner i the OpenMP parallel-for --
 * writing rows [row0,row1) x n as uint32 into out (row-major).  bench.py times the second only. */
char *oracle_snp_dists_prepare(const char *seqs, long n, long L, int allchars, int keepcase) {
    char *p = (char *)malloc((size_t)n * (size_t)L + 1);
    if (!p) return NULL;
#pragma omp parallel for
    for (long i = 0; i < n; i++) {
        memcpy(p + (size_t)i * L, seqs + (size_t)i * L, (size_t)L);
        ref_prepare(p + (size_t)i * L, (size_t)L, allchars, keepcase);
    }
    return p;
}

void oracle_snp_dists_release(char *prepared) { free(prepared); }

int oracle_snp_dists_rows_prepared(const char *p, long n, long L, int threads, long row0, long row1, uint32_t *out) {
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#endif
    for (long j = row0; j < row1; j++) {
        uint32_t *d = out + (size_t)(j - row0) * n;
#pragma omp parallel for
        for (long i = 0; i < n; i++)
            d[i] = (uint32_t)ref_distance(p + (size_t)j * L, p + (size_t)i * L, (size_t)L);
    }
    return 0;
}

/* prepare + rows + release in one call (tests) */
int oracle_snp_dists_rows(const char *seqs, long n, long L, int allchars, int keepcase,
                          int threads, long row0, long row1, uint32_t *out) {
    char *p = oracle_snp_dists_prepare(seqs, n, L, allchars, keepcase);
    if (!p) return 2;
    int rc = oracle_snp_dists_rows_prepared(p, n, L, threads, row0, row1, out);
    free(p);
    return rc;
}

#ifndef ORACLE_NO_MAIN
static void usage(void) {
    fprintf(stderr, "SYNOPSIS\n  Pairwise SNP distance matrix from a FASTA alignment\n"
                    "USAGE\n  %s [opts] aligned.fasta[.gz] > matrix.tsv\n"
                    "OPTIONS\n  -h -v -j N -q -a -k -m -c -b\n", REF_EXENAME);
}

int main(int argc, char **argv) {
    int opt, quiet = 0, csv = 0, corner = 1, allchars = 0, keepcase = 0, molten = 0, threads = 1;
    while ((opt = getopt(argc, argv, "hj:qcakmbv")) != -1) {
        switch (opt) {
        case 'h': usage(); return 0;
        case 'j': threads = atoi(optarg); break;
        case 'q': quiet = 1; break;
        case 'c': csv = 1; break;
        case 'a': allchars = 1; break;
        case 'k': keepcase = 1; break;
        case 'm': molten = 1; break;
        case 'b': corner = 0; break;
        case 'v': printf("%s %s\n", REF_EXENAME, REF_VERSION); return 0;
        default: usage(); return 1;
        }
    }
    if (optind >= argc) { usage(); return 1; }
    const char *path = argv[optind];
    if (threads < 1) threads = 1;
    if (!quiet) {
        fprintf(stderr, "This is %s %s\n", REF_EXENAME, REF_VERSION);
        fprintf(stderr, "Will use %d threads.\n", threads);
    }
    fa_reader *r = fa_open(path);
    if (!r) { fprintf(stderr, "ERROR: Could not open filename '%s'\n", path); return 1; }
    char **seq = (char **)calloc(REF_MAX_SEQ, sizeof(char *));
    char **name = (char **)calloc(REF_MAX_SEQ, sizeof(char *));
    long l, L = -1, N = 0;
    while ((l = fa_read(r)) >= 0) {
        if (N >= REF_MAX_SEQ) {
            fprintf(stderr, "ERROR: %s can only handle %d sequences at most. Please change MAX_SEQ and recompile.\n",
                    REF_EXENAME, REF_MAX_SEQ);
            return 1;
        }
        if (L < 0) L = l;
        if (l != L) {
            fprintf(stderr, "ERROR: sequence #%ld '%s' has length %ld but expected %ld\n", N + 1, r->name, l, L);
            return 1;
        }
        seq[N] = (char *)malloc((size_t)L + 1);
        memcpy(seq[N], r->seq, (size_t)L); seq[N][L] = 0;
        name[N] = strdup(r->name);
        ref_prepare(seq[N], (size_t)L, allchars, keepcase);
        N++;
    }
    fa_close(r);
    if (N < 1) { fprintf(stderr, "ERROR: file contained no sequences\n"); return 1; }
    if (!quiet) fprintf(stderr, "Read %ld sequences of length %ld\n", N, L);
#ifdef _OPENMP
    omp_set_num_threads(threads);
#endif
    char sep = csv ? ',' : '\t';
    size_t *d = (size_t *)malloc(sizeof(size_t) * (size_t)N);
    if (!molten) {
        if (corner) printf("%s %s", REF_EXENAME, REF_VERSION);
        for (long j = 0; j < N; j++) printf("%c%s", sep, name[j]);
        printf("\n");
    }
    for (long j = 0; j < N; j++) {
#pragma omp parallel for
        for (long i = 0; i < N; i++) d[i] = ref_distance(seq[j], seq[i], (size_t)L);
        if (molten) {
            for (long i = 0; i < N; i++) printf("%s%c%s%c%zu\n", name[j], sep, name[i], sep, d[i]);
        } else {
            printf("%s", name[j]);
            for (long i = 0; i < N; i++) printf("%c%zu", sep, d[i]);
            printf("\n");
        }
    }
    fflush(stdout);
    return 0;
}
#endif
