/*
 * btb_phylo.h -- C-ABI of the B200-native phylogeny hot path (libbtbphylo.so).
 *
 * The reference (APHA-CSU/btb-phylo) has no FFI: its phylogeny stage is two subprocess calls,
 *     snp-sites <multi_fasta> -c -o <snps.fas>           btbphylo/phylogeny.py:133
 *     snp-dists -c -j <T> <snps.fas> > <snps.csv>        btbphylo/phylogeny.py:149
 * wrapped by snp_sites() (phylogeny.py:128-141) and build_snp_matrix() (phylogeny.py:144-150).
 * The entry points below are what a binding for that path attaches to.  Level 1 replaces the two
 * command lines one for one; level 2 is the same work on host buffers (no files); level 3 is
 * the individual GPU stages on device pointers, used by the tests and by bench.py.
 *
 * Conventions
 *   - plain C, no exceptions cross the boundary; every function returns 0 on success or a
 *     BTB_E* code, and btb_last_error() gives the text (thread-local);
 *   - paths are NUL-terminated; buffers are caller-allocated; sizes are element counts unless
 *     the name says bytes; `stream` is a cudaStream_t passed as void* (NULL = default stream);
 *   - `device` is a CUDA ordinal; level 1/2 calls pick device 0.. themselves from `n_gpus`;
 *   - there is no CPU fallback: without a usable sm_100 device every compute call fails with
 *     BTB_ENODEVICE.
 */
#ifndef BTB_PHYLO_H
#define BTB_PHYLO_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- error codes ---- */
#define BTB_OK          0
#define BTB_EINVAL      1   /* bad argument */
#define BTB_EIO         2   /* cannot open / read / write a file (snp-dists "Could not open filename") */
#define BTB_EFORMAT     3   /* unequal sequence lengths, no sequences, > BTB_MAX_SEQ sequences */
#define BTB_ENOSNPS     4   /* snp-sites: "No SNPs were detected" (upstream exits 1, no output file) */
#define BTB_ECUDA       5   /* a CUDA runtime/driver call failed */
#define BTB_ENODEVICE   6   /* no sm_100 device visible */
#define BTB_ENOMEM      7
#define BTB_ECOMM       8   /* multi-rank exchange: a peer failed, timed out, or has no peer-to-peer path */

#define BTB_MAX_SEQ     100000      /* snp-dists 0.8.2 MAX_SEQ */

/* distance engines: both are exact integer kernels for sm_100a */
#define BTB_ENGINE_AUTO (-1)
#define BTB_ENGINE_POPC   0   /* warp-tiled XOR/AND/POPC over bit-planes */
#define BTB_ENGINE_UMMA   1   /* tcgen05 GEMM: kind::mxf4 on 0/1 nibble codes (kind::i8 for very long rows), accumulators in TMEM */

/* `sigma` (number of symbol planes) takes this value when the alignment is "dense": at most four
 * distinct symbols and no byte that is ignored anywhere (always true for what snp-sites -c writes).
 * The engines then use their compact encodings: UMMA a single operand of 3 elements per site whose
 * symbols are the 0/1 vertices of a regular tetrahedron (d = c_i + c_j - dot), POPC two code planes
 * without a validity plane. */
#define BTB_SIGMA_DENSE4 (-4)

/* side length of one scheduler tile of the distance matrix (rows == columns) */
#define BTB_TILE        256

int         btb_last_error(char *buf, size_t cap);       /* copies the message, returns its length */
const char *btb_version(void);
int         btb_device_count(int *count);                 /* number of usable sm_100 devices */
/* Tuning knobs for tests, benchmarks and profilers (defaults in brackets).  Results never depend on them.
 *   distance kernels (tcgen05):
 *     "umma_fp4" [1]           kind::mxf4 kernels on packed e2m1 nibbles (0: the kind::i8 kernels)
 *     "umma_fp4_general" [1]   validity + one-hot nibble planes for alphabets with ignored bytes / -a (0: int8 A x B)
 *     "umma_fp4_pair" [1]      cta_group::2 CTA-pair kernel on 512 x 240 units (0: single-CTA 256 x 240 units)
 *     "umma_major_zero" [1]    per site, the most frequent symbol takes the all-zero vertex of the dense code
 *     "umma_tile_mode" [1]     full-tile kernels when a launch has >= 2 units per SM (0: half-tile units only)
 *     "umma_dense_simplex" [2] dense int8 code: 0 one-hot, 1 +-1 simplex, 2 0/1 tetrahedron
 *     "umma_lockstep" [1]      0 off, k >= 1: TMA producers of all CTAs start a unit at most k-1 rounds apart
 *     "umma_coop" [1]          0: CTA-pair kernel without the cooperative launch attribute (set automatically when
 *                              a profiler is attached: ncu cannot replay a cooperative cluster launch)
 *   extraction: "pack_wpt" [2] plane words per thread of k1 (1, 2, 4); "pack_gy" [0 = automatic] sample groups per launch;
 *               "ingest_predict" [1] predicted FASTA framing
 *   level 1/2:  "dense_pageable" [1] large snps.fas into pageable host memory; "host_stream" [0] +
 *               "host_stream_rows" [64] streamed level-2 call;
 *               "comm_band_rows" [16] block rows per launch + copy band of a rank's share (btb_dist_matrix_host_comm);
 *               "csv_mmap" [-1] snps.csv through a file mapping (-1: on tmpfs only); "csv_prealloc" [1]; "replace_unlink_first" [0];
 *               "multi_gpu_oversubscribe" [0] (tests) several workers per device when n_gpus exceeds the devices
 * btb_query also answers "ingest_last_predicted", "host_stream_last", "comm_last_path", "umma_max_active_clusters_cg1|2" and
 * "probe_cg<1|2>_e<k>" (tensor-pipe issue probe: cycles x 100 per stage of 8 MMAs 128 x 240 x 64 per SM, a commit every
 * k stages -- the measured peak bench.py's roofline divides by).
 * The environment variable BTB_OPTIONS="key=value,..." sets options for the Python binding. */
int         btb_set_option(const char *key, int value);
/* introspection: the option values, "umma_max_active_clusters_cg1" / "_cg2" (resident CTAs / CTA pairs) */
int         btb_query(const char *key, int64_t *value);

/* =====================================================================================
 * Level 1 -- file to file.  Replaces the two command lines.
 * ===================================================================================== */

/* snp-sites <in_fasta> [-c] -o <out_fasta>   (phylogeny.py:133).
 * pure_mode = 1 is `-c` (the only mode btb-phylo uses).  Writes one unwrapped record per
 * sample (phylogeny.py:136-140 and :158-160 rely on that).  BTB_ENOSNPS / BTB_EFORMAT mirror the
 * upstream exit-1 cases; nothing is written then.  host_threads is the reference's -j. */
int btb_snp_sites(const char *in_fasta, const char *out_fasta, int pure_mode,
                  int n_gpus, int host_threads,
                  int64_t *n_samples, int64_t *genome_len, int64_t *n_snps);

/* The same stage over SEVERAL input files taken as if concatenated in the given order -- what
 * btbphylo/phylogeny.py:48-88 build_multi_fasta produces by appending every sample's consensus
 * file to multi_fasta.fas before snp-sites reads it back: here the per-sample
 * files are staged straight into the pinned upload buffers and the multi-FASTA never has to be
 * written.  alignment_name (may be NULL) is the file name shown in the unequal-length message. */
int btb_snp_sites_files(const char *const *in_fastas, int64_t n_files, const char *alignment_name,
                        const char *out_fasta, int pure_mode, int n_gpus, int host_threads,
                        int64_t *n_samples, int64_t *genome_len, int64_t *n_snps);

/* Framing of a list of FASTA files taken as if concatenated, by prediction (equal-layout records: what consensus
 * files are), for callers that stage the record bytes themselves -- the plane store of
 * btb_phylo_b200/plane_store.py.  out[6 r + 0..5] = file index, offset of the first sequence byte
 * in that file, bytes of the sequence region, bases, bases per line (0 = one line), terminator bytes; names = the
 * record names joined by '\n'.  *n_records = -1 when the files are not predictable (gzip, FASTQ, mixed or
 * irregular layouts): go through btb_snp_sites_files then, which parses anything. */
int btb_fasta_frame_files(const char *const *in_fastas, int64_t n_files, int64_t *out, int64_t cap_records,
                          char *names, size_t names_cap, int64_t *n_records);

/* snp-dists [-a] [-k] [-m] [-c] [-b] [-q] -j T <in_fasta> > <out_path>   (phylogeny.py:149).
 * sep is ',' for -c else '\t'; corner = 0 is -b; the three banner lines go to stderr unless
 * quiet.  The file holds the full square, byte-identical with snp-dists 0.8.2. */
int btb_snp_dists(const char *in_fasta, const char *out_path,
                  int allchars, int keepcase, int molten, char sep, int corner, int quiet,
                  int n_gpus, int host_threads, int engine,
                  int64_t *n_samples, int64_t *seq_len);

/* =====================================================================================
 * Level 2 -- host buffers in, host buffers out (copies are inside the call).
 * ===================================================================================== */

/* Full n x n distance matrix of n sequences of `len` bytes (row i at seqs + i*row_stride).
 * out is row-major with out_ld elements per row, out_elem_bytes = 2 (uint16, needs len < 65536)
 * or 4 (uint32).  seqs/out may be pageable or pinned host memory.  Ranks of a multi-process
 * job pass (rank, world) to compute only their share (btb_dist_row_share block rows: the cells
 * (r, c >= r) of those rows and their mirrors); other cells are left untouched.
 * (world = 1: the whole matrix, both triangles.) */
int btb_dist_matrix_host(const uint8_t *seqs, int64_t n, int64_t len, int64_t row_stride,
                         int allchars, int keepcase, int engine, int device,
                         int rank, int world,
                         void *out, int out_elem_bytes, int64_t out_ld);

/* The same, upper triangle only: the cells (r, c >= r) are written (every unique pair once, the diagonal
 * zero), cells below the diagonal are left as they were or unspecified.  Half the device-to-host bytes --
 * the call is bound by that copy -- for consumers that take a symmetric matrix as its upper triangle.
 * (Measured: mirroring on the host instead, 16 threads transposing 256 x 256 tiles, is slower than copying
 * both triangles over PCIe, so the full matrix of btb_dist_matrix_host is mirrored on the device.) */
int btb_dist_matrix_host_upper(const uint8_t *seqs, int64_t n, int64_t len, int64_t row_stride,
                               int allchars, int keepcase, int engine, int device,
                               int rank, int world,
                               void *out, int out_elem_bytes, int64_t out_ld);

/* btb_dist_matrix_host keeps its device buffers between calls; this frees them. */
int btb_release_workspace(void);

/* ---- the same for the ranks of a one-node job, one process per GPU ----
 * btb_dist_matrix_host(rank, world) makes every rank upload and encode the whole alignment.  With a
 * btb_comm the ranks share that work: rank r uploads and encodes samples [n r / P, n (r+1) / P) only, the
 * encoded operand shards are exchanged GPU to GPU (CUDA IPC peer copies over NVLink, issued by the rank
 * that needs the rows; no NCCL, no host staging), and each rank computes its btb_dist_row_share block rows
 * and copies them into `out` band by band while later bands are computed.  `out` is normally ONE n x n
 * matrix shared by all ranks (a shared-memory segment every rank has mapped and page-locked): the ranks
 * write disjoint cells, and after the last rank returns it holds the whole matrix, both triangles.
 *
 * btb_comm_create: collective over the `world` ranks of the job.  `name` is a POSIX shared-memory name
 * unique to the job (rank 0 creates "/name", the others attach; it is unlinked again once all have).
 * A comm on which any call failed is dead (every later call returns BTB_ECOMM): destroy it.
 * Alphabets / engines without an exchange path (POPC bit-planes, int8 operands of very long rows) fall
 * back to the replicated upload; btb_query("comm_last_path") tells which path the last call took. */
typedef struct btb_comm btb_comm;
int btb_comm_create(const char *name, int rank, int world, int device, btb_comm **comm);
int btb_comm_destroy(btb_comm *comm);
int btb_dist_matrix_host_comm(btb_comm *comm, const uint8_t *seqs, int64_t n, int64_t len, int64_t row_stride,
                              int allchars, int keepcase, int engine,
                              void *out, int out_elem_bytes, int64_t out_ld);
/* wall-clock seconds of the phases of the last call on this rank: upload + presence scan, barrier A, buffers +
 * site counts, barrier B, encode, wait for peers + NVLink copies, distance kernels + D2H (7 values) */
int btb_comm_last_timings(const btb_comm *comm, double *seconds, int cap);

/* Variable columns of an n x G alignment held as dense rows: keep[k] = 1 where snp-sites keeps
 * column k; returns their count in *n_snps.  snps_out (may be NULL) receives n rows of *n_snps
 * upper-cased bytes, row stride snps_ld (>= number kept; call once with NULL to size it). */
int btb_extract_host(const uint8_t *seqs, int64_t n, int64_t G, int64_t row_stride,
                     int pure_mode, int device,
                     uint8_t *keep, int64_t *n_snps, uint8_t *snps_out, int64_t snps_ld);

/* =====================================================================================
 * Level 3 -- device pointers, asynchronous on `stream`.  (tests, bench.py, multi-rank drivers)
 * ===================================================================================== */

/* ---- extraction: k1 pack, k2 column mask, k3 compaction ---- */

/* words of 32 columns per sample and plane, padded for 128-bit access */
int64_t btb_plane_words(int64_t G);

/* k1: text -> bit-planes.  d_text holds text_bytes bytes of FASTA text (allocation padded to a
 * 16-byte boundary).  Sample i's bases start at d_text + rec_off[i]; its lines are line_w[i] bases
 * followed by eol[i] line-terminator bytes (line_w = 0: unwrapped; otherwise >= 32).  Writes three
 * planes per sample (code bit 0, code bit 1, ACGT validity; code = (byte >> 1) & 3), plane p of
 * sample i at d_planes + ((p * n_total + first + i) * words); column c = 32 w + cl is bit cl of
 * word w.  d_flags[i] is
 * OR-ed with 1 when a line terminator is not where the layout says or a control byte sits among the
 * bases (the caller then re-stages that record unwrapped), and with 2 when a line of the region starts
 * with '>', '@' or '+' (the region is then not one record: the caller must frame the file again). */
int btb_pack_planes(const uint8_t *d_text, int64_t text_bytes, const int64_t *d_rec_off,
                    const int32_t *d_line_w, const int32_t *d_eol, int64_t n_chunk, int64_t first,
                    int64_t n_total, int64_t G, uint32_t *d_planes, int32_t *d_flags, void *stream);

/* k1 with k2 fused in: as btb_pack_planes, and the column state of these samples (see btb_column_mask) is
 * OR-ed into d_colstate (5 x btb_plane_words(G) words, zeroed by the caller before the first chunk) in the
 * same pass -- the planes are never read back for it.  Launches that share a d_colstate must be issued on one
 * stream.  This is what btb_snp_sites and btb_extract_host use. */
int btb_pack_planes_state(const uint8_t *d_text, int64_t text_bytes, const int64_t *d_rec_off,
                          const int32_t *d_line_w, const int32_t *d_eol, int64_t n_chunk, int64_t first,
                          int64_t n_total, int64_t G, uint32_t *d_planes, int32_t *d_flags,
                          uint32_t *d_colstate, void *stream);

/* k2: per-column reduction over samples [0, n) -> 5 words per 32 columns in d_colstate:
 * seenA, seenC, seenG, seenT, anyInvalid.  accumulate != 0 ORs into existing state. */
int btb_column_mask(const uint32_t *d_planes, int64_t n, int64_t n_total, int64_t G,
                    uint32_t *d_colstate, int accumulate, void *stream);

/* k3a: column state -> keep bitmask (G bits, plain order: bit c & 31 of word c >> 5) + ascending kept-column indices; *d_count gets L.
 * d_idx must hold G entries (worst case).  pure_mode as in btb_snp_sites. */
int btb_select_columns(const uint32_t *d_colstate, int64_t G, int pure_mode,
                       uint32_t *d_keep, uint32_t *d_idx, uint32_t *d_count, void *stream);

/* k3b: gather the kept columns of samples [first, first + n_chunk) as upper-case ASCII:
 * d_snps[i * snps_ld + l] for l < L. */
int btb_compact_columns(const uint32_t *d_planes, int64_t first, int64_t n_chunk, int64_t n_total,
                        int64_t G, const uint32_t *d_idx, int64_t L,
                        uint8_t *d_snps, int64_t snps_ld, void *stream);

/* ---- distances: k4 ---- */

/* Symbol table of the alignment: table[b] = code 0..sigma-1, or 0xFF when byte b never counts
 * (everything but ACGT by default; only '.' with allchars).  With d_seqs the alignment is scanned
 * on the device (synchronises `stream`): allchars needs that to learn the alphabet, and either
 * mode reports *sigma = BTB_SIGMA_DENSE4 when every byte present counts and sigma <= 4.
 * Default mode with d_seqs == NULL skips the scan and returns sigma = 4. */
int btb_dist_alphabet(const uint8_t *d_seqs, int64_t n, int64_t len, int64_t row_stride,
                      int allchars, int keepcase, uint8_t table[256], int *sigma, void *stream);

/* bytes of operand storage the engine needs for this problem */
int64_t btb_dist_operand_bytes(int engine, int64_t n, int64_t len, int sigma);

/* alignment bytes -> engine operands (bit-planes for POPC; one-hot int8 A and masked-complement
 * int8 B for UMMA) */
int btb_dist_encode(int engine, const uint8_t *d_seqs, int64_t n, int64_t len, int64_t row_stride,
                    const uint8_t table[256], int sigma, void *d_operands, void *stream);

/* btb_dist_encode for the rows [row0, row1) only, written at their place in the operand of all n rows (the other
 * rows of d_operands are left as they are).  FP4 operand layouts; other layouts encode all rows. */
int btb_dist_encode_rows(int engine, const uint8_t *d_seqs, int64_t n, int64_t len, int64_t row_stride,
                         const uint8_t table[256], int sigma, void *d_operands, int64_t row0, int64_t row1,
                         void *stream);

/* number of BTB_TILE x BTB_TILE tiles in the upper triangle (diagonal included) */
int64_t btb_dist_tile_count(int64_t n);
/* tile index -> (block row I, block column J), I <= J; the order is L2-friendly bands */
int btb_dist_tile_coords(int64_t n, int64_t tile, int32_t *I, int32_t *J);

/* operands -> distances for tiles [tile_lo, tile_hi).
 * out_mode 0: d_out is the n x n matrix (out_ld elements per row); the tile and, when mirror != 0,
 *             its transpose are written.
 * out_mode 1: d_out is tile-packed: tile t at d_out + (t - tile_lo) * BTB_TILE * BTB_TILE elements,
 *             row-major inside the tile (cells beyond n are zero). */
int btb_distance_tiles(int engine, const void *d_operands, int64_t n, int64_t len, int sigma,
                       int64_t tile_lo, int64_t tile_hi,
                       void *d_out, int out_elem_bytes, int64_t out_ld, int out_mode, int mirror,
                       void *stream);

/* Multi-GPU unit: block rows.  btb_dist_block_rows = number of BTB_TILE-row blocks; btb_dist_row_share
 * gives rank `rank` of `world` a contiguous range of block rows holding (nearly) the same number of
 * upper-triangle tiles; btb_distance_rows computes every cell (r, c >= r) whose row r lies in the block
 * rows [row_block_lo, row_block_hi) into the n x n matrix d_out and, when mirror != 0, the transposed
 * cells as well (so the rows are complete once every earlier block row has been computed too). */
int64_t btb_dist_block_rows(int64_t n);
int btb_dist_row_share(int64_t n, int rank, int world, int64_t *row_block_lo, int64_t *row_block_hi);
/* btb_dist_row_share with the operand encode counted in: rank r only has to encode the rows from its first block
 * row on (btb_dist_encode_rows with row0 = max(0, lo * BTB_TILE - BTB_TILE)), which is (T - lo) / T of
 * encode_tiles, the cost of encoding all rows in units of one tile of distance work.  Equalises tiles + encode. */
int btb_dist_row_share_weighted(int64_t n, int rank, int world, double encode_tiles,
                                int64_t *row_block_lo, int64_t *row_block_hi);
int btb_distance_rows(int engine, const void *d_operands, int64_t n, int64_t len, int sigma,
                      int64_t row_block_lo, int64_t row_block_hi,
                      void *d_out, int out_elem_bytes, int64_t out_ld, int mirror, void *stream);

/* ---- serialisation: k5 ---- */

/* Formats rows [row0, row1) of a host matrix as snp-dists prints them (name SEP d SEP d ... \n)
 * into buf (cap bytes) using host_threads threads; *written gets the byte count. */
int btb_format_csv_rows(const void *mat, int elem_bytes, int64_t ld, int64_t n,
                        int64_t row0, int64_t row1, const char *const *names, char sep,
                        int host_threads, char *buf, size_t cap, size_t *written);

/* ---- sample labels: the step after snps.csv ----
 * btbphylo/phylogeny.py:172-187 post_process_snps_csv parses the O(N^2) snps.csv with pandas only to
 * rename its row and column labels (phylogeny.py:190-203) with process_sample_name
 * (phylogeny.py:206-216 + btbphylo/utils.py:110-119) and writes it back.  Here the labels are
 * printed by the CSV writer directly, or swapped in a finished file by a streaming pass that never
 * parses a number; both produce the bytes pandas' to_csv would. */

/* ASCII restatement of process_sample_name: strip "_consensus", "AF-" + first DD-DDDD[D]-DD match
 * (else the name), upper-cased.  out receives a NUL-terminated string. */
int btb_process_sample_name(const char *name, char *out, size_t cap);

/* Record names of a FASTA file / first field of every data row of a snps.csv, joined by '\n' into
 * buf (NUL-terminated).  With buf == NULL only *n and *needed (bytes incl. NUL) are set. */
int btb_fasta_names(const char *in_fasta, int host_threads, char *buf, size_t cap, int64_t *n, size_t *needed);
int btb_csv_row_names(const char *csv_path, char sep, char *buf, size_t cap, int64_t *n, size_t *needed);

/* btb_snp_dists printing `labels` (n_labels strings joined by '\n', one per sequence, in file order)
 * in the header row and as row names instead of the FASTA record names. */
int btb_snp_dists_labelled(const char *in_fasta, const char *out_path, int allchars, int keepcase, int molten,
                           char sep, int corner, int quiet, int n_gpus, int host_threads, int engine,
                           const char *labels, int64_t n_labels, int64_t *n_samples, int64_t *seq_len);

/* Rewrites in_csv to out_csv (may be the same path) with the header names and the row names
 * replaced by `labels`; the corner cell and every distance byte are copied through. */
int btb_relabel_csv(const char *in_csv, const char *out_csv, char sep, const char *labels, int64_t n_labels);

#ifdef __cplusplus
}
#endif
#endif /* BTB_PHYLO_H */
