// Host-side file framing for libbtbphylo.so: FASTA records in, CSV / FASTA text out.
// No sequence arithmetic happens here -- only parsing and printing (SURVEY.md App. A.1, A.3 output).
#pragma once
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

namespace btb {

// Whole input file in memory (mmap for plain files, inflated through zlib for gzip).
struct FileBytes {
    const uint8_t *data = nullptr;
    size_t size = 0;
    void *map_base = nullptr;       // munmap target, or nullptr when `owned` holds the bytes
    size_t map_size = 0;
    std::vector<uint8_t> owned;
    ~FileBytes();
    FileBytes() = default;
    FileBytes(const FileBytes &) = delete;
    FileBytes &operator=(const FileBytes &) = delete;
};
int load_file(const char *path, FileBytes &fb);   // BTB_EIO on failure

// One FASTA record as it lies in the file.
struct FastaRecord {
    std::string name;        // header text after '>' up to the first whitespace
    size_t seq_off = 0;      // offset of the first byte after the header line
    size_t raw_len = 0;      // bytes from seq_off to the next record marker (or EOF)
    int64_t seq_len = -1;    // bases, when the layout is regular (else computed by the exact walk)
    int32_t line_w = 0;      // bases per line (0 = single line)
    int32_t eol = 1;         // line terminator bytes (1 = "\n", 2 = "\r\n")
    bool regular = false;    // every line but the last has line_w bases followed by eol bytes
    int32_t file = 0;        // predicted framing: index of the SourceFile the offsets refer to
};

// Finds the records (kseq framing: a record starts at '>' or '@' that is the first byte of a line,
// the first one at the first marker anywhere).  Fills name / seq_off / raw_len and, by looking at
// the first line and the region length only, the tentative regular layout.
int index_fasta(const FileBytes &fb, std::vector<FastaRecord> &recs, int threads);

// ---- predicted framing (the fast ingest path of btb_snp_sites) ----------------------------------
// Consensus FASTA files are N records of one layout: same genome length, same line width, same
// terminators.  Instead of scanning every byte for record markers, frame the first record exactly
// and PREDICT the others: record i+1's header must start where record i's region ends, its region
// has the same byte length.  The host checks the header byte and the terminator at each predicted
// boundary; everything in between is checked by the pack kernel, which reads every byte anyway
// (flag 1: a terminator is missing or a control byte sits among the bases, flag 2: a line starts
// with '>', '@' or '+').  Any failed check sends the whole call through the exact path
// (load_file + index_fasta), so results never depend on the prediction.
struct SourceFile {
    std::string path;
    int fd = -1;
    size_t size = 0;
    SourceFile() = default;
    SourceFile(const SourceFile &) = delete;
    SourceFile &operator=(const SourceFile &) = delete;
    SourceFile(SourceFile &&o) noexcept : path(std::move(o.path)), fd(o.fd), size(o.size) { o.fd = -1; }
    ~SourceFile();
};
int open_sources(const char *const *paths, int64_t n_paths, std::vector<SourceFile> &files);   // BTB_EIO on failure
// *fast = false: not predictable (gzip, FASTQ, irregular or mixed layouts, ...) -- use the exact path.
int frame_predicted(const std::vector<SourceFile> &files, std::vector<FastaRecord> &recs, bool *fast);
// pread that loops until len bytes are in (BTB_EIO on a short file)
int read_exact(const SourceFile &f, size_t off, size_t len, uint8_t *dst);

// Exact kseq-style walk of one record's sequence region: appends the kept bytes to `out`
// (or only counts them when out == nullptr).  Returns the sequence length.
int64_t walk_record(const FileBytes &fb, const FastaRecord &r, uint8_t *out);

// Wall-clock stage report for the level-1 calls: BTB_TIMING=1 prints one line per stage on stderr.
struct StageTimer {
    const char *what;
    bool on;
    double t0, last;
    std::vector<std::pair<std::string, double>> marks;
    explicit StageTimer(const char *what);
    void mark(const char *stage);
    ~StageTimer();
};

// Output files are written under a unique temporary name in the target's directory (mkstemp) and moved
// over the target with one rename(): the replace is atomic, a failed run leaves the old file alone, and
// two runs writing the same path do not collide.  open_output_temp returns the descriptor (mode 0666 &
// ~umask) and the temporary path; commit_output renames (with `unlink_first` the target is removed first --
// "replace_unlink_first": ext4 flushes a file that replaces another one by rename synchronously, which
// costs seconds for a 12 GB snps.csv; off by default because it is not atomic); discard_output removes it.
int open_output_temp(const char *final_path, std::string &tmp_path, int *fd);
int commit_output(const std::string &tmp_path, const char *final_path, bool unlink_first);
void discard_output(const std::string &tmp_path);
// true when `path`'s directory is a tmpfs / ramfs (page cache only: mapped writes scale with threads there)
bool on_memory_fs(const char *path);

// ViewBovine label of one sample name (ASCII restatement of btbphylo/phylogeny.py:206-216).
std::string process_sample_name(const std::string &name);
// snps.csv helpers behind btb_csv_row_names / btb_relabel_csv.
int csv_row_names(const FileBytes &fb, char sep, std::vector<std::string> &names);
int relabel_csv(const FileBytes &fb, char sep, const std::vector<std::string> &labels, FILE *fo);

// Fast decimal formatting of one matrix row segment; returns bytes written.
size_t format_row(const void *row, int elem_bytes, int64_t n, char sep, char *dst);
// Exactly the number of bytes format_row writes for this row (no formatting done).
size_t row_text_len(const void *row, int elem_bytes, int64_t n);

}  // namespace btb
