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
};

// Finds the records (kseq framing: a record starts at '>' or '@' that is the first byte of a line,
// the first one at the first marker anywhere).  Fills name / seq_off / raw_len and, by looking at
// the first line and the region length only, the tentative regular layout.
int index_fasta(const FileBytes &fb, std::vector<FastaRecord> &recs, int threads);

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

// ViewBovine label of one sample name (ASCII restatement of btbphylo/phylogeny.py:206-216).
std::string process_sample_name(const std::string &name);
// snps.csv helpers behind btb_csv_row_names / btb_relabel_csv.
int csv_row_names(const FileBytes &fb, char sep, std::vector<std::string> &names);
int relabel_csv(const FileBytes &fb, char sep, const std::vector<std::string> &labels, FILE *fo);

// Fast decimal formatting of one matrix row segment; returns bytes written.
size_t format_row(const void *row, int elem_bytes, int64_t n, char sep, char *dst);

}  // namespace btb
