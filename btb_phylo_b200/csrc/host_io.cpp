// Host-side file framing: see host_io.h.  Parsing and printing only.
#include "host_io.h"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <sys/statfs.h>
#include <unistd.h>
#include <zlib.h>

#include <time.h>

#include <algorithm>
#include <cerrno>
#include <cstdlib>
#include <cstring>
#include <thread>

#include "common.cuh"

namespace btb {

static double now_s() {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return double(ts.tv_sec) + 1e-9 * double(ts.tv_nsec);
}
StageTimer::StageTimer(const char *w) : what(w), on(getenv("BTB_TIMING") != nullptr), t0(now_s()), last(t0) {}
void StageTimer::mark(const char *stage) {
    if (!on) return;
    const double t = now_s();
    marks.emplace_back(stage, t - last);
    last = t;
}
StageTimer::~StageTimer() {
    if (!on) return;
    fprintf(stderr, "[btb timing] %s: total %.3f s\n", what, now_s() - t0);
    for (auto &m : marks) fprintf(stderr, "[btb timing]   %-28s %8.3f s\n", m.first.c_str(), m.second);
}

int open_output_temp(const char *final_path, std::string &tmp_path, int *fd) {
    std::string dir = ".", base = final_path;
    const size_t slash = base.rfind('/');
    if (slash != std::string::npos) { dir = slash == 0 ? "/" : base.substr(0, slash); base = base.substr(slash + 1); }
    tmp_path = dir + "/." + base + ".XXXXXX";
    std::vector<char> name(tmp_path.begin(), tmp_path.end());
    name.push_back(0);
    const int f = mkstemp(name.data());
    if (f < 0) return fail(BTB_EIO, "cannot create a temporary file next to '%s': %s", final_path, strerror(errno));
    tmp_path = name.data();
    const mode_t um = umask(0);
    umask(um);
    fchmod(f, 0666 & ~um);
    *fd = f;
    return BTB_OK;
}

int commit_output(const std::string &tmp_path, const char *final_path, bool unlink_first) {
    // The file that is being replaced stays open across the rename and is closed by a detached thread: giving
    // back the pages of a 12 GB snps.csv takes half a second (tmpfs and ext4 alike) and nobody has to wait for it.
    const int old_fd = open(final_path, O_RDONLY | O_CLOEXEC);
    if (unlink_first) unlink(final_path);
    if (rename(tmp_path.c_str(), final_path) != 0) {
        const int e = errno;
        if (old_fd >= 0) close(old_fd);
        unlink(tmp_path.c_str());
        return fail(BTB_EIO, "cannot rename to '%s': %s", final_path, strerror(e));
    }
    if (old_fd >= 0) std::thread([old_fd] { close(old_fd); }).detach();
    return BTB_OK;
}

void discard_output(const std::string &tmp_path) {
    if (!tmp_path.empty()) unlink(tmp_path.c_str());
}

bool on_memory_fs(const char *path) {
    std::string dir = path;
    const size_t slash = dir.rfind('/');
    dir = slash == std::string::npos ? "." : (slash == 0 ? "/" : dir.substr(0, slash));
    struct statfs sf;
    if (statfs(dir.c_str(), &sf) != 0) return false;
    return sf.f_type == 0x01021994 /* TMPFS_MAGIC */ || sf.f_type == 0x858458f6 /* RAMFS_MAGIC */;
}

FileBytes::~FileBytes() {
    if (map_base) munmap(map_base, map_size);
}

int load_file(const char *path, FileBytes &fb) {
    int fd = open(path, O_RDONLY);
    if (fd < 0) return fail(BTB_EIO, "ERROR: Could not open filename '%s'", path);
    struct stat st;
    if (fstat(fd, &st) != 0 || !S_ISREG(st.st_mode)) {
        close(fd);
        return fail(BTB_EIO, "ERROR: Could not open filename '%s'", path);
    }
    unsigned char magic[2] = {0, 0};
    ssize_t got = pread(fd, magic, 2, 0);
    if (got == 2 && magic[0] == 0x1f && magic[1] == 0x8b) {       // gzip: inflate through zlib
        close(fd);
        gzFile gz = gzopen(path, "rb");
        if (!gz) return fail(BTB_EIO, "ERROR: Could not open filename '%s'", path);
        gzbuffer(gz, 1 << 20);
        size_t cap = size_t(st.st_size) * 4 + (1 << 20), len = 0;
        fb.owned.resize(cap);
        for (;;) {
            if (len == cap) { cap *= 2; fb.owned.resize(cap); }
            int r = gzread(gz, fb.owned.data() + len, unsigned(std::min<size_t>(cap - len, 1u << 30)));
            if (r < 0) { gzclose(gz); return fail(BTB_EIO, "ERROR: cannot inflate '%s'", path); }
            if (r == 0) break;
            len += size_t(r);
        }
        gzclose(gz);
        fb.owned.resize(len);
        fb.data = fb.owned.data();
        fb.size = len;
        return BTB_OK;
    }
    fb.size = size_t(st.st_size);
    if (fb.size) {
        void *m = mmap(nullptr, fb.size, PROT_READ, MAP_PRIVATE, fd, 0);
        if (m == MAP_FAILED) { close(fd); return fail(BTB_EIO, "ERROR: cannot map '%s'", path); }
        madvise(m, fb.size, MADV_SEQUENTIAL);
        fb.map_base = m;
        fb.map_size = fb.size;
        fb.data = static_cast<const uint8_t *>(m);
    }
    close(fd);
    return BTB_OK;
}

static inline bool is_space(uint8_t c) { return c == ' ' || (c >= '\t' && c <= '\r'); }

// positions in [lo, hi) holding `ch` as the first byte of a line
static void line_start_hits(const uint8_t *d, size_t lo, size_t hi, uint8_t ch, std::vector<size_t> &out) {
    size_t p = lo;
    while (p < hi) {
        const void *q = memchr(d + p, ch, hi - p);
        if (!q) break;
        size_t at = size_t(static_cast<const uint8_t *>(q) - d);
        if (at == 0 || d[at - 1] == '\n') out.push_back(at);
        p = at + 1;
    }
}

static void describe_layout(const FileBytes &fb, FastaRecord &r) {
    const uint8_t *d = fb.data + r.seq_off;
    size_t R = r.raw_len;
    while (R > 0 && (d[R - 1] == '\n' || d[R - 1] == '\r')) --R;       // trailing terminators / empty lines
    r.regular = true;
    r.eol = 1;
    r.line_w = 0;
    if (R == 0) { r.seq_len = 0; return; }
    const void *nl = memchr(d, '\n', R);
    if (!nl) { r.seq_len = int64_t(R); return; }                        // single line
    size_t first = size_t(static_cast<const uint8_t *>(nl) - d);
    if (first > 0 && d[first - 1] == '\r') { r.eol = 2; --first; }
    if (first == 0 || first > size_t(INT32_MAX)) { r.regular = false; return; }
    r.line_w = int32_t(first);
    const size_t pitch = first + size_t(r.eol);
    const size_t X = R + size_t(r.eol);
    const size_t q = X / pitch, rem = X % pitch;
    if (rem == 0) r.seq_len = int64_t(q * first);
    else if (rem > size_t(r.eol)) r.seq_len = int64_t(q * first + rem - size_t(r.eol));
    else r.regular = false;
}

int64_t walk_record(const FileBytes &fb, const FastaRecord &r, uint8_t *out) {
    const uint8_t *d = fb.data + r.seq_off;
    const size_t R = r.raw_len;
    int64_t len = 0;
    size_t p = 0;
    while (p < R) {
        if (d[p] == '\n') { ++p; continue; }                            // empty line
        const void *nl = memchr(d + p, '\n', R - p);
        size_t e = nl ? size_t(static_cast<const uint8_t *>(nl) - d) : R;
        size_t n = e - p;
        // one trailing CR per line is dropped (whole-string test, as the parser the tools share does).  It is not
        // copied either: `out` may be a slot of exactly seq_len bytes next to another thread's slot.
        if (len + int64_t(n) > 1 && d[e - 1] == '\r') --n;
        if (out) memcpy(out + len, d + p, n);
        len += int64_t(n);
        p = e + 1;
    }
    return len;
}

int index_fasta(const FileBytes &fb, std::vector<FastaRecord> &recs, int threads) {
    recs.clear();
    const uint8_t *d = fb.data;
    const size_t N = fb.size;
    if (N == 0) return BTB_OK;
    // first record: first '>' or '@' anywhere
    const void *g = memchr(d, '>', N), *a = memchr(d, '@', N);
    if (!g && !a) return BTB_OK;
    size_t first = N;
    if (g) first = std::min(first, size_t(static_cast<const uint8_t *>(g) - d));
    if (a) first = std::min(first, size_t(static_cast<const uint8_t *>(a) - d));
    // later records: '>' / '@' at the start of a line.  A '+' line start would open a FASTQ quality
    // block; those files take the sequential walk below.
    threads = std::max(1, std::min(threads, 64));
    std::vector<std::vector<size_t>> part(size_t(threads) * 3);
    {
        std::vector<std::thread> pool;
        const size_t lo0 = first + 1, span = (N - lo0 + size_t(threads) - 1) / size_t(threads);
        for (int t = 0; t < threads; ++t) {
            const size_t lo = lo0 + size_t(t) * span, hi = std::min(N, lo + span);
            if (lo >= hi) continue;
            auto job = [&, t, lo, hi] {
                line_start_hits(d, lo, hi, '>', part[size_t(t) * 3 + 0]);
                line_start_hits(d, lo, hi, '@', part[size_t(t) * 3 + 1]);
                line_start_hits(d, lo, hi, '+', part[size_t(t) * 3 + 2]);
            };
            if (threads == 1) job(); else pool.emplace_back(job);
        }
        for (auto &th : pool) th.join();
    }
    std::vector<size_t> starts{first};
    bool fastq = false;
    for (int t = 0; t < threads; ++t) {
        for (int k = 0; k < 2; ++k)
            for (size_t at : part[size_t(t) * 3 + size_t(k)]) starts.push_back(at);
        if (!part[size_t(t) * 3 + 2].empty()) fastq = true;
    }
    std::sort(starts.begin(), starts.end());

    if (fastq) {
        // sequential exact walk: record := marker line, sequence lines up to a line starting with
        // '>', '@' or '+'; after '+': skip that line, then consume >= seq_len quality bytes.
        size_t p = first;
        while (p < N) {
            FastaRecord r;
            const void *nl = memchr(d + p, '\n', N - p);
            size_t he = nl ? size_t(static_cast<const uint8_t *>(nl) - d) : N;
            size_t ne = p + 1;
            while (ne < he && !is_space(d[ne])) ++ne;
            r.name.assign(reinterpret_cast<const char *>(d + p + 1), ne - p - 1);
            r.seq_off = std::min(N, he + 1);
            size_t q = r.seq_off;
            while (q < N) {
                if (d[q] == '>' || d[q] == '@' || d[q] == '+') break;
                const void *l2 = memchr(d + q, '\n', N - q);
                q = l2 ? size_t(static_cast<const uint8_t *>(l2) - d) + 1 : N;
            }
            r.raw_len = q - r.seq_off;
            r.regular = false;
            r.seq_len = walk_record(fb, r, nullptr);
            recs.push_back(std::move(r));
            if (q < N && d[q] == '+') {
                const void *l2 = memchr(d + q, '\n', N - q);
                q = l2 ? size_t(static_cast<const uint8_t *>(l2) - d) + 1 : N;
                int64_t need = recs.back().seq_len, have = 0;
                while (q < N && have < need) { if (d[q] != '\n' && d[q] != '\r') ++have; ++q; }
                if (have > 0) {
                    const void *l3 = memchr(d + q, '\n', N - q);
                    q = l3 ? size_t(static_cast<const uint8_t *>(l3) - d) + 1 : N;
                }
                while (q < N && d[q] != '>' && d[q] != '@') ++q;     // scan to the next marker
            }
            p = q;
        }
        return BTB_OK;
    }

    recs.resize(starts.size());
    auto fill = [&](size_t lo, size_t hi) {
        for (size_t i = lo; i < hi; ++i) {
            FastaRecord &r = recs[i];
            const size_t p = starts[i], end = i + 1 < starts.size() ? starts[i + 1] : N;
            const void *nl = memchr(d + p, '\n', end - p);
            size_t he = nl ? size_t(static_cast<const uint8_t *>(nl) - d) : end;
            size_t ne = p + 1;
            while (ne < he && !is_space(d[ne])) ++ne;
            r.name.assign(reinterpret_cast<const char *>(d + p + 1), ne - p - 1);
            r.seq_off = std::min(end, he + 1);
            r.raw_len = end - r.seq_off;
            describe_layout(fb, r);
        }
    };
    if (threads == 1 || starts.size() < 64) {
        fill(0, starts.size());
    } else {
        std::vector<std::thread> pool;
        const size_t span = (starts.size() + size_t(threads) - 1) / size_t(threads);
        for (int t = 0; t < threads; ++t) {
            size_t lo = size_t(t) * span, hi = std::min(starts.size(), lo + span);
            if (lo < hi) pool.emplace_back(fill, lo, hi);
        }
        for (auto &th : pool) th.join();
    }
    return BTB_OK;
}

// ------------------------------------------------------------------------------ predicted framing
SourceFile::~SourceFile() { if (fd >= 0) close(fd); }

int open_sources(const char *const *paths, int64_t n_paths, std::vector<SourceFile> &files) {
    files.clear();
    files.reserve(size_t(n_paths));
    for (int64_t i = 0; i < n_paths; ++i) {
        SourceFile f;
        f.path = paths[i];
        f.fd = open(paths[i], O_RDONLY);
        struct stat st;
        if (f.fd < 0 || fstat(f.fd, &st) != 0) return fail(BTB_EIO, "ERROR: Could not open filename '%s'", paths[i]);
        f.size = size_t(st.st_size);
        files.push_back(std::move(f));
    }
    return BTB_OK;
}

int read_exact(const SourceFile &f, size_t off, size_t len, uint8_t *dst) {
    while (len) {
        const ssize_t k = pread(f.fd, dst, len, off_t(off));
        if (k <= 0) return fail(BTB_EIO, "ERROR: short read from '%s'", f.path.c_str());
        dst += k; off += size_t(k); len -= size_t(k);
    }
    return BTB_OK;
}

int frame_predicted(const std::vector<SourceFile> &files, std::vector<FastaRecord> &recs, bool *fast) {
    recs.clear();
    *fast = false;
    constexpr size_t kHead = 4096, kScan = size_t(1) << 20, kMaxFirst = size_t(1) << 30;
    bool have_tpl = false, open_end = false;      // open_end: the previous record had no final terminator
    FastaRecord tpl;
    size_t tpl_trail = 0;                          // terminator bytes that close the template record
    std::vector<uint8_t> head(kHead + 2), first;
    for (size_t fi = 0; fi < files.size(); ++fi) {
        const SourceFile &f = files[fi];
        size_t pos = 0;
        while (pos < f.size) {
            if (open_end) return BTB_OK;           // bytes follow an unterminated line: they would join it
            const size_t hn = std::min(kHead, f.size - pos);
            BTB_TRY(read_exact(f, pos, hn, head.data()));
            if (head[0] != '>') return BTB_OK;     // gzip, FASTQ, leading junk, ...
            const void *nl = memchr(head.data(), '\n', hn);
            if (!nl) return BTB_OK;
            const size_t hl = size_t(static_cast<const uint8_t *>(nl) - head.data());
            FastaRecord r;
            size_t ne = 1;
            while (ne < hl && !is_space(head[ne])) ++ne;
            r.name.assign(reinterpret_cast<const char *>(head.data() + 1), ne - 1);
            r.file = int32_t(fi);
            r.seq_off = pos + hl + 1;
            if (!have_tpl) {
                // the first record is framed exactly: read on until a line starts with a marker
                first.clear();
                size_t at = r.seq_off, end = f.size;
                bool found = false;
                while (at < f.size && !found) {
                    const size_t n = std::min(kScan, f.size - at), base = first.size();
                    if (base + n > kMaxFirst) return BTB_OK;
                    first.resize(base + n);
                    BTB_TRY(read_exact(f, at, n, first.data() + base));
                    for (size_t k = base; k < base + n; ++k) {
                        const uint8_t c = first[k];
                        if ((c == '>' || c == '@' || c == '+') && (k == 0 || first[k - 1] == '\n')) {
                            if (c != '>') return BTB_OK;           // FASTQ-style input
                            end = r.seq_off + k;
                            first.resize(k);
                            found = true;
                            break;
                        }
                    }
                    at += n;
                }
                r.raw_len = end - r.seq_off;
                FileBytes view;
                view.data = first.data();
                view.size = first.size();
                FastaRecord probe = r;
                probe.seq_off = 0;
                describe_layout(view, probe);
                if (!probe.regular || probe.seq_len <= 0 || (probe.line_w > 0 && probe.line_w < 32)) return BTB_OK;
                size_t R = r.raw_len;
                while (R > 0 && (first[R - 1] == '\n' || first[R - 1] == '\r')) --R;
                tpl_trail = r.raw_len - R;
                if (tpl_trail != 0 && tpl_trail != size_t(probe.eol)) return BTB_OK;     // blank lines at the end
                if (tpl_trail == 0 && end != f.size) return BTB_OK;
                r.line_w = probe.line_w; r.eol = probe.eol; r.seq_len = probe.seq_len; r.regular = true;
                tpl = r;
                tpl.raw_len = R;                   // template length without the closing terminator
                if (tpl_trail == 0) { tpl_trail = size_t(probe.eol); open_end = true; }
                have_tpl = true;
                pos = end;
            } else {
                size_t end = r.seq_off + tpl.raw_len + tpl_trail;
                if (end > f.size) {
                    if (r.seq_off + tpl.raw_len != f.size) return BTB_OK;
                    end = f.size;                  // last line not terminated
                    open_end = true;
                } else {
                    uint8_t t[3] = {0, 0, 0};      // closing terminator, then the next header (or the end)
                    const size_t tn = std::min<size_t>(tpl_trail + 1, f.size - (end - tpl_trail));
                    BTB_TRY(read_exact(f, end - tpl_trail, tn, t));
                    if (t[tpl_trail - 1] != '\n' || (tpl_trail == 2 && t[0] != '\r')) return BTB_OK;
                    if (tn > tpl_trail && t[tpl_trail] != '>') return BTB_OK;
                }
                r.raw_len = end - r.seq_off;
                r.line_w = tpl.line_w; r.eol = tpl.eol; r.seq_len = tpl.seq_len; r.regular = true;
                pos = end;
            }
            recs.push_back(std::move(r));
        }
    }
    *fast = !recs.empty();
    return BTB_OK;
}

// ------------------------------------------------------------------------------ decimal output
namespace {
struct DigitLut {
    char d[10000][4];
    DigitLut() {
        for (int v = 0; v < 10000; ++v) {
            d[v][0] = char('0' + v / 1000);
            d[v][1] = char('0' + v / 100 % 10);
            d[v][2] = char('0' + v / 10 % 10);
            d[v][3] = char('0' + v % 10);
        }
    }
};
const DigitLut kLut;

inline char *put_u32(char *p, uint32_t v) {
    if (v < 10000) {
        const char *s = kLut.d[v];
        const int skip = int(v < 10) + int(v < 100) + int(v < 1000);     // branch-free: digit counts are data dependent
        memcpy(p, s + skip, 4);                      // writes up to 4, advances by the real width
        return p + (4 - skip);
    }
    if (v < 100000000u) {
        const uint32_t hi = v / 10000, lo = v % 10000;
        p = put_u32(p, hi);
        memcpy(p, kLut.d[lo], 4);
        return p + 4;
    }
    const uint32_t hi = v / 100000000u, rest = v % 100000000u;
    p = put_u32(p, hi);
    memcpy(p, kLut.d[rest / 10000], 4);
    memcpy(p + 4, kLut.d[rest % 10000], 4);
    return p + 8;
}
}  // namespace

// bytes format_row will produce: one separator per cell plus the decimal digits
size_t row_text_len(const void *row, int elem_bytes, int64_t n) {
    size_t digits = 0;
    if (elem_bytes == 2) {
        const uint16_t *r = static_cast<const uint16_t *>(row);
        for (int64_t i = 0; i < n; ++i) {
            const uint32_t v = r[i];
            digits += 1u + (v >= 10) + (v >= 100) + (v >= 1000) + (v >= 10000);
        }
    } else {
        const uint32_t *r = static_cast<const uint32_t *>(row);
        for (int64_t i = 0; i < n; ++i) {
            const uint32_t v = r[i];
            digits += 1u + (v >= 10) + (v >= 100) + (v >= 1000) + (v >= 10000) + (v >= 100000) + (v >= 1000000) +
                      (v >= 10000000) + (v >= 100000000) + (v >= 1000000000);
        }
    }
    return size_t(n) + digits;
}

namespace {
// exact-width variant for the last cells of a row: never writes past the value's own digits
inline char *put_u32_exact(char *p, uint32_t v) {
    char tmp[16];
    char *e = put_u32(tmp, v);
    memcpy(p, tmp, size_t(e - tmp));
    return p + (e - tmp);
}
template <typename T>
inline char *format_cells(const T *r, int64_t n, char sep, char *p) {
    // put_u32 may write up to 3 bytes beyond the digits it keeps (overwritten by the next cell): the
    // last two cells use the exact variant so a row never touches bytes behind its own end
    const int64_t bulk = n > 2 ? n - 2 : 0;
    for (int64_t i = 0; i < bulk; ++i) { *p++ = sep; p = put_u32(p, r[i]); }
    for (int64_t i = bulk; i < n; ++i) { *p++ = sep; p = put_u32_exact(p, r[i]); }
    return p;
}
}  // namespace

size_t format_row(const void *row, int elem_bytes, int64_t n, char sep, char *dst) {
    char *p = elem_bytes == 2 ? format_cells(static_cast<const uint16_t *>(row), n, sep, dst)
                              : format_cells(static_cast<const uint32_t *>(row), n, sep, dst);
    return size_t(p - dst);
}

// ---- ViewBovine labels (btbphylo/phylogeny.py:206-216 process_sample_name, btbphylo/utils.py:110-119
// extract_submission_no).  ASCII rule: drop a trailing "_consensus"; the first substring of the
// form DD-DDDD[D]-DD (D = decimal digit, leftmost match, five middle digits preferred over four)
// becomes "AF-" + that substring, otherwise the name stays; the result is upper-cased.
namespace {
inline bool is_digit(char c) { return c >= '0' && c <= '9'; }

// Length of the match starting at s[i], or 0.
size_t submission_match(const std::string &s, size_t i) {
    const size_t n = s.size();
    if (i + 10 > n || !is_digit(s[i]) || !is_digit(s[i + 1]) || s[i + 2] != '-') return 0;
    for (size_t mid = 5; mid >= 4; --mid) {
        const size_t dash = i + 3 + mid;
        if (dash + 3 > n) continue;
        bool ok = true;
        for (size_t k = i + 3; k < dash && ok; ++k) ok = is_digit(s[k]);
        if (ok && s[dash] == '-' && is_digit(s[dash + 1]) && is_digit(s[dash + 2])) return dash + 3 - i;
    }
    return 0;
}
}  // namespace

std::string process_sample_name(const std::string &name) {
    static const char kSuffix[] = "_consensus";
    const size_t sl = sizeof(kSuffix) - 1;
    std::string s = name;
    if (s.size() >= sl && s.compare(s.size() - sl, sl, kSuffix) == 0) s.resize(s.size() - sl);
    for (size_t i = 0; i + 10 <= s.size(); ++i) {
        const size_t m = submission_match(s, i);
        if (m) { s = "AF-" + s.substr(i, m); break; }
    }
    for (char &c : s) if (c >= 'a' && c <= 'z') c = char(c - 'a' + 'A');
    return s;
}

// First field of every line after the header line.
int csv_row_names(const FileBytes &fb, char sep, std::vector<std::string> &names) {
    names.clear();
    const char *p = reinterpret_cast<const char *>(fb.data), *end = p + fb.size;
    const char *nl = static_cast<const char *>(memchr(p, '\n', size_t(end - p)));
    if (!nl) return BTB_OK;
    p = nl + 1;
    while (p < end) {
        nl = static_cast<const char *>(memchr(p, '\n', size_t(end - p)));
        const char *line_end = nl ? nl : end;
        if (line_end > p) {
            const char *f = static_cast<const char *>(memchr(p, sep, size_t(line_end - p)));
            names.emplace_back(p, f ? f : line_end);
        }
        p = line_end + 1;
    }
    return BTB_OK;
}

int relabel_csv(const FileBytes &fb, char sep, const std::vector<std::string> &labels, FILE *fo) {
    const char *p = reinterpret_cast<const char *>(fb.data), *end = p + fb.size;
    const char *nl = static_cast<const char *>(memchr(p, '\n', size_t(end - p)));
    const char *head_end = nl ? nl : end;
    int64_t fields = 1;
    for (const char *q = p; q < head_end; ++q) fields += *q == sep;
    if (fields - 1 != int64_t(labels.size()))
        return fail(BTB_EFORMAT, "Length mismatch: Expected axis has %lld elements, new values have %lld elements",
                    (long long)(fields - 1), (long long)labels.size());
    const char *corner_end = static_cast<const char *>(memchr(p, sep, size_t(head_end - p)));
    std::string head(p, corner_end ? corner_end : head_end);
    for (auto &l : labels) { head.push_back(sep); head += l; }
    head.push_back('\n');
    if (fwrite(head.data(), 1, head.size(), fo) != head.size()) return fail(BTB_EIO, "write failed");
    p = head_end + 1;
    size_t row = 0;
    while (p < end) {
        nl = static_cast<const char *>(memchr(p, '\n', size_t(end - p)));
        const char *line_end = nl ? nl : end;
        if (line_end > p) {
            if (row >= labels.size()) return fail(BTB_EFORMAT, "csv has more rows than labels (%zu)", labels.size());
            const char *f = static_cast<const char *>(memchr(p, sep, size_t(line_end - p)));
            if (!f) f = line_end;
            const std::string &l = labels[row++];
            if (fwrite(l.data(), 1, l.size(), fo) != l.size() ||
                fwrite(f, 1, size_t(line_end - f), fo) != size_t(line_end - f) || fputc('\n', fo) == EOF)
                return fail(BTB_EIO, "write failed");
        }
        p = line_end + 1;
    }
    if (row != labels.size()) return fail(BTB_EFORMAT, "csv has %zu rows but %zu labels were given", row, labels.size());
    return BTB_OK;
}

}  // namespace btb
