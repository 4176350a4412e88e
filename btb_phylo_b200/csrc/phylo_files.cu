// Levels 1 and 2 of the extraction stage: btb_snp_sites (file -> file) and btb_extract_host
// (host rows -> keep mask + SNP rows).  Host code only stages bytes and frames records; every
// per-base decision is taken by the kernels in extract.cu.
//
// Reference call site: btbphylo/phylogeny.py:133  `snp-sites <multi_fasta> -c -o <snps.fas>`.
#include <fcntl.h>
#include <sys/mman.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <cerrno>
#include <chrono>
#include <cstring>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "common.cuh"
#include "host_io.h"

namespace btb {
int require_device(int device);
std::vector<int> usable_devices();
std::vector<int> worker_devices(int n_gpus);
int pack_planes(const uint8_t *d_text, int64_t text_bytes, const int64_t *d_rec_off, const int32_t *d_line_w,
                const int32_t *d_eol, int64_t n_chunk, int64_t first, int64_t n_total, int64_t G, uint32_t *d_planes,
                int32_t *d_flags, uint32_t *d_colstate, cudaStream_t stream);
extern int g_replace_unlink_first;
int first_usable_device();
}
using namespace btb;

namespace {

struct DeviceBuf {
    void *p = nullptr;
    ~DeviceBuf() { if (p) cudaFree(p); }
    int alloc(size_t bytes) {
        BTB_CUDA(cudaMalloc(&p, std::max<size_t>(bytes, 16)));
        return BTB_OK;
    }
    template <typename T> T *as() { return static_cast<T *>(p); }
};
struct PinnedBuf {
    void *p = nullptr;
    ~PinnedBuf() { if (p) cudaFreeHost(p); }
    int alloc(size_t bytes) {
        BTB_CUDA(cudaMallocHost(&p, std::max<size_t>(bytes, 16)));
        return BTB_OK;
    }
    template <typename T> T *as() { return static_cast<T *>(p); }
};
struct Stream {
    cudaStream_t s = nullptr;
    ~Stream() { if (s) cudaStreamDestroy(s); }
    int create() { BTB_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking)); return BTB_OK; }
};

double wall_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

void parallel_for(int64_t n, int threads, const std::function<void(int64_t, int64_t)> &fn) {
    threads = int(std::max<int64_t>(1, std::min<int64_t>(std::min(threads, 64), n)));
    if (threads == 1) { fn(0, n); return; }
    std::vector<std::thread> pool;
    const int64_t span = (n + threads - 1) / threads;
    for (int t = 0; t < threads; ++t) {
        const int64_t lo = t * span, hi = std::min(n, lo + span);
        if (lo < hi) pool.emplace_back(fn, lo, hi);
    }
    for (auto &th : pool) th.join();
}

// After all samples are packed (and their column state accumulated by k1): select.  Leaves L in *L_out, the
// kept-column indices in d_idx, the keep bitmask in d_keep.
int select_only(int64_t G, int pure_mode, uint32_t *d_state, uint32_t *d_keep, uint32_t *d_idx, uint32_t *d_count,
                int64_t *L_out, cudaStream_t st) {
    BTB_TRY(btb_select_columns(d_state, G, pure_mode, d_keep, d_idx, d_count, st));
    uint32_t L = 0;
    BTB_CUDA(cudaMemcpyAsync(&L, d_count, 4, cudaMemcpyDeviceToHost, st));
    BTB_CUDA(cudaStreamSynchronize(st));
    *L_out = L;
    return BTB_OK;
}

}  // namespace

// =============================================================================== level 2
extern "C" int btb_extract_host(const uint8_t *seqs, int64_t n, int64_t G, int64_t row_stride, int pure_mode,
                                int device, uint8_t *keep, int64_t *n_snps, uint8_t *snps_out, int64_t snps_ld) {
    if (!seqs || n <= 0 || G <= 0 || row_stride < G || !n_snps) return fail(BTB_EINVAL, "btb_extract_host: bad argument");
    if (!pure_mode) return fail(BTB_EINVAL, "only snp-sites -c (pure mode) is implemented on the GPU");
    BTB_TRY(require_device(device));
    Stream st;
    BTB_TRY(st.create());
    const int64_t words = btb_plane_words(G);
    DeviceBuf planes, state, keepw, idx, count, text, off, lw, el, flags, snps;
    BTB_TRY(planes.alloc(size_t(3 * n * words * 4)));
    BTB_TRY(state.alloc(size_t(5 * words * 4)));
    BTB_TRY(keepw.alloc(size_t(words * 4)));
    BTB_TRY(idx.alloc(size_t(words * 32 * 4)));
    BTB_TRY(count.alloc(16));
    // dense rows = unwrapped records; stage them in chunks of <= 256 MB
    const int64_t d_stride = (G + 15) / 16 * 16;
    const int64_t chunk_rows = std::max<int64_t>(1, std::min<int64_t>(n, (int64_t(256) << 20) / d_stride));
    BTB_TRY(text.alloc(size_t(chunk_rows * d_stride + 16)));
    BTB_TRY(off.alloc(size_t(chunk_rows * 8)));
    BTB_TRY(lw.alloc(size_t(chunk_rows * 4)));
    BTB_TRY(el.alloc(size_t(chunk_rows * 4)));
    BTB_TRY(flags.alloc(size_t(n * 4)));
    BTB_CUDA(cudaMemsetAsync(flags.p, 0, size_t(n * 4), st.s));
    BTB_CUDA(cudaMemsetAsync(state.p, 0, size_t(5 * words * 4), st.s));
    BTB_CUDA(cudaMemsetAsync(lw.p, 0, size_t(chunk_rows * 4), st.s));
    BTB_CUDA(cudaMemsetAsync(el.p, 0, size_t(chunk_rows * 4), st.s));
    std::vector<int64_t> h_off(static_cast<size_t>(chunk_rows));
    for (int64_t r = 0; r < chunk_rows; ++r) h_off[size_t(r)] = r * d_stride;
    BTB_CUDA(cudaMemcpyAsync(off.p, h_off.data(), size_t(chunk_rows * 8), cudaMemcpyHostToDevice, st.s));
    for (int64_t r0 = 0; r0 < n; r0 += chunk_rows) {
        const int64_t rows = std::min(chunk_rows, n - r0);
        BTB_CUDA(cudaMemcpy2DAsync(text.p, size_t(d_stride), seqs + r0 * row_stride, size_t(row_stride), size_t(G),
                                   size_t(rows), cudaMemcpyHostToDevice, st.s));
        // k1 with the column state accumulated on the way (no second pass over the planes)
        BTB_TRY(pack_planes(text.as<uint8_t>(), rows * d_stride, off.as<int64_t>(), lw.as<int32_t>(), el.as<int32_t>(), rows, r0, n, G,
                            planes.as<uint32_t>(), flags.as<int32_t>() + r0, state.as<uint32_t>(), st.s));
    }
    int64_t L = 0;
    BTB_TRY(select_only(G, pure_mode, state.as<uint32_t>(), keepw.as<uint32_t>(), idx.as<uint32_t>(), count.as<uint32_t>(), &L, st.s));
    *n_snps = L;
    if (keep) {
        std::vector<uint32_t> hk(static_cast<size_t>(words));
        BTB_CUDA(cudaMemcpy(hk.data(), keepw.p, size_t(words * 4), cudaMemcpyDeviceToHost));
        for (int64_t c = 0; c < G; ++c) keep[c] = uint8_t((hk[size_t(c >> 5)] >> (c & 31)) & 1u);
    }
    if (snps_out && L > 0) {
        if (snps_ld < L) return fail(BTB_EINVAL, "btb_extract_host: snps_ld < number of SNPs (%lld)", (long long)L);
        const int64_t ld = (L + 15) / 16 * 16;
        BTB_TRY(snps.alloc(size_t(n * ld)));
        BTB_TRY(btb_compact_columns(planes.as<uint32_t>(), 0, n, n, G, idx.as<uint32_t>(), L, snps.as<uint8_t>(), ld, st.s));
        BTB_CUDA(cudaMemcpy2DAsync(snps_out, size_t(snps_ld), snps.p, size_t(ld), size_t(L), size_t(n),
                                   cudaMemcpyDeviceToHost, st.s));
        BTB_CUDA(cudaStreamSynchronize(st.s));
    }
    return BTB_OK;
}

namespace {

// Where the record bytes come from: the mapped / inflated file of the exact path, or the open files
// of the predicted framing (pread straight into the pinned staging buffers).
struct ByteSource {
    const FileBytes *fb = nullptr;
    const std::vector<SourceFile> *files = nullptr;
};
constexpr int kRetryExact = 1000;      // internal: the predicted framing did not hold, run the exact path

// One device's share of the extraction stage: samples [s0, s1) of the file, their planes and the
// staging pipeline (host threads copy record bytes into pinned chunks, H2D, k1).
struct ExtractCtx {
    int slot = 0, dev = 0, threads = 1, rc = BTB_OK;      // slot: worker index; dev: its CUDA ordinal
    std::string err;
    int64_t s0 = 0, s1 = 0, nloc = 0, L = 0;
    Stream st;
    DeviceBuf planes, state, keepw, idx, count, flags, d_off, d_lw, d_el, d_text[2], snps;
    PinnedBuf h_text[2];
    cudaEvent_t done[2] = {nullptr, nullptr};
    bool used[2] = {false, false};
    size_t chunk_cap = 0;
    std::vector<int64_t> h_off;
    std::vector<int32_t> h_lw, h_el;
    double t_wait = 0, t_fill = 0, t_enqueue = 0;      // BTB_TIMING: where run_chunk spends its time
    bool restaged = false;                             // a record was packed twice (see pack_all)

    ~ExtractCtx() {
        for (int k = 0; k < 2; ++k)
            if (done[k]) cudaEventDestroy(done[k]);
    }

    int setup(const std::vector<FastaRecord> &recs, int64_t G, int64_t words) {
        nloc = s1 - s0;
        BTB_TRY(st.create());
        BTB_TRY(planes.alloc(size_t(3 * nloc * words * 4)));
        BTB_TRY(state.alloc(size_t(5 * words * 4)));
        BTB_TRY(keepw.alloc(size_t(words * 4)));
        BTB_TRY(idx.alloc(size_t(words * 32 * 4)));
        BTB_TRY(count.alloc(16));
        BTB_TRY(flags.alloc(size_t(nloc * 4)));
        BTB_TRY(d_off.alloc(size_t(nloc * 8)));
        BTB_TRY(d_lw.alloc(size_t(nloc * 4)));
        BTB_TRY(d_el.alloc(size_t(nloc * 4)));
        BTB_CUDA(cudaMemsetAsync(flags.p, 0, size_t(nloc * 4), st.s));
        BTB_CUDA(cudaMemsetAsync(state.p, 0, size_t(5 * words * 4), st.s));
        size_t max_rec = 0;
        for (int64_t i = s0; i < s1; ++i) max_rec = std::max(max_rec, recs[size_t(i)].raw_len + 64);
        // 64 MB chunks keep the H2D pipeline busy; larger pinned buffers only cost allocation time
        chunk_cap = std::max<size_t>(size_t(64) << 20, max_rec + (size_t(G) + 64));
        for (int k = 0; k < 2; ++k) {
            BTB_TRY(h_text[k].alloc(chunk_cap + 64));
            BTB_TRY(d_text[k].alloc(chunk_cap + 64));
            BTB_CUDA(cudaEventCreateWithFlags(&done[k], cudaEventDisableTiming));
        }
        h_off.assign(size_t(nloc), 0);
        h_lw.assign(size_t(nloc), 0);
        h_el.assign(size_t(nloc), 0);
        return BTB_OK;
    }

    // stage records [i0, i1) (global indices) into buffer b; `force_walk` re-stages with the exact walk
    int run_chunk(const ByteSource &src, const std::vector<FastaRecord> &recs, int64_t G, int b, int64_t i0, int64_t i1, bool force_walk) {
        const double tw0 = wall_s();
        if (used[b]) BTB_CUDA(cudaEventSynchronize(done[b]));
        const double tw1 = wall_s();
        t_wait += tw1 - tw0;
        uint8_t *dst = h_text[b].as<uint8_t>();
        size_t pos = 0;
        std::vector<size_t> at(static_cast<size_t>(i1 - i0));
        for (int64_t i = i0; i < i1; ++i) {
            const FastaRecord &r = recs[size_t(i)];
            const bool walk = force_walk || !r.regular;
            at[size_t(i - i0)] = pos;
            h_off[size_t(i - s0)] = int64_t(pos);
            h_lw[size_t(i - s0)] = walk ? 0 : r.line_w;
            h_el[size_t(i - s0)] = walk ? 0 : r.eol;
            pos += walk ? size_t(G) : r.raw_len;
            pos = (pos + 15) & ~size_t(15);
        }
        const size_t total = pos;
        if (src.fb) {
            const FileBytes &fb = *src.fb;
            parallel_for(i1 - i0, threads, [&](int64_t lo, int64_t hi) {
                for (int64_t k = lo; k < hi; ++k) {
                    const FastaRecord &r = recs[size_t(i0 + k)];
                    if (force_walk || !r.regular) walk_record(fb, r, dst + at[size_t(k)]);
                    else memcpy(dst + at[size_t(k)], fb.data + r.seq_off, r.raw_len);
                }
            });
        } else {
            // predicted framing: the record regions are read straight into the pinned buffer, in
            // pieces of <= 2 MB so that every host thread has work whatever the record size
            struct Piece { int32_t file; size_t src, len, dst; };
            std::vector<Piece> pieces;
            constexpr size_t kPiece = size_t(2) << 20;
            for (int64_t k = 0; k < i1 - i0; ++k) {
                const FastaRecord &r = recs[size_t(i0 + k)];
                for (size_t o = 0; o < r.raw_len; o += kPiece)
                    pieces.push_back({r.file, r.seq_off + o, std::min(kPiece, r.raw_len - o), at[size_t(k)] + o});
            }
            std::atomic<int> bad{BTB_OK};
            std::string bad_text;
            std::mutex bad_mutex;
            parallel_for(int64_t(pieces.size()), threads, [&](int64_t lo, int64_t hi) {
                for (int64_t k = lo; k < hi; ++k) {
                    const Piece &pc = pieces[size_t(k)];
                    const int rc = read_exact((*src.files)[size_t(pc.file)], pc.src, pc.len, dst + pc.dst);
                    if (rc != BTB_OK) { std::lock_guard<std::mutex> lock(bad_mutex); bad = rc; bad_text = last_error(); }
                }
            });
            if (bad != BTB_OK) return fail(bad, "%s", bad_text.c_str());
        }
        const double tw2 = wall_s();
        t_fill += tw2 - tw1;
        const int64_t l0 = i0 - s0, cnt = i1 - i0;
        BTB_CUDA(cudaMemcpyAsync(d_text[b].p, dst, total + 16, cudaMemcpyHostToDevice, st.s));
        BTB_CUDA(cudaMemcpyAsync(d_off.as<int64_t>() + l0, h_off.data() + l0, size_t(cnt) * 8, cudaMemcpyHostToDevice, st.s));
        BTB_CUDA(cudaMemcpyAsync(d_lw.as<int32_t>() + l0, h_lw.data() + l0, size_t(cnt) * 4, cudaMemcpyHostToDevice, st.s));
        BTB_CUDA(cudaMemcpyAsync(d_el.as<int32_t>() + l0, h_el.data() + l0, size_t(cnt) * 4, cudaMemcpyHostToDevice, st.s));
        // k1 + k2 in one pass: the planes are written and the column state of these samples is OR-ed into `state`
        BTB_TRY(pack_planes(d_text[b].as<uint8_t>(), int64_t(total), d_off.as<int64_t>() + l0, d_lw.as<int32_t>() + l0,
                            d_el.as<int32_t>() + l0, cnt, l0, nloc, G, planes.as<uint32_t>(), flags.as<int32_t>() + l0,
                            state.as<uint32_t>(), st.s));
        BTB_CUDA(cudaEventRecord(done[b], st.s));
        used[b] = true;
        t_enqueue += wall_s() - tw2;
        return BTB_OK;
    }

    // packs every sample of this device; *bad gets the global index of a record whose exact length
    // turned out different from G (reported by the caller in file order).  With predicted framing any
    // flagged record means the prediction was wrong somewhere: kRetryExact.
    int pack_all(const ByteSource &src, std::vector<FastaRecord> &recs, int64_t G, int64_t *bad) {
        int b = 0;
        for (int64_t i0 = s0; i0 < s1;) {
            size_t bytes = 0;
            int64_t i1 = i0;
            while (i1 < s1) {
                const FastaRecord &r = recs[size_t(i1)];
                const size_t need = (!r.regular ? size_t(G) : r.raw_len) + 16;
                if (i1 > i0 && bytes + need > chunk_cap) break;
                bytes += need;
                ++i1;
            }
            BTB_TRY(run_chunk(src, recs, G, b, i0, i1, false));
            b ^= 1;
            i0 = i1;
        }
        // records whose terminators were not where their first line promised: exact walk, pack again
        std::vector<int32_t> h_flags(static_cast<size_t>(nloc));
        BTB_CUDA(cudaMemcpyAsync(h_flags.data(), flags.p, size_t(nloc * 4), cudaMemcpyDeviceToHost, st.s));
        BTB_CUDA(cudaStreamSynchronize(st.s));
        for (int64_t i = s0; i < s1; ++i) {
            if (!h_flags[size_t(i - s0)]) continue;
            if (!src.fb) return kRetryExact;
            restaged = true;
            FastaRecord &r = recs[size_t(i)];
            r.seq_len = walk_record(*src.fb, r, nullptr);
            r.regular = false;
            if (r.seq_len != G) { *bad = i; return BTB_OK; }
            BTB_TRY(run_chunk(src, recs, G, b, i, i + 1, true));
            b ^= 1;
        }
        if (restaged) {
            // the first packing of those records (with the layout their first line promised) has already been OR-ed
            // into the column state, which cannot be undone: recompute it from the final planes (k2)
            BTB_TRY(btb_column_mask(planes.as<uint32_t>(), nloc, nloc, G, state.as<uint32_t>(), 0, st.s));
        }
        BTB_CUDA(cudaStreamSynchronize(st.s));
        if (getenv("BTB_TIMING"))
            fprintf(stderr, "[btb timing]     (device %d staging: wait for a free buffer %.3f s, fill pinned chunks %.3f s, enqueue H2D + k1 %.3f s)\n",
                    dev, t_wait, t_fill, t_enqueue);
        return BTB_OK;
    }
};

}  // namespace

// =============================================================================== level 1
namespace {

int g_ingest_predict = 1;      // "ingest_predict": try the predicted framing first (see host_io.h)
int g_ingest_last = -1;        // "ingest_last_predicted": 1 if the last btb_snp_sites call completed on the predicted framing

int no_snps() {
    fprintf(stderr, "Warning: No SNPs were detected so there is nothing to output.\n");
    return fail(BTB_ENOSNPS, "Warning: No SNPs were detected so there is nothing to output.");
}

// Framed records -> snps.fas.  Shared by the exact and the predicted framing.
int extract_pipeline(const ByteSource &src, std::vector<FastaRecord> &recs, const char *in_fasta, const char *out_fasta,
                     int pure_mode, int n_gpus, int threads, StageTimer &timer, int64_t *n_samples,
                     int64_t *genome_len, int64_t *n_snps) {
    const int64_t n = int64_t(recs.size());
    const int64_t G = recs[0].seq_len;
    auto unequal = [&](const FastaRecord &r) {
        fprintf(stderr, "Alignment %s contains sequences of unequal length. Expected length is %i but got %i in sequence %s\n\n",
                in_fasta, int(G), int(r.seq_len), r.name.c_str());
        return fail(BTB_EFORMAT, "Alignment %s contains sequences of unequal length. Expected length is %i but got %i in sequence %s",
                    in_fasta, int(G), int(r.seq_len), r.name.c_str());
    };
    for (auto &r : recs)
        if (r.seq_len != G) return unequal(r);
    if (n_samples) *n_samples = n;
    if (genome_len) *genome_len = G;
    if (G == 0) return no_snps();

    // ---- devices: samples are dealt to the GPUs as contiguous record ranges of equal bytes ---------
    const std::vector<int> devs = worker_devices(n_gpus);
    const int n_dev = int(std::max<int64_t>(1, std::min<int64_t>(int64_t(devs.size()), n)));
    const int64_t words = btb_plane_words(G);
    std::vector<ExtractCtx> ctx(static_cast<size_t>(n_dev));
    {
        size_t total_bytes = 0;
        for (auto &r : recs) total_bytes += r.raw_len;
        size_t acc = 0;
        int d = 0;
        ctx[0].s0 = 0;
        for (int64_t i = 0; i < n; ++i) {
            acc += recs[size_t(i)].raw_len;
            while (d + 1 < n_dev && acc >= total_bytes * size_t(d + 1) / size_t(n_dev) && i + 1 < n && (n - i - 1) >= (n_dev - d - 1)) {
                ctx[size_t(d)].s1 = i + 1;
                ctx[size_t(++d)].s0 = i + 1;
            }
        }
        ctx[size_t(d)].s1 = n;
        for (int k = d + 1; k < n_dev; ++k) ctx[size_t(k)].s0 = ctx[size_t(k)].s1 = n;
    }
    for (int d = 0; d < n_dev; ++d) { ctx[size_t(d)].slot = d; ctx[size_t(d)].dev = devs.empty() ? 0 : devs[size_t(d)]; ctx[size_t(d)].threads = std::max(1, threads / n_dev); }
    auto for_each_device = [&](const std::function<int(ExtractCtx &)> &fn) -> int {
        if (n_dev == 1) { int rc = fn(ctx[0]); return rc; }
        std::vector<std::thread> pool;
        for (int d = 0; d < n_dev; ++d)
            pool.emplace_back([&, d] { ctx[size_t(d)].rc = fn(ctx[size_t(d)]); if (ctx[size_t(d)].rc) ctx[size_t(d)].err = last_error(); });
        for (auto &t : pool) t.join();
        for (int d = 0; d < n_dev; ++d)
            if (ctx[size_t(d)].rc) return fail(ctx[size_t(d)].rc, "%s", ctx[size_t(d)].err.c_str());
        return BTB_OK;
    };

    // phase 1: pack this device's samples, reduce them to a local column state, copy it to the host
    std::vector<std::vector<uint32_t>> h_state(static_cast<size_t>(n_dev), std::vector<uint32_t>(size_t(5 * words), 0));
    int bad_record = -1;
    std::mutex bad_mutex;
    BTB_TRY(for_each_device([&](ExtractCtx &c) -> int {
        if (c.s1 <= c.s0) return BTB_OK;
        BTB_TRY(require_device(c.dev));
        BTB_TRY(c.setup(recs, G, words));
        int64_t bad = -1;
        BTB_TRY(c.pack_all(src, recs, G, &bad));
        if (bad >= 0) { std::lock_guard<std::mutex> lock(bad_mutex); if (bad_record < 0 || bad < bad_record) bad_record = int(bad); return BTB_OK; }
        BTB_CUDA(cudaMemcpyAsync(h_state[size_t(c.slot)].data(), c.state.p, size_t(5 * words * 4), cudaMemcpyDeviceToHost, c.st.s));
        BTB_CUDA(cudaStreamSynchronize(c.st.s));
        return BTB_OK;
    }));
    if (bad_record >= 0) return unequal(recs[size_t(bad_record)]);
    timer.mark("stage + H2D + k1 pack (column state fused in)");

    // exchange: OR the per-device column states on the host (5 words per 32 columns, a few MB)
    for (int d = 1; d < n_dev; ++d)
        for (size_t k = 0; k < h_state[0].size(); ++k) h_state[0][k] |= h_state[size_t(d)][k];

    // phase 2: every device selects the kept columns from the combined state
    int64_t L = 0;
    BTB_TRY(for_each_device([&](ExtractCtx &c) -> int {
        if (c.s1 <= c.s0) return BTB_OK;
        BTB_TRY(require_device(c.dev));
        if (n_dev > 1) BTB_CUDA(cudaMemcpyAsync(c.state.p, h_state[0].data(), size_t(5 * words * 4), cudaMemcpyHostToDevice, c.st.s));
        BTB_TRY(btb_select_columns(c.state.as<uint32_t>(), G, pure_mode, c.keepw.as<uint32_t>(), c.idx.as<uint32_t>(), c.count.as<uint32_t>(), c.st.s));
        uint32_t l = 0;
        BTB_CUDA(cudaMemcpyAsync(&l, c.count.p, 4, cudaMemcpyDeviceToHost, c.st.s));
        BTB_CUDA(cudaStreamSynchronize(c.st.s));
        c.L = l;
        return BTB_OK;
    }));
    L = ctx[0].L;
    timer.mark("exchange + k3 select");
    if (n_snps) *n_snps = L;
    if (L == 0) return no_snps();

    // phase 3 + write: every record (">name\n" + L bases + "\n", one line per sequence) has its place in the
    // file before anything is copied: each device compacts its samples, sends them to the host in chunks
    // through the pinned staging buffers the ingest already owns (no n x L pinned matrix is allocated:
    // page-locking 0.7 GB costs more than the whole stage), and its host threads put the rows, with their
    // headers, straight where they belong -- into a mapping of the output file on tmpfs, else into one image
    // in memory that a single write() pushes out.
    const int64_t ld = (L + 15) / 16 * 16;
    std::vector<size_t> rec_at(size_t(n) + 1, 0);
    for (int64_t r = 0; r < n; ++r) rec_at[size_t(r) + 1] = rec_at[size_t(r)] + recs[size_t(r)].name.size() + 2 + size_t(L) + 1;
    const size_t total = rec_at[size_t(n)];
    std::string tmp_path;
    int tfd = -1;
    BTB_TRY(open_output_temp(out_fasta, tmp_path, &tfd));
    char *image = nullptr;
    bool mapped = false;
    if (on_memory_fs(out_fasta) && ftruncate(tfd, off_t(total)) == 0) {
        if (fallocate(tfd, 0, 0, off_t(total)) != 0) { /* sparse is fine too */ }
        void *m = mmap(nullptr, total, PROT_READ | PROT_WRITE, MAP_SHARED, tfd, 0);
        if (m != MAP_FAILED) { image = static_cast<char *>(m); mapped = true; }
    }
    if (!image) image = static_cast<char *>(malloc(std::max<size_t>(total, 1)));
    auto drop_image = [&]() { if (mapped) munmap(image, total); else free(image); image = nullptr; };
    if (!image) { close(tfd); discard_output(tmp_path); return fail(BTB_ENOMEM, "out of host memory"); }
    int rc = for_each_device([&](ExtractCtx &c) -> int {
        if (c.s1 <= c.s0) return BTB_OK;
        BTB_TRY(require_device(c.dev));
        BTB_TRY(c.snps.alloc(size_t(c.nloc * ld)));
        BTB_TRY(btb_compact_columns(c.planes.as<uint32_t>(), 0, c.nloc, c.nloc, G, c.idx.as<uint32_t>(), L, c.snps.as<uint8_t>(), ld, c.st.s));
        const int64_t chunk_rows = std::max<int64_t>(1, int64_t(c.chunk_cap) / ld);
        auto fetch = [&](int64_t r0, int b) -> cudaError_t {         // rows [r0, r0 + chunk_rows) of this device -> staging buffer b
            const int64_t rows = std::min(chunk_rows, c.nloc - r0);
            cudaError_t e = cudaMemcpyAsync(c.h_text[b].p, c.snps.as<uint8_t>() + r0 * ld, size_t(rows * ld), cudaMemcpyDeviceToHost, c.st.s);
            return e != cudaSuccess ? e : cudaEventRecord(c.done[b], c.st.s);
        };
        BTB_CUDA(fetch(0, 0));
        int b = 0;
        for (int64_t r0 = 0; r0 < c.nloc; r0 += chunk_rows, b ^= 1) {
            const int64_t rows = std::min(chunk_rows, c.nloc - r0);
            BTB_CUDA(cudaEventSynchronize(c.done[b]));
            if (r0 + chunk_rows < c.nloc) BTB_CUDA(fetch(r0 + chunk_rows, b ^ 1));     // the next chunk travels while this one is placed
            const char *src = c.h_text[b].as<char>();
            parallel_for(rows, c.threads, [&](int64_t lo, int64_t hi) {
                for (int64_t k = lo; k < hi; ++k) {
                    const int64_t r = c.s0 + r0 + k;
                    const std::string &nm = recs[size_t(r)].name;
                    char *q = image + rec_at[size_t(r)];
                    *q++ = '>';
                    memcpy(q, nm.data(), nm.size());
                    q += nm.size();
                    *q++ = '\n';
                    memcpy(q, src + k * ld, size_t(L));
                    q[L] = '\n';
                }
            });
        }
        return BTB_OK;
    });
    timer.mark("k3 compact + D2H + records placed");
    if (rc == BTB_OK && !mapped) {
        size_t off = 0;
        while (off < total) {
            const ssize_t k = write(tfd, image + off, total - off);
            if (k < 0 && errno == EINTR) continue;
            if (k <= 0) { rc = fail(BTB_EIO, "write to '%s' failed", tmp_path.c_str()); break; }
            off += size_t(k);
        }
    }
    drop_image();
    if (close(tfd) != 0 && rc == BTB_OK) rc = fail(BTB_EIO, "closing '%s' failed", tmp_path.c_str());
    timer.mark("write snps.fas");
    if (rc != BTB_OK) { discard_output(tmp_path); return rc; }
    return commit_output(tmp_path, out_fasta, g_replace_unlink_first != 0);
}

}  // namespace

extern "C" int btb_snp_sites_files(const char *const *in_fastas, int64_t n_files, const char *alignment_name,
                                   const char *out_fasta, int pure_mode, int n_gpus, int host_threads,
                                   int64_t *n_samples, int64_t *genome_len, int64_t *n_snps) {
    if (!in_fastas || n_files <= 0 || !out_fasta) return fail(BTB_EINVAL, "btb_snp_sites: NULL path");
    for (int64_t i = 0; i < n_files; ++i)
        if (!in_fastas[i]) return fail(BTB_EINVAL, "btb_snp_sites: NULL path");
    if (!pure_mode) return fail(BTB_EINVAL, "only snp-sites -c (pure mode) is implemented on the GPU");
    const char *shown = alignment_name ? alignment_name : in_fastas[0];
    const int threads = std::max(1, host_threads);
    StageTimer timer("btb_snp_sites");
    BTB_TRY(require_device(first_usable_device()));
    timer.mark("device init");
    std::vector<FastaRecord> recs;
    if (g_ingest_predict) {
        std::vector<SourceFile> files;
        BTB_TRY(open_sources(in_fastas, n_files, files));
        bool fast = false;
        BTB_TRY(frame_predicted(files, recs, &fast));
        timer.mark("open + predicted framing");
        if (fast) {
            ByteSource src;
            src.files = &files;
            const int rc = extract_pipeline(src, recs, shown, out_fasta, pure_mode, n_gpus, threads, timer, n_samples, genome_len, n_snps);
            if (rc != kRetryExact) { g_ingest_last = 1; return rc; }
            timer.mark("predicted framing rejected by the device checks");
        }
    }
    // exact framing: every byte is scanned for record markers (several files: as if concatenated)
    g_ingest_last = 0;
    FileBytes fb;
    if (n_files == 1) {
        BTB_TRY(load_file(in_fastas[0], fb));
    } else {
        for (int64_t i = 0; i < n_files; ++i) {
            FileBytes one;
            BTB_TRY(load_file(in_fastas[i], one));
            fb.owned.insert(fb.owned.end(), one.data, one.data + one.size);
        }
        fb.data = fb.owned.data();
        fb.size = fb.owned.size();
    }
    BTB_TRY(index_fasta(fb, recs, threads));
    timer.mark("map + index records");
    if (recs.empty()) return no_snps();
    // Records the GPU cannot address with (line width, terminator length) go through the exact walk.
    // A record's layout (and with it its length) is only inferred from its first line and its byte count:
    // the first record, which sets the expected length, is always walked, and so is every record whose
    // inferred length disagrees with it -- a blank line inside a record must not read as "unequal length".
    auto needs_walk = [](const FastaRecord &r) { return !r.regular || (r.line_w > 0 && r.line_w < 32); };
    {
        const int64_t exact0 = walk_record(fb, recs[0], nullptr);
        if (exact0 != recs[0].seq_len) recs[0].regular = false;
        recs[0].seq_len = exact0;
    }
    for (auto &r : recs)
        if (needs_walk(r) || r.seq_len != recs[0].seq_len) { r.seq_len = walk_record(fb, r, nullptr); r.regular = false; }
    ByteSource src;
    src.fb = &fb;
    return extract_pipeline(src, recs, shown, out_fasta, pure_mode, n_gpus, threads, timer, n_samples, genome_len, n_snps);
}

extern "C" int btb_snp_sites(const char *in_fasta, const char *out_fasta, int pure_mode, int n_gpus,
                             int host_threads, int64_t *n_samples, int64_t *genome_len, int64_t *n_snps) {
    if (!in_fasta || !out_fasta) return fail(BTB_EINVAL, "btb_snp_sites: NULL path");
    return btb_snp_sites_files(&in_fasta, 1, in_fasta, out_fasta, pure_mode, n_gpus, host_threads, n_samples, genome_len, n_snps);
}

// Framing of a list of FASTA files taken as if concatenated, by prediction (host_io.h): for callers that stage the
// record bytes themselves (btb_phylo_b200/plane_store.py).  out[6 r + 0..5] = file index, offset of the first
// sequence byte in that file, bytes of the sequence region, bases, bases per line (0 = one line), terminator bytes;
// names = the record names joined by '\n'.  *n_records = -1 when the files are not predictable (gzip, FASTQ, mixed
// or irregular layouts): the caller then has to go through btb_snp_sites_files, which parses anything.
extern "C" int btb_fasta_frame_files(const char *const *in_fastas, int64_t n_files, int64_t *out, int64_t cap_records,
                                     char *names, size_t names_cap, int64_t *n_records) {
    if (!in_fastas || n_files <= 0 || !n_records) return fail(BTB_EINVAL, "btb_fasta_frame_files: bad argument");
    *n_records = -1;
    std::vector<SourceFile> files;
    BTB_TRY(open_sources(in_fastas, n_files, files));
    std::vector<FastaRecord> recs;
    bool fast = false;
    BTB_TRY(frame_predicted(files, recs, &fast));
    if (!fast) return BTB_OK;
    size_t need = 1;
    for (auto &r : recs) need += r.name.size() + 1;
    if (int64_t(recs.size()) > cap_records || (names && need > names_cap) || !out)
        return fail(BTB_EINVAL, "btb_fasta_frame_files: %zu records / %zu name bytes do not fit the buffers", recs.size(), need);
    char *p = names;
    for (size_t i = 0; i < recs.size(); ++i) {
        const FastaRecord &r = recs[i];
        int64_t *o = out + 6 * int64_t(i);
        o[0] = r.file; o[1] = int64_t(r.seq_off); o[2] = int64_t(r.raw_len); o[3] = r.seq_len; o[4] = r.line_w; o[5] = r.eol;
        if (names) {
            if (i) *p++ = '\n';
            memcpy(p, r.name.data(), r.name.size());
            p += r.name.size();
        }
    }
    if (names) *p = 0;
    *n_records = int64_t(recs.size());
    return BTB_OK;
}

namespace btb {
void set_ingest_predict(int v) { g_ingest_predict = v ? 1 : 0; }
int get_ingest_predict() { return g_ingest_predict; }
int get_ingest_last() { return g_ingest_last; }
}  // namespace btb
