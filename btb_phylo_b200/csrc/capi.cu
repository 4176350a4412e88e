// C-ABI of libbtbphylo.so: error plumbing, distance stage (levels 1-3).  See include/btb_phylo.h.
#include <fcntl.h>
#include <sys/mman.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <cerrno>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <cstdarg>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "common.cuh"
#include "host_io.h"

namespace btb {

std::string &last_error() {
    thread_local std::string e;
    return e;
}

int fail(int code, const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    last_error() = buf;
    return code;
}

int distance_tiles_popc(const void *, int64_t, int64_t, int, int64_t, int64_t, void *, int, int64_t, int, int,
                        cudaStream_t, const int2 *);
int distance_tiles_umma(const void *, int64_t, int64_t, int, int64_t, int64_t, void *, int, int64_t, int, int,
                        cudaStream_t, const int2 *);
int distance_rows_popc(const void *, int64_t, int64_t, int, int64_t, int64_t, void *, int, int64_t, int, cudaStream_t);
int distance_rows_umma(const void *, int64_t, int64_t, int, int64_t, int64_t, void *, int, int64_t, int, cudaStream_t);
extern int g_umma_lockstep;
extern int g_umma_fp4_pair;
extern int g_umma_coop;
extern int g_umma_major_zero;
void set_ingest_predict(int v);
int get_ingest_predict();
int get_ingest_last();
extern int g_umma_tile_mode;
extern int g_pack_wpt;
extern int g_pack_gy;
int umma_max_active_clusters(int cg, int *out);
int umma_probe(int cg, int every, int64_t *cycles_x100);
int dist_presence_async(const uint8_t *, int64_t, int64_t, int64_t, uint32_t *, cudaStream_t);
int dist_encode_fp4_dense_rows(const uint8_t *, int64_t, int64_t, int64_t, const uint8_t *, int64_t, uint8_t *, int32_t *,
                               cudaStream_t);
extern int g_comm_band_rows;
extern int g_comm_last_path;
struct CommShared;
CommShared *comm_shared_new(int world);
void comm_shared_free(CommShared *sh);
int comm_create_local(CommShared *sh, int rank, int world, int device, btb_comm **out);
int g_csv_prealloc = 1;          // "csv_prealloc": write() mode allocates the file's blocks before writing (see snp_dists_impl)
int g_replace_unlink_first = 0;  // "replace_unlink_first": remove an existing output before the rename (see host_io.h)
int g_csv_mmap = -1;            // "csv_mmap": -1 auto (mapped writes on tmpfs, write() elsewhere), 0 never, 1 always.
            // "csv_mmap": write snps.csv through a file mapping, rows formatted in place by all host threads.
                               // Off: measured 3.5-3.9 s for the 12.2 GB C4 file against 2.7-3.1 s for write() -- shared-mapping
                               // write faults cost more than the buffered write path, and formatting (0.4 s) hides behind either.
int g_dense_pageable = 1;      // "dense_pageable": alignments above 256 MB are read into pageable memory
int g_host_stream_last = -1;    // "host_stream_last": 1 if the last level-2 call completed on the streamed path
int g_host_stream_rows = 64;    // "host_stream_rows": block rows per streamed band (narrow strips copy slowly)
int g_host_stream = 0;         // "host_stream": level 2 uploads, computes and downloads band by band (see streamed_matrix).
                               // Off by default: measured 0.126-0.136 s against 0.131-0.140 s per C4 call -- the call is bound
                               // by the 5 GB download either way, and an upload running beside it slows the download down.

int require_device(int device) {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        cudaGetLastError();
        return fail(BTB_ENODEVICE, "no CUDA device visible (%s); this library has no CPU fallback",
                    e == cudaSuccess ? "count = 0" : cudaGetErrorString(e));
    }
    if (device < 0 || device >= count) return fail(BTB_EINVAL, "device %d out of range (0..%d)", device, count - 1);
    cudaDeviceProp prop;
    BTB_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) return fail(BTB_ENODEVICE, "device %d is sm_%d%d; libbtbphylo is built for sm_100a only", device, prop.major, prop.minor);
    BTB_CUDA(cudaSetDevice(device));
    return BTB_OK;
}

// engine choice for AUTO: the tensor-core engine whenever the HBM it still has to allocate
// (`extra_bytes`: its one-hot operands plus whatever else the caller has not allocated yet) is free
int pick_engine(int engine, int64_t len, int64_t extra_bytes) {
    if (engine == BTB_ENGINE_POPC || engine == BTB_ENGINE_UMMA) return engine;
    if (len == 0) return BTB_ENGINE_POPC;
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) return BTB_ENGINE_POPC;
    return extra_bytes + (int64_t(1) << 30) < int64_t(free_b) ? BTB_ENGINE_UMMA : BTB_ENGINE_POPC;
}

// rank r of `world` owns tiles [lo, hi) of the band order: equal counts, contiguous for L2 locality
void tile_share(int64_t tiles, int rank, int world, int64_t &lo, int64_t &hi) {
    lo = tiles * rank / world;
    hi = tiles * (rank + 1) / world;
}

}  // namespace btb

using namespace btb;

extern "C" int btb_last_error(char *buf, size_t cap) {
    const std::string &e = last_error();
    if (buf && cap) {
        size_t n = std::min(cap - 1, e.size());
        memcpy(buf, e.data(), n);
        buf[n] = 0;
    }
    return int(e.size());
}

extern "C" const char *btb_version(void) { return "btb-phylo-b200 0.1 (snp-sites 2.5.1 -c / snp-dists 0.8.2 semantics, sm_100a)"; }

namespace btb {
// CUDA ordinals of the usable (sm_100) devices, ascending: worker k of a multi-GPU call runs on ordinal [k]
std::vector<int> usable_devices() {
    std::vector<int> out;
    int c = 0;
    if (cudaGetDeviceCount(&c) != cudaSuccess) { cudaGetLastError(); c = 0; }
    for (int d = 0; d < c; ++d) {
        int major = 0;
        if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, d) == cudaSuccess && major == 10) out.push_back(d);
    }
    return out;
}
// the ordinals the n workers of a multi-GPU call run on: the first n usable devices; with the test knob
// "multi_gpu_oversubscribe" a box with fewer devices runs several workers per device
int g_oversubscribe = 0;
std::vector<int> worker_devices(int n_gpus) {
    std::vector<int> d = usable_devices();
    const size_t have = d.size();
    if (g_oversubscribe && have > 0)
        for (size_t k = have; k < size_t(std::max(1, n_gpus)) && k < 16; ++k) d.push_back(d[k % have]);
    if (d.size() > size_t(std::max(1, n_gpus))) d.resize(size_t(std::max(1, n_gpus)));
    return d;
}
int first_usable_device() {
    const std::vector<int> d = usable_devices();
    return d.empty() ? 0 : d[0];          // 0: require_device(0) then reports why nothing is usable
}
}  // namespace btb

extern "C" int btb_device_count(int *count) {
    if (count) *count = int(usable_devices().size());
    return BTB_OK;
}

extern "C" int btb_set_option(const char *key, int value) {
    if (key && !strcmp(key, "pack_gy")) { g_pack_gy = value < 0 ? 0 : value; return BTB_OK; }
    if (key && !strcmp(key, "pack_wpt")) { g_pack_wpt = value == 4 ? 4 : (value == 1 ? 1 : 2); return BTB_OK; }
    if (key && !strcmp(key, "umma_fp4")) { g_umma_fp4 = value ? 1 : 0; return BTB_OK; }
    if (key && !strcmp(key, "csv_mmap")) { g_csv_mmap = value < 0 ? -1 : (value ? 1 : 0); return BTB_OK; }
    if (key && !strcmp(key, "csv_prealloc")) { g_csv_prealloc = value ? 1 : 0; return BTB_OK; }
    if (key && !strcmp(key, "replace_unlink_first")) { g_replace_unlink_first = value ? 1 : 0; return BTB_OK; }
    if (key && !strcmp(key, "dense_pageable")) { g_dense_pageable = value ? 1 : 0; return BTB_OK; }
    if (key && !strcmp(key, "host_stream")) { g_host_stream = value ? 1 : 0; return BTB_OK; }
    if (key && !strcmp(key, "multi_gpu_oversubscribe")) { g_oversubscribe = value ? 1 : 0; return BTB_OK; }
    if (key && !strcmp(key, "comm_band_rows")) { g_comm_band_rows = value < 2 ? 2 : value; return BTB_OK; }
    if (key && !strcmp(key, "host_stream_rows")) { g_host_stream_rows = value < 8 ? 8 : value; return BTB_OK; }
    if (key && !strcmp(key, "ingest_predict")) { set_ingest_predict(value); return BTB_OK; }
    if (key && !strcmp(key, "umma_fp4_general")) { g_umma_fp4_general = value ? 1 : 0; return BTB_OK; }
    if (key && !strcmp(key, "umma_coop")) { g_umma_coop = value ? 1 : 0; return BTB_OK; }
    if (key && !strcmp(key, "umma_major_zero")) { g_umma_major_zero = value ? 1 : 0; return BTB_OK; }
    if (key && !strcmp(key, "umma_fp4_pair")) { g_umma_fp4_pair = value ? 1 : 0; return BTB_OK; }
    if (key && !strcmp(key, "umma_tile_mode")) { g_umma_tile_mode = value ? 1 : 0; return BTB_OK; }
    if (key && !strcmp(key, "umma_lockstep")) { g_umma_lockstep = value < 0 ? 0 : value; return BTB_OK; }
    if (key && !strcmp(key, "umma_dense_simplex")) { g_umma_dense_simplex = value < 0 || value > 2 ? 0 : value; return BTB_OK; }
    return fail(BTB_EINVAL, "unknown option");
}

extern "C" int btb_query(const char *key, int64_t *value) {
    if (!key || !value) return fail(BTB_EINVAL, "btb_query: NULL argument");
    int v = 0;
    if (!strncmp(key, "probe_cg", 8) && strlen(key) >= 12) {       // probe_cg<1|2>_e<every>
        BTB_TRY(umma_probe(key[8] - '0', atoi(key + 11), value));
        return BTB_OK;
    }
    if (!strcmp(key, "umma_max_active_clusters_cg2")) { BTB_TRY(umma_max_active_clusters(2, &v)); *value = v; return BTB_OK; }
    if (!strcmp(key, "umma_max_active_clusters_cg1")) { BTB_TRY(umma_max_active_clusters(1, &v)); *value = v; return BTB_OK; }
    if (!strcmp(key, "umma_dense_simplex")) { *value = g_umma_dense_simplex; return BTB_OK; }
    if (!strcmp(key, "umma_tile_mode")) { *value = g_umma_tile_mode; return BTB_OK; }
    if (!strcmp(key, "umma_fp4")) { *value = g_umma_fp4; return BTB_OK; }
    if (!strcmp(key, "umma_fp4_pair")) { *value = g_umma_fp4_pair; return BTB_OK; }
    if (!strcmp(key, "umma_major_zero")) { *value = g_umma_major_zero; return BTB_OK; }
    if (!strcmp(key, "ingest_predict")) { *value = get_ingest_predict(); return BTB_OK; }
    if (!strcmp(key, "host_stream")) { *value = g_host_stream; return BTB_OK; }
    if (!strcmp(key, "comm_last_path")) { *value = g_comm_last_path; return BTB_OK; }
    if (!strcmp(key, "comm_band_rows")) { *value = g_comm_band_rows; return BTB_OK; }
    if (!strcmp(key, "host_stream_last")) { *value = g_host_stream_last; return BTB_OK; }
    if (!strcmp(key, "ingest_last_predicted")) { *value = get_ingest_last(); return BTB_OK; }
    return fail(BTB_EINVAL, "btb_query: unknown key '%s'", key);
}

extern "C" int64_t btb_dist_tile_count(int64_t n) { return n > 0 ? tile_count(n) : 0; }

extern "C" int btb_dist_tile_coords(int64_t n, int64_t tile, int32_t *I, int32_t *J) {
    if (n <= 0 || tile < 0 || tile >= tile_count(n)) return fail(BTB_EINVAL, "tile index out of range");
    int i, j;
    tile_coords(tiles_per_dim(n), tile, i, j);
    if (I) *I = i;
    if (J) *J = j;
    return BTB_OK;
}

extern "C" int btb_distance_tiles(int engine, const void *d_operands, int64_t n, int64_t len, int sigma,
                                  int64_t tile_lo, int64_t tile_hi, void *d_out, int out_elem_bytes,
                                  int64_t out_ld, int out_mode, int mirror, void *stream) {
    if (!d_operands || !d_out || n <= 0 || len < 0) return fail(BTB_EINVAL, "btb_distance_tiles: bad argument");
    if (out_elem_bytes != 2 && out_elem_bytes != 4) return fail(BTB_EINVAL, "out_elem_bytes must be 2 or 4");
    if (out_elem_bytes == 2 && len > 65535) return fail(BTB_EINVAL, "uint16 output needs len < 65536");
    if (tile_lo < 0 || tile_hi > tile_count(n) || tile_lo > tile_hi) return fail(BTB_EINVAL, "tile range out of bounds");
    if (out_mode == 0 && out_ld < n) return fail(BTB_EINVAL, "out_ld < n");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (engine == BTB_ENGINE_POPC)
        return distance_tiles_popc(d_operands, n, len, sigma, tile_lo, tile_hi, d_out, out_elem_bytes, out_ld, out_mode, mirror, s, nullptr);
    if (engine == BTB_ENGINE_UMMA)
        return distance_tiles_umma(d_operands, n, len, sigma, tile_lo, tile_hi, d_out, out_elem_bytes, out_ld, out_mode, mirror, s, nullptr);
    return fail(BTB_EINVAL, "btb_distance_tiles: engine must be POPC or UMMA");
}

extern "C" int64_t btb_dist_block_rows(int64_t n) { return n > 0 ? tiles_per_dim(n) : 0; }

// block rows [lo, hi) of rank `rank`: contiguous, (nearly) equal numbers of upper-triangle tiles
extern "C" int btb_dist_row_share(int64_t n, int rank, int world, int64_t *lo, int64_t *hi) {
    if (n <= 0 || world < 1 || rank < 0 || rank >= world || !lo || !hi) return fail(BTB_EINVAL, "btb_dist_row_share: bad argument");
    const int64_t T = tiles_per_dim(n), total = tile_count(n);
    auto boundary = [&](int r) -> int64_t {                 // first block row of rank r
        if (r <= 0) return 0;
        if (r >= world) return T;
        const int64_t want = total * r / world;
        int64_t acc = 0, I = 0;
        while (I < T && acc + (T - I) <= want) { acc += T - I; ++I; }
        // shares start on even block rows when there is room: the CTA-pair kernel works on 512-row units
        if ((I & 1) && I < T && T >= 4 * int64_t(world)) ++I;
        return I;
    };
    *lo = boundary(rank);
    *hi = boundary(rank + 1);
    return BTB_OK;
}

// The same with the cost of preparing the operands counted in: a rank whose block rows start at `lo` encodes
// the rows from lo on (btb_dist_encode_rows), i.e. (T - lo) / T of what encoding all rows costs, `encode_tiles`
// (in units of one tile of distance work).  Rank 0 reads every row, the last rank a third of them: contiguous
// ranges whose  tiles + encode share  are (nearly) equal, found by bisection on the largest cost.
extern "C" int btb_dist_row_share_weighted(int64_t n, int rank, int world, double encode_tiles, int64_t *lo, int64_t *hi) {
    if (n <= 0 || world < 1 || rank < 0 || rank >= world || !lo || !hi || encode_tiles < 0) return fail(BTB_EINVAL, "btb_dist_row_share_weighted: bad argument");
    const int64_t T = tiles_per_dim(n);
    auto cost = [&](int64_t a, int64_t b) -> double {     // block rows [a, b)
        const double tiles = double(b - a) * double(T - a) - double(b - a) * double(b - a - 1) / 2.0;
        return tiles + encode_tiles * double(T - a) / double(T);
    };
    std::vector<int64_t> bound(size_t(world) + 1, T);
    auto fits = [&](double cap, std::vector<int64_t> *out) -> bool {
        int64_t a = 0;
        for (int r = 0; r < world; ++r) {
            int64_t b = a;
            while (b < T && cost(a, b + 1) <= cap) ++b;
            if (b == a && a < T) return false;             // not even one block row fits
            // ranks start on even block rows when there is room (512-row units of the CTA-pair kernel)
            if ((b & 1) && b < T && T >= 4 * int64_t(world) && b - 1 > a) --b;
            if (out) (*out)[size_t(r)] = a;
            a = b;
        }
        if (out) (*out)[size_t(world)] = T;
        return a >= T;
    };
    double lo_c = 0, hi_c = cost(0, T);
    for (int it = 0; it < 60; ++it) {
        const double mid = 0.5 * (lo_c + hi_c);
        if (fits(mid, nullptr)) hi_c = mid; else lo_c = mid;
    }
    fits(hi_c, &bound);
    bound[size_t(world)] = T;
    *lo = bound[size_t(rank)];
    *hi = rank + 1 == world ? T : bound[size_t(rank) + 1];
    return BTB_OK;
}

extern "C" int btb_distance_rows(int engine, const void *d_operands, int64_t n, int64_t len, int sigma,
                                 int64_t row_block_lo, int64_t row_block_hi, void *d_out, int out_elem_bytes,
                                 int64_t out_ld, int mirror, void *stream) {
    if (!d_operands || !d_out || n <= 0 || len < 0) return fail(BTB_EINVAL, "btb_distance_rows: bad argument");
    if (out_elem_bytes != 2 && out_elem_bytes != 4) return fail(BTB_EINVAL, "out_elem_bytes must be 2 or 4");
    if (out_elem_bytes == 2 && len > 65535) return fail(BTB_EINVAL, "uint16 output needs len < 65536");
    if (row_block_lo < 0 || row_block_hi > tiles_per_dim(n) || row_block_lo > row_block_hi || out_ld < n)
        return fail(BTB_EINVAL, "btb_distance_rows: block row range or out_ld out of bounds");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (engine == BTB_ENGINE_POPC)
        return distance_rows_popc(d_operands, n, len, sigma, row_block_lo, row_block_hi, d_out, out_elem_bytes, out_ld, mirror, s);
    if (engine == BTB_ENGINE_UMMA)
        return distance_rows_umma(d_operands, n, len, sigma, row_block_lo, row_block_hi, d_out, out_elem_bytes, out_ld, mirror, s);
    return fail(BTB_EINVAL, "btb_distance_rows: engine must be POPC or UMMA");
}

// =============================================================================== level 2
namespace {
// Device buffers of btb_dist_matrix_host are kept between calls (one set per device): a weekly
// production run calls once, a benchmark loop many times; cudaMalloc of ~20 GB is not free.
struct Workspace {
    void *p[4] = {nullptr, nullptr, nullptr, nullptr};      // sequences, operands, matrix, small scratch
    size_t cap[4] = {0, 0, 0, 0};
    cudaStream_t st = nullptr, st2 = nullptr, st_in = nullptr;
    cudaEvent_t ev = nullptr;
    std::vector<cudaEvent_t> band_ev;     // streamed path: chunk uploaded + scanned
    uint32_t *h_present = nullptr;        // pinned: one 256-bit byte-presence map per band
    size_t h_present_bands = 0;
    int ensure(int k, size_t bytes, void **out) {
        bytes = std::max<size_t>(bytes, 16);
        if (cap[k] < bytes) {
            if (p[k]) cudaFree(p[k]);
            p[k] = nullptr; cap[k] = 0;
            BTB_CUDA(cudaMalloc(&p[k], bytes));
            cap[k] = bytes;
        }
        *out = p[k];
        return BTB_OK;
    }
    void release() {
        for (int k = 0; k < 4; ++k) { if (p[k]) cudaFree(p[k]); p[k] = nullptr; cap[k] = 0; }
        if (st) cudaStreamDestroy(st);
        if (st2) cudaStreamDestroy(st2);
        if (st_in) cudaStreamDestroy(st_in);
        if (ev) cudaEventDestroy(ev);
        for (auto e : band_ev) cudaEventDestroy(e);
        band_ev.clear();
        if (h_present) cudaFreeHost(h_present);
        h_present = nullptr; h_present_bands = 0;
        st = st2 = st_in = nullptr;
        ev = nullptr;
    }
};
Workspace g_ws[16];
std::mutex g_ws_mutex;
}  // namespace

extern "C" int btb_release_workspace(void) {
    std::lock_guard<std::mutex> lock(g_ws_mutex);
    int cur = 0;
    cudaGetDevice(&cur);
    for (int d = 0; d < 16; ++d)
        if (g_ws[d].p[0] || g_ws[d].p[1] || g_ws[d].p[2] || g_ws[d].st || g_ws[d].st2) { cudaSetDevice(d); g_ws[d].release(); }
    cudaSetDevice(cur);
    return BTB_OK;
}

namespace {
// Level 2, one GPU, default mode, dense input (every snp-sites -c output): upload, GEMM and download
// overlap.  The bands of block rows are computed LAST TO FIRST: the last band only needs the last
// samples, every earlier band needs its own samples plus the ones already on the device, so the
// samples are uploaded from the end of the alignment backwards while the tensor cores already run.
// After band b the square of all rows and columns >= its first row is final (mirrored writes), and
// the two newly finished strips are copied to the host while earlier bands are still computed.
// The input is assumed dense (ACGTacgt only) -- what the per-band byte-presence maps check; the
// first band that is not sends the whole call down the general path (*done = false).
int streamed_matrix(Workspace &ws, const uint8_t *seqs, int64_t n, int64_t len, int64_t row_stride, void *out,
                    int out_elem_bytes, int64_t out_ld, bool *done) {
    *done = false;
    const int sigma = BTB_SIGMA_DENSE4;
    const int64_t T = tiles_per_dim(n);
    const int64_t SB = std::max<int64_t>(kBand, int64_t(g_host_stream_rows) / kBand * kBand);   // block rows per streamed band
    const int64_t bands = (T + SB - 1) / SB;
    const int64_t d_stride = (len + 15) / 16 * 16, d_ld = (n + 7) / 8 * 8;
    const int64_t kb = btb_dist_operand_bytes(BTB_ENGINE_UMMA, 1, len, sigma) - 4;      // operand bytes per row
    const size_t eb = size_t(out_elem_bytes);
    uint8_t *d_seqs = nullptr, *d_ops = nullptr, *d_out = nullptr, *d_small = nullptr;
    BTB_TRY(ws.ensure(0, size_t(n * d_stride), reinterpret_cast<void **>(&d_seqs)));
    BTB_TRY(ws.ensure(1, size_t(btb_dist_operand_bytes(BTB_ENGINE_UMMA, n, len, sigma)), reinterpret_cast<void **>(&d_ops)));
    BTB_TRY(ws.ensure(2, size_t(n * d_ld) * eb, reinterpret_cast<void **>(&d_out)));
    BTB_TRY(ws.ensure(3, size_t(256 + bands * 32), reinterpret_cast<void **>(&d_small)));
    if (!ws.st2) BTB_CUDA(cudaStreamCreateWithFlags(&ws.st2, cudaStreamNonBlocking));
    if (!ws.st_in) BTB_CUDA(cudaStreamCreateWithFlags(&ws.st_in, cudaStreamNonBlocking));
    if (!ws.ev) BTB_CUDA(cudaEventCreateWithFlags(&ws.ev, cudaEventDisableTiming));
    while (int64_t(ws.band_ev.size()) < bands) {
        cudaEvent_t e;
        BTB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        ws.band_ev.push_back(e);
    }
    if (ws.h_present_bands < size_t(bands)) {
        if (ws.h_present) cudaFreeHost(ws.h_present);
        ws.h_present = nullptr;
        BTB_CUDA(cudaMallocHost(reinterpret_cast<void **>(&ws.h_present), size_t(bands) * 32));
        ws.h_present_bands = size_t(bands);
    }
    uint8_t table[256];
    int s4 = 0;
    BTB_TRY(btb_dist_alphabet(nullptr, n, len, d_stride, 0, 0, table, &s4, ws.st));        // default-mode table, no scan
    uint8_t *d_table = d_small;
    uint32_t *d_present = reinterpret_cast<uint32_t *>(d_small + 256);
    int32_t *rowsum = reinterpret_cast<int32_t *>(d_ops + n * kb);
    BTB_CUDA(cudaMemcpyAsync(d_table, table, 256, cudaMemcpyHostToDevice, ws.st_in));
    BTB_CUDA(cudaMemsetAsync(d_present, 0, size_t(bands) * 32, ws.st_in));
    BTB_CUDA(cudaMemsetAsync(rowsum, 0, size_t(n) * 4, ws.st_in));
    // every upload is queued now, last band first; each is followed by its presence scan and encode
    for (int64_t b = bands - 1; b >= 0; --b) {
        const int64_t row0 = b * SB * BTB_TILE, row1 = std::min<int64_t>(n, row0 + SB * BTB_TILE);
        BTB_CUDA(cudaMemcpy2DAsync(d_seqs + row0 * d_stride, size_t(d_stride), seqs + row0 * row_stride, size_t(row_stride),
                                   size_t(len), size_t(row1 - row0), cudaMemcpyHostToDevice, ws.st_in));
        BTB_TRY(dist_presence_async(d_seqs + row0 * d_stride, row1 - row0, len, d_stride, d_present + b * 8, ws.st_in));
        BTB_CUDA(cudaMemcpyAsync(ws.h_present + b * 8, d_present + b * 8, 32, cudaMemcpyDeviceToHost, ws.st_in));
        BTB_TRY(dist_encode_fp4_dense_rows(d_seqs + row0 * d_stride, row1 - row0, len, d_stride, d_table, kb,
                                           d_ops + row0 * kb, rowsum + row0, ws.st_in));
        BTB_CUDA(cudaEventRecord(ws.band_ev[size_t(b)], ws.st_in));
    }
    char *ho = static_cast<char *>(out);
    auto copy2d = [&](int64_t r0, int64_t r1, int64_t c0, int64_t c1) -> cudaError_t {
        if (r1 <= r0 || c1 <= c0) return cudaSuccess;
        return cudaMemcpy2DAsync(ho + (r0 * out_ld + c0) * eb, size_t(out_ld) * eb, d_out + (r0 * d_ld + c0) * eb,
                                 size_t(d_ld) * eb, size_t(c1 - c0) * eb, size_t(r1 - r0), cudaMemcpyDeviceToHost, ws.st2);
    };
    for (int64_t b = bands - 1; b >= 0; --b) {
        const int64_t row0 = b * SB * BTB_TILE, row1 = std::min<int64_t>(n, row0 + SB * BTB_TILE);
        BTB_CUDA(cudaEventSynchronize(ws.band_ev[size_t(b)]));
        bool dense = true;                         // only bytes the default-mode table maps to a base
        for (int v = 0; v < 256 && dense; ++v)
            if ((ws.h_present[b * 8 + (v >> 5)] >> (v & 31) & 1u) && table[v] == 0xff) dense = false;
        if (!dense) {
            BTB_CUDA(cudaStreamSynchronize(ws.st_in));
            BTB_CUDA(cudaStreamSynchronize(ws.st));
            BTB_CUDA(cudaStreamSynchronize(ws.st2));
            return BTB_OK;
        }
        BTB_CUDA(cudaStreamWaitEvent(ws.st, ws.band_ev[size_t(b)], 0));
        BTB_TRY(btb_distance_rows(BTB_ENGINE_UMMA, d_ops, n, len, sigma, b * SB, std::min<int64_t>(T, (b + 1) * SB),
                                  d_out, out_elem_bytes, d_ld, 1, ws.st));
        BTB_CUDA(cudaEventRecord(ws.ev, ws.st));
        BTB_CUDA(cudaStreamWaitEvent(ws.st2, ws.ev, 0));
        BTB_CUDA(copy2d(row0, row1, row0, n));     // the band's rows from the diagonal on
        BTB_CUDA(copy2d(row1, n, row0, row1));     // their mirror images below
    }
    BTB_CUDA(cudaStreamSynchronize(ws.st2));
    BTB_CUDA(cudaStreamSynchronize(ws.st));
    *done = true;
    return BTB_OK;
}
}  // namespace

static int dist_matrix_host_impl(const uint8_t *seqs, int64_t n, int64_t len, int64_t row_stride,
                                 int allchars, int keepcase, int engine, int device, int rank, int world,
                                 void *out, int out_elem_bytes, int64_t out_ld, bool upper_only) {
    if (!seqs || !out || n <= 0 || len < 0 || row_stride < len || out_ld < n || world < 1 || rank < 0 || rank >= world)
        return fail(BTB_EINVAL, "btb_dist_matrix_host: bad argument");
    if (out_elem_bytes != 2 && out_elem_bytes != 4) return fail(BTB_EINVAL, "out_elem_bytes must be 2 or 4");
    if (out_elem_bytes == 2 && len > 65535) return fail(BTB_EINVAL, "uint16 output needs len < 65536");
    BTB_TRY(require_device(device));
    if (device >= 16) return fail(BTB_EINVAL, "device ordinal too large");
    std::lock_guard<std::mutex> lock(g_ws_mutex);
    Workspace &ws = g_ws[device];
    if (!ws.st) BTB_CUDA(cudaStreamCreateWithFlags(&ws.st, cudaStreamNonBlocking));
    cudaStream_t st = ws.st;
    int rc = BTB_OK;
    g_host_stream_last = 0;
    if (world == 1 && g_host_stream && !upper_only && !allchars && !keepcase && engine != BTB_ENGINE_POPC && len > 0 &&
        umma_fp4_dense(BTB_SIGMA_DENSE4, len) && tiles_per_dim(n) >= 2 * kBand) {
        size_t free_b = 0, total_b = 0;
        const size_t need = size_t(btb_dist_operand_bytes(BTB_ENGINE_UMMA, n, len, BTB_SIGMA_DENSE4)) + size_t(n) * size_t((len + 15) / 16 * 16) +
                            size_t(n) * size_t((n + 7) / 8 * 8) * size_t(out_elem_bytes);
        if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess && need < free_b + ws.cap[0] + ws.cap[1] + ws.cap[2]) {
            bool done = false;
            BTB_TRY(streamed_matrix(ws, seqs, n, len, row_stride, out, out_elem_bytes, out_ld, &done));
            if (done) { g_host_stream_last = 1; return BTB_OK; }
        }
    }
    uint8_t *d_seqs = nullptr, *d_ops = nullptr, *d_out = nullptr;
    const int64_t d_stride = (len + 15) / 16 * 16;
    const int64_t d_ld = (n + 7) / 8 * 8;
    auto cleanup = [&]() {};
#define L2_CUDA(call)                                                                                     \
    do {                                                                                                  \
        cudaError_t e_ = (call);                                                                          \
        if (e_ != cudaSuccess) {                                                                          \
            rc = fail(BTB_ECUDA, "%s failed: %s", #call, cudaGetErrorString(e_));                         \
            cleanup();                                                                                    \
            return rc;                                                                                    \
        }                                                                                                 \
    } while (0)
#define L2_TRY(call) do { rc = (call); if (rc != BTB_OK) { cleanup(); return rc; } } while (0)
    L2_TRY(ws.ensure(0, size_t(std::max<int64_t>(1, n * d_stride)), reinterpret_cast<void **>(&d_seqs)));
    if (len > 0)
        L2_CUDA(cudaMemcpy2DAsync(d_seqs, size_t(d_stride), seqs, size_t(row_stride), size_t(len), size_t(n),
                                  cudaMemcpyHostToDevice, st));
    uint8_t table[256];
    int sigma = 4;
    L2_TRY(btb_dist_alphabet(d_seqs, n, len, d_stride, allchars, keepcase, table, &sigma, st));
    L2_TRY(ws.ensure(2, size_t(n * d_ld * out_elem_bytes), reinterpret_cast<void **>(&d_out)));
    const int eng = pick_engine(engine, len, std::max<int64_t>(0, btb_dist_operand_bytes(BTB_ENGINE_UMMA, n, len, sigma) - int64_t(ws.cap[1])));
    L2_TRY(ws.ensure(1, size_t(std::max<int64_t>(16, btb_dist_operand_bytes(eng, n, len, sigma))), reinterpret_cast<void **>(&d_ops)));
    L2_TRY(btb_dist_encode(eng, d_seqs, n, len, d_stride, table, sigma, d_ops, st));
    const size_t eb = size_t(out_elem_bytes);
    const int64_t T = tiles_per_dim(n);
    if (world == 1) {
        // one launch per band of block rows; a finished band's matrix rows are final (row-band tile
        // order + mirrored writes), so their copy to the host overlaps the next bands' GEMM
        if (!ws.st2) L2_CUDA(cudaStreamCreateWithFlags(&ws.st2, cudaStreamNonBlocking));
        if (!ws.ev) L2_CUDA(cudaEventCreateWithFlags(&ws.ev, cudaEventDisableTiming));
        int64_t first = 0;
        for (int64_t r0 = 0; r0 < T; r0 += kBand) {
            const int64_t w = std::min<int64_t>(kBand, T - r0), cnt = band_tiles(T, r0, w);
            L2_TRY(btb_distance_tiles(eng, d_ops, n, len, sigma, first, first + cnt, d_out, out_elem_bytes, d_ld, 0, upper_only ? 0 : 1, st));
            L2_CUDA(cudaEventRecord(ws.ev, st));
            L2_CUDA(cudaStreamWaitEvent(ws.st2, ws.ev, 0));
            const int64_t row0 = r0 * BTB_TILE, row1 = std::min<int64_t>(n, (r0 + w) * BTB_TILE);
            const int64_t col0 = upper_only ? row0 : 0;                // upper triangle only: nothing left of the band's diagonal block
            L2_CUDA(cudaMemcpy2DAsync(static_cast<char *>(out) + size_t(row0 * out_ld + col0) * eb, size_t(out_ld) * eb,
                                      d_out + size_t(row0 * d_ld + col0) * eb, size_t(d_ld) * eb, size_t(n - col0) * eb,
                                      size_t(row1 - row0), cudaMemcpyDeviceToHost, ws.st2));
            first += cnt;
        }
        L2_CUDA(cudaStreamSynchronize(ws.st2));
    } else {
        // this rank's contiguous block rows: every cell (r, c >= r) of those rows plus the mirrors, copied
        // back as two wide 2-D copies (the rows from their first column on, and the mirrored strip below)
        int64_t lo = 0, hi = 0;
        L2_TRY(btb_dist_row_share(n, rank, world, &lo, &hi));
        L2_TRY(btb_distance_rows(eng, d_ops, n, len, sigma, lo, hi, d_out, out_elem_bytes, d_ld, upper_only ? 0 : 1, st));
        char *ho = static_cast<char *>(out);
        auto copy2d = [&](int64_t r0, int64_t r1, int64_t c0, int64_t c1) -> cudaError_t {
            if (r1 <= r0 || c1 <= c0) return cudaSuccess;
            return cudaMemcpy2DAsync(ho + (r0 * out_ld + c0) * eb, size_t(out_ld) * eb, d_out + (r0 * d_ld + c0) * eb,
                                     size_t(d_ld) * eb, size_t(c1 - c0) * eb, size_t(r1 - r0), cudaMemcpyDeviceToHost, st);
        };
        const int64_t row0 = std::min<int64_t>(n, lo * BTB_TILE), row1 = std::min<int64_t>(n, hi * BTB_TILE);
        L2_CUDA(copy2d(row0, row1, row0, n));
        if (!upper_only) L2_CUDA(copy2d(row1, n, row0, row1));
    }
    L2_CUDA(cudaStreamSynchronize(st));
    cleanup();
    return BTB_OK;
#undef L2_CUDA
#undef L2_TRY
}

extern "C" int btb_dist_matrix_host(const uint8_t *seqs, int64_t n, int64_t len, int64_t row_stride,
                                    int allchars, int keepcase, int engine, int device, int rank, int world,
                                    void *out, int out_elem_bytes, int64_t out_ld) {
    return dist_matrix_host_impl(seqs, n, len, row_stride, allchars, keepcase, engine, device, rank, world, out, out_elem_bytes, out_ld, false);
}

extern "C" int btb_dist_matrix_host_upper(const uint8_t *seqs, int64_t n, int64_t len, int64_t row_stride,
                                          int allchars, int keepcase, int engine, int device, int rank, int world,
                                          void *out, int out_elem_bytes, int64_t out_ld) {
    return dist_matrix_host_impl(seqs, n, len, row_stride, allchars, keepcase, engine, device, rank, world, out, out_elem_bytes, out_ld, true);
}

// =============================================================================== level 1: snp-dists
namespace {

struct DenseFasta {
    std::vector<std::string> names;
    int64_t n = 0, len = 0, stride = 0;
    uint8_t *rows = nullptr;          // pinned
    bool pinned = true;               // large alignments use pageable memory: pinning 1.5 GB costs more than the staged upload
    ~DenseFasta() { release(); }
    void release() { if (rows) { if (pinned) cudaFreeHost(rows); else free(rows); } rows = nullptr; }
    int alloc(size_t bytes) {
        pinned = bytes <= (size_t(256) << 20) || !g_dense_pageable;
        if (pinned) { BTB_CUDA(cudaMallocHost(reinterpret_cast<void **>(&rows), bytes)); }
        else if (!(rows = static_cast<uint8_t *>(aligned_alloc(4096, (bytes + 4095) & ~size_t(4095))))) return fail(BTB_ENOMEM, "out of host memory");
        return BTB_OK;
    }
};

// File -> dense pinned rows.  Error texts follow snp-dists 0.8.2.
// `on_framed` (optional) runs as soon as the record names and the sequence count are known -- before the
// sequence bytes are read -- so that the caller can start preparing its output beside the read.
int read_dense(const char *path, int threads, DenseFasta &df, const std::function<void()> *on_framed = nullptr) {
    // Fast path for what snp-sites writes (one unwrapped line per record, equal lengths): frame by
    // prediction (host_io.h) and pread every record straight into the pinned rows.  The host checks
    // that no line break hides inside a record; anything unexpected takes the exact path below.
    if (get_ingest_predict()) {
        std::vector<SourceFile> files;
        const char *one[1] = {path};
        std::vector<FastaRecord> recs;
        bool fast = false;
        if (open_sources(one, 1, files) == BTB_OK && frame_predicted(files, recs, &fast) == BTB_OK && fast &&
            recs[0].line_w == 0 && recs[0].eol == 1 && recs[0].seq_len > 0 && int64_t(recs.size()) <= BTB_MAX_SEQ) {
            df.n = int64_t(recs.size());
            df.len = recs[0].seq_len;
            df.stride = std::max<int64_t>(16, (df.len + 15) / 16 * 16);
            df.names.resize(recs.size());
            for (size_t i = 0; i < recs.size(); ++i) df.names[i] = recs[i].name;
            if (on_framed) (*on_framed)();
            BTB_TRY(df.alloc(size_t(df.n * df.stride)));
            std::atomic<bool> ok{true};
            auto work = [&](int64_t lo, int64_t hi) {
                for (int64_t i = lo; i < hi && ok; ++i) {
                    uint8_t *dst = df.rows + i * df.stride;
                    if (read_exact(files[0], recs[size_t(i)].seq_off, size_t(df.len), dst) != BTB_OK ||
                        memchr(dst, '\n', size_t(df.len)) || dst[df.len - 1] == '\r') { ok = false; return; }
                    memset(dst + df.len, '.', size_t(df.stride - df.len));
                }
            };
            const int T = int(std::max<int64_t>(1, std::min<int64_t>(std::min(threads, 64), df.n / 16)));
            if (T == 1) work(0, df.n);
            else {
                std::vector<std::thread> pool;
                const int64_t span = (df.n + T - 1) / T;
                for (int t = 0; t < T; ++t) {
                    const int64_t lo = t * span, hi = std::min(df.n, lo + span);
                    if (lo < hi) pool.emplace_back(work, lo, hi);
                }
                for (auto &th : pool) th.join();
            }
            if (ok) return BTB_OK;
            df.release();
            on_framed = nullptr;                    // the exact path below sees the same names: the output is already being prepared
        }
    }
    FileBytes fb;
    BTB_TRY(load_file(path, fb));
    std::vector<FastaRecord> recs;
    BTB_TRY(index_fasta(fb, recs, threads));
    if (recs.empty()) return fail(BTB_EFORMAT, "ERROR: file contained no sequences");
    if (int64_t(recs.size()) > BTB_MAX_SEQ)
        return fail(BTB_EFORMAT, "ERROR: snp-dists can only handle %d sequences at most. Please change MAX_SEQ and recompile.", BTB_MAX_SEQ);
    // the exact walk gives every record's length whatever its line layout
    df.n = int64_t(recs.size());
    df.len = walk_record(fb, recs[0], nullptr);
    df.stride = std::max<int64_t>(16, (df.len + 15) / 16 * 16);
    df.names.resize(recs.size());
    for (size_t i = 0; i < recs.size(); ++i) df.names[i] = recs[i].name;
    if (on_framed) (*on_framed)();
    BTB_TRY(df.alloc(size_t(df.n * df.stride)));
    std::vector<int64_t> lens(recs.size(), 0);
    auto work = [&](int64_t lo, int64_t hi) {
        std::vector<uint8_t> tmp;
        for (int64_t i = lo; i < hi; ++i) {
            const FastaRecord &r = recs[size_t(i)];
            df.names[size_t(i)] = r.name;
            uint8_t *dst = df.rows + i * df.stride;
            tmp.resize(r.raw_len + 1);
            const int64_t l = walk_record(fb, r, tmp.data());
            lens[size_t(i)] = l;
            if (l != df.len) continue;
            memcpy(dst, tmp.data(), size_t(l));
            memset(dst + l, '.', size_t(df.stride - l));
        }
    };
    threads = std::max(1, std::min<int>(threads, 64));
    if (threads == 1 || df.n < 64) {
        work(0, df.n);
    } else {
        std::vector<std::thread> pool;
        const int64_t span = (df.n + threads - 1) / threads;
        for (int t = 0; t < threads; ++t) {
            int64_t lo = t * span, hi = std::min(df.n, lo + span);
            if (lo < hi) pool.emplace_back(work, lo, hi);
        }
        for (auto &th : pool) th.join();
    }
    for (int64_t i = 0; i < df.n; ++i)
        if (lens[size_t(i)] != df.len)
            return fail(BTB_EFORMAT, "ERROR: sequence #%lld '%s' has length %lld but expected %lld", (long long)(i + 1),
                        recs[size_t(i)].name.c_str(), (long long)lens[size_t(i)], (long long)df.len);
    return BTB_OK;
}

}  // namespace

namespace {
// Formats `rows` matrix rows (name, then sep + decimal per cell, then \n) into generously spaced
// slots of `work`: row r starts at off[r] and is used[r] bytes long.  host_threads workers.
size_t csv_slot_bytes(const char *name, int64_t n, int elem_bytes) { return strlen(name) + size_t(n) * (elem_bytes == 2 ? 6 : 11) + 1 + 8; }

void format_rows_spaced(const void *mat, int elem_bytes, int64_t ld, int64_t n, int64_t rows, const char *const *names,
                        char sep, int host_threads, char *work, const std::vector<size_t> &off, std::vector<size_t> &used) {
    auto job = [&](int64_t lo, int64_t hi) {
        for (int64_t r = lo; r < hi; ++r) {
            char *p = work + off[size_t(r)];
            const size_t nl = strlen(names[r]);
            memcpy(p, names[r], nl);
            p += nl;
            p += format_row(static_cast<const char *>(mat) + size_t(r) * size_t(ld) * size_t(elem_bytes), elem_bytes, n, sep, p);
            *p++ = '\n';
            used[size_t(r)] = size_t(p - (work + off[size_t(r)]));
        }
    };
    const int threads = std::max(1, std::min<int>(host_threads, 64));
    if (threads == 1 || rows < 2 * threads) { job(0, rows); return; }
    std::vector<std::thread> pool;
    const int64_t span = (rows + threads - 1) / threads;
    for (int t = 0; t < threads; ++t) {
        int64_t lo = t * span, hi = std::min(rows, lo + span);
        if (lo < hi) pool.emplace_back(job, lo, hi);
    }
    for (auto &th : pool) th.join();
}
// copies the formatted rows from their slots to their final, contiguous places dst + pos[r]
void gather_rows(char *dst, const char *work, const std::vector<size_t> &off, const std::vector<size_t> &used,
                 const std::vector<size_t> &pos, int64_t rows, int host_threads) {
    auto job = [&](int64_t lo, int64_t hi) {
        for (int64_t r = lo; r < hi; ++r) memcpy(dst + pos[size_t(r)], work + off[size_t(r)], used[size_t(r)]);
    };
    const int threads = std::max(1, std::min<int>(host_threads, 64));
    if (threads == 1 || rows < 2 * threads) { job(0, rows); return; }
    std::vector<std::thread> pool;
    const int64_t span = (rows + threads - 1) / threads;
    for (int t = 0; t < threads; ++t) {
        int64_t lo = t * span, hi = std::min(rows, lo + span);
        if (lo < hi) pool.emplace_back(job, lo, hi);
    }
    for (auto &th : pool) th.join();
}
}  // namespace

extern "C" int btb_format_csv_rows(const void *mat, int elem_bytes, int64_t ld, int64_t n, int64_t row0,
                                   int64_t row1, const char *const *names, char sep, int host_threads,
                                   char *buf, size_t cap, size_t *written) {
    if (!mat || !buf || !names || row0 < 0 || row1 < row0 || (elem_bytes != 2 && elem_bytes != 4))
        return fail(BTB_EINVAL, "btb_format_csv_rows: bad argument");
    const int64_t rows = row1 - row0;
    std::vector<size_t> off(size_t(rows) + 1, 0);
    for (int64_t r = 0; r < rows; ++r) off[size_t(r) + 1] = off[size_t(r)] + csv_slot_bytes(names[row0 + r], n, elem_bytes);
    std::vector<size_t> used(size_t(rows), 0), pos(size_t(rows), 0);
    // format into generously spaced slots (all threads), then gather the rows into place (all threads)
    std::vector<char> tmp(off[size_t(rows)]);
    format_rows_spaced(mat, elem_bytes, ld, n, rows, names + row0, sep, host_threads, tmp.data(), off, used);
    size_t w = 0;
    for (int64_t r = 0; r < rows; ++r) { pos[size_t(r)] = w; w += used[size_t(r)]; }
    if (w > cap) return fail(BTB_EINVAL, "btb_format_csv_rows: buffer too small (%zu needed)", w);
    gather_rows(buf, tmp.data(), off, used, pos, rows, host_threads);
    if (written) *written = w;
    return BTB_OK;
}

namespace {

// Multi-GPU distance stage inside one process: one host thread per device, the
// devices form an in-process btb_comm group -- each uploads and encodes 1/P of the samples, the operand
// shards travel GPU to GPU (peer copies), every device computes its block rows and copies them into the
// matrix H (n x ld elements) band by band.  H is host memory or -- level 1 -- the matrix buffer of the first
// device, which every device can write over NVLink: the CSV writer then serialises it exactly as in the
// one-GPU case and no n x n host matrix is ever allocated (page-locking 5 GB costs 2 s, a pageable
// target slows the copies down 6-fold; both measured).  No collective library.  See dist_comm.cu.
int multi_gpu_matrix(const DenseFasta &df, const std::vector<int> &devs, int allchars, int keepcase, int engine, int eb,
                     int64_t ld, void *H) {
    const int n_dev = int(devs.size());
    CommShared *sh = comm_shared_new(n_dev);
    if (!sh) return fail(BTB_ENOMEM, "out of host memory");
    std::vector<int> rcs(static_cast<size_t>(n_dev), BTB_OK);
    std::vector<std::string> errs(static_cast<size_t>(n_dev));
    std::vector<btb_comm *> comms(static_cast<size_t>(n_dev), nullptr);
    auto worker = [&](int k) {
        int rc = comm_create_local(sh, k, n_dev, devs[size_t(k)], &comms[size_t(k)]);
        if (rc == BTB_OK)
            rc = btb_dist_matrix_host_comm(comms[size_t(k)], df.rows, df.n, df.len, df.stride, allchars, keepcase, engine, H, eb, ld);
        rcs[size_t(k)] = rc;
        if (rc) errs[size_t(k)] = last_error();
    };
    std::vector<std::thread> pool;
    for (int k = 0; k < n_dev; ++k) pool.emplace_back(worker, k);
    for (auto &t : pool) t.join();
    pool.clear();
    // teardown waits for the peers, so it is collective as well
    for (int k = 0; k < n_dev; ++k) pool.emplace_back([&, k] { btb_comm_destroy(comms[size_t(k)]); });
    for (auto &t : pool) t.join();
    comm_shared_free(sh);
    for (int k = 0; k < n_dev; ++k)
        if (rcs[size_t(k)]) return fail(rcs[size_t(k)], "%s", errs[size_t(k)].c_str());
    return BTB_OK;
}

}  // namespace

namespace {
// '\n'-separated blob -> exactly `count` strings.
int split_labels(const char *labels, int64_t count, std::vector<std::string> &out) {
    out.clear();
    if (count <= 0) return labels && *labels ? fail(BTB_EINVAL, "labels given for zero rows") : BTB_OK;
    const char *p = labels;
    for (;;) {
        const char *nl = strchr(p, '\n');
        out.emplace_back(p, nl ? size_t(nl - p) : strlen(p));
        if (!nl) break;
        p = nl + 1;
    }
    if (int64_t(out.size()) != count)
        return fail(BTB_EINVAL, "%lld labels expected, %zu given", (long long)count, out.size());
    return BTB_OK;
}

int join_names(const std::vector<std::string> &names, char *buf, size_t cap, int64_t *n, size_t *needed) {
    size_t total = 1;
    for (auto &s : names) total += s.size() + 1;
    if (n) *n = int64_t(names.size());
    if (needed) *needed = total;
    if (!buf) return BTB_OK;
    if (cap < total) return fail(BTB_EINVAL, "name buffer too small: %zu bytes needed", total);
    char *p = buf;
    for (size_t i = 0; i < names.size(); ++i) {
        if (i) *p++ = '\n';
        memcpy(p, names[i].data(), names[i].size());
        p += names[i].size();
    }
    *p = 0;
    return BTB_OK;
}
}  // namespace

extern "C" int btb_process_sample_name(const char *name, char *out, size_t cap) {
    if (!name || !out) return fail(BTB_EINVAL, "btb_process_sample_name: NULL argument");
    const std::string l = process_sample_name(name);
    if (l.size() + 1 > cap) return fail(BTB_EINVAL, "label buffer too small: %zu bytes needed", l.size() + 1);
    memcpy(out, l.c_str(), l.size() + 1);
    return BTB_OK;
}

extern "C" int btb_fasta_names(const char *in_fasta, int host_threads, char *buf, size_t cap, int64_t *n, size_t *needed) {
    if (!in_fasta) return fail(BTB_EINVAL, "btb_fasta_names: NULL path");
    FileBytes fb;
    BTB_TRY(load_file(in_fasta, fb));
    std::vector<FastaRecord> recs;
    BTB_TRY(index_fasta(fb, recs, std::max(1, host_threads)));
    std::vector<std::string> names(recs.size());
    for (size_t i = 0; i < recs.size(); ++i) names[i] = recs[i].name;
    return join_names(names, buf, cap, n, needed);
}

extern "C" int btb_csv_row_names(const char *csv_path, char sep, char *buf, size_t cap, int64_t *n, size_t *needed) {
    if (!csv_path) return fail(BTB_EINVAL, "btb_csv_row_names: NULL path");
    FileBytes fb;
    BTB_TRY(load_file(csv_path, fb));
    std::vector<std::string> names;
    BTB_TRY(csv_row_names(fb, sep, names));
    return join_names(names, buf, cap, n, needed);
}

extern "C" int btb_relabel_csv(const char *in_csv, const char *out_csv, char sep, const char *labels, int64_t n_labels) {
    if (!in_csv || !out_csv || !labels) return fail(BTB_EINVAL, "btb_relabel_csv: NULL argument");
    std::vector<std::string> lab;
    BTB_TRY(split_labels(labels, n_labels, lab));
    std::string tmp;
    int rc;
    {
        FileBytes fb;                                    // unmapped before the rename below
        BTB_TRY(load_file(in_csv, fb));
        int fd = -1;
        BTB_TRY(open_output_temp(out_csv, tmp, &fd));
        FILE *fo = fdopen(fd, "wb");
        if (!fo) { close(fd); discard_output(tmp); return fail(BTB_EIO, "cannot open '%s' for writing", tmp.c_str()); }
        setvbuf(fo, nullptr, _IOFBF, 1 << 22);
        rc = relabel_csv(fb, sep, lab, fo);
        if (fclose(fo) != 0 && rc == BTB_OK) rc = fail(BTB_EIO, "closing '%s' failed", tmp.c_str());
    }
    if (rc != BTB_OK) { discard_output(tmp); return rc; }
    return commit_output(tmp, out_csv, g_replace_unlink_first != 0);     // in == out (post_process_snps_csv): one atomic rename
}


namespace {

bool write_all(int fd, const char *p, size_t left) {
    while (left) {
        const ssize_t k = write(fd, p, left);
        if (k < 0 && errno == EINTR) continue;
        if (k <= 0) return false;
        p += k;
        left -= size_t(k);
    }
    return true;
}

void parallel_rows(int64_t rows, int threads, const std::function<void(int64_t, int64_t)> &fn) {
    const int T = int(std::max<int64_t>(1, std::min<int64_t>(std::min(threads, 64), rows)));
    if (T == 1) { fn(0, rows); return; }
    std::vector<std::thread> pool;
    const int64_t span = (rows + T - 1) / T;
    for (int k = 0; k < T; ++k) {
        const int64_t lo = k * span, hi = std::min(rows, lo + span);
        if (lo < hi) pool.emplace_back(fn, lo, hi);
    }
    for (auto &th : pool) th.join();
}

// write() mode of the CSV writer: the formatting threads fill block b into slot b % kSlots while the one
// writer thread pushes earlier blocks out, in order (a file has one inode lock: more writers do not help).
struct TextRing {
    static constexpr int kSlots = 3;
    std::vector<char> buf[kSlots];
    size_t len[kSlots] = {0, 0, 0};
    int64_t filled = 0, written = 0;          // blocks handed to / finished by the writer
    bool failed = false, stop = false;
    std::mutex m;
    std::condition_variable cv;
    void init(size_t slot_bytes) { for (auto &b : buf) b.resize(slot_bytes); }
    char *acquire(int64_t b) {                // the slot of block b, once the writer is done with block b - kSlots
        std::unique_lock<std::mutex> lock(m);
        cv.wait(lock, [&] { return written > b - kSlots || failed || stop; });
        return (failed || stop) ? nullptr : buf[b % kSlots].data();
    }
    void submit(int64_t b, size_t bytes) {
        std::lock_guard<std::mutex> lock(m);
        len[b % kSlots] = bytes;
        filled = b + 1;
        cv.notify_all();
    }
    void run(int fd, int64_t blocks) {
        for (int64_t b = 0; b < blocks; ++b) {
            {
                std::unique_lock<std::mutex> lock(m);
                cv.wait(lock, [&] { return filled > b || stop; });
                if (filled <= b) return;
            }
            const bool ok = write_all(fd, buf[b % kSlots].data(), len[b % kSlots]);
            std::lock_guard<std::mutex> lock(m);
            if (!ok) failed = true;
            written = b + 1;
            cv.notify_all();
            if (!ok) return;
        }
    }
    bool finish(int64_t blocks) {
        std::unique_lock<std::mutex> lock(m);
        cv.wait(lock, [&] { return written >= blocks || failed || stop; });
        return !failed;
    }
    void abort() {
        std::lock_guard<std::mutex> lock(m);
        stop = true;
        cv.notify_all();
    }
};

// k5 on the device: bytes of text the cells of one matrix row take (one separator + the decimal digits per
// cell) -- one CTA per row, 16-byte loads.  HBM-bound: n * n * elem bytes read once.
template <typename T>
__global__ void __launch_bounds__(256) csv_row_len_kernel(const T *__restrict__ mat, int64_t ld, int64_t n, unsigned long long *__restrict__ out) {
    constexpr int kPer = 16 / int(sizeof(T));
    const T *row = mat + int64_t(blockIdx.x) * ld;
    unsigned int s = 0;
    for (int64_t c = int64_t(threadIdx.x) * kPer; c < n; c += 256 * kPer) {
        const uint4 v = *reinterpret_cast<const uint4 *>(row + c);          // ld is a multiple of 8 elements: 16-byte aligned, padded
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < kPer; ++k) {
            if (c + k >= n) break;
            uint32_t x;
            if constexpr (sizeof(T) == 2) x = (w[k >> 1] >> (16 * (k & 1))) & 0xffffu; else x = w[k];
            s += 1u + (x > 9u) + (x > 99u) + (x > 999u) + (x > 9999u);
            if constexpr (sizeof(T) == 4) s += (x > 99999u) + (x > 999999u) + (x > 9999999u) + (x > 99999999u) + (x > 999999999u);
        }
    }
    __shared__ unsigned int part[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        for (int k = 0; k < 8; ++k) t += part[k];
        out[blockIdx.x] = t + (unsigned long long)n;                          // + one separator per cell
    }
}

int csv_row_lengths(const void *d_mat, int eb, int64_t ld, int64_t n, unsigned long long *d_len, cudaStream_t st) {
    if (n <= 0) return BTB_OK;
    if (eb == 2) csv_row_len_kernel<uint16_t><<<unsigned(n), 256, 0, st>>>(static_cast<const uint16_t *>(d_mat), ld, n, d_len);
    else csv_row_len_kernel<uint32_t><<<unsigned(n), 256, 0, st>>>(static_cast<const uint32_t *>(d_mat), ld, n, d_len);
    BTB_CUDA(cudaGetLastError());
    return BTB_OK;
}

}  // namespace

static int snp_dists_impl(const char *in_fasta, const char *out_path, int allchars, int keepcase, int molten,
                          char sep, int corner, int quiet, int n_gpus, int host_threads, int engine,
                          const char *labels, int64_t n_labels, int64_t *n_samples, int64_t *seq_len);

extern "C" int btb_snp_dists(const char *in_fasta, const char *out_path, int allchars, int keepcase, int molten,
                             char sep, int corner, int quiet, int n_gpus, int host_threads, int engine,
                             int64_t *n_samples, int64_t *seq_len) {
    return snp_dists_impl(in_fasta, out_path, allchars, keepcase, molten, sep, corner, quiet, n_gpus, host_threads, engine,
                          nullptr, 0, n_samples, seq_len);
}

extern "C" int btb_snp_dists_labelled(const char *in_fasta, const char *out_path, int allchars, int keepcase, int molten,
                                      char sep, int corner, int quiet, int n_gpus, int host_threads, int engine,
                                      const char *labels, int64_t n_labels, int64_t *n_samples, int64_t *seq_len) {
    if (!labels) return fail(BTB_EINVAL, "btb_snp_dists_labelled: NULL labels");
    return snp_dists_impl(in_fasta, out_path, allchars, keepcase, molten, sep, corner, quiet, n_gpus, host_threads, engine,
                          labels, n_labels, n_samples, seq_len);
}

static int snp_dists_impl(const char *in_fasta, const char *out_path, int allchars, int keepcase, int molten,
                          char sep, int corner, int quiet, int n_gpus, int host_threads, int engine,
                          const char *labels, int64_t n_labels, int64_t *n_samples, int64_t *seq_len) {
    if (!in_fasta || !out_path) return fail(BTB_EINVAL, "btb_snp_dists: NULL path");
    const int threads = std::max(1, host_threads);
    if (!quiet) {
        fprintf(stderr, "This is snp-dists 0.8.2\n");
        fprintf(stderr, "Will use %d threads.\n", threads);
    }
    StageTimer timer("btb_snp_dists");
    BTB_TRY(require_device(first_usable_device()));
    timer.mark("device init");
    int rc = BTB_OK;
    DenseFasta df;
    cudaStream_t st = nullptr, st2 = nullptr;
    BTB_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    BTB_CUDA(cudaStreamCreateWithFlags(&st2, cudaStreamNonBlocking));
    uint8_t *d_seqs = nullptr, *d_ops = nullptr, *d_out = nullptr;
    unsigned long long *d_rowlen = nullptr;
    char *h_stage[2] = {nullptr, nullptr};
    const bool to_stdout = !strcmp(out_path, "-") || !strcmp(out_path, "/dev/stdout");
    int fd = to_stdout ? STDOUT_FILENO : -1;
    std::string tmp_path;
    char *map = nullptr;               // snps.csv written through a mapping (tmpfs, see below)
    size_t map_bytes = 0;
    std::thread prealloc;              // fallocate of the mapped file, beside the read and the distance stage
    std::atomic<size_t> prealloc_limit{0};
    TextRing ring;                     // write() mode: formatted blocks on their way to the writer thread
    std::thread writer;
    cudaEvent_t ev[2] = {nullptr, nullptr};
    auto stop_prealloc = [&]() {
        prealloc_limit = 0;
        if (prealloc.joinable()) prealloc.join();
    };
    auto cleanup = [&]() {
        ring.abort();
        if (writer.joinable()) writer.join();
        stop_prealloc();
        if (map) { munmap(map, map_bytes); map = nullptr; }
        cudaFree(d_seqs); cudaFree(d_ops); cudaFree(d_out); cudaFree(d_rowlen);
        if (h_stage[0]) cudaFreeHost(h_stage[0]);
        if (h_stage[1]) cudaFreeHost(h_stage[1]);
        if (ev[0]) cudaEventDestroy(ev[0]);
        if (ev[1]) cudaEventDestroy(ev[1]);
        cudaStreamDestroy(st); cudaStreamDestroy(st2);
        if (fd >= 0 && !to_stdout) { close(fd); fd = -1; discard_output(tmp_path); }
    };
#define L1_CUDA(call)                                                                                     \
    do {                                                                                                  \
        cudaError_t e_ = (call);                                                                          \
        if (e_ != cudaSuccess) {                                                                          \
            rc = fail(BTB_ECUDA, "%s failed: %s", #call, cudaGetErrorString(e_));                         \
            cleanup();                                                                                    \
            return rc;                                                                                    \
        }                                                                                                 \
    } while (0)
#define L1_TRY(call) do { rc = (call); if (rc != BTB_OK) { cleanup(); return rc; } } while (0)
#define L1_FAIL(...) do { rc = fail(__VA_ARGS__); cleanup(); return rc; } while (0)

    // ---- the output file is opened (and, on tmpfs, sized, mapped and preallocated by a helper thread) as
    //      soon as the record names are known: beside the read of the sequences and the distance stage.
    // Buffered write()s to one file serialise on its inode whatever the number of writers (measured on
    // the B200 boxes, tools/io_probe.cpp: 5.4 GB/s on the ext4 disk, 3.2 GB/s on tmpfs, for 1 or 16 writers;
    // 16 threads writing 16 FILES reach 32-36 GB/s).  On tmpfs a shared mapping of the preallocated file
    // takes 8.8 GB/s from 16 threads (page faults need no inode lock); on ext4 mapped writes are slower
    // than write() (4.5-5.1 GB/s: page_mkwrite journals every page).  So: tmpfs -> the file is sized to an
    // upper bound of the text, mapped, and preallocated by a helper thread (fallocate in chunks, stopping
    // at the exact text size once that is known); when the helper is done all host threads format their rows
    // straight into the mapping (formatting WHILE fallocate runs is three times slower, measured: 1.6 s
    // instead of 0.5 s -- page faults and fallocate contend -- so the two are not overlapped); the file is
    // cut to its real length at the end.  Anything else -> formatted blocks go through a ring to one writer
    // thread and write().
    std::string head;                  // what the output was prepared with
    int64_t prepared_n = -1;
    int prep_rc = BTB_OK;
    std::string prep_err;
    std::vector<std::string> label_names;
    if (labels) BTB_TRY(split_labels(labels, n_labels, label_names));
    auto prepare_output = [&]() -> int {
        const std::vector<std::string> &names = labels ? label_names : df.names;
        const int64_t n = df.n;
        const int eb = df.len < 65536 ? 2 : 4;
        if (labels && int64_t(label_names.size()) != n) return BTB_OK;          // reported by the caller below
        if (!to_stdout && fd < 0) BTB_TRY(open_output_temp(out_path, tmp_path, &fd));
        head.clear();
        size_t names_total = 0;
        if (!molten) {
            head = corner ? "snp-dists 0.8.2" : "";
            for (int64_t i = 0; i < n; ++i) { head.push_back(sep); head += names[size_t(i)]; names_total += names[size_t(i)].size(); }
            head.push_back('\n');
            if (!write_all(fd, head.data(), head.size())) return fail(BTB_EIO, "write to '%s' failed", to_stdout ? "stdout" : tmp_path.c_str());
        }
        prepared_n = n;
        const bool want_map = !to_stdout && !molten && n >= 1024 && (g_csv_mmap == 1 || (g_csv_mmap < 0 && on_memory_fs(out_path)));
        if (!want_map) return BTB_OK;
        const size_t bound = head.size() + names_total + size_t(n) + size_t(n) * size_t(n) * size_t(eb == 2 ? 6 : 11);
        if (ftruncate(fd, off_t(bound)) != 0) return BTB_OK;
        void *m = mmap(nullptr, bound, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
        if (m == MAP_FAILED) return ftruncate(fd, off_t(head.size())) == 0 ? BTB_OK : fail(BTB_EIO, "cannot size '%s'", tmp_path.c_str());
        map = static_cast<char *>(m);
        map_bytes = bound;
        prealloc_limit = bound;
        const int pfd = fd;
        const size_t from = head.size();
        prealloc = std::thread([pfd, from, &prealloc_limit] {
            constexpr size_t kChunk = size_t(128) << 20;
            for (size_t done = from; done < prealloc_limit.load(std::memory_order_acquire);) {
                const size_t len = std::min(kChunk, prealloc_limit.load(std::memory_order_acquire) - done);
                if (fallocate(pfd, 0, off_t(done), off_t(len)) != 0) break;       // unsupported: the formatters fault the pages in
                done += len;
            }
        });
        return BTB_OK;
    };
    const std::function<void()> on_framed = [&]() { prep_rc = prepare_output(); if (prep_rc != BTB_OK) prep_err = last_error(); };
    rc = read_dense(in_fasta, threads, df, &on_framed);
    timer.mark("read + parse snps.fas");
    if (rc != BTB_OK) {
        const std::string msg = last_error();
        if (!quiet) fprintf(stderr, "%s\n", msg.c_str());
        cleanup();
        return fail(rc, "%s", msg.c_str());
    }
    if (prep_rc != BTB_OK) { rc = fail(prep_rc, "%s", prep_err.c_str()); cleanup(); return rc; }
    if (!quiet) fprintf(stderr, "Read %lld sequences of length %lld\n", (long long)df.n, (long long)df.len);
    if (n_samples) *n_samples = df.n;
    if (seq_len) *seq_len = df.len;
    if (labels) {                                        // printed instead of the record names
        if (n_labels != df.n) L1_FAIL(BTB_EINVAL, "%lld labels given for %lld sequences", (long long)n_labels, (long long)df.n);
        df.names = label_names;
    }
    const int64_t n = df.n, len = df.len;
    const int eb = len < 65536 ? 2 : 4;
    const int64_t ld = (n + 7) / 8 * 8;
    {
        // the output was prepared from the names of the framing pass; if the exact parse that followed a
        // rejected prediction saw other records (it never should), prepare it again from scratch
        std::string want = molten ? std::string() : std::string(corner ? "snp-dists 0.8.2" : "");
        if (!molten) { for (int64_t i = 0; i < n; ++i) { want.push_back(sep); want += df.names[size_t(i)]; } want.push_back('\n'); }
        if (prepared_n != n || want != head) {
            stop_prealloc();
            if (map) { munmap(map, map_bytes); map = nullptr; map_bytes = 0; }
            if (!to_stdout && fd >= 0 && (ftruncate(fd, 0) != 0 || lseek(fd, 0, SEEK_SET) != 0)) L1_FAIL(BTB_EIO, "cannot rewind '%s'", tmp_path.c_str());
            if (to_stdout && prepared_n >= 0 && !head.empty()) L1_FAIL(BTB_EFORMAT, "the alignment changed while it was read");
            L1_TRY(prepare_output());
        }
    }
    std::vector<size_t> name_len(static_cast<size_t>(n));
    for (int64_t i = 0; i < n; ++i) name_len[size_t(i)] = df.names[size_t(i)].size();

    // ---- distance stage -------------------------------------------------------------------------------------
    std::vector<int> devs = worker_devices(n_gpus);
    const int n_dev = int(std::max<int64_t>(1, std::min<int64_t>(int64_t(devs.size()), tiles_per_dim(n) / kBand)));
    devs.resize(size_t(n_dev));
    std::vector<unsigned long long> cells_len(static_cast<size_t>(n), 0);       // text bytes of row r's cells (separators + digits)
    L1_CUDA(cudaMalloc(reinterpret_cast<void **>(&d_out), size_t(n * ld * eb)));
    L1_CUDA(cudaMalloc(reinterpret_cast<void **>(&d_rowlen), size_t(n) * 8));
    if (n_dev > 1) {
        // every device computes its block rows and writes them into the first device's matrix (NVLink)
        rc = multi_gpu_matrix(df, devs, allchars, keepcase, engine, eb, ld, d_out);
        if (rc != BTB_OK) { cleanup(); return rc; }
        L1_TRY(require_device(devs[0]));
        timer.mark("sharded H2D + encode + exchange + distance kernels + assembly on device 0");
    } else {
        uint8_t table[256];
        int sigma = 4;
        L1_CUDA(cudaMalloc(reinterpret_cast<void **>(&d_seqs), size_t(n * df.stride)));
        L1_CUDA(cudaMemcpyAsync(d_seqs, df.rows, size_t(n * df.stride), cudaMemcpyHostToDevice, st));
        L1_TRY(btb_dist_alphabet(d_seqs, n, len, df.stride, allchars, keepcase, table, &sigma, st));
        const int eng = pick_engine(engine, len, btb_dist_operand_bytes(BTB_ENGINE_UMMA, n, len, sigma));
        L1_CUDA(cudaMalloc(reinterpret_cast<void **>(&d_ops), size_t(std::max<int64_t>(16, btb_dist_operand_bytes(eng, n, len, sigma)))));
        L1_TRY(btb_dist_encode(eng, d_seqs, n, len, df.stride, table, sigma, d_ops, st));
        L1_TRY(btb_distance_tiles(eng, d_ops, n, len, sigma, 0, tile_count(n), d_out, eb, ld, 0, 1, st));
    }
    // k5, first half on the device: the exact text length of every row (one pass over the matrix at HBM
    // speed) -- every row's place in the file is known before a single byte is formatted
    L1_TRY(csv_row_lengths(d_out, eb, ld, n, d_rowlen, st));
    L1_CUDA(cudaMemcpyAsync(cells_len.data(), d_rowlen, size_t(n) * 8, cudaMemcpyDeviceToHost, st));
    L1_CUDA(cudaStreamSynchronize(st));
    if (n_dev == 1) timer.mark("H2D + encode + distance kernel + row text lengths");
    cudaFree(d_seqs); d_seqs = nullptr;
    cudaFree(d_ops); d_ops = nullptr;
    std::vector<size_t> row_pos(static_cast<size_t>(n) + 1);                      // file offset of every row
    row_pos[0] = head.size();
    for (int64_t r = 0; r < n; ++r) row_pos[size_t(r) + 1] = row_pos[size_t(r)] + name_len[size_t(r)] + size_t(cells_len[size_t(r)]) + 1;
    const size_t total = row_pos[size_t(n)];
    if (map) {
        if (total > map_bytes) L1_FAIL(BTB_EIO, "internal error: snps.csv text exceeds its bound");
        prealloc_limit = total;                         // the helper stops at the real size
    } else if (!to_stdout && !molten && g_csv_prealloc && total > (size_t(64) << 20)) {
        // ext4 & co: blocks allocated up front (unwritten extents) -- no delayed allocation is pending when the
        // file is renamed over an existing snps.csv, which would otherwise make the rename flush all of it
        if (fallocate(fd, 0, 0, off_t(total)) != 0) { /* not supported here: fine */ }
    }

    // ---- serialise: D2H row blocks (double-buffered) -> all host threads format rows in their final place
    const int64_t block_rows = std::max<int64_t>(1, std::min<int64_t>(n, (int64_t(64) << 20) / std::max<int64_t>(1, ld * eb)));
    const int64_t blocks = (n + block_rows - 1) / block_rows;
    L1_CUDA(cudaMallocHost(reinterpret_cast<void **>(&h_stage[0]), size_t(block_rows * ld * eb)));
    L1_CUDA(cudaMallocHost(reinterpret_cast<void **>(&h_stage[1]), size_t(block_rows * ld * eb)));
    L1_CUDA(cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming));
    L1_CUDA(cudaEventCreateWithFlags(&ev[1], cudaEventDisableTiming));
    if (!molten && !map) {
        size_t max_block = 0;
        for (int64_t b = 0; b < blocks; ++b)
            max_block = std::max(max_block, row_pos[size_t(std::min(n, (b + 1) * block_rows))] - row_pos[size_t(b * block_rows)]);
        ring.init(max_block + 64);
        const int wfd = fd;
        writer = std::thread([&ring, wfd, blocks] { ring.run(wfd, blocks); });
    }
    double t_wait = 0, t_format = 0, t_ring = 0;
    auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    auto issue = [&](int64_t b) -> cudaError_t {
        const int64_t r0 = b * block_rows, r1 = std::min(n, r0 + block_rows);
        cudaError_t e = cudaMemcpyAsync(h_stage[b & 1], d_out + size_t(r0 * ld * eb), size_t((r1 - r0) * ld * eb),
                                        cudaMemcpyDeviceToHost, st2);
        if (e != cudaSuccess) return e;
        return cudaEventRecord(ev[b & 1], st2);
    };
    L1_CUDA(issue(0));
    for (int64_t b = 0; b < blocks; ++b) {
        const int64_t r0 = b * block_rows, r1 = std::min(n, r0 + block_rows), rows = r1 - r0;
        { const double tw = now(); L1_CUDA(cudaEventSynchronize(ev[b & 1])); t_wait += now() - tw; }
        if (b + 1 < blocks) L1_CUDA(issue(b + 1));                // into the other staging buffer: block b - 1 is formatted
        const char *src_rows = h_stage[b & 1];
        if (molten) {
            std::string text;
            for (int64_t r = r0; r < r1; ++r)
                for (int64_t i = 0; i < n; ++i) {
                    const uint32_t v = eb == 2 ? reinterpret_cast<const uint16_t *>(src_rows)[(r - r0) * ld + i]
                                               : reinterpret_cast<const uint32_t *>(src_rows)[(r - r0) * ld + i];
                    text += df.names[size_t(r)]; text.push_back(sep); text += df.names[size_t(i)]; text.push_back(sep);
                    text += std::to_string(v); text.push_back('\n');
                    if (text.size() > (size_t(8) << 20)) { if (!write_all(fd, text.data(), text.size())) L1_FAIL(BTB_EIO, "write to '%s' failed", out_path); text.clear(); }
                }
            if (!write_all(fd, text.data(), text.size())) L1_FAIL(BTB_EIO, "write to '%s' failed", out_path);
            continue;
        }
        char *dst = nullptr;                                      // where file offset row_pos[r0] lives
        if (map) {
            dst = map + row_pos[size_t(r0)];
            if (prealloc.joinable()) { const double tr = now(); prealloc.join(); t_ring += now() - tr; }    // see above: not overlapped
        } else {
            const double tr = now();
            dst = ring.acquire(b);
            t_ring += now() - tr;
            if (!dst) L1_FAIL(BTB_EIO, "write to '%s' failed", to_stdout ? "stdout" : tmp_path.c_str());
        }
        const double tf = now();
        std::atomic<bool> bad_len{false};
        parallel_rows(rows, threads, [&](int64_t lo, int64_t hi) {
            for (int64_t r = lo; r < hi; ++r) {
                const size_t at = row_pos[size_t(r0 + r)] - row_pos[size_t(r0)];
                char *q = dst + at;
                memcpy(q, df.names[size_t(r0 + r)].data(), name_len[size_t(r0 + r)]);
                q += name_len[size_t(r0 + r)];
                q += format_row(src_rows + size_t(r) * size_t(ld) * size_t(eb), eb, n, sep, q);
                *q++ = '\n';
                if (size_t(q - dst) != row_pos[size_t(r0 + r) + 1] - row_pos[size_t(r0)]) bad_len = true;
            }
        });
        t_format += now() - tf;
        if (bad_len) L1_FAIL(BTB_EIO, "internal error: a formatted row does not have its predicted length");
        if (!map) ring.submit(b, row_pos[size_t(r1)] - row_pos[size_t(r0)]);
    }
    if (!molten && !map) {
        const double tr = now();
        const bool ok = ring.finish(blocks);
        t_ring += now() - tr;
        if (writer.joinable()) writer.join();
        if (!ok) L1_FAIL(BTB_EIO, "write to '%s' failed", to_stdout ? "stdout" : tmp_path.c_str());
    }
    if (map) {
        stop_prealloc();
        {   // tearing down 3 million page-table entries takes 0.1-0.2 s and nothing waits for it
            char *m = map;
            const size_t mb = map_bytes;
            map = nullptr;
            std::thread([m, mb] { munmap(m, mb); }).detach();
        }
        if (ftruncate(fd, off_t(total)) != 0) L1_FAIL(BTB_EIO, "cannot size '%s'", tmp_path.c_str());
    }
    timer.mark("D2H + format + write snps.csv");
    if (timer.on) fprintf(stderr, "[btb timing]     (%s: wait for D2H %.3f s, format rows in place %.3f s, wait for the %s %.3f s)\n",
                          map_bytes ? "mapped file" : "write()", t_wait, t_format, map_bytes ? "preallocating helper" : "writer thread", t_ring);
    if (!to_stdout) {
        const int cfd = fd;
        fd = -1;
        if (close(cfd) != 0) { discard_output(tmp_path); L1_FAIL(BTB_EIO, "closing '%s' failed", tmp_path.c_str()); }
        rc = commit_output(tmp_path, out_path, g_replace_unlink_first != 0);
        if (rc != BTB_OK) { cleanup(); return rc; }
    }
    timer.mark("close + rename");
    cleanup();
    timer.mark("free device + pinned buffers");
    return BTB_OK;
#undef L1_CUDA
#undef L1_TRY
#undef L1_FAIL
}
