// Multi-GPU distance stage with a sharded upload: the ranks of one node --
// one process per GPU (shared-memory rendezvous + CUDA IPC) or one thread per GPU inside a process --
// each upload and encode 1/P of the samples, exchange the encoded operand shards GPU to GPU over
// NVLink (peer copies issued by the consumer: no NCCL, no host staging), compute their block rows of
// the matrix and copy them to the host band by band while later bands are still being computed.
//
// Reference call site: btbphylo/phylogeny.py:149 `snp-dists -c -j T snps.fas > snps.csv` -- the load
// loop (here: upload + encode, sharded) and the O(N^2 L) loop nest (here: the tcgen05 kernels of
// dist_umma.cu on this rank's share).
//
// Protocol of one call (epoch e), rank r of P, shared control block `sh`:
//   wait until every peer has finished reading my buffer in epoch e-1          (slot.pulled >= e-1)
//   H2D of my sample shard, byte-presence scan                 -> slot.present -> barrier A
//   all ranks derive the same symbol table / sigma from the OR of the presence maps
//   (re)allocate my exchange buffer [operands | row sums | site counts], publish its IPC handle,
//   per-site symbol counts of my shard                          -> barrier B
//   map the peers' buffers, majority symbol per site from the SUM of all ranks' counts (read over
//   NVLink), encode my shard into my buffer                    -> slot.ready = e
//   for every peer q that holds rows I need: wait slot[q].ready >= e, copy its rows (NVLink)
//   slot.pulled = e; distance kernels band by band, each band's rows copied to the host matrix
#include <fcntl.h>
#include <sched.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstring>
#include <string>
#include <vector>

#include "common.cuh"

namespace btb {

int require_device(int device);
int pick_engine(int engine, int64_t len, int64_t extra_bytes);
int dist_presence_async(const uint8_t *, int64_t, int64_t, int64_t, uint32_t *, cudaStream_t);
int dist_alphabet_from_presence(const uint32_t present[8], bool scanned, int allchars, int keepcase, uint8_t table[256], int *sigma);
int dist_site_counts(const uint8_t *d_seqs, int64_t rows, int64_t len, int64_t row_stride, const uint8_t *d_table,
                     uint32_t *d_counts, cudaStream_t stream);
int dist_site_major_sum(const CountPtrs &cp, int64_t len, uint8_t *d_major, cudaStream_t stream);
int dist_encode_fp4_rows(const uint8_t *d_seqs_rows, int64_t rows, int64_t len, int64_t row_stride, const uint8_t table[256],
                         const uint8_t *d_table, int sigma, uint8_t *A_rows, int32_t *rowsum_rows, const uint8_t *d_major,
                         cudaStream_t stream);
extern int g_umma_major_zero;

int g_comm_band_rows = 16;        // "comm_band_rows": block rows per launch + D2H band of a rank's share
int g_comm_last_path = -1;        // "comm_last_path": 1 sharded upload + GPU-to-GPU exchange, 0 replicated upload (fallback)

constexpr uint32_t kCommMagic = 0x43425442u;   // "BTBC"
constexpr int kCommMaxRanks = 16;

struct alignas(128) CommSlot {
    cudaIpcMemHandle_t handle;       // of this rank's exchange buffer (inter-process mode)
    void *local_ptr;                 // the same buffer as a plain device pointer (in-process mode)
    uint64_t buf_gen;                // bumped whenever the buffer is reallocated: peers map it again
    uint64_t buf_bytes;
    std::atomic<uint64_t> ready;     // last epoch whose encoded shard is complete in the buffer
    std::atomic<uint64_t> pulled;    // last epoch in which this rank finished reading its peers' buffers
    uint32_t present[8];             // byte values that occur in this rank's shard
    int32_t device;
};

struct CommShared {
    std::atomic<uint32_t> magic;
    uint32_t world;
    std::atomic<uint32_t> bar_arrived;
    std::atomic<uint32_t> bar_gen;
    std::atomic<uint32_t> aborted;   // a rank failed: everybody waiting gives up
    std::atomic<uint32_t> attached;
    CommSlot slot[kCommMaxRanks];
};

static double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

}  // namespace btb

using namespace btb;

struct btb_comm {
    int rank = 0, world = 1, device = 0;
    bool ipc = false;                  // shm + CUDA IPC (processes) or plain pointers (threads of one process)
    CommShared *sh = nullptr;
    std::string shm_name;
    bool owns_shared = false;          // in-process mode: rank 0 frees the control block
    uint64_t epoch = 0;
    double timeout_s = 120.0;
    // my exchange buffer and what is retired (freed once no peer can still have it mapped)
    uint8_t *buf = nullptr;
    size_t buf_bytes = 0;
    std::vector<void *> retired;
    // the peers' buffers as mapped here
    void *peer_ptr[kCommMaxRanks] = {nullptr};
    uint64_t peer_gen[kCommMaxRanks] = {0};
    // private device buffers
    uint8_t *d_seqs = nullptr; size_t seqs_cap = 0;
    uint8_t *d_out = nullptr; size_t out_cap = 0;
    uint8_t *d_small = nullptr;        // table 256 | presence 32 | major (padded len)
    size_t small_cap = 0;
    uint32_t *h_present = nullptr;     // pinned
    cudaStream_t st = nullptr, st_out = nullptr, st_pull[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev_band = nullptr, ev_pull[4] = {nullptr, nullptr, nullptr, nullptr};
    double t_phase[8] = {0, 0, 0, 0, 0, 0, 0, 0};
};

namespace {

void cpu_relax(int spins) {
    if (spins < 2000) { __builtin_ia32_pause(); return; }
    if (spins < 20000) { sched_yield(); return; }
    usleep(50);
}

int comm_abort(btb_comm *c, int rc) {
    if (c && c->sh) c->sh->aborted.store(1, std::memory_order_release);
    return rc;
}

// waits until pred() holds; fails when a peer aborted or the timeout expires
template <typename Pred>
int comm_wait(btb_comm *c, const char *what, Pred pred) {
    const double t0 = now_s();
    for (int spins = 0;; ++spins) {
        if (pred()) return BTB_OK;
        if (c->sh->aborted.load(std::memory_order_acquire)) return fail(BTB_ECOMM, "rank %d: a peer rank failed while waiting for %s", c->rank, what);
        if ((spins & 1023) == 1023 && now_s() - t0 > c->timeout_s)
            return comm_abort(c, fail(BTB_ECOMM, "rank %d: timed out after %.0f s waiting for %s", c->rank, c->timeout_s, what));
        cpu_relax(spins);
    }
}

int comm_barrier(btb_comm *c, const char *what) {
    CommShared *sh = c->sh;
    const uint32_t gen = sh->bar_gen.load(std::memory_order_acquire);
    if (sh->bar_arrived.fetch_add(1, std::memory_order_acq_rel) + 1 == uint32_t(c->world)) {
        sh->bar_arrived.store(0, std::memory_order_relaxed);
        sh->bar_gen.store(gen + 1, std::memory_order_release);
        return BTB_OK;
    }
    return comm_wait(c, what, [&] { return sh->bar_gen.load(std::memory_order_acquire) != gen; });
}

int ensure(uint8_t **p, size_t *cap, size_t bytes) {
    bytes = std::max<size_t>(bytes, 256);
    if (*cap >= bytes) return BTB_OK;
    if (*p) cudaFree(*p);
    *p = nullptr; *cap = 0;
    if (cudaMalloc(reinterpret_cast<void **>(p), bytes) != cudaSuccess) {
        cudaGetLastError();
        return fail(BTB_ENOMEM, "out of device memory (%zu bytes)", bytes);
    }
    *cap = bytes;
    return BTB_OK;
}

int comm_init_streams(btb_comm *c) {
    BTB_TRY(require_device(c->device));
    BTB_CUDA(cudaStreamCreateWithFlags(&c->st, cudaStreamNonBlocking));
    BTB_CUDA(cudaStreamCreateWithFlags(&c->st_out, cudaStreamNonBlocking));
    BTB_CUDA(cudaEventCreateWithFlags(&c->ev_band, cudaEventDisableTiming));
    for (int k = 0; k < 4; ++k) {
        BTB_CUDA(cudaStreamCreateWithFlags(&c->st_pull[k], cudaStreamNonBlocking));
        BTB_CUDA(cudaEventCreateWithFlags(&c->ev_pull[k], cudaEventDisableTiming));
    }
    BTB_CUDA(cudaMallocHost(reinterpret_cast<void **>(&c->h_present), 32));
    if (const char *t = getenv("BTB_COMM_TIMEOUT_S")) c->timeout_s = std::max(1.0, atof(t));
    return BTB_OK;
}

void comm_unmap_peers(btb_comm *c) {
    for (int q = 0; q < c->world; ++q) {
        if (c->ipc && c->peer_ptr[q] && q != c->rank) cudaIpcCloseMemHandle(c->peer_ptr[q]);
        c->peer_ptr[q] = nullptr;
        c->peer_gen[q] = 0;
    }
}

// the rows [r0, r1) of every per-row array of the FP4 operand layout: operand rows, then (dense) the int32 row sums
struct RowArray { size_t offset, pitch; };
int operand_arrays(int64_t n, int64_t len, int sigma, RowArray out[2]) {
    const int64_t kb = umma_k_bytes(len, sigma);
    out[0] = {0, size_t(kb)};
    if (sigma_dense(sigma)) { out[1] = {size_t(n * kb), 4}; return 2; }
    return 1;
}

}  // namespace

namespace btb {

// in-process group: `world` members that share one heap-allocated control block (multi_gpu_matrix)
CommShared *comm_shared_new(int world) {
    void *mem = nullptr;
    if (posix_memalign(&mem, 128, sizeof(CommShared)) != 0) return nullptr;
    memset(mem, 0, sizeof(CommShared));
    CommShared *sh = new (mem) CommShared;
    sh->world = uint32_t(world);
    sh->magic.store(kCommMagic, std::memory_order_release);
    return sh;
}
void comm_shared_free(CommShared *sh) { free(sh); }

int comm_create_local(CommShared *sh, int rank, int world, int device, btb_comm **out) {
    if (!sh || world < 1 || world > kCommMaxRanks || rank < 0 || rank >= world) return fail(BTB_EINVAL, "comm_create_local: bad argument");
    btb_comm *c = new btb_comm;
    c->rank = rank; c->world = world; c->device = device; c->ipc = false; c->sh = sh;
    int rc = comm_init_streams(c);
    if (rc != BTB_OK) { delete c; return rc; }
    sh->slot[rank].device = device;
    *out = c;
    return BTB_OK;
}

}  // namespace btb

extern "C" int btb_comm_create(const char *name, int rank, int world, int device, btb_comm **out) {
    if (!name || !*name || !out || world < 1 || world > kCommMaxRanks || rank < 0 || rank >= world)
        return fail(BTB_EINVAL, "btb_comm_create: bad argument (1 <= world <= %d)", kCommMaxRanks);
    BTB_TRY(require_device(device));
    std::string shm = name[0] == '/' ? std::string(name) : "/" + std::string(name);
    int fd = -1;
    const double t0 = now_s();
    double create_timeout = 120.0;
    if (const char *t = getenv("BTB_COMM_TIMEOUT_S")) create_timeout = std::max(1.0, atof(t));
    if (rank == 0) {
        shm_unlink(shm.c_str());                      // a stale segment of a crashed job with the same name
        fd = shm_open(shm.c_str(), O_CREAT | O_EXCL | O_RDWR, 0600);
        if (fd < 0) return fail(BTB_EIO, "btb_comm_create: shm_open('%s') failed: %s", shm.c_str(), strerror(errno));
        if (ftruncate(fd, off_t(sizeof(CommShared))) != 0) { close(fd); shm_unlink(shm.c_str()); return fail(BTB_EIO, "btb_comm_create: cannot size '%s'", shm.c_str()); }
    } else {
        for (;;) {                                    // rank 0 creates it; the others wait for it to appear, sized
            fd = shm_open(shm.c_str(), O_RDWR, 0600);
            struct stat sb;
            if (fd >= 0 && fstat(fd, &sb) == 0 && size_t(sb.st_size) >= sizeof(CommShared)) break;
            if (fd >= 0) { close(fd); fd = -1; }
            if (now_s() - t0 > create_timeout) return fail(BTB_ECOMM, "btb_comm_create: rank %d never saw '%s' (is rank 0 running?)", rank, shm.c_str());
            usleep(1000);
        }
    }
    void *mem = mmap(nullptr, sizeof(CommShared), PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (mem == MAP_FAILED) return fail(BTB_EIO, "btb_comm_create: mmap of '%s' failed", shm.c_str());
    CommShared *sh = static_cast<CommShared *>(mem);
    if (rank == 0) {
        sh->world = uint32_t(world);
        sh->magic.store(kCommMagic, std::memory_order_release);
    }
    btb_comm *c = new btb_comm;
    c->rank = rank; c->world = world; c->device = device; c->ipc = true; c->sh = sh; c->shm_name = shm;
    int rc = comm_init_streams(c);
    if (rc == BTB_OK)
        rc = comm_wait(c, "rank 0 to initialise the control block", [&] { return sh->magic.load(std::memory_order_acquire) == kCommMagic; });
    if (rc == BTB_OK && int(sh->world) != world) rc = fail(BTB_EINVAL, "btb_comm_create: '%s' belongs to a job of %u ranks", shm.c_str(), sh->world);
    if (rc == BTB_OK) {
        sh->slot[rank].device = device;
        sh->attached.fetch_add(1, std::memory_order_acq_rel);
        // everybody attached: the name can go (the mapping lives on), so nothing is left behind in /dev/shm
        rc = comm_wait(c, "all ranks to attach", [&] { return sh->attached.load(std::memory_order_acquire) >= uint32_t(world); });
        if (rank == 0) shm_unlink(shm.c_str());
    }
    if (rc != BTB_OK) { munmap(mem, sizeof(CommShared)); delete c; return rc; }
    *out = c;
    return BTB_OK;
}

extern "C" int btb_comm_destroy(btb_comm *c) {
    if (!c) return BTB_OK;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    comm_unmap_peers(c);
    // nobody may still be reading my buffer: the peers' last pulls finished before they set `pulled`
    if (c->sh && c->world > 1 && !c->sh->aborted.load())
        comm_wait(c, "the peers to finish the last epoch", [&] {
            for (int q = 0; q < c->world; ++q)
                if (c->sh->slot[q].pulled.load(std::memory_order_acquire) < c->epoch) return false;
            return true;
        });
    for (void *p : c->retired) cudaFree(p);
    if (c->buf) cudaFree(c->buf);
    if (c->d_seqs) cudaFree(c->d_seqs);
    if (c->d_out) cudaFree(c->d_out);
    if (c->d_small) cudaFree(c->d_small);
    if (c->h_present) cudaFreeHost(c->h_present);
    if (c->st) cudaStreamDestroy(c->st);
    if (c->st_out) cudaStreamDestroy(c->st_out);
    if (c->ev_band) cudaEventDestroy(c->ev_band);
    for (int k = 0; k < 4; ++k) {
        if (c->st_pull[k]) cudaStreamDestroy(c->st_pull[k]);
        if (c->ev_pull[k]) cudaEventDestroy(c->ev_pull[k]);
    }
    if (c->ipc && c->sh) munmap(c->sh, sizeof(CommShared));
    delete c;
    return BTB_OK;
}

extern "C" int btb_comm_last_timings(const btb_comm *c, double *seconds, int cap) {
    if (!c || !seconds) return fail(BTB_EINVAL, "btb_comm_last_timings: NULL argument");
    for (int k = 0; k < cap && k < 8; ++k) seconds[k] = c->t_phase[k];
    return BTB_OK;
}

namespace btb {

// This rank's block rows [lo, hi) of the matrix: distance kernels band by band, every finished band's
// cells -- the rows from the diagonal on and the mirrored strip below -- copied to the matrix `out`
// while the next band is computed.  d_out is the rank's full-size device matrix; `out` is host memory
// (level 2: the job's one host matrix) or a peer-accessible device matrix (level 1, several GPUs in one
// process: the shares are assembled on the first device, over NVLink, and serialised from there).
int share_compute_and_copy(int eng, const void *d_ops, int64_t n, int64_t len, int sigma, int64_t lo, int64_t hi,
                           uint8_t *d_out, int eb, int64_t d_ld, void *out, int64_t out_ld, cudaStream_t st,
                           cudaStream_t st_out, cudaEvent_t ev) {
    char *ho = static_cast<char *>(out);
    auto copy2d = [&](int64_t r0, int64_t r1, int64_t c0, int64_t c1) -> cudaError_t {
        if (r1 <= r0 || c1 <= c0) return cudaSuccess;
        return cudaMemcpy2DAsync(ho + size_t(r0 * out_ld + c0) * size_t(eb), size_t(out_ld) * size_t(eb),
                                 d_out + size_t(r0 * d_ld + c0) * size_t(eb), size_t(d_ld) * size_t(eb),
                                 size_t(c1 - c0) * size_t(eb), size_t(r1 - r0), cudaMemcpyDefault, st_out);
    };
    const int64_t step = std::max<int64_t>(2, int64_t(g_comm_band_rows) / 2 * 2);
    for (int64_t b0 = lo; b0 < hi; b0 += step) {
        const int64_t b1 = std::min(hi, b0 + step);
        BTB_TRY(btb_distance_rows(eng, d_ops, n, len, sigma, b0, b1, d_out, eb, d_ld, 1, st));
        BTB_CUDA(cudaEventRecord(ev, st));
        BTB_CUDA(cudaStreamWaitEvent(st_out, ev, 0));
        const int64_t row0 = std::min<int64_t>(n, b0 * BTB_TILE), row1 = std::min<int64_t>(n, b1 * BTB_TILE);
        BTB_CUDA(copy2d(row0, row1, row0, n));
        BTB_CUDA(copy2d(row1, n, row0, row1));
    }
    BTB_CUDA(cudaStreamSynchronize(st_out));
    BTB_CUDA(cudaStreamSynchronize(st));
    return BTB_OK;
}

}  // namespace btb

static int comm_matrix_impl(btb_comm *c, const uint8_t *seqs, int64_t n, int64_t len, int64_t row_stride, int allchars,
                            int keepcase, int engine, void *out, int eb, int64_t out_ld, bool *done) {
    *done = false;
    CommShared *sh = c->sh;
    const int P = c->world, r = c->rank;
    const uint64_t e = ++c->epoch;
    double t0 = now_s();
    auto lap = [&](int k) { const double t = now_s(); c->t_phase[k] = t - t0; t0 = t; };
    BTB_TRY(require_device(c->device));

    // nobody may still be reading the buffer I am about to overwrite (or free)
    BTB_TRY(comm_wait(c, "the peers to finish the previous epoch", [&] {
        for (int q = 0; q < P; ++q)
            if (sh->slot[q].pulled.load(std::memory_order_acquire) + 1 < e) return false;
        return true;
    }));
    for (void *p : c->retired) cudaFree(p);
    c->retired.clear();

    // ---- my sample shard: upload + byte presence ----------------------------------------------------
    const int64_t s0 = n * r / P, s1 = n * (r + 1) / P, rows = s1 - s0;
    const int64_t d_stride = (len + 15) / 16 * 16, d_ld = (n + 7) / 8 * 8;
    BTB_TRY(ensure(&c->d_seqs, &c->seqs_cap, size_t(std::max<int64_t>(1, rows) * d_stride)));
    const int64_t major_pad = (len + 31) / 32 * 32;
    BTB_TRY(ensure(&c->d_small, &c->small_cap, size_t(512 + major_pad)));
    uint8_t *d_table = c->d_small;
    uint32_t *d_present = reinterpret_cast<uint32_t *>(c->d_small + 256);
    uint8_t *d_major = c->d_small + 512;
    BTB_CUDA(cudaMemsetAsync(d_present, 0, 32, c->st));
    if (rows > 0)
        BTB_CUDA(cudaMemcpy2DAsync(c->d_seqs, size_t(d_stride), seqs + s0 * row_stride, size_t(row_stride), size_t(len),
                                   size_t(rows), cudaMemcpyHostToDevice, c->st));
    BTB_TRY(dist_presence_async(c->d_seqs, rows, len, d_stride, d_present, c->st));
    BTB_CUDA(cudaMemcpyAsync(c->h_present, d_present, 32, cudaMemcpyDeviceToHost, c->st));
    BTB_CUDA(cudaStreamSynchronize(c->st));
    memcpy(sh->slot[r].present, c->h_present, 32);
    lap(0);
    BTB_TRY(comm_barrier(c, "barrier A (byte presence)"));
    lap(1);

    // ---- the same table and sigma on every rank -------------------------------------------------------
    uint32_t present[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int q = 0; q < P; ++q)
        for (int k = 0; k < 8; ++k) present[k] |= sh->slot[q].present[k];
    uint8_t table[256];
    int sigma = 4;
    BTB_TRY(dist_alphabet_from_presence(present, n * len > 0, allchars, keepcase, table, &sigma));
    if (engine == BTB_ENGINE_POPC || len == 0 || !umma_use_fp4(sigma, len)) {
        // layouts this path does not exchange (bit-planes, int8 operand pairs): replicated upload.  Every
        // rank takes this branch together (same sigma, same len); nothing of the buffers was touched.
        sh->slot[r].pulled.store(e, std::memory_order_release);
        return BTB_OK;
    }
    const int64_t kb = umma_k_bytes(len, sigma);
    const size_t ops_bytes = size_t(btb_dist_operand_bytes(BTB_ENGINE_UMMA, n, len, sigma));
    const size_t counts_off = (ops_bytes + 255) & ~size_t(255);
    const bool dense = sigma_dense(sigma);
    const bool use_major = dense && g_umma_major_zero && n >= 64;
    const size_t need = counts_off + (use_major ? size_t(len) * 16 : 0) + 256;
    if (c->buf_bytes < need) {
        if (c->buf) c->retired.push_back(c->buf);      // a peer may still have it mapped: freed next epoch
        c->buf = nullptr; c->buf_bytes = 0;
        if (cudaMalloc(reinterpret_cast<void **>(&c->buf), need) != cudaSuccess) {
            cudaGetLastError();
            return comm_abort(c, fail(BTB_ENOMEM, "rank %d: out of device memory for %zu bytes of operands", r, need));
        }
        c->buf_bytes = need;
        if (c->ipc) {
            cudaError_t ce = cudaIpcGetMemHandle(&sh->slot[r].handle, c->buf);
            if (ce != cudaSuccess) return comm_abort(c, fail(BTB_ECUDA, "cudaIpcGetMemHandle failed: %s", cudaGetErrorString(ce)));
        }
        sh->slot[r].local_ptr = c->buf;
        sh->slot[r].buf_bytes = need;
        sh->slot[r].buf_gen += 1;
    }
    BTB_TRY(ensure(&c->d_out, &c->out_cap, size_t(n * d_ld) * size_t(eb)));
    BTB_CUDA(cudaMemcpyAsync(d_table, table, 256, cudaMemcpyHostToDevice, c->st));
    uint32_t *my_counts = reinterpret_cast<uint32_t *>(c->buf + counts_off);
    if (use_major) BTB_TRY(dist_site_counts(c->d_seqs, rows, len, d_stride, d_table, my_counts, c->st));
    BTB_CUDA(cudaStreamSynchronize(c->st));
    lap(2);
    BTB_TRY(comm_barrier(c, "barrier B (buffers + site counts)"));
    lap(3);

    // ---- map the peers' buffers ----------------------------------------------------------------------------
    for (int q = 0; q < P; ++q) {
        if (q == r) { c->peer_ptr[q] = c->buf; continue; }
        const uint64_t gen = sh->slot[q].buf_gen;
        if (c->peer_gen[q] == gen && c->peer_ptr[q]) continue;
        if (c->ipc) {
            if (c->peer_ptr[q]) cudaIpcCloseMemHandle(c->peer_ptr[q]);
            c->peer_ptr[q] = nullptr;
            cudaError_t ce = cudaIpcOpenMemHandle(&c->peer_ptr[q], sh->slot[q].handle, cudaIpcMemLazyEnablePeerAccess);
            if (ce != cudaSuccess) {
                cudaGetLastError();
                return comm_abort(c, fail(BTB_ECOMM, "rank %d: cudaIpcOpenMemHandle of rank %d's buffer failed: %s", r, q, cudaGetErrorString(ce)));
            }
        } else {
            const int pd = sh->slot[q].device;
            if (pd != c->device) {
                int can = 0;
                cudaDeviceCanAccessPeer(&can, c->device, pd);
                if (!can) return comm_abort(c, fail(BTB_ECOMM, "device %d cannot access device %d (no peer-to-peer path)", c->device, pd));
                cudaError_t ce = cudaDeviceEnablePeerAccess(pd, 0);
                if (ce != cudaSuccess && ce != cudaErrorPeerAccessAlreadyEnabled)
                    return comm_abort(c, fail(BTB_ECOMM, "cudaDeviceEnablePeerAccess(%d) failed: %s", pd, cudaGetErrorString(ce)));
                cudaGetLastError();
            }
            c->peer_ptr[q] = sh->slot[q].local_ptr;
        }
        c->peer_gen[q] = gen;
    }

    // ---- encode my shard (majority symbols from all ranks' counts, read over NVLink) ----------------------
    if (use_major) {
        CountPtrs cp;
        cp.n = P;
        for (int q = 0; q < P; ++q) cp.p[q] = reinterpret_cast<const uint32_t *>(static_cast<uint8_t *>(c->peer_ptr[q]) + counts_off);
        BTB_CUDA(cudaMemsetAsync(d_major, 0, size_t(major_pad), c->st));
        BTB_TRY(dist_site_major_sum(cp, len, d_major, c->st));
    }
    int32_t *rowsum = reinterpret_cast<int32_t *>(c->buf + n * kb);
    if (dense && rows > 0) BTB_CUDA(cudaMemsetAsync(rowsum + s0, 0, size_t(rows) * 4, c->st));
    BTB_TRY(dist_encode_fp4_rows(c->d_seqs, rows, len, d_stride, table, d_table, sigma, c->buf + s0 * kb, dense ? rowsum + s0 : nullptr,
                                 use_major ? d_major : nullptr, c->st));
    BTB_CUDA(cudaStreamSynchronize(c->st));
    sh->slot[r].ready.store(e, std::memory_order_release);
    lap(4);

    // ---- pull the rows my block rows meet: everything from (a little before) my first row on -----------------
    int64_t lo = 0, hi = 0;
    BTB_TRY(btb_dist_row_share(n, r, P, &lo, &hi));
    const int64_t need0 = std::max<int64_t>(0, lo * BTB_TILE - BTB_TILE);      // 240-column units start up to 239 rows early
    RowArray arr[2];
    const int n_arr = operand_arrays(n, len, sigma, arr);
    int used_streams = 0;
    if (hi > lo) {
        // nearest shards first: the first band only needs the rows near the diagonal
        for (int q = 0; q < P; ++q) {
            if (q == r) continue;
            const int64_t q0 = n * q / P, q1 = n * (q + 1) / P;
            const int64_t a = std::max(q0, need0), b = q1;
            if (b <= a) continue;
            BTB_TRY(comm_wait(c, "a peer's encoded shard", [&] { return sh->slot[q].ready.load(std::memory_order_acquire) >= e; }));
            cudaStream_t ps = c->st_pull[used_streams & 3];
            ++used_streams;
            for (int k = 0; k < n_arr; ++k) {
                const size_t off = arr[k].offset + size_t(a) * arr[k].pitch, bytes = size_t(b - a) * arr[k].pitch;
                BTB_CUDA(cudaMemcpyAsync(c->buf + off, static_cast<const uint8_t *>(c->peer_ptr[q]) + off, bytes, cudaMemcpyDefault, ps));
            }
        }
    }
    for (int k = 0; k < std::min(used_streams, 4); ++k) {
        BTB_CUDA(cudaEventRecord(c->ev_pull[k], c->st_pull[k]));
        BTB_CUDA(cudaStreamWaitEvent(c->st, c->ev_pull[k], 0));
    }
    for (int k = 0; k < std::min(used_streams, 4); ++k) BTB_CUDA(cudaStreamSynchronize(c->st_pull[k]));
    sh->slot[r].pulled.store(e, std::memory_order_release);
    lap(5);

    // ---- my block rows, band by band, each band's cells on their way to the host while the next is computed ----
    BTB_TRY(share_compute_and_copy(BTB_ENGINE_UMMA, c->buf, n, len, sigma, lo, hi, c->d_out, eb, d_ld, out, out_ld, c->st, c->st_out, c->ev_band));
    lap(6);
    *done = true;
    return BTB_OK;
}

extern "C" int btb_dist_matrix_host_comm(btb_comm *c, const uint8_t *seqs, int64_t n, int64_t len, int64_t row_stride,
                                         int allchars, int keepcase, int engine, void *out, int out_elem_bytes, int64_t out_ld) {
    if (!c || !seqs || !out || n <= 0 || len < 0 || row_stride < len || out_ld < n) return fail(BTB_EINVAL, "btb_dist_matrix_host_comm: bad argument");
    if (out_elem_bytes != 2 && out_elem_bytes != 4) return fail(BTB_EINVAL, "out_elem_bytes must be 2 or 4");
    if (out_elem_bytes == 2 && len > 65535) return fail(BTB_EINVAL, "uint16 output needs len < 65536");
    bool done = false;
    int rc = comm_matrix_impl(c, seqs, n, len, row_stride, allchars, keepcase, engine, out, out_elem_bytes, out_ld, &done);
    if (rc != BTB_OK) return comm_abort(c, rc);
    g_comm_last_path = done ? 1 : 0;
    if (done) return BTB_OK;
    // layouts without an exchange path: every rank uploads the whole alignment (btb_dist_matrix_host)
    return btb_dist_matrix_host(seqs, n, len, row_stride, allchars, keepcase, engine, c->device, c->rank, c->world, out,
                                out_elem_bytes, out_ld);
}
