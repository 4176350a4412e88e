// k4, UMMA engine: the SNP distance matrix as an exact GEMM on the 5th-gen tensor cores.
//
// Replaces the O(N^2 L) byte-compare loop of snp-dists 0.8.2, reached in
// the reference through btbphylo/phylogeny.py:149.  With the operands of dist_encode.cu (0/1 codes),
//     dense alignments:   d(i,j) = c_i + c_j - sum_k X[i][k] X[j][k]      (tetrahedron code, c = row sums)
//     general alphabets:  d(i,j) = sum_k V[i][k] V[j][k] - sum_s sum_k X_s[i][k] X_s[j][k]
// evaluated exactly by tcgen05.mma kind::mxf4 on e2m1 nibbles (fp32 accumulators hold the integers)
// or, for very long rows / umma_fp4 = 0, by kind::i8 with int32 accumulators
//     d(i,j) = sum_k A[i][k] * B[j][k]          (A one-hot, B masked complement).
//
// Kernels in this file (sm_100a only; all persistent, warp 0 = TMA producer, warp 1 = MMA issuer with
// one elected lane, warp 2 = TMEM allocator, warps 4.. = epilogue; full/empty mbarrier rings):
//   dist_umma_fp4_pair_kernel   default: CTA pairs (cta_group::2), 512 x 240 units, B split across the pair
//   dist_umma_fp4_tile_kernel   one CTA per 256 x 240 unit (two halves share every staged B block)
//   dist_umma_fp4_kernel        128-row half tiles (small launches, explicit tile lists)
//   dist_umma_tile_kernel       int8, 256 x 256 tiles as two halves
//   dist_umma_kernel            int8 half tiles
//   umma_probe_kernel           MMA issue-rate probe (no data movement): the measured tensor-pipe peak bench.py's
//                               roofline divides by
// Bound: tensor pipe (at the board's power cap).  DESIGN.md section 4 has the measurements.
#include <cuda.h>

#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <mutex>
#include <vector>

#include "common.cuh"

namespace btb {

constexpr int kBlockK = 128;          // K bytes per pipeline stage = one 128B swizzle atom row
constexpr int kUmmaK = 32;            // K per tcgen05.mma for 8-bit operands
constexpr int kThreads = 256;
constexpr int kThreadsWide = 384;     // FP4 full-tile kernels: a second group of epilogue warps, one group per accumulator half
constexpr int kAccCols = 256;         // accumulator columns per stage (N of the MMA)

// ------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive_local(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// arrive on the barrier at the same smem offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar, uint32_t cta) {
    asm volatile(
        "{\n\t.reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}"
        ::"r"(bar), "r"(cta) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t}"
        ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
// one lane of a converged warp; the compiler knows the guarded region is single-lane, so the
// uniform-datapath instructions inside (UTMALDG, UTCIMMA, UTCBAR) need no per-lane election loops
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
        "elect.sync rx|px, %1;\n\t"
        "@px mov.s32 %0, 1;\n\t}"
        : "+r"(pred) : "r"(0xFFFFFFFFu));
    return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

template <int CG>
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1) {
    if constexpr (CG == 1) {
        asm volatile(
            "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
            ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
    } else {
        // both CTAs of the pair signal the leader's barrier: clear the peer bit of the address
        asm volatile(
            "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
            ::"r"(dst), "l"(map), "r"(bar & 0xFEFFFFFFu), "r"(c0), "r"(c1) : "memory");
    }
}

template <int CG>
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols) {
    if constexpr (CG == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    } else {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
}
template <int CG>
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
    if constexpr (CG == 1)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols));
    else
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols));
}

// D[tmem] (+)= A[smem] * B[smem], 8-bit integer operands, int32 accumulate
template <int CG>
__device__ __forceinline__ void umma_i8(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                        uint32_t accumulate) {
    if constexpr (CG == 1) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
    } else {
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
    }
}

// tcgen05.commit: arrive on `bar` when every MMA issued so far by this thread has completed
template <int CG>
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    if constexpr (CG == 1) {
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
    } else {
        const uint16_t mask = 3;      // the barrier at this offset in both CTAs of the pair
        asm volatile(
            "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
            ::"r"(bar), "h"(mask) : "memory");
    }
}

struct UmmaParams;
__device__ __forceinline__ void tile_of(const int2 *sq_list, int64_t T, int64_t tile, int &I, int &J) {
    if (sq_list) { const int2 w = sq_list[tile]; I = w.x; J = w.y; }
    else tile_coords(T, tile, I, J);
}

__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major operand tile in shared memory, 128-byte swizzle: rows of 128 bytes, 8-row atoms of 1024 B.
// (descriptor bit layout: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48),
//  layout type [61,64) with 2 = SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= uint64_t((smem_addr >> 4) & 0x3FFF);
    d |= uint64_t(1024 >> 4) << 32;
    d |= uint64_t(1) << 46;
    d |= uint64_t(2) << 61;
    return d;
}

// instruction descriptor, kind::i8: D = S32 (2 at [4,6)), A/B = signed 8-bit (1 at [7,10) / [10,13)),
// both K-major (0 at 15 / 16), N>>3 at [17,23), M>>4 at [24,29)
__host__ __device__ constexpr uint32_t make_i8_idesc(int M, int N) {
    return (2u << 4) | (1u << 7) | (1u << 10) | (uint32_t(N >> 3) << 17) | (uint32_t(M >> 4) << 24);
}

struct UmmaParams {
    int64_t n;            // sequences
    int64_t k_blocks;     // K bytes / 128
    int64_t tile_lo;      // first tile of this launch
    int64_t units;        // work units in this launch (tiles for CG=2, half tiles for CG=1)
    void *out;
    int64_t ld;
    int out_mode;         // 0 = n x n matrix, 1 = tile-packed
    int mirror;
    int dense;            // 0: out = acc; 1: out = bias - acc (one-hot matches); 2: out = (bias - acc) >> 2 (simplex)
    int32_t bias;         // len or 3 * len
    int fp32acc;          // accumulators are fp32 (kind::mxf4): convert to integers first
    const int32_t *rowsum; // dense == 3: c_i, out = c_row + c_col - acc
    const int2 *worklist;      // FP4 full-tile kernel: (block row I, 240-column block j) per unit
    const int2 *sq_list;       // optional explicit list of square tiles (I, J) replacing the band enumeration
    unsigned int *lockstep;    // optional grid-wide arrival counter: producers start every unit together
    int lock_slack = 0;        // a producer may run this many rounds ahead of the slowest one (0 = strict)
    int64_t neg_from = INT64_MAX;   // kind::mxf4: K blocks >= neg_from are subtracted (general alphabets: V blocks - X blocks)
    int64_t b_shift = 0;       // 0: subtract with the instruction descriptor's negate-A bit
};

// Epilogue of one 128-row x 256-column accumulator buffer: this thread owns matrix row `row`
// (TMEM lane of `taddr`), reads 32 columns at a time and writes the tile (16-byte stores), its
// mirror (lane-coalesced 2/4-byte stores) or the tile-packed slot.  The TMEM loads are software
// pipelined (the next 32 columns are in flight while the current ones are converted and stored).
// `valid_cols` < 256 and `tri` = true serve the rectangular 256 x 240 tiles of the FP4 full-tile
// kernels: a cell is written where column >= row (mirrored where column > row), whatever tile it is
// in.  `colsum` (shared memory, optional): the row sums of the unit's columns, staged by the caller.
__device__ __forceinline__ void tmem_ld_32x32_issue(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

template <typename OutT>
__device__ __forceinline__ void emit_chunk(const UmmaParams &p, uint32_t (&v)[32], int c0, int64_t row, int64_t n0,
                                           int row_in_tile, int64_t tile_slot, bool do_mirror, int valid_cols, bool tri,
                                           int32_t cr, const int32_t *colsum) {
    OutT *out = static_cast<OutT *>(p.out);
    if (p.fp32acc) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = uint32_t(__float2int_rn(__uint_as_float(v[j])));
    }
    if (p.dense == 3) {
        if (colsum) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = uint32_t(cr + colsum[c0 + j] - int32_t(v[j]));
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const int64_t cj = n0 + c0 + j;
                v[j] = uint32_t(cr + (cj < p.n ? p.rowsum[cj] : 0) - int32_t(v[j]));
            }
        }
    } else if (p.dense) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = uint32_t(p.bias - int32_t(v[j])) >> (p.dense == 2 ? 2 : 0);
    }
    const int64_t col = n0 + c0;
    if (tri) {
        // rectangular tiles straddle the diagonal anywhere: per-cell rule
        const int lim = valid_cols - c0 < 32 ? valid_cols - c0 : 32;
        if (row < p.n) {
            OutT *dst = out + row * p.ld + col;
            if (lim == 32 && col >= row && col + 32 <= p.n && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
                uint4 *d4 = reinterpret_cast<uint4 *>(dst);
                if constexpr (sizeof(OutT) == 2) {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        d4[j] = make_uint4(v[8 * j] | (v[8 * j + 1] << 16), v[8 * j + 2] | (v[8 * j + 3] << 16),
                                           v[8 * j + 4] | (v[8 * j + 5] << 16), v[8 * j + 6] | (v[8 * j + 7] << 16));
                } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j) d4[j] = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                }
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    if (j < lim && col + j >= row && col + j < p.n) dst[j] = OutT(v[j]);
            }
            if (do_mirror) {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    if (j < lim && col + j > row && col + j < p.n) out[(col + j) * p.ld + row] = OutT(v[j]);
            }
        }
        return;
    }
    if (p.out_mode == 1) {
        OutT *dst = out + tile_slot * (BTB_TILE * BTB_TILE) +
                    (int64_t)row_in_tile * BTB_TILE + c0;
#pragma unroll
        for (int j = 0; j < 32; ++j) dst[j] = (row < p.n && col + j < p.n) ? OutT(v[j]) : OutT(0);
    } else {
        if (row < p.n) {
            OutT *dst = out + row * p.ld + col;
            if (col + 32 <= p.n && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
                // interior: 32 consecutive outputs of this row as 16-byte stores
                uint4 *d4 = reinterpret_cast<uint4 *>(dst);
                if constexpr (sizeof(OutT) == 2) {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        d4[j] = make_uint4(v[8 * j] | (v[8 * j + 1] << 16), v[8 * j + 2] | (v[8 * j + 3] << 16),
                                           v[8 * j + 4] | (v[8 * j + 5] << 16), v[8 * j + 6] | (v[8 * j + 7] << 16));
                } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j) d4[j] = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                }
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    if (col + j < p.n) dst[j] = OutT(v[j]);
            }
        }
        if (do_mirror && row < p.n) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (col + j < p.n) out[(col + j) * p.ld + row] = OutT(v[j]);
        }
    }
}

template <typename OutT>
__device__ __forceinline__ void drain_accumulator(const UmmaParams &p, uint32_t taddr, int64_t row, int64_t n0,
                                                  int row_in_tile, int64_t tile_slot, bool do_mirror,
                                                  int valid_cols = kAccCols, bool tri = false,
                                                  const int32_t *colsum = nullptr) {
    const int32_t cr = (p.dense == 3 && row < p.n) ? p.rowsum[row] : 0;
    uint32_t va[32], vb[32];
    tmem_ld_32x32_issue(taddr, va);
#pragma unroll 1
    for (int c0 = 0; c0 < valid_cols; c0 += 64) {
        tmem_ld_wait();
        const bool more = c0 + 32 < valid_cols;
        if (more) tmem_ld_32x32_issue(taddr + uint32_t(c0 + 32), vb);
        emit_chunk<OutT>(p, va, c0, row, n0, row_in_tile, tile_slot, do_mirror, valid_cols, tri, cr, colsum);
        if (more) {
            tmem_ld_wait();
            if (c0 + 64 < valid_cols) tmem_ld_32x32_issue(taddr + uint32_t(c0 + 64), va);
            emit_chunk<OutT>(p, vb, c0 + 32, row, n0, row_in_tile, tile_slot, do_mirror, valid_cols, tri, cr, colsum);
        }
    }
}

// int8 half-tile kernel: one CTA per 128-row x 256-column unit (UMMA 128 x 256 x 32, kind::i8), 4 stages of
// A 16 KB + B 32 KB, two accumulator buffers so that the epilogue of a unit overlaps the K loop of the next.
struct UmmaCfg {
    static constexpr int kARows = 128;                               // A rows loaded per stage
    static constexpr int kBRows = 256;                               // B rows loaded per stage
    static constexpr int kABytes = kARows * kBlockK;
    static constexpr int kBBytes = kBRows * kBlockK;
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kStages = 4;
    static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
};

// lockstep: re-align the producers of all CTAs at every unit boundary.  CTAs that stream the same operand panels
// then sit at the same K offset, so a panel block is fetched from HBM once and hit in L2 by everyone else (needs
// all CTAs co-resident: cooperative launch).  Round r expects every worker that has a unit in round r;
// `per_worker` arrivals per worker and round (2 for the CTAs of a pair).
__device__ __forceinline__ void lockstep_wait(const UmmaParams &p, int64_t u, int64_t worker, int64_t n_workers, int per_worker) {
    const int64_t round = (u - worker) / n_workers;
    const int64_t full_rounds = p.units / n_workers;                  // rounds in which every worker has a unit
    const int64_t last_count = p.units - full_rounds * n_workers;
    int64_t target = (round <= full_rounds - 1 ? round : full_rounds - 1) * n_workers * per_worker;
    if (round > full_rounds - 1) target += last_count * per_worker;
    target -= int64_t(p.lock_slack) * n_workers * per_worker;
    __threadfence();
    atomicAdd(p.lockstep, 1u);
    while (int64_t(*reinterpret_cast<volatile unsigned int *>(p.lockstep)) < target) __nanosleep(64);
}

template <typename OutT>
__global__ void __launch_bounds__(kThreads, 1)
dist_umma_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                 const UmmaParams p) {
    using Cfg = UmmaCfg;
    constexpr int kStages = Cfg::kStages;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;    // 1024-byte alignment for the 128B swizzle atoms
    const uint32_t bar_base = smem_base + kStages * Cfg::kStageBytes;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
    auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * kStages + a); };
    auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * kStages + 2 + a); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * kStages + 4);
    auto smem_a = [&](int s) { return smem_base + s * Cfg::kStageBytes; };
    auto smem_b = [&](int s) { return smem_base + s * Cfg::kStageBytes + Cfg::kABytes; };

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t worker = blockIdx.x, n_workers = gridDim.x;
    const int64_t T = tiles_per_dim(p.n);

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 4); }   // tempty: one arrive per epilogue warp
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc<1>(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    // unit -> rows of A (m0), rows of B (n0)
    auto unit_coords = [&](int64_t u, int64_t &m0, int64_t &n0, int &I, int &J, int &half) {
        half = int(u & 1);
        tile_of(p.sq_list, T, p.tile_lo + u / 2, I, J);
        m0 = int64_t(I) * BTB_TILE + half * 128;
        n0 = int64_t(J) * BTB_TILE;
    };

    if (warp == 0) {
        // ================================================================= TMA producer
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            for (int64_t u = worker; u < p.units; u += n_workers) {
                int64_t m0, n0; int I, J, half;
                unit_coords(u, m0, n0, I, J, half);
                if (p.lockstep && u != worker) lockstep_wait(p, u, worker, n_workers, 1);
                for (int64_t kb = 0; kb < p.k_blocks; ++kb) {
                    mbar_wait(empty_bar(stage), phase ^ 1);
                    mbar_expect_tx(full_bar(stage), Cfg::kStageBytes);
                    const int k0 = int(kb * kBlockK);                    // beyond K: zero-filled by TMA
                    tma_load_2d<1>(smem_a(stage), &map_a, full_bar(stage), k0, int(m0));
                    tma_load_2d<1>(smem_b(stage), &map_b, full_bar(stage), k0, int(n0));
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ================================================================= MMA issuer
        // the whole warp walks the pipeline (warp-uniform control flow keeps descriptors in uniform registers);
        // one elected lane issues the tensor-core instructions
        constexpr uint32_t idesc = make_i8_idesc(128, kAccCols);
        int stage = 0;
        uint32_t phase = 0;
        int64_t it = 0;
        for (int64_t u = worker; u < p.units; u += n_workers, ++it) {
            const int acc = int(it & 1);
            const uint32_t acc_phase = uint32_t(it >> 1) & 1;
            mbar_wait(tempty_bar(acc), acc_phase ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * kAccCols;
            for (int64_t kb = 0; kb < p.k_blocks; ++kb) {
                mbar_wait(full_bar(stage), phase);
                tc_fence_after();
                if (elect_one()) {
                    const uint64_t a_desc = make_kmajor_sw128_desc(smem_a(stage));
                    const uint64_t b_desc = make_kmajor_sw128_desc(smem_b(stage));
#pragma unroll
                    for (int k = 0; k < kBlockK / kUmmaK; ++k)
                        umma_i8<1>(d_tmem, a_desc + uint64_t(k * kUmmaK >> 4), b_desc + uint64_t(k * kUmmaK >> 4),
                                   idesc, (kb > 0 || k > 0) ? 1u : 0u);
                    umma_commit<1>(empty_bar(stage));
                    if (kb == p.k_blocks - 1) umma_commit<1>(tfull_bar(acc));
                }
                __syncwarp();
                if (++stage == kStages) { stage = 0; phase ^= 1; }
            }
        }
        __syncwarp();
    } else if (warp >= 4) {
        // ================================================================= epilogue
        const int quarter = warp & 3;                       // TMEM lanes 32*quarter .. +31
        int64_t it = 0;
        for (int64_t u = worker; u < p.units; u += n_workers, ++it) {
            const int acc = int(it & 1);
            const uint32_t acc_phase = uint32_t(it >> 1) & 1;
            int64_t m0, n0; int I, J, half;
            unit_coords(u, m0, n0, I, J, half);
            mbar_wait(tfull_bar(acc), acc_phase);
            tc_fence_after();
            const int64_t row = m0 + quarter * 32 + lane;
            const bool do_mirror = p.mirror && p.out_mode == 0 && I != J;
            drain_accumulator<OutT>(p, tmem_base + (uint32_t(quarter * 32) << 16) + uint32_t(acc * kAccCols), row, n0,
                                    half * 128 + quarter * 32 + lane, u / 2, do_mirror);
            // accumulator buffer drained: hand it back to the MMA issuer
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_local(tempty_bar(acc));
        }
    }

    // teardown: nobody leaves (or frees TMEM) while a warp may still signal or read
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc<1>(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------
// Full-tile variant: one CTA computes a whole 256 x 256 tile as two 128-row UMMA halves that share
// every staged B block.  Per 128 K-bytes the CTA pulls 64 KB from L2 for 1024 tensor cycles
// (64 B/clk instead of 96 B/clk for the half-tile kernel), so the same 192 KB of shared memory
// covers 3072 instead of 2048 cycles of L2 latency.  The two TMEM buffers are the two halves of
// the current tile; each has its own full/empty barrier, so the epilogue of half h of tile t
// overlaps the K loop of tile t+1 as soon as that half is drained.
struct TileCfg {
    static constexpr int kABytes = 256 * kBlockK;                     // 32 KB: both halves of the row panel
    static constexpr int kBBytes = 256 * kBlockK;                     // 32 KB
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kStages = 3;
    static constexpr int kSmemBytes = kStages * kStageBytes + 1024 + 256;
};

template <typename OutT>
__global__ void __launch_bounds__(kThreads, 1)
dist_umma_tile_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                      const UmmaParams p) {
    using Cfg = TileCfg;
    constexpr int kStages = Cfg::kStages;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_base = smem_base + kStages * Cfg::kStageBytes;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
    auto tfull_bar = [&](int h) { return bar_base + 8u * (2 * kStages + h); };
    auto tempty_bar = [&](int h) { return bar_base + 8u * (2 * kStages + 2 + h); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * kStages + 4);
    auto smem_a = [&](int s) { return smem_base + s * Cfg::kStageBytes; };
    auto smem_b = [&](int s) { return smem_base + s * Cfg::kStageBytes + Cfg::kABytes; };

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t worker = blockIdx.x, n_workers = gridDim.x;
    const int64_t T = tiles_per_dim(p.n);

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int h = 0; h < 2; ++h) { mbar_init(tfull_bar(h), 1); mbar_init(tempty_bar(h), 4); }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc<1>(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    if (warp == 0) {
        // ================================================================= TMA producer
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            for (int64_t u = worker; u < p.units; u += n_workers) {
                int I, J;
                tile_of(p.sq_list, T, p.tile_lo + u, I, J);
                const int m0 = I * BTB_TILE, n0 = J * BTB_TILE;
                if (p.lockstep && u != worker) lockstep_wait(p, u, worker, n_workers, 1);
                for (int64_t kb = 0; kb < p.k_blocks; ++kb) {
                    mbar_wait(empty_bar(stage), phase ^ 1);
                    mbar_expect_tx(full_bar(stage), Cfg::kStageBytes);
                    tma_load_2d<1>(smem_a(stage), &map_a, full_bar(stage), int(kb * kBlockK), m0);
                    tma_load_2d<1>(smem_b(stage), &map_b, full_bar(stage), int((kb + (kb >= p.neg_from ? p.b_shift : 0)) * kBlockK), n0);
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ================================================================= MMA issuer
        constexpr uint32_t idesc = make_i8_idesc(128, kAccCols);
        int stage = 0;
        uint32_t phase = 0;
        int64_t it = 0;
        for (int64_t u = worker; u < p.units; u += n_workers, ++it) {
            const uint32_t drained = (uint32_t(it) & 1u) ^ 1u;       // parity of the previous tile's drain
            for (int64_t kb = 0; kb < p.k_blocks; ++kb) {
                mbar_wait(full_bar(stage), phase);
                if (kb == 0) mbar_wait(tempty_bar(0), drained);
                tc_fence_after();
                if (elect_one()) {
                    const uint64_t a_desc = make_kmajor_sw128_desc(smem_a(stage));
                    const uint64_t b_desc = make_kmajor_sw128_desc(smem_b(stage));
#pragma unroll
                    for (int k = 0; k < kBlockK / kUmmaK; ++k)
                        umma_i8<1>(tmem_base, a_desc + uint64_t(k * kUmmaK >> 4), b_desc + uint64_t(k * kUmmaK >> 4), idesc,
                                   (kb > 0 || k > 0) ? 1u : 0u);
                }
                __syncwarp();
                if (kb == 0) { mbar_wait(tempty_bar(1), drained); tc_fence_after(); }
                if (elect_one()) {
                    // second half: rows 128..255 of the staged A block (16 KB further), accumulator columns 256..511
                    const uint64_t a_desc = make_kmajor_sw128_desc(smem_a(stage) + 128 * kBlockK);
                    const uint64_t b_desc = make_kmajor_sw128_desc(smem_b(stage));
#pragma unroll
                    for (int k = 0; k < kBlockK / kUmmaK; ++k)
                        umma_i8<1>(tmem_base + kAccCols, a_desc + uint64_t(k * kUmmaK >> 4), b_desc + uint64_t(k * kUmmaK >> 4),
                                   idesc, (kb > 0 || k > 0) ? 1u : 0u);
                    umma_commit<1>(empty_bar(stage));
                    if (kb == p.k_blocks - 1) { umma_commit<1>(tfull_bar(0)); umma_commit<1>(tfull_bar(1)); }
                }
                __syncwarp();
                if (++stage == kStages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp >= 4) {
        // ================================================================= epilogue
        const int quarter = warp & 3;
        int64_t it = 0;
        for (int64_t u = worker; u < p.units; u += n_workers, ++it) {
            int I, J;
            tile_of(p.sq_list, T, p.tile_lo + u, I, J);
            const int64_t n0 = int64_t(J) * BTB_TILE;
            const bool do_mirror = p.mirror && p.out_mode == 0 && I != J;
#pragma unroll 1
            for (int h = 0; h < 2; ++h) {
                mbar_wait(tfull_bar(h), uint32_t(it) & 1u);
                tc_fence_after();
                const int row_in_tile = h * 128 + quarter * 32 + lane;
                drain_accumulator<OutT>(p, tmem_base + (uint32_t(quarter * 32) << 16) + uint32_t(h * kAccCols),
                                        int64_t(I) * BTB_TILE + row_in_tile, n0, row_in_tile, u, do_mirror);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_local(tempty_bar(h));
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc<1>(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------
// FP4 variant (experimental, dense alignments only): the 0/1 tetrahedron code stored as packed
// e2m1 nibbles and multiplied with tcgen05.mma kind::mxf4.block_scale (K = 64 per instruction,
// twice the int8 rate).  All block scale factors are 1.0 (UE8M0 0x7F): a 16-column TMEM region is
// filled with 0x7F7F7F7F once, so its layout does not matter.  fp32 accumulators hold exact
// integers (sums of 0/1 products < 2^24).  Unit of work = one 128-row half tile (UMMA 128x256x64),
// one accumulator buffer (columns 0..255) + the scale-factor columns: the epilogue of a unit is not
// overlapped with the next unit's K loop.
__device__ __forceinline__ void umma_mxf4(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t sfa, uint32_t sfb, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::mxf4.block_scale.block32 [%0], %1, %2, %3, [%5], [%6], p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(sfa), "r"(sfb) : "memory");
}
__device__ __forceinline__ void umma_mxf4_cg2(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                              uint32_t sfa, uint32_t sfb, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::mxf4.block_scale.block32 [%0], %1, %2, %3, [%5], [%6], p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(sfa), "r"(sfb) : "memory");
}
__device__ __forceinline__ void tmem_st_32x32_x16(uint32_t taddr, uint32_t v) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};"
        ::"r"(taddr), "r"(v) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
// block-scaled instruction descriptor, kind::mxf4: A/B = E2M1 (1 at [7,10) / [10,13)), K-major,
// N>>3 at [17,23), scale format UE8M0 (1 at bit 23), M>>4 at [24,29), scale-factor ids 0, K64 dense
constexpr uint32_t kNegateA = 1u << 13;                     // instruction descriptor: use -A
__host__ __device__ constexpr uint32_t make_mxf4_idesc(int M, int N) {
    return (1u << 7) | (1u << 10) | (uint32_t(N >> 3) << 17) | (1u << 23) | (uint32_t(M >> 4) << 24);
}

template <typename OutT>
__global__ void __launch_bounds__(kThreads, 1)
dist_umma_fp4_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                     const UmmaParams p) {
    using Cfg = UmmaCfg;                                     // A 128 rows + B 256 rows per 128-byte K atom, 4 stages
    constexpr int kStages = Cfg::kStages;
    constexpr uint32_t kSfCol = 256;                         // scale factors: TMEM columns 256..271
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_base = smem_base + kStages * Cfg::kStageBytes;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
    const uint32_t tfull_bar = bar_base + 8u * (2 * kStages);
    const uint32_t tempty_bar = bar_base + 8u * (2 * kStages + 1);
    const uint32_t tmem_slot = bar_base + 8u * (2 * kStages + 4);
    auto smem_a = [&](int s) { return smem_base + s * Cfg::kStageBytes; };
    auto smem_b = [&](int s) { return smem_base + s * Cfg::kStageBytes + Cfg::kABytes; };

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t worker = blockIdx.x, n_workers = gridDim.x;
    const int64_t T = tiles_per_dim(p.n);

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        mbar_init(tfull_bar, 1);
        mbar_init(tempty_bar, 4);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc<1>(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
    if (warp >= 4) tmem_st_32x32_x16(tmem_base + (uint32_t((warp & 3) * 32) << 16) + kSfCol, 0x7F7F7F7Fu);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    auto unit_coords = [&](int64_t u, int64_t &m0, int64_t &n0, int &I, int &J, int &half) {
        half = int(u & 1);
        tile_of(p.sq_list, T, p.tile_lo + u / 2, I, J);
        m0 = int64_t(I) * BTB_TILE + half * 128;
        n0 = int64_t(J) * BTB_TILE;
    };

    if (warp == 0) {
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            for (int64_t u = worker; u < p.units; u += n_workers) {
                int64_t m0, n0; int I, J, half;
                unit_coords(u, m0, n0, I, J, half);
                if (p.lockstep && u != worker) lockstep_wait(p, u, worker, n_workers, 1);
                for (int64_t kb = 0; kb < p.k_blocks; ++kb) {
                    mbar_wait(empty_bar(stage), phase ^ 1);
                    mbar_expect_tx(full_bar(stage), Cfg::kStageBytes);
                    tma_load_2d<1>(smem_a(stage), &map_a, full_bar(stage), int(kb * kBlockK), int(m0));
                    tma_load_2d<1>(smem_b(stage), &map_b, full_bar(stage), int((kb + (kb >= p.neg_from ? p.b_shift : 0)) * kBlockK), int(n0));
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        constexpr uint32_t idesc0 = make_mxf4_idesc(128, kAccCols);
        int stage = 0;
        uint32_t phase = 0;
        int64_t it = 0;
        for (int64_t u = worker; u < p.units; u += n_workers, ++it) {
            mbar_wait(tempty_bar, (uint32_t(it) & 1u) ^ 1u);
            tc_fence_after();
            for (int64_t kb = 0; kb < p.k_blocks; ++kb) {
                mbar_wait(full_bar(stage), phase);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t idesc = idesc0 | ((kb >= p.neg_from && p.b_shift == 0) ? kNegateA : 0u);
                    const uint64_t a_desc = make_kmajor_sw128_desc(smem_a(stage));
                    const uint64_t b_desc = make_kmajor_sw128_desc(smem_b(stage));
#pragma unroll
                    for (int k = 0; k < 4; ++k)               // 4 x 64 nibbles = 128 bytes
                        umma_mxf4(tmem_base, a_desc + uint64_t(k * 2), b_desc + uint64_t(k * 2), idesc,
                                  tmem_base + kSfCol, tmem_base + kSfCol + 4, (kb > 0 || k > 0) ? 1u : 0u);
                    umma_commit<1>(empty_bar(stage));
                    if (kb == p.k_blocks - 1) umma_commit<1>(tfull_bar);
                }
                __syncwarp();
                if (++stage == kStages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp >= 4) {
        const int quarter = warp & 3;
        int64_t it = 0;
        for (int64_t u = worker; u < p.units; u += n_workers, ++it) {
            int64_t m0, n0; int I, J, half;
            unit_coords(u, m0, n0, I, J, half);
            mbar_wait(tfull_bar, uint32_t(it) & 1u);
            tc_fence_after();
            const bool do_mirror = p.mirror && p.out_mode == 0 && I != J;
            drain_accumulator<OutT>(p, tmem_base + (uint32_t(quarter * 32) << 16), m0 + quarter * 32 + lane, n0,
                                    half * 128 + quarter * 32 + lane, u / 2, do_mirror);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_local(tempty_bar);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc<1>(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------
// FP4 full-tile variant: one CTA computes 256 rows x 240 columns as two 128-row halves that share
// every staged B block (64 instead of 96 B/clk of L2 traffic per SM, as in dist_umma_tile_kernel).
// The tiles are rectangular because two 256-column fp32 accumulators would fill all 512 TMEM
// columns and leave no room for the (unit) block scale factors: 2 x 240 + 16 columns fit.  Units
// come from a host-built work list in row-band order; cells are written where column >= row and
// mirrored where column > row, so any set of whole block rows covers exactly the same cells as the
// square-tile enumeration.
// (Measured and dropped, see DESIGN.md: six 31 KB stages with 64-byte swizzle, an L2 prefetch running 2-64 K blocks
// ahead, two stages -- none changed the time by more than noise, or lost: one SM takes in ~53 B/clk of TMA traffic.)
constexpr int kUnitColMask = 0xffff;                                  // work-list entry .y: the 240-column block
struct Fp4TileCfg {
    static constexpr int kBN = 240;
    static constexpr int kABytes = 256 * kBlockK;                     // 32 KB
    static constexpr int kBBytes = kBN * kBlockK;                     // 30 KB
    static constexpr int kStageBytes = kABytes + kBBytes;             // 62 KB (a multiple of 1024)
    static constexpr int kStages = 3;
    static constexpr int kSmemBytes = kStages * kStageBytes + 1024 + 256 + 1024;   // + align slack, barriers, column sums
};

template <typename OutT>
__global__ void __launch_bounds__(kThreadsWide, 1)
dist_umma_fp4_tile_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                          const UmmaParams p) {
    using Cfg = Fp4TileCfg;
    constexpr int kStages = Cfg::kStages;
    constexpr int BK = kBlockK;
    auto make_desc = [](uint32_t addr) { return make_kmajor_sw128_desc(addr); };
    constexpr uint32_t kHalf1Col = 256, kSfCol = 496;         // accumulators at 0..239 and 256..495, scale factors 496..511
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_base = smem_base + kStages * Cfg::kStageBytes;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
    auto tfull_bar = [&](int h) { return bar_base + 8u * (2 * kStages + h); };
    auto tempty_bar = [&](int h) { return bar_base + 8u * (2 * kStages + 2 + h); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * kStages + 4);
    auto smem_a = [&](int s) { return smem_base + s * Cfg::kStageBytes; };
    auto smem_b = [&](int s) { return smem_base + s * Cfg::kStageBytes + Cfg::kABytes; };

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t worker = blockIdx.x, n_workers = gridDim.x;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int h = 0; h < 2; ++h) { mbar_init(tfull_bar(h), 1); mbar_init(tempty_bar(h), 4); }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc<1>(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
    if (warp >= 4) tmem_st_32x32_x16(tmem_base + (uint32_t((warp & 3) * 32) << 16) + kSfCol, 0x7F7F7F7Fu);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    if (warp == 0) {
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            for (int64_t u = worker; u < p.units; u += n_workers) {
                const int2 w = p.worklist[u];
                const int m0 = w.x * BTB_TILE, n0 = (w.y & kUnitColMask) * Cfg::kBN;
                if (p.lockstep && u != worker) lockstep_wait(p, u, worker, n_workers, 1);
                for (int64_t kb = 0; kb < p.k_blocks; ++kb) {
                    mbar_wait(empty_bar(stage), phase ^ 1);
                    mbar_expect_tx(full_bar(stage), Cfg::kStageBytes);
                    tma_load_2d<1>(smem_a(stage), &map_a, full_bar(stage), int(kb * BK), m0);
                    tma_load_2d<1>(smem_b(stage), &map_b, full_bar(stage), int((kb + (kb >= p.neg_from ? p.b_shift : 0)) * BK), n0);
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        constexpr uint32_t idesc0 = make_mxf4_idesc(128, Cfg::kBN);
        int stage = 0;
        uint32_t phase = 0;
        int64_t it = 0;
        for (int64_t u = worker; u < p.units; u += n_workers, ++it) {
            const uint32_t drained = (uint32_t(it) & 1u) ^ 1u;
            for (int64_t kb = 0; kb < p.k_blocks; ++kb) {
                const uint32_t idesc = idesc0 | ((kb >= p.neg_from && p.b_shift == 0) ? kNegateA : 0u);
                mbar_wait(full_bar(stage), phase);
                if (kb == 0) mbar_wait(tempty_bar(0), drained);
                tc_fence_after();
                if (elect_one()) {
                    const uint64_t a_desc = make_desc(smem_a(stage));
                    const uint64_t b_desc = make_desc(smem_b(stage));
#pragma unroll
                    for (int k = 0; k < BK / 32; ++k)
                        umma_mxf4(tmem_base, a_desc + uint64_t(k * 2), b_desc + uint64_t(k * 2), idesc,
                                  tmem_base + kSfCol, tmem_base + kSfCol + 4, (kb > 0 || k > 0) ? 1u : 0u);
                }
                __syncwarp();
                if (kb == 0) { mbar_wait(tempty_bar(1), drained); tc_fence_after(); }
                if (elect_one()) {
                    const uint64_t a_desc = make_desc(smem_a(stage) + 128 * BK);
                    const uint64_t b_desc = make_desc(smem_b(stage));
#pragma unroll
                    for (int k = 0; k < BK / 32; ++k)
                        umma_mxf4(tmem_base + kHalf1Col, a_desc + uint64_t(k * 2), b_desc + uint64_t(k * 2), idesc,
                                  tmem_base + kSfCol, tmem_base + kSfCol + 4, (kb > 0 || k > 0) ? 1u : 0u);
                    umma_commit<1>(empty_bar(stage));
                    if (kb == p.k_blocks - 1) { umma_commit<1>(tfull_bar(0)); umma_commit<1>(tfull_bar(1)); }
                }
                __syncwarp();
                if (++stage == kStages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp >= 4) {
        const int quarter = warp & 3;
        int64_t it = 0;
        for (int64_t u = worker; u < p.units; u += n_workers, ++it) {
            const int2 w = p.worklist[u];
            const int64_t n0 = int64_t(w.y & kUnitColMask) * Cfg::kBN;
            // the row sums of this unit's columns, staged while the K loop still runs
            int32_t *colsum = reinterpret_cast<int32_t *>(smem_raw + (bar_base - smem_u32(smem_raw)) + 256);
            if (p.dense == 3) {
                for (int c = int(threadIdx.x) - 128; c < 256; c += 256)
                    colsum[c] = (c < Cfg::kBN && n0 + c < p.n) ? p.rowsum[n0 + c] : 0;
                asm volatile("bar.sync 1, 256;" ::: "memory");
            }
            {
                const int h = (warp - 4) >> 2;                 // warps 4-7 drain half 0, warps 8-11 half 1, concurrently
                mbar_wait(tfull_bar(h), uint32_t(it) & 1u);
                tc_fence_after();
                const int row_in_tile = h * 128 + quarter * 32 + lane;
                drain_accumulator<OutT>(p, tmem_base + (uint32_t(quarter * 32) << 16) + uint32_t(h) * kHalf1Col,
                                        int64_t(w.x) * BTB_TILE + row_in_tile, n0, row_in_tile, 0, p.mirror != 0, Cfg::kBN, true,
                                        p.dense == 3 ? colsum : nullptr);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_local(tempty_bar(h));
            }
            if (p.dense == 3) asm volatile("bar.sync 1, 256;" ::: "memory");      // everyone is done with colsum
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc<1>(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------
// FP4 CTA-pair variant (cta_group::2): the single-CTA full-tile kernel is bound by how fast one SM
// can pull operands (measured ~53 B/clk; it needs 66 B/clk to keep the tensor pipe busy).  Here a
// pair of CTAs computes 512 rows x 240 columns: each CTA stages its own 256 A rows but only HALF of
// the B block (120 rows) -- the pair's MMAs (M = 256: 128 rows from each CTA, N = 240: 120 B rows
// from each CTA) read the other half from the partner's shared memory -- 47 KB instead of 62 KB per
// SM per 960 tensor cycles (49 B/clk).  Protocol: every TMA load of both CTAs signals the LEADER's
// full barrier (leader: arrive.expect_tx for both CTAs' bytes); the leader's MMA warp issues for
// the pair and releases the stage in both CTAs with one multicast tcgen05.commit; each CTA drains
// its own 2 x 128 accumulator rows and reports to the leader's tempty barriers.
struct Fp4PairCfg {
    static constexpr int kBN = 240;
    static constexpr int kABytes = 256 * kBlockK;                     // 32 KB
    static constexpr int kBBytes = (kBN / 2) * kBlockK;               // 15 KB
    static constexpr int kStageBytes = kABytes + kBBytes;             // 47 KB (multiple of 1024)
    static constexpr int kStages = 4;
    static constexpr int kSmemBytes = kStages * kStageBytes + 1024 + 256 + 1024;   // + align slack, barriers, column sums
};

__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t cta) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(cta));
    return r;
}
// TMA load of this CTA's tile, completion bytes reported to `remote_bar` (a shared::cluster address)
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap *map, uint32_t remote_bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(remote_bar), "r"(c0), "r"(c1) : "memory");
}

template <typename OutT>
__global__ void __launch_bounds__(kThreadsWide, 1)
dist_umma_fp4_pair_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                          const UmmaParams p) {
    using Cfg = Fp4PairCfg;
    constexpr int kStages = Cfg::kStages;
    constexpr uint32_t kHalf1Col = 256, kSfCol = 496;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_base = smem_base + kStages * Cfg::kStageBytes;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
    auto tfull_bar = [&](int h) { return bar_base + 8u * (2 * kStages + h); };
    auto tempty_bar = [&](int h) { return bar_base + 8u * (2 * kStages + 2 + h); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * kStages + 4);
    auto smem_a = [&](int s) { return smem_base + s * Cfg::kStageBytes; };
    auto smem_b = [&](int s) { return smem_base + s * Cfg::kStageBytes + Cfg::kABytes; };

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t cta_rank = cluster_ctarank();
    const bool leader = cta_rank == 0;
    const int64_t worker = blockIdx.x / 2, n_workers = gridDim.x / 2;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int h = 0; h < 2; ++h) { mbar_init(tfull_bar(h), 1); mbar_init(tempty_bar(h), 8); }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc<2>(tmem_slot, 512);
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
    if (warp >= 4) tmem_st_32x32_x16(tmem_base + (uint32_t((warp & 3) * 32) << 16) + kSfCol, 0x7F7F7F7Fu);
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();

    if (warp == 0) {
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            for (int64_t u = worker; u < p.units; u += n_workers) {
                const int2 w = p.worklist[u];
                const int m0 = w.x * 512 + int(cta_rank) * 256;
                const int n0 = (w.y & kUnitColMask) * Cfg::kBN + int(cta_rank) * (Cfg::kBN / 2);
                if (p.lockstep && u != worker) lockstep_wait(p, u, worker, n_workers, 2);
                for (int64_t kb = 0; kb < p.k_blocks; ++kb) {
                    mbar_wait(empty_bar(stage), phase ^ 1);
                    if (leader) mbar_expect_tx(full_bar(stage), 2 * Cfg::kStageBytes);
                    const uint32_t lead_full = mapa_u32(full_bar(stage), 0);
                    tma_load_2d_pair(smem_a(stage), &map_a, lead_full, int(kb * kBlockK), m0);
                    tma_load_2d_pair(smem_b(stage), &map_b, lead_full, int((kb + (kb >= p.neg_from ? p.b_shift : 0)) * kBlockK), n0);
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (leader) {
            constexpr uint32_t idesc0 = make_mxf4_idesc(256, Cfg::kBN);
            int stage = 0;
            uint32_t phase = 0;
            int64_t it = 0;
            for (int64_t u = worker; u < p.units; u += n_workers, ++it) {
                const uint32_t drained = (uint32_t(it) & 1u) ^ 1u;
                for (int64_t kb = 0; kb < p.k_blocks; ++kb) {
                    const uint32_t idesc = idesc0 | ((kb >= p.neg_from && p.b_shift == 0) ? kNegateA : 0u);
                    mbar_wait(full_bar(stage), phase);
                    if (kb == 0) mbar_wait(tempty_bar(0), drained);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint64_t a_desc = make_kmajor_sw128_desc(smem_a(stage));
                        const uint64_t b_desc = make_kmajor_sw128_desc(smem_b(stage));
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma_mxf4_cg2(tmem_base, a_desc + uint64_t(k * 2), b_desc + uint64_t(k * 2), idesc,
                                          tmem_base + kSfCol, tmem_base + kSfCol + 4, (kb > 0 || k > 0) ? 1u : 0u);
                    }
                    __syncwarp();
                    if (kb == 0) { mbar_wait(tempty_bar(1), drained); tc_fence_after(); }
                    if (elect_one()) {
                        const uint64_t a_desc = make_kmajor_sw128_desc(smem_a(stage) + 128 * kBlockK);
                        const uint64_t b_desc = make_kmajor_sw128_desc(smem_b(stage));
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma_mxf4_cg2(tmem_base + kHalf1Col, a_desc + uint64_t(k * 2), b_desc + uint64_t(k * 2), idesc,
                                          tmem_base + kSfCol, tmem_base + kSfCol + 4, (kb > 0 || k > 0) ? 1u : 0u);
                        umma_commit<2>(empty_bar(stage));
                        if (kb == p.k_blocks - 1) { umma_commit<2>(tfull_bar(0)); umma_commit<2>(tfull_bar(1)); }
                    }
                    __syncwarp();
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
        __syncwarp();
    } else if (warp >= 4) {
        const int quarter = warp & 3;
        int64_t it = 0;
        for (int64_t u = worker; u < p.units; u += n_workers, ++it) {
            const int2 w = p.worklist[u];
            const int64_t n0 = int64_t(w.y & kUnitColMask) * Cfg::kBN;
            const int64_t m0 = int64_t(w.x) * 512 + int64_t(cta_rank) * 256;
            int32_t *colsum = reinterpret_cast<int32_t *>(smem_raw + (bar_base - smem_u32(smem_raw)) + 256);
            if (p.dense == 3) {
                for (int c = int(threadIdx.x) - 128; c < 256; c += 256)
                    colsum[c] = (c < Cfg::kBN && n0 + c < p.n) ? p.rowsum[n0 + c] : 0;
                asm volatile("bar.sync 1, 256;" ::: "memory");
            }
            {
                const int h = (warp - 4) >> 2;
                mbar_wait(tfull_bar(h), uint32_t(it) & 1u);
                tc_fence_after();
                const int row_in_tile = h * 128 + quarter * 32 + lane;
                drain_accumulator<OutT>(p, tmem_base + (uint32_t(quarter * 32) << 16) + uint32_t(h) * kHalf1Col,
                                        m0 + row_in_tile, n0, row_in_tile, 0, p.mirror != 0, Cfg::kBN, true,
                                        p.dense == 3 ? colsum : nullptr);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(tempty_bar(h), 0);
            }
            if (p.dense == 3) asm volatile("bar.sync 1, 256;" ::: "memory");
        }
    }
    tc_fence_before();
    cluster_sync_all();
    if (warp == 2) tmem_dealloc<2>(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------
// Probe (no data movement): how fast does one SM / one SM pair issue kind::mxf4 MMAs when the stage
// release (tcgen05.commit) comes every `every` stages?  Shared memory is whatever it holds; a ring
// of 4 barriers stands in for the empty barriers.  Reports cycles per stage (8 MMAs of 128x240x64
// per CTA), averaged over `iters` stages, in *cycles_x100.
template <int CG>
__global__ void __launch_bounds__(kThreads, 1) umma_probe_kernel(int iters, int every, unsigned long long *out) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t a_smem = smem_base, b_smem = smem_base + 256 * 128;
    const uint32_t bar_base = smem_base + 64 * 1024;
    auto bar = [&](int k) { return bar_base + 8u * k; };
    const uint32_t tmem_slot = bar_base + 64;
    const int warp = threadIdx.x >> 5;
    const uint32_t cta_rank = CG == 2 ? cluster_ctarank() : 0;
    if (warp == 1 && (threadIdx.x & 31) == 0) {
        for (int k = 0; k < 4; ++k) mbar_init(bar(k), 1);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc<CG>(tmem_slot, 512);
    tc_fence_before();
    if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
    if (warp >= 4) tmem_st_32x32_x16(tmem_base + (uint32_t((warp & 3) * 32) << 16) + 496, 0x7F7F7F7Fu);
    tc_fence_before();
    if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    if (warp == 1 && cta_rank == 0) {
        const uint32_t idesc = make_mxf4_idesc(CG == 2 ? 256 : 128, 240);
        long long t0 = 0;
        uint32_t phase[4] = {0, 0, 0, 0};
        int commits = 0;
        for (int it = 0; it < iters + 8; ++it) {
            if (it == 8) t0 = clock64();
            if (elect_one()) {
                const uint64_t a_desc = make_kmajor_sw128_desc(a_smem), a2_desc = make_kmajor_sw128_desc(a_smem + 128 * 128);
                const uint64_t b_desc = make_kmajor_sw128_desc(b_smem);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if constexpr (CG == 1) umma_mxf4(tmem_base, a_desc + uint64_t(k * 2), b_desc + uint64_t(k * 2), idesc, tmem_base + 496, tmem_base + 500, 1u);
                    else umma_mxf4_cg2(tmem_base, a_desc + uint64_t(k * 2), b_desc + uint64_t(k * 2), idesc, tmem_base + 496, tmem_base + 500, 1u);
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if constexpr (CG == 1) umma_mxf4(tmem_base + 256, a2_desc + uint64_t(k * 2), b_desc + uint64_t(k * 2), idesc, tmem_base + 496, tmem_base + 500, 1u);
                    else umma_mxf4_cg2(tmem_base + 256, a2_desc + uint64_t(k * 2), b_desc + uint64_t(k * 2), idesc, tmem_base + 496, tmem_base + 500, 1u);
                }
                if (it % every == every - 1) umma_commit<CG>(bar(commits & 3));
            }
            __syncwarp();
            if (it % every == every - 1) {
                ++commits;
                if (commits >= 3) {                                  // stay at most 3 commits ahead
                    const int k = (commits - 3) & 3;
                    mbar_wait(bar(k), phase[k]);
                    phase[k] ^= 1;
                }
            }
        }
        // drain
        if (elect_one()) umma_commit<CG>(bar(commits & 3));
        __syncwarp();
        for (int c = commits >= 3 ? commits - 2 : 0; c <= commits; ++c) { const int k = c & 3; mbar_wait(bar(k), phase[k]); phase[k] ^= 1; }
        const long long t1 = clock64();
        if ((threadIdx.x & 31) == 0) out[blockIdx.x] = (unsigned long long)((t1 - t0) * 100 / iters);
    }
    tc_fence_before();
    if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
    if (warp == 2) tmem_dealloc<CG>(tmem_base, 512);
}

int umma_probe(int cg, int every, int64_t *cycles_x100) {
    const int iters = 4000;
    unsigned long long *d = nullptr;
    BTB_CUDA(cudaMalloc(reinterpret_cast<void **>(&d), 148 * 8));
    BTB_CUDA(cudaMemset(d, 0, 148 * 8));
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    cfg.gridDim = dim3(148);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = 64 * 1024 + 1024 + 256;
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cg;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (cg == 2) {
        BTB_CUDA(cudaFuncSetAttribute(umma_probe_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(cfg.dynamicSmemBytes)));
        BTB_CUDA(cudaLaunchKernelEx(&cfg, umma_probe_kernel<2>, iters, every, d));
    } else {
        BTB_CUDA(cudaFuncSetAttribute(umma_probe_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(cfg.dynamicSmemBytes)));
        BTB_CUDA(cudaLaunchKernelEx(&cfg, umma_probe_kernel<1>, iters, every, d));
    }
    BTB_CUDA(cudaDeviceSynchronize());
    unsigned long long h[148];
    BTB_CUDA(cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost));
    cudaFree(d);
    unsigned long long sum = 0, cnt = 0;
    for (int i = 0; i < 148; ++i) if (h[i]) { sum += h[i]; ++cnt; }
    *cycles_x100 = cnt ? int64_t(sum / cnt) : 0;
    return BTB_OK;
}

// ------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

static int get_encode_fn(EncodeTiledFn *fn) {
    static EncodeTiledFn cached = nullptr;
    if (!cached) {
        void *sym = nullptr;
        cudaDriverEntryPointQueryResult q;
        BTB_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q));
        if (!sym || q != cudaDriverEntryPointSuccess) return fail(BTB_ECUDA, "cuTensorMapEncodeTiled not available");
        cached = reinterpret_cast<EncodeTiledFn>(sym);
    }
    *fn = cached;
    return BTB_OK;
}

// 2-D uint8 tensor [rows][kbytes], box = 128 K-bytes x box_rows rows, 128B swizzle, OOB rows read as 0
static int make_operand_map(CUtensorMap *map, const void *base, int64_t rows, int64_t kbytes, int box_rows) {
    constexpr int box_k = kBlockK;
    EncodeTiledFn enc;
    BTB_TRY(get_encode_fn(&enc));
    cuuint64_t dims[2] = {cuuint64_t(kbytes), cuuint64_t(rows)};
    cuuint64_t strides[1] = {cuuint64_t(kbytes)};
    cuuint32_t box[2] = {cuuint32_t(box_k), cuuint32_t(box_rows)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void *>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(BTB_ECUDA, "cuTensorMapEncodeTiled failed: %d", int(r));
    return BTB_OK;
}

int g_umma_dense_simplex = 2;  // dense code: 0 one-hot, 1 +-1 simplex, 2 0/1 tetrahedron (default, see common.cuh)
int g_umma_fp4 = 1;            // kind::mxf4 paths (see common.cuh); 0 = int8 only
int g_umma_fp4_general = 1;    // kind::mxf4 for alphabets with ignored bytes / more than four symbols
// "umma_coop": 0 launches the CTA-pair kernel without the cooperative attribute -- only for profilers (ncu cannot
// replay a cooperative cluster launch); the grid is one CTA per SM, so all CTAs are co-resident whenever the GPU
// is otherwise idle.  The default is 1 unless a profiler has injected itself into the process (Nsight Compute
// sets these variables for its target), so that a profiled run lists this kernel like any other.
static bool profiler_attached() {
    return getenv("NV_NSIGHT_INJECTION_PORT_BASE") || getenv("NV_NSIGHT_INJECTION_TRANSPORT_TYPE") || getenv("CUDA_INJECTION64_PATH") ||
           getenv("NV_COMPUTE_PROFILER_PERFWORKS_DIR");
}
int g_umma_coop = profiler_attached() ? 0 : 1;
int g_umma_fp4_pair = 1;       // "umma_fp4_pair": cta_group::2 version of the FP4 full-tile kernel (0: single-CTA kernel)
int g_umma_tile_mode = 1;      // 1: one CTA per full 256x256 tile (shared B block), 0: half-tile units
int g_umma_lockstep = 1;       // 0: off; k >= 1: producers of all CTAs start each unit at most k-1 rounds apart (L2 reuse), cooperative launch

// Arrival counters of the lockstep barrier.  A launch OWNS its counter until the kernel has finished: the slot is
// taken from a per-device pool, cleared on the launching stream, and an event recorded behind the kernel marks
// when it may be handed out again -- so a launch queued on another stream can never clear a counter that a
// running (or still pending) kernel waits on.  When all slots are pending the oldest one is waited for.
namespace {
struct LockSlot { unsigned int *ptr = nullptr; cudaEvent_t done = nullptr; bool pending = false; };
struct LockPool { unsigned int *mem = nullptr; std::vector<LockSlot> slots; size_t next = 0; };
std::mutex g_lock_mutex;
std::vector<LockPool> g_lock_pools;            // indexed by device ordinal
constexpr int kLockSlots = 64;

int lock_acquire(unsigned int **ptr, int *slot_index, cudaStream_t stream) {
    int dev = 0;
    BTB_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(g_lock_mutex);
    if (int(g_lock_pools.size()) <= dev) g_lock_pools.resize(size_t(dev) + 1);
    LockPool &pool = g_lock_pools[size_t(dev)];
    if (!pool.mem) {
        BTB_CUDA(cudaMalloc(reinterpret_cast<void **>(&pool.mem), kLockSlots * 128));
        pool.slots.resize(kLockSlots);
        for (int k = 0; k < kLockSlots; ++k) {
            pool.slots[size_t(k)].ptr = pool.mem + k * 32;          // 128 bytes apart
            BTB_CUDA(cudaEventCreateWithFlags(&pool.slots[size_t(k)].done, cudaEventDisableTiming));
        }
    }
    int pick = -1;
    for (int k = 0; k < kLockSlots && pick < 0; ++k) {
        LockSlot &sl = pool.slots[(pool.next + size_t(k)) % kLockSlots];
        if (!sl.pending || cudaEventQuery(sl.done) == cudaSuccess) pick = int((pool.next + size_t(k)) % kLockSlots);
    }
    if (pick < 0) {                                                  // every slot belongs to a kernel still in flight
        pick = int(pool.next % kLockSlots);
        BTB_CUDA(cudaEventSynchronize(pool.slots[size_t(pick)].done));
    }
    cudaGetLastError();                                              // cudaErrorNotReady from the queries is not an error
    pool.next = size_t(pick) + 1;
    pool.slots[size_t(pick)].pending = false;
    BTB_CUDA(cudaMemsetAsync(pool.slots[size_t(pick)].ptr, 0, 4, stream));
    *ptr = pool.slots[size_t(pick)].ptr;
    *slot_index = pick;
    return BTB_OK;
}
// behind the kernel launch, on its stream
int lock_commit(int slot_index, cudaStream_t stream) {
    int dev = 0;
    BTB_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(g_lock_mutex);
    LockSlot &sl = g_lock_pools[size_t(dev)].slots[size_t(slot_index)];
    BTB_CUDA(cudaEventRecord(sl.done, stream));
    sl.pending = true;
    return BTB_OK;
}

// cooperative launch of a kernel whose producers run in lockstep (when `lockstep` and the launch has more units
// than workers); a plain launch otherwise
template <typename Kernel>
int launch_lockstep(Kernel kernel, dim3 grid, dim3 block, int smem_bytes, int cluster, bool want_lockstep, bool cooperative,
                    const CUtensorMap &ma, const CUtensorMap &mb, UmmaParams p, cudaStream_t stream) {
    BTB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[2];
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = size_t(smem_bytes);
    cfg.stream = stream;
    cfg.attrs = attr;
    cfg.numAttrs = 0;
    if (cluster > 1) {
        attr[cfg.numAttrs].id = cudaLaunchAttributeClusterDimension;
        attr[cfg.numAttrs].val.clusterDim.x = unsigned(cluster);
        attr[cfg.numAttrs].val.clusterDim.y = 1;
        attr[cfg.numAttrs].val.clusterDim.z = 1;
        ++cfg.numAttrs;
    }
    int slot = -1;
    if (want_lockstep) {
        unsigned int *counter = nullptr;
        BTB_TRY(lock_acquire(&counter, &slot, stream));
        p.lockstep = counter;
        p.lock_slack = g_umma_lockstep - 1;
        if (cooperative) {
            attr[cfg.numAttrs].id = cudaLaunchAttributeCooperative;
            attr[cfg.numAttrs].val.cooperative = 1;
            ++cfg.numAttrs;
        }
    }
    BTB_CUDA(cudaLaunchKernelEx(&cfg, kernel, ma, mb, p));
    if (slot >= 0) BTB_TRY(lock_commit(slot, stream));
    return BTB_OK;
}
}  // namespace

static int device_sms(int *sms) {
    int dev = 0;
    BTB_CUDA(cudaGetDevice(&dev));
    BTB_CUDA(cudaDeviceGetAttribute(sms, cudaDevAttrMultiProcessorCount, dev));
    return BTB_OK;
}

// how many CTA pairs of the FP4 pair kernel (cg = 2) / CTAs of the int8 kernel (cg = 1) the device keeps resident
int umma_max_active_clusters(int cg, int *out) {
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    cfg.gridDim = dim3(148);
    cfg.blockDim = dim3(cg == 2 ? kThreadsWide : kThreads);
    cfg.dynamicSmemBytes = cg == 2 ? Fp4PairCfg::kSmemBytes : UmmaCfg::kSmemBytes;
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cg == 2 ? 2 : 1;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (cg == 2) {
        BTB_CUDA(cudaFuncSetAttribute(dist_umma_fp4_pair_kernel<uint16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, Fp4PairCfg::kSmemBytes));
        BTB_CUDA(cudaOccupancyMaxActiveClusters(out, dist_umma_fp4_pair_kernel<uint16_t>, &cfg));
    } else {
        BTB_CUDA(cudaFuncSetAttribute(dist_umma_kernel<uint16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, UmmaCfg::kSmemBytes));
        BTB_CUDA(cudaOccupancyMaxActiveClusters(out, dist_umma_kernel<uint16_t>, &cfg));
    }
    return BTB_OK;
}

// ---- explicit work lists (device resident, cached per device) ----------------------------------
// kind 0: square tiles (I, J >= I) of block rows [lo, hi); kind 1: rectangular 256 x 240 tiles (I, j);
// kind 2: CTA-pair units (P, j) of 512 rows x 240 columns.  All in band order relative to `lo`: bands of
// kBand block rows, column-major inside a band.
struct DevList { int64_t n, lo, hi; int kind; int2 *d; int64_t count; };
static std::vector<std::vector<DevList>> g_lists;            // indexed by device ordinal
static std::mutex g_lists_mutex;

int get_rows_list(int64_t n, int64_t lo, int64_t hi, int kind, const int2 **d_out, int64_t *count) {
    int dev = 0;
    BTB_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(g_lists_mutex);
    if (int(g_lists.size()) <= dev) g_lists.resize(size_t(dev) + 1);
    std::vector<DevList> &lists = g_lists[size_t(dev)];
    for (auto &e : lists)
        if (e.n == n && e.lo == lo && e.hi == hi && e.kind == kind) { *d_out = e.d; *count = e.count; return BTB_OK; }
    if (lists.size() > 512) {                            // (never reached by one alignment: a list per band and kind)
        BTB_CUDA(cudaDeviceSynchronize());
        for (auto &e : lists) cudaFree(e.d);
        lists.clear();
    }
    const int64_t T = tiles_per_dim(n);
    std::vector<int2> h;
    for (int64_t r0 = lo; r0 < hi; r0 += kBand) {
        const int64_t r1 = std::min(hi, r0 + kBand);
        if (kind == 0) {
            for (int64_t J = r0; J < T; ++J)
                for (int64_t I = r0; I < r1 && I <= J; ++I) h.push_back(make_int2(int(I), int(J)));
        } else if (kind == 1) {
            const int64_t NJ = (n + Fp4TileCfg::kBN - 1) / Fp4TileCfg::kBN;
            for (int64_t j = (r0 * BTB_TILE) / Fp4TileCfg::kBN; j < NJ; ++j)
                for (int64_t I = r0; I < r1; ++I)
                    if ((I * BTB_TILE) / Fp4TileCfg::kBN <= j) h.push_back(make_int2(int(I), int(j)));
        } else {
            const int64_t NJ = (n + Fp4PairCfg::kBN - 1) / Fp4PairCfg::kBN;
            for (int64_t j = (r0 * BTB_TILE) / Fp4PairCfg::kBN; j < NJ; ++j)
                for (int64_t P = r0 / 2; P < (r1 + 1) / 2; ++P)
                    if ((P * 512) / Fp4PairCfg::kBN <= j) h.push_back(make_int2(int(P), int(j)));
        }
    }
    DevList e{n, lo, hi, kind, nullptr, int64_t(h.size())};
    BTB_CUDA(cudaMalloc(reinterpret_cast<void **>(&e.d), std::max<size_t>(h.size(), 1) * sizeof(int2)));
    if (!h.empty()) {
        // A plain cudaMemcpy from pageable memory may return while its DMA is still in flight on the
        // legacy stream, which the non-blocking launch streams do not wait for: copy on a stream of
        // our own and wait for it, so the list is complete for every later launch on any stream.
        cudaStream_t cs;
        BTB_CUDA(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
        cudaError_t ce = cudaMemcpyAsync(e.d, h.data(), h.size() * sizeof(int2), cudaMemcpyHostToDevice, cs);
        if (ce == cudaSuccess) ce = cudaStreamSynchronize(cs);
        cudaStreamDestroy(cs);
        if (ce != cudaSuccess) { cudaFree(e.d); return fail(BTB_ECUDA, "work list upload failed: %s", cudaGetErrorString(ce)); }
    }
    lists.push_back(e);
    *d_out = e.d;
    *count = e.count;
    return BTB_OK;
}

static UmmaParams base_params(const uint8_t *A, int64_t n, int64_t kb, int64_t neg_from, void *d_out, int64_t out_ld, int out_mode, int mirror) {
    UmmaParams p;
    p.n = n; p.k_blocks = kb / kBlockK; p.tile_lo = 0; p.units = 0;
    p.out = d_out; p.ld = out_ld; p.out_mode = out_mode; p.mirror = mirror;
    p.lockstep = nullptr; p.fp32acc = 1; p.worklist = nullptr; p.sq_list = nullptr;
    p.dense = neg_from < 0 ? 3 : 0; p.bias = 0;             // tetrahedron code (row sums) / V - X planes
    p.rowsum = reinterpret_cast<const int32_t *>(A + n * kb);
    if (neg_from >= 0) p.neg_from = neg_from;
    return p;
}

// FP4 full-tile kernels over the block rows [row_lo, row_hi); *launched = 0 when the range is too
// small to fill the GPU twice (the caller then uses the half-tile kernel)
static int launch_fp4_rows(const uint8_t *A, int64_t n, int64_t kb, int64_t neg_from, int64_t row_lo, int64_t row_hi, void *d_out,
                           int out_elem_bytes, int64_t out_ld, int mirror, cudaStream_t stream, int *launched) {
    *launched = 0;
    int sms = 0;
    BTB_TRY(device_sms(&sms));
    const int2 *d_units = nullptr;
    int64_t n_units = 0;
    CUtensorMap ma, mb;
    if (g_umma_fp4_pair && out_elem_bytes == 2 && row_lo % 2 == 0 && (row_hi % 2 == 0 || row_hi == tiles_per_dim(n))) {
        // CTA pairs: 512-row units, B split between the two SMs of a pair
        BTB_TRY(get_rows_list(n, row_lo, row_hi, 2, &d_units, &n_units));
        if (n_units >= 2 * int64_t(sms / 2)) {
            BTB_TRY(make_operand_map(&ma, A, n, kb, 256));
            BTB_TRY(make_operand_map(&mb, A, n, kb, Fp4PairCfg::kBN / 2));
            UmmaParams p = base_params(A, n, kb, neg_from, d_out, out_ld, 0, mirror);
            p.units = n_units; p.worklist = d_units;
            // the lockstep barrier needs every pair resident: never launch more pairs than the device can hold
            int max_pairs = 0;
            BTB_TRY(umma_max_active_clusters(2, &max_pairs));
            const int64_t pairs = std::min<int64_t>(std::min<int64_t>(sms / 2, std::max(1, max_pairs)), p.units);
            BTB_TRY(launch_lockstep(dist_umma_fp4_pair_kernel<uint16_t>, dim3(unsigned(pairs * 2)), dim3(kThreadsWide), Fp4PairCfg::kSmemBytes, 2,
                                    g_umma_lockstep && p.units > pairs, g_umma_coop != 0, ma, mb, p, stream));
            *launched = 1;
            return BTB_OK;
        }
    }
    BTB_TRY(get_rows_list(n, row_lo, row_hi, 1, &d_units, &n_units));
    if (n_units < 2 * int64_t(sms)) return BTB_OK;
    BTB_TRY(make_operand_map(&ma, A, n, kb, 256));
    BTB_TRY(make_operand_map(&mb, A, n, kb, Fp4TileCfg::kBN));
    UmmaParams p = base_params(A, n, kb, neg_from, d_out, out_ld, 0, mirror);
    p.units = n_units; p.worklist = d_units;
    const unsigned ctas = unsigned(std::min<int64_t>(sms, p.units));
    if (out_elem_bytes == 2)
        BTB_TRY(launch_lockstep(dist_umma_fp4_tile_kernel<uint16_t>, dim3(ctas), dim3(kThreadsWide), Fp4TileCfg::kSmemBytes, 1,
                                g_umma_lockstep && p.units > ctas, true, ma, mb, p, stream));
    else
        BTB_TRY(launch_lockstep(dist_umma_fp4_tile_kernel<uint32_t>, dim3(ctas), dim3(kThreadsWide), Fp4TileCfg::kSmemBytes, 1,
                                g_umma_lockstep && p.units > ctas, true, ma, mb, p, stream));
    *launched = 1;
    return BTB_OK;
}

int distance_tiles_umma(const void *d_operands, int64_t n, int64_t len, int sigma, int64_t tile_lo,
                        int64_t tile_hi, void *d_out, int out_elem_bytes, int64_t out_ld, int out_mode,
                        int mirror, cudaStream_t stream, const int2 *sq_list) {
    const int64_t tiles = tile_hi - tile_lo;
    if (tiles <= 0) return BTB_OK;
    const int64_t kb = umma_k_bytes(len, sigma);
    const uint8_t *A = static_cast<const uint8_t *>(d_operands);
    const uint8_t *B = sigma_dense(sigma) ? A : A + n * kb;   // dense: one operand serves both sides
    if (kb == 0) return fail(BTB_EINVAL, "UMMA engine needs len > 0");
    const int64_t neg_from = umma_fp4_general(sigma, len) ? umma_fp4_plane_bytes(len) / kBlockK : -1;
    int sms = 0;
    BTB_TRY(device_sms(&sms));
    CUtensorMap ma, mb;
    if (!sq_list && umma_use_fp4(sigma, len) && g_umma_tile_mode == 1 && out_mode == 0) {
        // whole bands of block rows (the full matrix, the level-1/2 band launches): rectangular full tiles
        const int64_t Tt = tiles_per_dim(n);
        int64_t first = 0, rb0 = -1, rb1 = -1;
        for (int64_t r0 = 0; r0 < Tt + kBand; r0 += kBand) {       // band starts, plus one step past the end
            if (first == tile_lo) rb0 = r0;
            if (first == tile_hi) rb1 = std::min(r0, Tt);
            if (r0 < Tt) first += band_tiles(Tt, r0, std::min<int64_t>(kBand, Tt - r0));
        }
        if (rb0 >= 0 && rb1 > rb0) {
            int launched = 0;
            BTB_TRY(launch_fp4_rows(A, n, kb, neg_from, rb0, rb1, d_out, out_elem_bytes, out_ld, mirror, stream, &launched));
            if (launched) return BTB_OK;
        }
    }
    if (umma_use_fp4(sigma, len)) {
        // FP4 half-tile units: small launches and explicit tile lists
        BTB_TRY(make_operand_map(&ma, A, n, kb, 128));
        BTB_TRY(make_operand_map(&mb, A, n, kb, 256));
        UmmaParams p = base_params(A, n, kb, neg_from, d_out, out_ld, out_mode, mirror);
        p.tile_lo = tile_lo; p.units = tiles * 2; p.sq_list = sq_list;
        const unsigned ctas = unsigned(std::min<int64_t>(sms, p.units));
        if (out_elem_bytes == 2)
            return launch_lockstep(dist_umma_fp4_kernel<uint16_t>, dim3(ctas), dim3(kThreads), UmmaCfg::kSmemBytes, 1,
                                   g_umma_lockstep && p.units > ctas, true, ma, mb, p, stream);
        return launch_lockstep(dist_umma_fp4_kernel<uint32_t>, dim3(ctas), dim3(kThreads), UmmaCfg::kSmemBytes, 1,
                               g_umma_lockstep && p.units > ctas, true, ma, mb, p, stream);
    }
    // kind::i8 (rows beyond the fp32-exact range of the FP4 paths, or "umma_fp4" = 0)
    UmmaParams p;
    p.n = n; p.k_blocks = kb / kBlockK; p.tile_lo = tile_lo;
    p.out = d_out; p.ld = out_ld; p.out_mode = out_mode; p.mirror = mirror;
    p.lockstep = nullptr; p.fp32acc = 0; p.worklist = nullptr; p.sq_list = sq_list;
    p.dense = sigma_dense(sigma) ? (g_umma_dense_simplex == 2 ? 3 : g_umma_dense_simplex ? 2 : 1) : 0;
    p.bias = int32_t(g_umma_dense_simplex ? 3 * len : len);
    p.rowsum = reinterpret_cast<const int32_t *>(A + n * kb);
    // full tiles only pay off once every SM gets at least two of them; small problems keep the
    // half-tile units (twice the parallelism)
    if (g_umma_tile_mode == 1 && tiles >= 2 * int64_t(sms)) {
        BTB_TRY(make_operand_map(&ma, A, n, kb, 256));
        BTB_TRY(make_operand_map(&mb, B, n, kb, 256));
        p.units = tiles;
        const unsigned ctas = unsigned(std::min<int64_t>(sms, p.units));
        if (out_elem_bytes == 2)
            return launch_lockstep(dist_umma_tile_kernel<uint16_t>, dim3(ctas), dim3(kThreads), TileCfg::kSmemBytes, 1,
                                   g_umma_lockstep && p.units > ctas, true, ma, mb, p, stream);
        return launch_lockstep(dist_umma_tile_kernel<uint32_t>, dim3(ctas), dim3(kThreads), TileCfg::kSmemBytes, 1,
                               g_umma_lockstep && p.units > ctas, true, ma, mb, p, stream);
    }
    BTB_TRY(make_operand_map(&ma, A, n, kb, 128));
    BTB_TRY(make_operand_map(&mb, B, n, kb, 256));
    p.units = tiles * 2;
    const unsigned ctas = unsigned(std::min<int64_t>(sms, p.units));
    if (out_elem_bytes == 2)
        return launch_lockstep(dist_umma_kernel<uint16_t>, dim3(ctas), dim3(kThreads), UmmaCfg::kSmemBytes, 1,
                               g_umma_lockstep && p.units > ctas, true, ma, mb, p, stream);
    return launch_lockstep(dist_umma_kernel<uint32_t>, dim3(ctas), dim3(kThreads), UmmaCfg::kSmemBytes, 1,
                           g_umma_lockstep && p.units > ctas, true, ma, mb, p, stream);
}


// every cell (r, c >= r) with r in the block rows [row_lo, row_hi), mirrored when asked: n x n output
int distance_rows_umma(const void *d_operands, int64_t n, int64_t len, int sigma, int64_t row_lo, int64_t row_hi,
                       void *d_out, int out_elem_bytes, int64_t out_ld, int mirror, cudaStream_t stream) {
    if (row_hi <= row_lo) return BTB_OK;
    const int64_t kb = umma_k_bytes(len, sigma);
    const uint8_t *A = static_cast<const uint8_t *>(d_operands);
    if (umma_use_fp4(sigma, len) && g_umma_tile_mode == 1) {
        // the CTA-pair kernel works on pairs of block rows: an odd end (other than the matrix end) is
        // split off, so that no row outside the requested range is ever written
        if (g_umma_fp4_pair && out_elem_bytes == 2 && row_lo % 2 == 0 && row_hi % 2 == 1 && row_hi < tiles_per_dim(n) &&
            row_hi - 1 > row_lo) {
            BTB_TRY(distance_rows_umma(d_operands, n, len, sigma, row_lo, row_hi - 1, d_out, out_elem_bytes, out_ld, mirror, stream));
            return distance_rows_umma(d_operands, n, len, sigma, row_hi - 1, row_hi, d_out, out_elem_bytes, out_ld, mirror, stream);
        }
        int launched = 0;
        const int64_t neg_from = umma_fp4_general(sigma, len) ? umma_fp4_plane_bytes(len) / kBlockK : -1;
        BTB_TRY(launch_fp4_rows(A, n, kb, neg_from, row_lo, row_hi, d_out, out_elem_bytes, out_ld, mirror, stream, &launched));
        if (launched) return BTB_OK;
    }
    const int2 *d_list = nullptr;
    int64_t count = 0;
    BTB_TRY(get_rows_list(n, row_lo, row_hi, 0, &d_list, &count));
    return distance_tiles_umma(d_operands, n, len, sigma, 0, count, d_out, out_elem_bytes, out_ld, 0, mirror, stream, d_list);
}

}  // namespace btb
