// Shared helpers for libbtbphylo.so: error plumbing, tile order, small device utilities.
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <string>

#include "../../include/btb_phylo.h"

namespace btb {

// ---- thread-local error text (btb_last_error) -------------------------------------------
std::string &last_error();
int fail(int code, const char *fmt, ...);

#define BTB_CUDA(call)                                                                      \
    do {                                                                                    \
        cudaError_t e_ = (call);                                                            \
        if (e_ != cudaSuccess)                                                              \
            return btb::fail(BTB_ECUDA, "%s failed at %s:%d: %s", #call, __FILE__, __LINE__, \
                             cudaGetErrorString(e_));                                       \
    } while (0)

#define BTB_TRY(call)                \
    do {                             \
        int rc_ = (call);            \
        if (rc_ != BTB_OK) return rc_; \
    } while (0)

// ---- upper-triangle tile order ------------------------------------------------------------
// Tiles are BTB_TILE x BTB_TILE blocks (I, J) with I <= J.  They are numbered band by band: a
// band is kBand consecutive block ROWS; inside a band the order is column-major (J ascending,
// then I).  Two properties follow:
//   * 74-148 CTAs working on consecutive indices touch the band's <= kBand row panels plus ~10-19
//     column panels at a time, which keeps the operand panels L2-resident;
//   * with mirrored writes, every matrix row of a band is final as soon as the band is done (its
//     left part was written by the mirrors of earlier bands), so finished row bands can be copied
//     to the host and formatted while later bands are still being computed.
constexpr int kBand = 8;

__host__ __device__ inline int64_t tiles_per_dim(int64_t n) { return (n + BTB_TILE - 1) / BTB_TILE; }

__host__ __device__ inline int64_t tile_count(int64_t n) {
    int64_t t = tiles_per_dim(n);
    return t * (t + 1) / 2;
}

// tiles in the band of block rows [r0, r0 + w)
__host__ __device__ inline int64_t band_tiles(int64_t T, int64_t r0, int64_t w) {
    return w * (w + 1) / 2 + (T - r0 - w) * w;
}

__host__ __device__ inline void tile_coords(int64_t T, int64_t idx, int &I, int &J) {
    for (int64_t r0 = 0; r0 < T; r0 += kBand) {
        const int64_t w = r0 + kBand < T ? kBand : T - r0;
        const int64_t tri = w * (w + 1) / 2;            // the band's diagonal block: column r0+c holds c+1 tiles
        if (idx < tri) {
            for (int64_t c = 0; c < w; ++c) {
                if (idx < c + 1) { I = int(r0 + idx); J = int(r0 + c); return; }
                idx -= c + 1;
            }
        }
        idx -= tri;
        const int64_t rect = (T - r0 - w) * w;          // columns right of the diagonal block: w tiles each
        if (idx < rect) { I = int(r0 + idx % w); J = int(r0 + w + idx / w); return; }
        idx -= rect;
    }
    I = J = -1;
}

// first tile index of the band that starts at block row r0 (r0 a multiple of kBand)
__host__ __device__ inline int64_t band_first_tile(int64_t T, int64_t r0) {
    int64_t idx = 0;
    for (int64_t b = 0; b < r0; b += kBand) idx += band_tiles(T, b, b + kBand < T ? kBand : T - b);
    return idx;
}

// ---- operand layouts of the two distance engines (see dist_encode.cu) -----------------------
__host__ __device__ inline bool sigma_dense(int sigma) { return sigma == BTB_SIGMA_DENSE4; }
__host__ __device__ inline int popc_code_planes(int sigma) {
    if (sigma_dense(sigma)) return 2;
    int q = 1;
    while ((1 << q) < sigma) ++q;
    return q;
}
__host__ __device__ inline int64_t popc_words(int64_t len) {      // words per row, multiple of 16
    int64_t w = (len + 31) / 32;
    return (w + 15) / 16 * 16;
}
__host__ __device__ inline int sigma4_of(int sigma) { return sigma <= 4 ? 4 : (sigma + 3) / 4 * 4; }
// dense UMMA operand (alignments without ignored bytes; btb_set_option("umma_dense_simplex", 0|1|2)):
//   0 = one-hot, single operand, 4 bytes per site (d = len - matches);
//   1 = +-1 regular simplex, 3 bytes per site (d = (3 len - dot) / 4): 25 % fewer MMAs, but dense +-1
//       data drives the tensor pipe into the power cap (measured: SM clock 1965 -> 1470-1770 MHz);
//   2 = 0/1 regular tetrahedron (0,0,0) (1,1,0) (1,0,1) (0,1,1), 3 bytes per site: any two different
//       vertices are at squared distance 2, so d(i,j) = c_i + c_j - dot(i,j) with c_i = number of
//       non-(0,0,0) sites of row i (int32 kept behind the operand rows).  Same 25 % saving, sparse
//       0/1 data, no throttling (measured 571 W, 1965 MHz): the default.
extern int g_umma_dense_simplex;
// g_umma_fp4 ("umma_fp4", default on): dense alignments with the 0/1 tetrahedron code are stored as
// packed 4-bit e2m1 values (0 -> 0x0, 1.0 -> 0x2), 1.5 bytes per site, and multiplied with
// tcgen05.mma kind::mxf4 (unit block scales) at twice the int8 rate.  The fp32 accumulators hold
// exact integers as long as a row's dot product stays below 2^24, i.e. 3 len < 2^24.
extern int g_umma_fp4;
inline bool umma_fp4_dense(int sigma, int64_t len) {
    return sigma_dense(sigma) && g_umma_fp4 && g_umma_dense_simplex == 2 && 3 * len < (int64_t(1) << 24);
}
// g_umma_fp4_general ("umma_fp4_general", default on): every other alphabet (ignored bytes present,
// or -a with more than four symbols) as sigma + 1 nibble planes per row -- validity V, then the
// one-hot planes X_s -- in ONE operand that serves both sides: d = V_i.V_j - sum_s X_s,i.X_s,j, the
// X blocks multiplied with the descriptor's negate-A bit.  (sigma + 1) / 2 bytes per site instead of
// 2 * sigma int8 bytes, and the mxf4 rate.
extern int g_umma_fp4_general;
inline bool umma_fp4_general(int sigma, int64_t len) {
    return !sigma_dense(sigma) && sigma >= 1 && g_umma_fp4 && g_umma_fp4_general && len < (int64_t(1) << 24);
}
inline bool umma_use_fp4(int sigma, int64_t len) { return umma_fp4_dense(sigma, len) || umma_fp4_general(sigma, len); }
inline int64_t umma_fp4_plane_bytes(int64_t len) { return (len + 255) / 256 * 128; }   // one nibble plane, whole K blocks
inline int umma_bytes_per_site(int sigma) { return sigma_dense(sigma) ? (g_umma_dense_simplex ? 3 : 4) : sigma4_of(sigma); }
inline int64_t umma_k_bytes(int64_t len, int sigma) {   // row length, multiple of 128
    if (umma_fp4_general(sigma, len)) return (sigma + 1) * umma_fp4_plane_bytes(len);
    int64_t k = umma_fp4_dense(sigma, len) ? (len * 3 + 1) / 2 : len * umma_bytes_per_site(sigma);
    return (k + 127) / 128 * 128;
}

// per-site symbol counts of several ranks (sharded encode, dist_comm.cu); the pointers may be peer memory
struct CountPtrs { const uint32_t *p[16]; int n; };

// ---- device helpers -------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

template <typename T>
__device__ __forceinline__ T ld_stream(const T *p) { return __ldcs(p); }
#endif

}  // namespace btb
