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
// Tiles are BTB_TILE x BTB_TILE blocks (I, J) with I <= J.  They are numbered band by band:
// a band is kBand consecutive block columns; inside a band the order is row-major (I ascending,
// then J).  74 CTA pairs working on consecutive indices therefore touch ~kBand column panels
// and ~10 row panels at a time, which is what keeps the operand panels L2-resident.
constexpr int kBand = 8;

__host__ __device__ inline int64_t tiles_per_dim(int64_t n) { return (n + BTB_TILE - 1) / BTB_TILE; }

__host__ __device__ inline int64_t tile_count(int64_t n) {
    int64_t t = tiles_per_dim(n);
    return t * (t + 1) / 2;
}

__host__ __device__ inline void tile_coords(int64_t T, int64_t idx, int &I, int &J) {
    for (int64_t c0 = 0; c0 < T; c0 += kBand) {
        int64_t c1 = c0 + kBand < T ? c0 + kBand : T;
        int64_t w = c1 - c0;
        int64_t rect = c0 * w;                 // rows above the band's diagonal block: w tiles each
        int64_t tri = w * (w + 1) / 2;         // the band's diagonal block
        if (idx < rect) { I = int(idx / w); J = int(c0 + idx % w); return; }
        idx -= rect;
        if (idx < tri) {
            for (int64_t r = 0; r < w; ++r) {  // row c0 + r holds w - r tiles
                if (idx < w - r) { I = int(c0 + r); J = int(c0 + r + idx); return; }
                idx -= w - r;
            }
        }
        idx -= tri;
    }
    I = J = -1;
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
__host__ __device__ inline int64_t umma_k_bytes(int64_t len, int sigma) {   // row length, multiple of 128
    int64_t k = len * (sigma_dense(sigma) ? 3 : sigma4_of(sigma));
    return (k + 127) / 128 * 128;
}

// ---- device helpers -------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

template <typename T>
__device__ __forceinline__ T ld_stream(const T *p) { return __ldcs(p); }
#endif

}  // namespace btb
