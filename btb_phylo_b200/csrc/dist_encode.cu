// k4 operand preparation: SNP alignment bytes -> what the distance engines consume.
//
// Replaces the load loop of snp-dists 0.8.2 (toupper unless -k, non-ACGT -> '.'
// unless -a), reached in the reference through btbphylo/phylogeny.py:149.  A 256-entry symbol table
// carries those rules: table[byte] = code 0..sigma-1, or 0xFF for bytes that never count.
//
//   POPC engine : Q = ceil(log2 sigma) code bit-planes + 1 validity plane, 32 sites per word.
//                 Bit order inside a row is a fixed permutation of the sites (a lane reads 4
//                 consecutive bytes, a ballot gathers byte b of all lanes), identical for every
//                 row -- a distance is a sum over sites, so any common order is exact.
//   UMMA engine, kind::mxf4 (default): packed e2m1 nibbles (1.0 = 0x2), ONE operand for both sides.
//                 dense (BTB_SIGMA_DENSE4: <= 4 symbols, nothing ignored -- every snp-sites -c output):
//                   3 nibbles per site, the 0/1 vertices of a regular tetrahedron (0,0,0) (1,1,0)
//                   (1,0,1) (0,1,1) + an int32 row sum c_i:  d = c_i + c_j - dot.  Per site the most
//                   frequent symbol takes the zero vertex (site_counts / site_major kernels).
//                 general: sigma + 1 nibble planes per row (validity V, one-hot X_s), each padded
//                   to whole 128-byte K blocks:  d = V.V - sum_s X_s.X_s  (negate-A descriptor bit).
//   UMMA engine, kind::i8 (umma_fp4 = 0 or rows beyond the fp32-exact range): K-major int8,
//                 sigma4 = roundup(sigma, 4) bytes per site:
//                 A[i][site*sigma4 + b] = [valid and code == b]            (one-hot)
//                 B[i][site*sigma4 + b] = [valid and code != b, b < sigma] (masked complement)
//                 so that  sum_b A_i * B_j = [valid_i][valid_j][code_i != code_j]  = snp-dists' test;
//                 dense: one operand, 3 bytes per site (tetrahedron code as above, or the +-1 simplex
//                 (1,1,1) (1,-1,-1) (-1,1,-1) (-1,-1,1) with d = (3 len - dot) / 4, or one-hot).
//   POPC, dense : the validity plane is dropped.
//
// HBM-bound streaming kernels: algorithmic bytes = n*len read + operand bytes written.
#include <algorithm>

#include "common.cuh"

namespace btb {

int g_umma_major_zero = 1;     // "umma_major_zero": per site, the most frequent symbol takes the all-zero vertex of the dense code

// ------------------------------------------------------------------ alphabet scan (-a mode)
__global__ void presence_kernel(const uint8_t *__restrict__ seqs, int64_t n, int64_t len,
                                int64_t row_stride, uint32_t *__restrict__ present /*[8]*/) {
    // one flag byte per byte value in shared memory: marking a value is a single (racy but idempotent)
    // store, so the scan runs at load speed; 16-byte loads when the rows allow it
    __shared__ uint8_t seen[256];
    seen[threadIdx.x] = 0;                                  // blockDim.x == 256
    __syncthreads();
    const bool vec = (reinterpret_cast<uintptr_t>(seqs) % 16 == 0) && (row_stride % 16 == 0);
    const int64_t chunks = vec ? len / 16 : 0;
    for (int64_t row = blockIdx.y; row < n; row += gridDim.y) {
        const uint8_t *p = seqs + row * row_stride;
        for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < chunks;
             k += (int64_t)gridDim.x * blockDim.x) {
            const uint4 v = *reinterpret_cast<const uint4 *>(p + 16 * k);
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                seen[w[q] & 0xff] = 1; seen[(w[q] >> 8) & 0xff] = 1;
                seen[(w[q] >> 16) & 0xff] = 1; seen[w[q] >> 24] = 1;
            }
        }
        for (int64_t k = chunks * 16 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < len;
             k += (int64_t)gridDim.x * blockDim.x)
            seen[p[k]] = 1;
    }
    __syncthreads();
    if (seen[threadIdx.x]) atomicOr(&present[threadIdx.x >> 5], 1u << (threadIdx.x & 31));
}

// ------------------------------------------------------------------ POPC planes
// One warp packs 1024 sites (32 words) of one row.  Lane l of iteration it reads the 4 bytes of
// sites 128*it + 4*l .. +3; word 4*it + b holds, at bit l, site 128*it + 4*l + b.
template <bool ALIGNED>
__global__ void __launch_bounds__(256)
encode_planes_kernel(const uint8_t *__restrict__ seqs, int64_t n, int64_t len, int64_t row_stride,
                     const uint8_t *__restrict__ table_g, int q_planes, int has_valid, int64_t words,
                     uint32_t *__restrict__ planes) {
    __shared__ uint8_t table[256];
    table[threadIdx.x] = table_g[threadIdx.x];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t blocks_per_row = words / 32 + (words % 32 != 0);
    const int64_t row = warp / blocks_per_row;
    if (row >= n) return;
    const int64_t w0 = (warp % blocks_per_row) * 32;
    const uint8_t *p = seqs + row * row_stride;
    uint32_t keep[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};   // lane j keeps word w0 + j of every plane
#pragma unroll 1
    for (int it = 0; it < 8; ++it) {
        const int64_t site = (w0 + 4 * it) * 32 + 4 * lane;
        uint32_t raw = 0x2e2e2e2eu;                    // '.' never counts under any table (dense: code 0 everywhere)
        if (ALIGNED && site + 3 < len) {
            raw = *reinterpret_cast<const uint32_t *>(p + site);
        } else {
#pragma unroll
            for (int b = 0; b < 4; ++b)
                if (site + b < len) raw = (raw & ~(0xffu << (8 * b))) | (uint32_t(p[site + b]) << (8 * b));
        }
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const uint32_t code = table[(raw >> (8 * b)) & 0xff];
            const bool valid = code != 0xff;
            const int j = 4 * it + b;
            const uint32_t v = __ballot_sync(0xffffffffu, valid);
            if (lane == j) keep[8] = v;
#pragma unroll
            for (int pl = 0; pl < 8; ++pl) {
                if (pl < q_planes) {                       // warp-uniform
                    const uint32_t bits = __ballot_sync(0xffffffffu, valid && ((code >> pl) & 1));
                    if (lane == j) keep[pl] = bits;
                }
            }
        }
    }
    const int64_t w = w0 + lane;
    if (w < words) {
        const int64_t plane_stride = n * words;
#pragma unroll
        for (int pl = 0; pl < 8; ++pl)
            if (pl < q_planes) planes[pl * plane_stride + row * words + w] = keep[pl];
        if (has_valid) planes[q_planes * plane_stride + row * words + w] = keep[8];
    }
}

// ------------------------------------------------------------------ UMMA one-hot operands
// One thread per 4 sites when sigma4 == 4 (a 16-byte store to A and to B); generic path per site.
__global__ void __launch_bounds__(256)
encode_onehot4_kernel(const uint8_t *__restrict__ seqs, int64_t n, int64_t len, int64_t row_stride,
                      const uint8_t *__restrict__ table_g, int64_t kbytes,
                      uint8_t *__restrict__ A, uint8_t *__restrict__ B, int aligned) {
    __shared__ uint8_t table[256];
    table[threadIdx.x] = table_g[threadIdx.x];
    __syncthreads();
    const int64_t quads = kbytes / 16;                     // 4 sites -> 16 operand bytes
    for (int64_t row = blockIdx.y; row < n; row += gridDim.y) {
    const uint8_t *p = seqs + row * row_stride;
    for (int64_t qd = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; qd < quads;
         qd += (int64_t)gridDim.x * blockDim.x) {
        const int64_t site = qd * 4;
        uint32_t raw = 0x2e2e2e2eu;
        if (aligned && site + 3 < len) {
            raw = *reinterpret_cast<const uint32_t *>(p + site);
        } else {
#pragma unroll
            for (int b = 0; b < 4; ++b)
                if (site + b < len) raw = (raw & ~(0xffu << (8 * b))) | (uint32_t(p[site + b]) << (8 * b));
        }
        uint4 a, c;
        uint32_t *aw = &a.x, *cw = &c.x;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const uint32_t code = table[(raw >> (8 * b)) & 0xff];
            const uint32_t hot = code < 4 ? (1u << (8 * code)) : 0u;
            aw[b] = hot;
            cw[b] = code < 4 ? (0x01010101u ^ hot) : 0u;
        }
        *reinterpret_cast<uint4 *>(A + row * kbytes + qd * 16) = a;
        if (B) *reinterpret_cast<uint4 *>(B + row * kbytes + qd * 16) = c;
    }
    }
}

__global__ void __launch_bounds__(256)
encode_onehot_generic_kernel(const uint8_t *__restrict__ seqs, int64_t n, int64_t len,
                             int64_t row_stride, const uint8_t *__restrict__ table_g, int sigma,
                             int sigma4, int64_t kbytes, uint8_t *__restrict__ A,
                             uint8_t *__restrict__ B) {
    __shared__ uint8_t table[256];
    table[threadIdx.x] = table_g[threadIdx.x];
    __syncthreads();
    const int wps = sigma4 / 4;                            // words per site
    const int64_t total_words = kbytes / 4;
    for (int64_t row = blockIdx.y; row < n; row += gridDim.y) {
    const uint8_t *p = seqs + row * row_stride;
    uint32_t *Aw = reinterpret_cast<uint32_t *>(A + row * kbytes);
    uint32_t *Bw = reinterpret_cast<uint32_t *>(B + row * kbytes);
    for (int64_t w = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; w < total_words;
         w += (int64_t)gridDim.x * blockDim.x) {
        const int64_t site = w / wps;
        const int sub = int(w % wps);
        uint32_t hot = 0, comp = 0;
        if (site < len) {
            const uint32_t code = table[p[site]];
            if (code != 0xff) {
                if (int(code >> 2) == sub) hot = 1u << (8 * (code & 3));
                uint32_t live = 0;                          // planes of this word that exist (< sigma)
#pragma unroll
                for (int b = 0; b < 4; ++b)
                    if (sub * 4 + b < sigma) live |= 1u << (8 * b);
                comp = live ^ hot;
            }
        }
        Aw[w] = hot;
        Bw[w] = comp;
    }
    }
}

// dense 3-byte-per-site operand (+-1 simplex or 0/1 tetrahedron): one thread per 16 sites -> one
// 16-byte load, three 16-byte stores (48 operand bytes)
__global__ void __launch_bounds__(256)
encode_simplex_kernel(const uint8_t *__restrict__ seqs, int64_t n, int64_t len, int64_t row_stride,
                      const uint8_t *__restrict__ table_g, int64_t kbytes, uint8_t *__restrict__ A, int aligned,
                      int tetra01, int32_t *__restrict__ rowsum) {
    __shared__ uint32_t vertex[256];                        // symbol -> its 3 operand bytes (little endian)
    {
        const uint32_t code = table_g[threadIdx.x];
        uint32_t t;
        if (tetra01) t = code == 1 ? 0x000101u : code == 2 ? 0x010001u : code == 3 ? 0x010100u : 0u;
        else t = code == 0 ? 0x010101u : code == 1 ? 0xFFFF01u : code == 2 ? 0xFF01FFu : code == 3 ? 0x01FFFFu : 0u;
        vertex[threadIdx.x] = t;
    }
    __syncthreads();
    const int64_t groups = (kbytes + 47) / 48;              // 16 sites = 48 operand bytes
    const bool vec_in = aligned && (row_stride % 16 == 0) && (reinterpret_cast<uintptr_t>(seqs) % 16 == 0);
    for (int64_t row = blockIdx.y; row < n; row += gridDim.y) {
        const uint8_t *p = seqs + row * row_stride;
        uint32_t *dst = reinterpret_cast<uint32_t *>(A + row * kbytes);
        const int64_t total = kbytes / 4;
        int nonzero = 0;
        for (int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; g < groups;
             g += (int64_t)gridDim.x * blockDim.x) {
            const int64_t site = g * 16;
            uint32_t raw[4] = {0x2e2e2e2eu, 0x2e2e2e2eu, 0x2e2e2e2eu, 0x2e2e2e2eu};   // '.' = ignored
            if (vec_in && site + 15 < len) {
                const uint4 v = *reinterpret_cast<const uint4 *>(p + site);
                raw[0] = v.x; raw[1] = v.y; raw[2] = v.z; raw[3] = v.w;
            } else {
                for (int b = 0; b < 16; ++b)
                    if (site + b < len) raw[b >> 2] = (raw[b >> 2] & ~(0xffu << (8 * (b & 3)))) | (uint32_t(p[site + b]) << (8 * (b & 3)));
            }
            uint32_t out[12];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                uint32_t t[4];
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    const uint32_t byte = (raw[q] >> (8 * b)) & 0xff;
                    t[b] = vertex[byte];
                    if (tetra01) nonzero += t[b] != 0;
                }
                out[3 * q + 0] = t[0] | (t[1] << 24);
                out[3 * q + 1] = (t[1] >> 8) | (t[2] << 16);
                out[3 * q + 2] = (t[2] >> 16) | (t[3] << 8);
            }
            const int64_t wi = g * 12;
            if (wi + 12 <= total) {
                uint4 *d4 = reinterpret_cast<uint4 *>(dst + wi);       // 48-byte groups: 16-byte aligned
                d4[0] = make_uint4(out[0], out[1], out[2], out[3]);
                d4[1] = make_uint4(out[4], out[5], out[6], out[7]);
                d4[2] = make_uint4(out[8], out[9], out[10], out[11]);
            } else {
                for (int k = 0; k < 12; ++k)
                    if (wi + k < total) dst[wi + k] = out[k];
            }
        }
        if (tetra01) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) nonzero += __shfl_xor_sync(0xffffffffu, nonzero, o);
            if ((threadIdx.x & 31) == 0 && nonzero) atomicAdd(&rowsum[row], nonzero);
        }
    }
}

// ---- per-site symbol statistics for the dense codes -----------------------------------------------
// d(i,j) only counts sites where the two rows differ, so every site may use its own bijection of
// the four symbols onto the tetrahedron vertices.  Mapping each site's MOST FREQUENT symbol to the
// all-zero vertex makes the operand as sparse as the data allows (a SNP alignment is mostly major
// alleles): the tensor cores multiply mostly zeros, which is what keeps the kernel under the power
// cap (the data-dependence of tensor power is measured: dense +-1 operands cost 25 % of the clock).
// counts: thread = 4 adjacent sites over a strip of rows, four 16-bit counters packed per site.
constexpr int kMajorStripRows = 2048, kMajorRowStep = 8;
__global__ void __launch_bounds__(256)
site_counts_kernel(const uint8_t *__restrict__ seqs, int64_t n, int64_t len, int64_t row_stride,
                   const uint8_t *__restrict__ table_g, uint32_t *__restrict__ counts /*[len][4]*/) {
    __shared__ uint8_t table[256];
    table[threadIdx.x] = table_g[threadIdx.x];
    __syncthreads();
    const int64_t site = (int64_t(blockIdx.x) * 256 + threadIdx.x) * 4;
    if (site >= len) return;
    const int64_t r0 = int64_t(blockIdx.y) * kMajorStripRows, r1 = r0 + kMajorStripRows < n ? r0 + kMajorStripRows : n;
    const bool vec = (reinterpret_cast<uintptr_t>(seqs) % 4 == 0) && (row_stride % 4 == 0) && site + 3 < len;
    uint64_t c[4] = {0, 0, 0, 0};
    for (int64_t r = r0; r < r1; r += kMajorRowStep) {      // a sample of the rows is enough: any choice is exact, only sparsity is at stake
        const uint8_t *p = seqs + r * row_stride + site;
        uint32_t w;
        if (vec) w = *reinterpret_cast<const uint32_t *>(p);
        else {
            w = 0;
            for (int b = 0; b < 4; ++b) w |= uint32_t(site + b < len ? p[b] : 'A') << (8 * b);
        }
#pragma unroll
        for (int b = 0; b < 4; ++b) c[b] += 1ull << (16 * (table[(w >> (8 * b)) & 0xff] & 3));
    }
#pragma unroll
    for (int b = 0; b < 4; ++b)
        if (site + b < len) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t v = uint32_t(c[b] >> (16 * k)) & 0xffffu;
                if (v) atomicAdd(&counts[(site + b) * 4 + k], v);
            }
        }
}
__global__ void __launch_bounds__(256)
site_major_kernel(const uint32_t *__restrict__ counts, int64_t len, uint8_t *__restrict__ major) {
    const int64_t site = int64_t(blockIdx.x) * 256 + threadIdx.x;
    if (site >= len) return;
    const uint4 c = *reinterpret_cast<const uint4 *>(counts + site * 4);
    uint32_t best = c.x;
    uint8_t m = 0;
    if (c.y > best) { best = c.y; m = 1; }
    if (c.z > best) { best = c.z; m = 2; }
    if (c.w > best) { best = c.w; m = 3; }
    major[site] = m;
}

// the same from the counts of several ranks (sharded upload, dist_comm.cu): the pointers may be peer memory
__global__ void __launch_bounds__(256)
site_major_sum_kernel(const CountPtrs cp, int64_t len, uint8_t *__restrict__ major) {
    const int64_t site = int64_t(blockIdx.x) * 256 + threadIdx.x;
    if (site >= len) return;
    uint4 c = make_uint4(0, 0, 0, 0);
    for (int q = 0; q < cp.n; ++q) {
        const uint4 v = *reinterpret_cast<const uint4 *>(cp.p[q] + site * 4);
        c.x += v.x; c.y += v.y; c.z += v.z; c.w += v.w;
    }
    uint32_t best = c.x;
    uint8_t m = 0;
    if (c.y > best) { best = c.y; m = 1; }
    if (c.z > best) { best = c.z; m = 2; }
    if (c.w > best) { best = c.w; m = 3; }
    major[site] = m;
}

// dense tetrahedron code as packed e2m1 nibbles (1.0 = 0x2): one thread per 16 sites -> 48 nibbles = 24 bytes
__global__ void __launch_bounds__(256)
encode_fp4_kernel(const uint8_t *__restrict__ seqs, int64_t n, int64_t len, int64_t row_stride,
                  const uint8_t *__restrict__ table_g, int64_t kbytes, uint8_t *__restrict__ A, int aligned,
                  int32_t *__restrict__ rowsum, const uint8_t *__restrict__ major /* per site, or nullptr */) {
    // one thread per 32 sites: two 16-byte loads -> 96 nibbles = three 16-byte stores
    __shared__ uint8_t code_of[256];                        // byte -> symbol code 0..3 (anything else: 0)
    code_of[threadIdx.x] = table_g[threadIdx.x] & 3;
    __syncthreads();
    constexpr uint64_t kVertex = 0x0220020200220000ull;     // vertex t -> its 3 nibbles (x | y << 4 | z << 8), 16 bits each
    const int64_t groups = (kbytes + 47) / 48;
    const bool vec_in = aligned && (row_stride % 16 == 0) && (reinterpret_cast<uintptr_t>(seqs) % 16 == 0);
    for (int64_t row = blockIdx.y; row < n; row += gridDim.y) {
        const uint8_t *p = seqs + row * row_stride;
        uint32_t *dst = reinterpret_cast<uint32_t *>(A + row * kbytes);
        const int64_t total = kbytes / 4;
        int nonzero = 0;
        for (int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; g < groups;
             g += (int64_t)gridDim.x * blockDim.x) {
            const int64_t site = g * 32;
            uint32_t raw[8], maj[8];                        // maj: the site's zero-coded symbol (padded to 32 sites by the host)
#pragma unroll
            for (int q = 0; q < 8; ++q) { raw[q] = 0x2e2e2e2eu; maj[q] = 0; }
            if (major && site < len) {
                const uint4 m0 = *reinterpret_cast<const uint4 *>(major + site);
                const uint4 m1 = *reinterpret_cast<const uint4 *>(major + site + 16);
                maj[0] = m0.x; maj[1] = m0.y; maj[2] = m0.z; maj[3] = m0.w;
                maj[4] = m1.x; maj[5] = m1.y; maj[6] = m1.z; maj[7] = m1.w;
            }
            if (vec_in && site + 31 < len) {
                const uint4 v0 = *reinterpret_cast<const uint4 *>(p + site);
                const uint4 v1 = *reinterpret_cast<const uint4 *>(p + site + 16);
                raw[0] = v0.x; raw[1] = v0.y; raw[2] = v0.z; raw[3] = v0.w;
                raw[4] = v1.x; raw[5] = v1.y; raw[6] = v1.z; raw[7] = v1.w;
            } else {
                for (int b = 0; b < 32; ++b)
                    if (site + b < len) raw[b >> 2] = (raw[b >> 2] & ~(0xffu << (8 * (b & 3)))) | (uint32_t(p[site + b]) << (8 * (b & 3)));
            }
            uint32_t out[12];
#pragma unroll
            for (int h = 0; h < 4; ++h) {                   // 8 sites -> 96 bits
                uint64_t lo = 0, hi = 0;                    // bits 0..63, 64..95
#pragma unroll
                for (int b = 0; b < 8; ++b) {
                    const uint32_t byte = (raw[2 * h + (b >> 2)] >> (8 * (b & 3))) & 0xff;
                    const uint32_t mj = (maj[2 * h + (b >> 2)] >> (8 * (b & 3))) & 3;
                    const bool live = site + 8 * h + b < len;
                    const uint32_t vtx = live ? (code_of[byte] ^ mj) : 0u;        // XOR: a bijection that sends the major symbol to 0
                    const uint64_t t = (kVertex >> (16 * vtx)) & 0xfffu;
                    nonzero += vtx != 0;
                    const int sh = 12 * b;
                    if (sh < 64) lo |= t << sh;
                    if (sh + 12 > 64) hi |= sh >= 64 ? t << (sh - 64) : t >> (64 - sh);
                }
                out[3 * h + 0] = uint32_t(lo);
                out[3 * h + 1] = uint32_t(lo >> 32);
                out[3 * h + 2] = uint32_t(hi);
            }
            const int64_t wi = g * 12;
            if (wi + 12 <= total) {
                uint4 *d4 = reinterpret_cast<uint4 *>(dst + wi);       // 48-byte groups: 16-byte aligned
                d4[0] = make_uint4(out[0], out[1], out[2], out[3]);
                d4[1] = make_uint4(out[4], out[5], out[6], out[7]);
                d4[2] = make_uint4(out[8], out[9], out[10], out[11]);
            } else {
                for (int k = 0; k < 12; ++k)
                    if (wi + k < total) dst[wi + k] = out[k];
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) nonzero += __shfl_xor_sync(0xffffffffu, nonzero, o);
        if ((threadIdx.x & 31) == 0 && nonzero) atomicAdd(&rowsum[row], nonzero);
    }
}

// The same operand for the default-mode table (ACGTacgt -> four codes, nothing else present in a dense
// alignment), four sites at a time: code = (byte >> 1) & 3 is a bijection of the four bases (A 0, C 1,
// T 2, G 3; any bijection gives the same distances), XOR with the site's major code, one multiply
// gathers the four 2-bit codes into a byte, and a 256-entry table holds the 48 nibble bits of four
// sites.  ~3 instructions per site instead of ~12: the encoder becomes a streaming kernel.
__global__ void __launch_bounds__(256)
encode_fp4_acgt_kernel(const uint8_t *__restrict__ seqs, int64_t n, int64_t len, int64_t row_stride,
                       int64_t kbytes, uint8_t *__restrict__ A, int32_t *__restrict__ rowsum,
                       const uint8_t *__restrict__ major /* per site, padded to 32, or nullptr */) {
    __shared__ uint64_t quad[256];                          // index = c3 | c2 << 2 | c1 << 4 | c0 << 6 -> 12 nibbles, site 0 lowest
    {
        constexpr uint64_t kVertex = 0x0220020200220000ull;
        const uint32_t i = threadIdx.x;
        uint64_t bits = 0;
        for (int k = 0; k < 4; ++k) {
            const uint32_t c = (i >> (2 * (3 - k))) & 3u;    // code of site k
            bits |= ((kVertex >> (16 * c)) & 0xfffull) << (12 * k);
        }
        quad[i] = bits;
    }
    __syncthreads();
    const int64_t groups = (kbytes + 47) / 48;              // 32 sites -> 48 operand bytes
    const bool vec_in = (row_stride % 16 == 0) && (reinterpret_cast<uintptr_t>(seqs) % 16 == 0);
    for (int64_t row = blockIdx.y; row < n; row += gridDim.y) {
        const uint8_t *p = seqs + row * row_stride;
        uint32_t *dst = reinterpret_cast<uint32_t *>(A + row * kbytes);
        const int64_t total = kbytes / 4;
        int nonzero = 0;
        for (int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; g < groups;
             g += (int64_t)gridDim.x * blockDim.x) {
            const int64_t site = g * 32;
            uint32_t raw[8], maj[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) { raw[q] = 0; maj[q] = 0; }
            if (major && site < len) {
                const uint4 m0 = *reinterpret_cast<const uint4 *>(major + site);
                const uint4 m1 = *reinterpret_cast<const uint4 *>(major + site + 16);
                maj[0] = m0.x; maj[1] = m0.y; maj[2] = m0.z; maj[3] = m0.w;
                maj[4] = m1.x; maj[5] = m1.y; maj[6] = m1.z; maj[7] = m1.w;
                // the majority codes come in table order (A C G T = 0 1 2 3); here G = 3 and T = 2
#pragma unroll
                for (int q = 0; q < 8; ++q) maj[q] ^= (maj[q] >> 1) & 0x01010101u;
            }
            if (vec_in && site + 31 < len) {
                const uint4 v0 = *reinterpret_cast<const uint4 *>(p + site);
                const uint4 v1 = *reinterpret_cast<const uint4 *>(p + site + 16);
                raw[0] = v0.x; raw[1] = v0.y; raw[2] = v0.z; raw[3] = v0.w;
                raw[4] = v1.x; raw[5] = v1.y; raw[6] = v1.z; raw[7] = v1.w;
            } else {
                // the tail of a row: sites beyond len must encode as the zero vertex, i.e. code == major
                for (int b = 0; b < 32; ++b) {
                    const uint32_t byte = site + b < len ? uint32_t(p[site + b]) : (((maj[b >> 2] >> (8 * (b & 3))) & 3u) << 1);
                    raw[b >> 2] |= byte << (8 * (b & 3));
                }
            }
            uint32_t out[12];
#pragma unroll
            for (int h = 0; h < 4; ++h) {                   // 8 sites -> two table entries -> 96 bits
                const uint32_t c0 = ((raw[2 * h] >> 1) & 0x03030303u) ^ maj[2 * h];
                const uint32_t c1 = ((raw[2 * h + 1] >> 1) & 0x03030303u) ^ maj[2 * h + 1];
                nonzero += __popc((c0 | (c0 >> 1)) & 0x01010101u) + __popc((c1 | (c1 >> 1)) & 0x01010101u);
                const uint64_t lo = quad[(c0 * 0x40100401u) >> 24], hi = quad[(c1 * 0x40100401u) >> 24];
                out[3 * h + 0] = uint32_t(lo);
                out[3 * h + 1] = uint32_t(lo >> 32) | (uint32_t(hi) << 16);
                out[3 * h + 2] = uint32_t(hi >> 16);
            }
            const int64_t wi = g * 12;
            if (wi + 12 <= total) {
                uint4 *d4 = reinterpret_cast<uint4 *>(dst + wi);
                d4[0] = make_uint4(out[0], out[1], out[2], out[3]);
                d4[1] = make_uint4(out[4], out[5], out[6], out[7]);
                d4[2] = make_uint4(out[8], out[9], out[10], out[11]);
            } else {
                for (int k = 0; k < 12; ++k)
                    if (wi + k < total) dst[wi + k] = out[k];
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) nonzero += __shfl_xor_sync(0xffffffffu, nonzero, o);
        if ((threadIdx.x & 31) == 0 && nonzero) atomicAdd(&rowsum[row], nonzero);
    }
}

// general alphabets as packed e2m1 nibbles (1.0 = 0x2): plane 0 = validity (site not ignored), planes
// 1..sigma = one-hot of the symbol code; every plane is `plane_bytes` long (len nibbles padded to a
// whole 128-byte K block), planes follow each other along K.  d = dot(V_i, V_j) - sum_s dot(X_s,i, X_s,j):
// the kernel multiplies the X blocks with the negate-A bit of the instruction descriptor.
// One thread per 32 sites -> one 16-byte store per plane.
__global__ void __launch_bounds__(256)
encode_fp4_planes_kernel(const uint8_t *__restrict__ seqs, int64_t n, int64_t len, int64_t row_stride,
                         const uint8_t *__restrict__ table_g, int sigma, int64_t plane_bytes, int64_t kbytes,
                         uint8_t *__restrict__ A, int aligned) {
    __shared__ uint8_t table[256];
    table[threadIdx.x] = table_g[threadIdx.x];
    __syncthreads();
    const int64_t groups = plane_bytes / 16;
    const bool vec_in = aligned && (row_stride % 16 == 0) && (reinterpret_cast<uintptr_t>(seqs) % 16 == 0);
    for (int64_t row = blockIdx.y; row < n; row += gridDim.y) {
        const uint8_t *p = seqs + row * row_stride;
        uint8_t *dst = A + row * kbytes;
        for (int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; g < groups;
             g += (int64_t)gridDim.x * blockDim.x) {
            const int64_t site = g * 32;
            uint32_t raw[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) raw[q] = 0x2e2e2e2eu;                    // '.' = ignored
            if (vec_in && site + 31 < len) {
                const uint4 v0 = *reinterpret_cast<const uint4 *>(p + site);
                const uint4 v1 = *reinterpret_cast<const uint4 *>(p + site + 16);
                raw[0] = v0.x; raw[1] = v0.y; raw[2] = v0.z; raw[3] = v0.w;
                raw[4] = v1.x; raw[5] = v1.y; raw[6] = v1.z; raw[7] = v1.w;
            } else {
                for (int b = 0; b < 32; ++b)
                    if (site + b < len) raw[b >> 2] = (raw[b >> 2] & ~(0xffu << (8 * (b & 3)))) | (uint32_t(p[site + b]) << (8 * (b & 3)));
            }
            uint32_t code[8];                               // 4 symbol codes per word
#pragma unroll
            for (int q = 0; q < 8; ++q)
                code[q] = uint32_t(table[raw[q] & 0xff]) | (uint32_t(table[(raw[q] >> 8) & 0xff]) << 8) |
                          (uint32_t(table[(raw[q] >> 16) & 0xff]) << 16) | (uint32_t(table[raw[q] >> 24]) << 24);
            for (int plane = 0; plane <= sigma; ++plane) {
                uint32_t w[4] = {0, 0, 0, 0};
#pragma unroll
                for (int b = 0; b < 32; ++b) {
                    const uint32_t c = (code[b >> 2] >> (8 * (b & 3))) & 0xff;
                    const bool hot = plane == 0 ? c != 0xff : c == uint32_t(plane - 1);
                    w[b >> 3] |= (hot ? 2u : 0u) << (4 * (b & 7));
                }
                *reinterpret_cast<uint4 *>(dst + plane * plane_bytes + g * 16) = make_uint4(w[0], w[1], w[2], w[3]);
            }
        }
    }
}

// the table btb_dist_alphabet builds in default mode without -k: A a -> 0, C c -> 1, G g -> 2, T t -> 3, rest ignored
static bool is_default_acgt_table(const uint8_t table[256]) {
    for (int b = 0; b < 256; ++b) {
        const int u = (b >= 'a' && b <= 'z') ? b - 32 : b;
        const uint8_t want = u == 'A' ? 0 : u == 'C' ? 1 : u == 'G' ? 2 : u == 'T' ? 3 : 0xff;
        if (table[b] != want) return false;
    }
    return true;
}

// ---- pieces of the C-ABI calls, for the streamed level-2 path (capi.cu) ------------------------
// ORs the byte values that occur in rows [0, n) into the 256-bit map d_present8
int dist_presence_async(const uint8_t *d_seqs, int64_t n, int64_t len, int64_t row_stride, uint32_t *d_present8,
                        cudaStream_t stream) {
    if (n <= 0 || len <= 0) return BTB_OK;
    dim3 grid((unsigned)std::min<int64_t>((len + 255) / 256, 64), (unsigned)std::min<int64_t>(n, 2048));
    presence_kernel<<<grid, 256, 0, stream>>>(d_seqs, n, len, row_stride, d_present8);
    BTB_CUDA(cudaGetLastError());
    return BTB_OK;
}

// dense FP4 operand rows (tetrahedron nibbles) of `rows` sequences; rowsum_rows must be zero
int dist_encode_fp4_dense_rows(const uint8_t *d_seqs, int64_t rows, int64_t len, int64_t row_stride,
                               const uint8_t *d_table, int64_t kb, uint8_t *A_rows, int32_t *rowsum_rows,
                               cudaStream_t stream) {
    if (rows <= 0) return BTB_OK;
    const bool aligned = (reinterpret_cast<uintptr_t>(d_seqs) % 4 == 0) && (row_stride % 4 == 0);
    dim3 grid((unsigned)std::min<int64_t>((kb / 48 + 256) / 256, 1024), (unsigned)std::min<int64_t>(rows, 65535));
    encode_fp4_kernel<<<grid, 256, 0, stream>>>(d_seqs, rows, len, row_stride, d_table, kb, A_rows, aligned ? 1 : 0, rowsum_rows, nullptr);
    BTB_CUDA(cudaGetLastError());
    return BTB_OK;
}

// ---- pieces for the sharded multi-rank path (dist_comm.cu): a rank encodes only its own rows -------
// per-site symbol counts of `rows` sequences into d_counts[len][4] (cleared here)
int dist_site_counts(const uint8_t *d_seqs, int64_t rows, int64_t len, int64_t row_stride, const uint8_t *d_table,
                     uint32_t *d_counts, cudaStream_t stream) {
    BTB_CUDA(cudaMemsetAsync(d_counts, 0, size_t(len) * 16, stream));
    if (rows <= 0 || len <= 0) return BTB_OK;
    dim3 cgrid((unsigned)((len + 1023) / 1024), (unsigned)((rows + kMajorStripRows - 1) / kMajorStripRows));
    site_counts_kernel<<<cgrid, 256, 0, stream>>>(d_seqs, rows, len, row_stride, d_table, d_counts);
    BTB_CUDA(cudaGetLastError());
    return BTB_OK;
}

int dist_site_major_sum(const CountPtrs &cp, int64_t len, uint8_t *d_major, cudaStream_t stream) {
    if (len <= 0) return BTB_OK;
    site_major_sum_kernel<<<(unsigned)((len + 255) / 256), 256, 0, stream>>>(cp, len, d_major);
    BTB_CUDA(cudaGetLastError());
    return BTB_OK;
}

// FP4 operand rows (dense tetrahedron nibbles + row sums, or the general validity / one-hot planes) of
// `rows` sequences; A_rows / rowsum_rows point at the first of these rows inside the full operand;
// rowsum_rows must be zero.  d_major: the per-site zero-coded symbol (padded to 32 sites) or nullptr.
int dist_encode_fp4_rows(const uint8_t *d_seqs, int64_t rows, int64_t len, int64_t row_stride, const uint8_t table[256],
                         const uint8_t *d_table, int sigma, uint8_t *A_rows, int32_t *rowsum_rows, const uint8_t *d_major,
                         cudaStream_t stream) {
    if (rows <= 0) return BTB_OK;
    const int64_t kb = umma_k_bytes(len, sigma);
    const bool aligned = (reinterpret_cast<uintptr_t>(d_seqs) % 4 == 0) && (row_stride % 4 == 0);
    if (umma_fp4_general(sigma, len)) {
        const int64_t pb = umma_fp4_plane_bytes(len);
        dim3 grid((unsigned)std::min<int64_t>((pb / 16 + 255) / 256, 1024), (unsigned)std::min<int64_t>(rows, 65535));
        encode_fp4_planes_kernel<<<grid, 256, 0, stream>>>(d_seqs, rows, len, row_stride, d_table, sigma, pb, kb, A_rows, aligned ? 1 : 0);
    } else if (umma_fp4_dense(sigma, len)) {
        dim3 grid((unsigned)std::min<int64_t>((kb / 48 + 256) / 256, 1024), (unsigned)std::min<int64_t>(rows, 65535));
        if (is_default_acgt_table(table))
            encode_fp4_acgt_kernel<<<grid, 256, 0, stream>>>(d_seqs, rows, len, row_stride, kb, A_rows, rowsum_rows, d_major);
        else
            encode_fp4_kernel<<<grid, 256, 0, stream>>>(d_seqs, rows, len, row_stride, d_table, kb, A_rows, aligned ? 1 : 0, rowsum_rows, d_major);
    } else {
        return fail(BTB_EINVAL, "dist_encode_fp4_rows: not an FP4 operand layout");
    }
    BTB_CUDA(cudaGetLastError());
    return BTB_OK;
}

// Symbol table and sigma from the byte values that occur (`present`, 256 bits; scanned = false: unknown).
// snp-dists load rules: upper-case unless -k; default mode counts ACGT only; -a counts
// everything but '.'.
int dist_alphabet_from_presence(const uint32_t present[8], bool scanned, int allchars, int keepcase, uint8_t table[256], int *sigma) {
    for (int b = 0; b < 256; ++b) table[b] = 0xff;
    auto up = [&](int b) { return (!keepcase && b >= 'a' && b <= 'z') ? b - 32 : b; };
    int s = 0;
    if (!allchars) {
        const char *acgt = "ACGT";
        for (int b = 0; b < 256; ++b)
            for (int c = 0; c < 4; ++c)
                if (up(b) == acgt[c]) table[b] = uint8_t(c);
        s = 4;
    } else {
        // symbols that occur after the load rules (upper-casing), '.' excluded, in byte order
        bool occurs[256] = {false};
        for (int b = 0; b < 256; ++b)
            if (present[b >> 5] >> (b & 31) & 1) occurs[up(b)] = true;
        occurs[int('.')] = false;
        int code_of[256];
        for (int b = 0; b < 256; ++b) code_of[b] = occurs[b] ? s++ : -1;
        if (s > 255) return fail(BTB_EINVAL, "alphabet too large");      // codes 0..254, 0xff = ignored: every byte value but '.' fits
        for (int b = 0; b < 256; ++b) {
            int u = up(b);
            if (u != '.' && code_of[u] >= 0) table[b] = uint8_t(code_of[u]);
        }
        if (s == 0) s = 1;
    }
    // dense: nothing that occurs is ignored, and at most four symbols
    bool dense = scanned && s <= 4;
    for (int b = 0; b < 256 && dense; ++b)
        if ((present[b >> 5] >> (b & 31) & 1) && table[b] == 0xff) dense = false;
    *sigma = dense ? BTB_SIGMA_DENSE4 : s;
    return BTB_OK;
}

}  // namespace btb

using namespace btb;

// =============================================================================== C-ABI
extern "C" int btb_dist_alphabet(const uint8_t *d_seqs, int64_t n, int64_t len, int64_t row_stride,
                                 int allchars, int keepcase, uint8_t table[256], int *sigma,
                                 void *stream_) {
    if (!table || !sigma) return fail(BTB_EINVAL, "btb_dist_alphabet: NULL output");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    uint32_t *d_present = nullptr;
    uint32_t present[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const bool scan = d_seqs && n * len > 0;
    if (allchars && !d_seqs && n * len > 0) return fail(BTB_EINVAL, "btb_dist_alphabet: allchars needs the alignment");
    if (scan) {
        BTB_CUDA(cudaMallocAsync(reinterpret_cast<void **>(&d_present), sizeof(present), stream));
        BTB_CUDA(cudaMemsetAsync(d_present, 0, sizeof(present), stream));
        dim3 grid((unsigned)std::min<int64_t>((len + 255) / 256, 64), (unsigned)std::min<int64_t>(n, 2048));
        presence_kernel<<<grid, 256, 0, stream>>>(d_seqs, n, len, row_stride, d_present);
        BTB_CUDA(cudaGetLastError());
        BTB_CUDA(cudaMemcpyAsync(present, d_present, sizeof(present), cudaMemcpyDeviceToHost, stream));
        BTB_CUDA(cudaStreamSynchronize(stream));
        BTB_CUDA(cudaFreeAsync(d_present, stream));
    }
    return dist_alphabet_from_presence(present, scan, allchars, keepcase, table, sigma);
}

extern "C" int64_t btb_dist_operand_bytes(int engine, int64_t n, int64_t len, int sigma) {
    if (engine == BTB_ENGINE_POPC) return (popc_code_planes(sigma) + (sigma_dense(sigma) ? 0 : 1)) * n * popc_words(len) * 4;
    if (engine == BTB_ENGINE_UMMA) return (sigma_dense(sigma) || umma_use_fp4(sigma, len) ? 1 : 2) * n * umma_k_bytes(len, sigma) + n * 4;
    return -1;
}

// Operand rows [row0, row1) only, at their place in the operand of all n rows: a rank of a multi-GPU job whose
// block rows start at row r never reads the rows before r - 256 (btb_distance_rows), so it need not encode them.
// FP4 layouts; anything else encodes everything.
extern "C" int btb_dist_encode_rows(int engine, const uint8_t *d_seqs, int64_t n, int64_t len, int64_t row_stride,
                                    const uint8_t table[256], int sigma, void *d_operands, int64_t row0, int64_t row1,
                                    void *stream_) {
    if (n <= 0 || len < 0 || !d_operands || !table || !d_seqs || row0 < 0 || row1 > n || row0 > row1)
        return fail(BTB_EINVAL, "btb_dist_encode_rows: bad argument");
    if (engine != BTB_ENGINE_UMMA || !umma_use_fp4(sigma, len) || (row0 == 0 && row1 == n))
        return btb_dist_encode(engine, d_seqs, n, len, row_stride, table, sigma, d_operands, stream_);
    if (row0 == row1) return BTB_OK;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const int64_t kb = umma_k_bytes(len, sigma), rows = row1 - row0;
    uint8_t *A = static_cast<uint8_t *>(d_operands);
    const uint8_t *seqs = d_seqs + row0 * row_stride;
    const bool dense = sigma_dense(sigma), use_major = dense && g_umma_major_zero && n >= 64;
    const int64_t padded = (len + 31) / 32 * 32;
    uint8_t *d_small = nullptr;                          // table 256 | major (padded) | counts (len x 16)
    const size_t counts_off = size_t(256 + padded + 255) & ~size_t(255);
    BTB_CUDA(cudaMallocAsync(reinterpret_cast<void **>(&d_small), counts_off + (use_major ? size_t(len) * 16 : 0) + 16, stream));
    BTB_CUDA(cudaMemcpyAsync(d_small, table, 256, cudaMemcpyHostToDevice, stream));
    uint8_t *d_major = nullptr;
    if (use_major) {                                     // any choice of the zero-coded symbol per site is exact: these rows decide
        d_major = d_small + 256;
        uint32_t *d_counts = reinterpret_cast<uint32_t *>(d_small + counts_off);
        BTB_CUDA(cudaMemsetAsync(d_major, 0, size_t(padded), stream));
        BTB_TRY(dist_site_counts(seqs, rows, len, row_stride, d_small, d_counts, stream));
        site_major_kernel<<<(unsigned)((len + 255) / 256), 256, 0, stream>>>(d_counts, len, d_major);
    }
    int32_t *rowsum = reinterpret_cast<int32_t *>(A + n * kb);
    if (dense) BTB_CUDA(cudaMemsetAsync(rowsum + row0, 0, size_t(rows) * 4, stream));
    BTB_TRY(dist_encode_fp4_rows(seqs, rows, len, row_stride, table, d_small, sigma, A + row0 * kb, dense ? rowsum + row0 : nullptr, d_major, stream));
    BTB_CUDA(cudaFreeAsync(d_small, stream));
    return BTB_OK;
}

extern "C" int btb_dist_encode(int engine, const uint8_t *d_seqs, int64_t n, int64_t len,
                               int64_t row_stride, const uint8_t table[256], int sigma,
                               void *d_operands, void *stream_) {
    if (n <= 0 || len < 0 || !d_operands || !table) return fail(BTB_EINVAL, "btb_dist_encode: bad argument");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    uint8_t *d_table = nullptr;
    BTB_CUDA(cudaMallocAsync(reinterpret_cast<void **>(&d_table), 256, stream));
    BTB_CUDA(cudaMemcpyAsync(d_table, table, 256, cudaMemcpyHostToDevice, stream));
    const bool aligned = (reinterpret_cast<uintptr_t>(d_seqs) % 4 == 0) && (row_stride % 4 == 0);
    if (engine == BTB_ENGINE_POPC) {
        const int q = popc_code_planes(sigma);
        if (q > 8) return fail(BTB_EINVAL, "alphabet too large for the POPC engine");
        const int64_t words = popc_words(len);
        const int64_t blocks_per_row = words / 32 + (words % 32 != 0);
        const int64_t warps = n * blocks_per_row;
        const unsigned grid = unsigned((warps + 7) / 8);
        if (aligned)
            encode_planes_kernel<true><<<grid, 256, 0, stream>>>(d_seqs, n, len, row_stride, d_table, q, sigma_dense(sigma) ? 0 : 1,
                                                                 words, static_cast<uint32_t *>(d_operands));
        else
            encode_planes_kernel<false><<<grid, 256, 0, stream>>>(d_seqs, n, len, row_stride, d_table, q, sigma_dense(sigma) ? 0 : 1,
                                                                  words, static_cast<uint32_t *>(d_operands));
    } else if (engine == BTB_ENGINE_UMMA) {
        const int64_t kb = umma_k_bytes(len, sigma);
        uint8_t *A = static_cast<uint8_t *>(d_operands);
        uint8_t *B = A + n * kb;
        if (umma_fp4_general(sigma, len)) {
            const int64_t pb = umma_fp4_plane_bytes(len);
            dim3 grid((unsigned)std::min<int64_t>((pb / 16 + 255) / 256, 1024), (unsigned)std::min<int64_t>(n, 65535));
            encode_fp4_planes_kernel<<<grid, 256, 0, stream>>>(d_seqs, n, len, row_stride, d_table, sigma, pb, kb, A, aligned ? 1 : 0);
        } else if (umma_fp4_dense(sigma, len)) {
            dim3 grid((unsigned)std::min<int64_t>((kb / 48 + 256) / 256, 1024), (unsigned)std::min<int64_t>(n, 65535));
            int32_t *rowsum = reinterpret_cast<int32_t *>(A + n * kb);
            BTB_CUDA(cudaMemsetAsync(rowsum, 0, size_t(n) * 4, stream));
            uint8_t *d_major = nullptr;
            uint32_t *d_counts = nullptr;
            if (g_umma_major_zero && n >= 64) {            // per-site majority symbol -> zero vertex (sparser operand)
                const int64_t padded = (len + 31) / 32 * 32;
                BTB_CUDA(cudaMallocAsync(reinterpret_cast<void **>(&d_counts), size_t(len) * 16, stream));
                BTB_CUDA(cudaMallocAsync(reinterpret_cast<void **>(&d_major), size_t(padded), stream));
                BTB_CUDA(cudaMemsetAsync(d_counts, 0, size_t(len) * 16, stream));
                BTB_CUDA(cudaMemsetAsync(d_major, 0, size_t(padded), stream));
                dim3 cgrid((unsigned)((len + 1023) / 1024), (unsigned)((n + kMajorStripRows - 1) / kMajorStripRows));
                site_counts_kernel<<<cgrid, 256, 0, stream>>>(d_seqs, n, len, row_stride, d_table, d_counts);
                site_major_kernel<<<(unsigned)((len + 255) / 256), 256, 0, stream>>>(d_counts, len, d_major);
            }
            if (is_default_acgt_table(table))
                encode_fp4_acgt_kernel<<<grid, 256, 0, stream>>>(d_seqs, n, len, row_stride, kb, A, rowsum, d_major);
            else
                encode_fp4_kernel<<<grid, 256, 0, stream>>>(d_seqs, n, len, row_stride, d_table, kb, A, aligned ? 1 : 0, rowsum, d_major);
            if (d_counts) BTB_CUDA(cudaFreeAsync(d_counts, stream));
            if (d_major) BTB_CUDA(cudaFreeAsync(d_major, stream));
        } else if (sigma_dense(sigma) && g_umma_dense_simplex) {
            dim3 grid((unsigned)std::min<int64_t>((kb / 48 + 256) / 256, 1024), (unsigned)std::min<int64_t>(n, 65535));
            int32_t *rowsum = reinterpret_cast<int32_t *>(A + n * kb);
            if (g_umma_dense_simplex == 2) BTB_CUDA(cudaMemsetAsync(rowsum, 0, size_t(n) * 4, stream));
            encode_simplex_kernel<<<grid, 256, 0, stream>>>(d_seqs, n, len, row_stride, d_table, kb, A, aligned ? 1 : 0,
                                                            g_umma_dense_simplex == 2 ? 1 : 0, rowsum);
        } else if (sigma_dense(sigma) || sigma4_of(sigma) == 4) {
            dim3 grid((unsigned)std::min<int64_t>((kb / 16 + 255) / 256, 1024), (unsigned)std::min<int64_t>(n, 65535));
            encode_onehot4_kernel<<<grid, 256, 0, stream>>>(d_seqs, n, len, row_stride, d_table, kb, A,
                                                            sigma_dense(sigma) ? nullptr : B, aligned ? 1 : 0);
        } else {
            dim3 grid((unsigned)std::min<int64_t>((kb / 4 + 255) / 256, 1024), (unsigned)std::min<int64_t>(n, 65535));
            encode_onehot_generic_kernel<<<grid, 256, 0, stream>>>(d_seqs, n, len, row_stride, d_table,
                                                                   sigma, sigma4_of(sigma), kb, A, B);
        }
    } else {
        return fail(BTB_EINVAL, "btb_dist_encode: unknown engine %d", engine);
    }
    BTB_CUDA(cudaGetLastError());
    BTB_CUDA(cudaFreeAsync(d_table, stream));
    return BTB_OK;
}
