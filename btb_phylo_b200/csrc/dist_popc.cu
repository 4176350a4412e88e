// k4, POPC engine: warp-tiled XOR/AND/POPC distance kernel over bit-planes.
//
// Replaces the O(N^2 L) byte-compare loop of snp-dists 0.8.2, reached in the
// reference through btbphylo/phylogeny.py:149:
//     d(i,j) = popc( OR_q (code_q,i XOR code_q,j)  AND valid_i AND valid_j )   summed over words.
// Exact integer arithmetic.  Bound: the SM integer pipes (one POPC + Q+1 LOP3 + one IADD per 32
// sites per pair), not HBM -- every operand word is reused 128 times from shared memory.
//
// A CTA of 256 threads computes one 128 x 128 quadrant of a BTB_TILE scheduler tile; a thread owns
// an 8 x 8 register tile (rows ty*8+r, columns tx+16*c).  Operand words are staged with 4-byte
// cp.async into a double-buffered [plane][128 rows][KC+1 words] layout (odd stride: the 16 distinct
// column rows of a warp hit 16 different banks, the 2 distinct row rows broadcast).
#include "common.cuh"

namespace btb {

constexpr int kPT = 128;        // CTA tile side

template <typename OutT>
__device__ __forceinline__ void store_tile_from_smem(const uint32_t *s_acc /*[128][129]*/, int64_t row0,
                                                     int64_t col0, int64_t n, OutT *out, int64_t ld,
                                                     int out_mode, int64_t tile_slot, int qr, int qc,
                                                     bool mirror) {
    // direct: 128 rows x 128 columns, coalesced along columns
    for (int e = threadIdx.x; e < kPT * kPT; e += blockDim.x) {
        const int r = e >> 7, c = e & 127;
        const uint32_t v = s_acc[r * 129 + c];
        if (out_mode == 1) {
            const bool live = row0 + r < n && col0 + c < n;
            out[tile_slot * (BTB_TILE * BTB_TILE) + (int64_t)(qr * kPT + r) * BTB_TILE + qc * kPT + c] =
                live ? OutT(v) : OutT(0);
        } else if (row0 + r < n && col0 + c < n) {
            out[(row0 + r) * ld + col0 + c] = OutT(v);
        }
    }
    if (mirror && out_mode == 0) {
        for (int e = threadIdx.x; e < kPT * kPT; e += blockDim.x) {
            const int c = e >> 7, r = e & 127;             // consecutive threads -> consecutive rows
            if (row0 + r < n && col0 + c < n) out[(col0 + c) * ld + row0 + r] = OutT(s_acc[r * 129 + c]);
        }
    }
}

template <int Q, int kKC, bool HASV, typename OutT>
__global__ void __launch_bounds__(256)
dist_popc_kernel(const uint32_t *__restrict__ planes, int64_t n, int64_t words, int64_t tile_lo,
                 OutT *__restrict__ out, int64_t ld, int out_mode, int mirror, const int2 *__restrict__ sq_list) {
    constexpr int P = Q + (HASV ? 1 : 0);                   // code planes (+ validity unless dense)
    constexpr int kStride = kKC + 1;                        // odd: conflict-free column-row reads
    extern __shared__ uint32_t smem[];
    constexpr int kOperand = P * kPT * kStride;             // words of one operand in one stage
    auto sA = [&](int buf) { return smem + (2 * buf) * kOperand; };
    auto sB = [&](int buf) { return smem + (2 * buf + 1) * kOperand; };

    const int64_t T = tiles_per_dim(n);
    const int64_t tile = tile_lo + blockIdx.x / 4;
    const int quad = blockIdx.x & 3, qr = quad >> 1, qc = quad & 1;
    int I, J;
    if (sq_list) { const int2 w = sq_list[tile]; I = w.x; J = w.y; }
    else tile_coords(T, tile, I, J);
    const int64_t row0 = (int64_t)I * BTB_TILE + qr * kPT;
    const int64_t col0 = (int64_t)J * BTB_TILE + qc * kPT;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int64_t plane_stride = n * words;

    uint32_t acc[8][8];
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[r][c] = 0;

    auto stage = [&](int buf, int64_t w0) {
        // P planes x 128 rows x 16 words for A and for B: 4-byte async copies, rows beyond n -> zero
        for (int e = threadIdx.x; e < P * kPT * kKC; e += 256) {
            const int k = e & (kKC - 1), r = (e / kKC) & (kPT - 1), p = e / (kKC * kPT);
            const int64_t ra = row0 + r, rb = col0 + r;
            uint32_t *da = &sA(buf)[(p * kPT + r) * kStride + k];
            uint32_t *db = &sB(buf)[(p * kPT + r) * kStride + k];
            if (ra < n) {
                const uint32_t *src = planes + p * plane_stride + ra * words + w0 + k;
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(da)), "l"(src));
            } else *da = 0;
            if (rb < n) {
                const uint32_t *src = planes + p * plane_stride + rb * words + w0 + k;
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(db)), "l"(src));
            } else *db = 0;
        }
        asm volatile("cp.async.commit_group;");
    };

    const int64_t chunks = words / kKC;
    if (row0 < n && col0 < n && chunks > 0) {
        stage(0, 0);
        for (int64_t ch = 0; ch < chunks; ++ch) {
            const int buf = int(ch & 1);
            if (ch + 1 < chunks) {
                stage(buf ^ 1, (ch + 1) * kKC);
                asm volatile("cp.async.wait_group 1;");
            } else {
                asm volatile("cp.async.wait_group 0;");
            }
            __syncthreads();
            const uint32_t *a = sA(buf) + (ty * 8) * kStride;
            const uint32_t *b = sB(buf) + tx * kStride;
#pragma unroll 2
            for (int k = 0; k < kKC; ++k) {
                uint32_t av[P][8], bv[P][8];
#pragma unroll
                for (int p = 0; p < P; ++p)
#pragma unroll
                    for (int r = 0; r < 8; ++r) {
                        av[p][r] = a[(p * kPT + r) * kStride + k];
                        bv[p][r] = b[(p * kPT + 16 * r) * kStride + k];
                    }
#pragma unroll
                for (int r = 0; r < 8; ++r)
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        uint32_t d = av[0][r] ^ bv[0][c];
#pragma unroll
                        for (int q = 1; q < Q; ++q) d |= av[q][r] ^ bv[q][c];
                        if constexpr (HASV) d &= av[Q][r] & bv[Q][c];
                        acc[r][c] += __popc(d);
                    }
            }
            __syncthreads();
        }
    }
    // epilogue through shared memory so that both the tile and its transpose are written coalesced
    uint32_t *s_acc = smem;
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int c = 0; c < 8; ++c) s_acc[(ty * 8 + r) * 129 + tx + 16 * c] = acc[r][c];
    __syncthreads();
    store_tile_from_smem<OutT>(s_acc, row0, col0, n, out, ld, out_mode, tile - tile_lo, qr, qc,
                               mirror && I != J);
}

template <int Q, bool HASV, typename OutT>
static int launch_popc(const uint32_t *planes, int64_t n, int64_t words, int64_t tile_lo, int64_t tiles,
                       void *out, int64_t ld, int out_mode, int mirror, cudaStream_t stream, const int2 *sq_list) {
    constexpr int P = Q + (HASV ? 1 : 0);
    constexpr int kKC = P <= 5 ? 16 : 8;                    // staged words per chunk: keep 4 buffers < 227 KB
    size_t smem = size_t(4) * P * kPT * (kKC + 1) * 4;
    const size_t epi = size_t(kPT) * 129 * 4;
    if (smem < epi) smem = epi;
    BTB_CUDA(cudaFuncSetAttribute(dist_popc_kernel<Q, kKC, HASV, OutT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dist_popc_kernel<Q, kKC, HASV, OutT><<<unsigned(tiles * 4), 256, smem, stream>>>(
        planes, n, words, tile_lo, static_cast<OutT *>(out), ld, out_mode, mirror, sq_list);
    BTB_CUDA(cudaGetLastError());
    return BTB_OK;
}

int distance_tiles_popc(const void *d_operands, int64_t n, int64_t len, int sigma, int64_t tile_lo,
                        int64_t tile_hi, void *d_out, int out_elem_bytes, int64_t out_ld, int out_mode,
                        int mirror, cudaStream_t stream, const int2 *sq_list) {
    const int q = popc_code_planes(sigma);
    const int64_t words = popc_words(len);
    const uint32_t *planes = static_cast<const uint32_t *>(d_operands);
    const int64_t tiles = tile_hi - tile_lo;
    if (tiles <= 0) return BTB_OK;
#define BTB_POPC_CASE(QQ)                                                                             \
    case QQ:                                                                                          \
        return out_elem_bytes == 2                                                                    \
                   ? launch_popc<QQ, true, uint16_t>(planes, n, words, tile_lo, tiles, d_out, out_ld, out_mode, mirror, stream, sq_list) \
                   : launch_popc<QQ, true, uint32_t>(planes, n, words, tile_lo, tiles, d_out, out_ld, out_mode, mirror, stream, sq_list);
    if (sigma_dense(sigma))
        return out_elem_bytes == 2
                   ? launch_popc<2, false, uint16_t>(planes, n, words, tile_lo, tiles, d_out, out_ld, out_mode, mirror, stream, sq_list)
                   : launch_popc<2, false, uint32_t>(planes, n, words, tile_lo, tiles, d_out, out_ld, out_mode, mirror, stream, sq_list);
    switch (q) {
        BTB_POPC_CASE(1)
        BTB_POPC_CASE(2)
        BTB_POPC_CASE(3)
        BTB_POPC_CASE(4)
        BTB_POPC_CASE(5)
        BTB_POPC_CASE(6)
        BTB_POPC_CASE(7)
        BTB_POPC_CASE(8)
    }
#undef BTB_POPC_CASE
    return fail(BTB_EINVAL, "POPC engine: unsupported plane count %d", q);
}

int get_rows_list(int64_t n, int64_t lo, int64_t hi, int kind, const int2 **d_out, int64_t *count);

int distance_rows_popc(const void *d_operands, int64_t n, int64_t len, int sigma, int64_t row_lo, int64_t row_hi,
                       void *d_out, int out_elem_bytes, int64_t out_ld, int mirror, cudaStream_t stream) {
    if (row_hi <= row_lo) return BTB_OK;
    const int2 *d_list = nullptr;
    int64_t count = 0;
    BTB_TRY(get_rows_list(n, row_lo, row_hi, 0, &d_list, &count));
    return distance_tiles_popc(d_operands, n, len, sigma, 0, count, d_out, out_elem_bytes, out_ld, 0, mirror, stream, d_list);
}

}  // namespace btb
