// k1-k3: variable-column extraction (the snp-sites stage) on the GPU.
//
// Replaces `snp-sites <multi_fasta> -c -o <snps.fas>` (btbphylo/phylogeny.py:133; behaviour:
// SURVEY.md App. A.2): a column is kept iff every sample's byte is one of ACGTacgt and at least
// two distinct bases occur; kept columns are emitted upper-cased in genome order.
//
//   k1 pack_planes   : FASTA text -> 3 bit-planes per sample (code bit 0, code bit 1, validity),
//                      32 columns per word (bit order: see plane_bit).  code = (byte >> 1) & 3
//                      (A=0 C=1 T=2 G=3).  4 bytes are classified at a time with SWAR logic.
//                      3 bits per base instead of 8 is what lets 50,000 x 4.35 Mb stay in 180 GB.
//   k2 column_mask   : OR/AND reduction over samples -> seenA, seenC, seenG, seenT, anyInvalid.
//   k3 select/compact: keep-mask, popc + shuffle-scan stream compaction of the kept column
//                      indices, then a gather of those columns for every sample.
// All three are HBM-bound streaming kernels: bytes = N*G (text) + 3/8 N*G (planes written, read
// once more by k2) + 2 N*L.
#include <algorithm>
#include <vector>

#include "common.cuh"

namespace btb {

constexpr int kPackThreads = 256;
constexpr int kMinLineW = 32;                            // shorter lines go through the host walk (<= 1 line end per 32 columns)
template <int WPT> struct PackCfg {                      // WPT = plane words (32 columns each) per thread
    static constexpr int kCols = kPackThreads * 32 * WPT;                     // columns per CTA
    static constexpr int kStageBytes = kCols + kCols / kMinLineW * 2 + 64;    // staged text incl. terminators
};
int g_pack_wpt = 4;                                     // plane words per thread: 4 (128 columns, 16-byte plane stores) or 1

__host__ __device__ inline int64_t plane_words(int64_t G) { return ((G + 31) / 32 + 3) / 4 * 4; }

// smem skew: 4 bytes of padding per 32 so that threads 32..36 bytes apart hit distinct banks.  Needed
// when a thread owns 32 columns (neighbours 8.25 words apart); with 128 columns per thread the
// neighbours are 33-34 words apart, which is conflict-free as it is.
template <bool SKEW>
__device__ __forceinline__ uint32_t skew(uint32_t p) { return SKEW ? p + ((p >> 5) << 2) : p; }

__device__ __forceinline__ uint4 ld_nc_v4(const uint8_t *p) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

// ------------------------------------------------------------------------------------------ k1
// Bit order inside a plane word: column c = 32 w + cl sits at bit 8 (cl & 3) + (cl >> 2), i.e.
// byte j of the word holds the columns 4 g + j, bit g.  That is the order in which a 4-byte SWAR
// classification produces its flags (one flag per byte lane), so k1 needs no bit transposition;
// k2 is bit-parallel and order-agnostic, k3 undoes the permutation once on the G-bit keep mask.
__host__ __device__ inline int plane_bit(int cl) { return 8 * (cl & 3) + (cl >> 2); }

// SWAR classification of 4 text bytes.  Flags come out at bit 0 of every byte lane:
//   valid <=> byte is one of ACGTacgt   (exact; checked exhaustively in tests/test_capi_host.py)
//   code  = (byte >> 1) & 3             (A 0, C 1, T 2, G 3)
// hi3 collects bit5|bit6|bit7 of every byte at bit 5: a byte < 0x20 (control character, e.g. a
// stray CR/LF inside a line) clears it.
struct Class4 { uint32_t v, b0, b1, hi3; };
// The bits of every byte lane are lined up at bit 7 with LEFT shifts (the compiler issues them as
// IMAD.SHL on the FMA pipe, leaving the ALU pipe to the LOP3s); the three flags are then moved to
// bit 0 of their lane.  hi3 keeps bit5|bit6|bit7 of every byte at bit 7.
__device__ __forceinline__ Class4 classify4(uint32_t w) {
    const uint32_t y0 = w << 7, y1 = w << 6, y2 = w << 5, y3 = w << 4, y4 = w << 3, y5 = w << 2, y6 = w << 1;
    const uint32_t t1 = y0 & (y1 | ~y2);
    const uint32_t t2 = y2 & ~y1 & ~y0;
    const uint32_t f = (~y4 & t1) | (y4 & t2);
    const uint32_t g = ~w & y6 & ~y3;
    const uint32_t v7 = f & g & 0x80808080u;
    Class4 c;
    c.v = v7 >> 7;
    c.b0 = (v7 & y1) >> 7;
    c.b1 = (v7 & y2) >> 7;
    c.hi3 = w | y6 | y5;
    return c;
}

template <int kPackWordsPerThread>
__global__ void __launch_bounds__(kPackThreads)
pack_planes_kernel(const uint8_t *__restrict__ text, int64_t text_bytes, const int64_t *__restrict__ rec_off,
                   const int32_t *__restrict__ line_w, const int32_t *__restrict__ eol_len, int64_t first,
                   int64_t n_total, int64_t G, int64_t words, uint32_t *__restrict__ planes,
                   int32_t *__restrict__ flags) {
    constexpr bool kSkew = kPackWordsPerThread == 1;
    constexpr int kPackCols = PackCfg<kPackWordsPerThread>::kCols;
    constexpr int kPackStageBytes = PackCfg<kPackWordsPerThread>::kStageBytes;
    __shared__ __align__(16) uint8_t stage[kPackStageBytes + kPackStageBytes / 8 + 64];
    const int64_t i = blockIdx.y;
    const int64_t c0 = int64_t(blockIdx.x) * kPackCols;
    const uint32_t W = uint32_t(line_w[i]), e = uint32_t(eol_len[i]);
    const int64_t rec = rec_off[i];
    const int64_t wbase = (c0 >> 5) + kPackWordsPerThread * threadIdx.x;     // first plane word of this thread
    uint32_t *p0 = planes + (0 * n_total + first + i) * words + wbase;
    uint32_t *p1 = planes + (1 * n_total + first + i) * words + wbase;
    uint32_t *pv = planes + (2 * n_total + first + i) * words + wbase;
    const bool owns = wbase < words;                      // words is a multiple of 4: all of this thread's words or none
    if (c0 >= G || (W && W < kMinLineW)) {                // padding words / layout the host must re-stage
        if (owns) {
#pragma unroll
            for (int k = 0; k < kPackWordsPerThread; ++k) p0[k] = p1[k] = pv[k] = 0;
        }
        if (c0 < G && threadIdx.x == 0) atomicOr(&flags[i], 1);
        return;
    }
    // genome lengths fit 32 bits (the tools use int): 32-bit divisions only
    const uint32_t G32 = uint32_t(G), cb = uint32_t(c0);
    const uint32_t ce = G32 - cb < uint32_t(kPackCols) ? G32 : cb + uint32_t(kPackCols);      // end of this CTA's columns
    auto off = [&](uint32_t c) -> int64_t { return W ? int64_t(c) + int64_t(c / W) * e : int64_t(c); };
    const int64_t lo = rec + off(cb);
    int64_t hi = rec + off(ce - 1) + 1 + e;               // incl. the terminator after the last column
    if (hi > text_bytes) hi = text_bytes;
    const int64_t alo = lo & ~int64_t(15);
    const int nchunk = int((hi - alo + 15) >> 4) + 1;      // +1: the funnel shift may touch one word beyond
    for (int j = threadIdx.x; j < nchunk; j += kPackThreads) {
        uint4 v = make_uint4(0x20202020u, 0x20202020u, 0x20202020u, 0x20202020u);
        if (alo + 16 * int64_t(j) < text_bytes) v = ld_nc_v4(text + alo + 16 * int64_t(j));
        uint32_t *d = reinterpret_cast<uint32_t *>(stage + skew<kSkew>(16u * j));
        d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
    }
    __syncthreads();
    const uint32_t *sw = reinterpret_cast<const uint32_t *>(stage);
    auto word_at = [&](uint32_t q) { return sw[kSkew ? q + (q >> 3) : q]; };     // q = unskewed word index
    const uint32_t c = cb + 32u * kPackWordsPerThread * threadIdx.x;
    uint32_t out_v[kPackWordsPerThread], out_0[kPackWordsPerThread], out_1[kPackWordsPerThread];
    uint32_t nonctl = 0x80808080u;
    bool irregular = false, marker = false;
    const uint32_t line = (W && c < ce) ? c / W : 0u;
    uint32_t pos = W ? c - line * W : 0u;                  // column inside the current line
    uint32_t o = uint32_t(rec + int64_t(c) + int64_t(line) * e - alo);      // byte offset inside the staged span
    // a line of the sequence region that starts with '>', '@' or '+' is not sequence at all for the
    // parser the tools share (it opens the next record / a quality block): the host must re-index
    auto is_marker = [](uint8_t b) { return b == '>' || b == '@' || b == '+'; };
    if (c < ce && pos == 0 && (W || c == 0)) marker = is_marker(stage[skew<kSkew>(o)]);
#pragma unroll
    for (int wi = 0; wi < kPackWordsPerThread; ++wi) {
        uint32_t fv = 0, f0 = 0, f1 = 0;
        const uint32_t cw = c + 32u * wi;
        if (cw < ce) {
            const uint32_t ncols = ce - cw < 32u ? ce - cw : 32u;
            // Branch-free for every lane: the 32 columns are the bytes o .. o+31, except that after the
            // `rem` bases left on the current line the text skips e terminator bytes (at most one line
            // end per word since W >= 32).  A[g] reads group g before the line end, B[g] after it.
            const uint32_t rem = W ? W - pos : 0xFFFFu;
            const uint32_t qa = o >> 2, sa = (o & 3u) * 8u, ob = o + e, qb = ob >> 2, sb = (ob & 3u) * 8u;
            // groups are visited last to first so that the plane words build up as F = 2 F + flags
            // (an IMAD on the FMA pipe): group g ends at bit g of its byte lane
            const uint32_t bg = rem >> 2, bm = (1u << (8u * (rem & 3u))) - 1u;       // break group / its byte mask
            const bool partial = ncols < 32u;              // only the last word of a row
            uint32_t tail[8];
            if (partial) {
#pragma unroll
                for (int g = 0; g < 8; ++g) {
                    const int have = int(ncols) - 4 * g;
                    tail[g] = have >= 4 ? 0xFFFFFFFFu : (have <= 0 ? 0u : (1u << (8 * have)) - 1u);
                }
            }
            uint32_t na = word_at(qa + 8), nb = word_at(qb + 8);
#pragma unroll
            for (int g = 7; g >= 0; --g) {
                const uint32_t ca = word_at(qa + g), cbw = word_at(qb + g);
                const uint32_t A = __funnelshift_r(ca, na, sa), B = __funnelshift_r(cbw, nb, sb);
                na = ca; nb = cbw;
                // bytes of this group that still belong to the current line come from A, the rest from B
                const uint32_t m = uint32_t(g) < bg ? 0xFFFFFFFFu : (uint32_t(g) == bg ? bm : 0u);
                uint32_t w = (A & m) | (B & ~m);
                if (partial) w = (w & tail[g]) | (0x20202020u & ~tail[g]);     // last word of the row: blank columns >= G
                const Class4 k = classify4(w);
                fv = fv * 2u + k.v; f0 = f0 * 2u + k.b0; f1 = f1 * 2u + k.b1;
                nonctl &= k.hi3;
            }
            if (rem <= ncols) {                             // a line ends inside (or exactly at the end of) this word
                if (cw + rem < G32) {                       // ... and more bases follow: the terminator must be there
                    const uint32_t t = o + rem;
                    if (e == 1) irregular |= stage[skew<kSkew>(t)] != '\n';
                    else irregular |= (stage[skew<kSkew>(t)] != '\r') | (stage[skew<kSkew>(t + 1)] != '\n');
                    marker |= is_marker(stage[skew<kSkew>(t + e)]);      // first byte of the next line
                }
                o += 32u + e;
                pos = 32u - rem;
            } else {
                o += 32u;
                pos += 32u;
            }
        }
        out_v[wi] = fv; out_0[wi] = f0; out_1[wi] = f1;
    }
    if (owns) {
        if constexpr (kPackWordsPerThread == 4) {
            *reinterpret_cast<uint4 *>(p0) = make_uint4(out_0[0], out_0[1], out_0[2], out_0[3]);
            *reinterpret_cast<uint4 *>(p1) = make_uint4(out_1[0], out_1[1], out_1[2], out_1[3]);
            *reinterpret_cast<uint4 *>(pv) = make_uint4(out_v[0], out_v[1], out_v[2], out_v[3]);
        } else {
#pragma unroll
            for (int k = 0; k < kPackWordsPerThread; ++k) { p0[k] = out_0[k]; p1[k] = out_1[k]; pv[k] = out_v[k]; }
        }
    }
    if (irregular || (nonctl & 0x80808080u) != 0x80808080u) atomicOr(&flags[i], 1);
    if (marker) atomicOr(&flags[i], 2);
}

// ------------------------------------------------------------------------------------------ k2
// thread = 4 consecutive words (128 columns), grid.y slices the samples; partial results are
// OR-ed into the 5 x words state with atomics (one per thread per plane, no contention to speak of)
__global__ void __launch_bounds__(256)
column_mask_kernel(const uint32_t *__restrict__ planes, int64_t n, int64_t n_total, int64_t words,
                   uint32_t *__restrict__ state) {
    const int64_t q = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;      // uint4 index in a row
    if (q * 4 >= words) return;
    const int64_t per = (n + gridDim.y - 1) / gridDim.y;
    const int64_t s0 = int64_t(blockIdx.y) * per, s1 = min(n, s0 + per);
    const uint4 *P0 = reinterpret_cast<const uint4 *>(planes);
    const uint4 *P1 = reinterpret_cast<const uint4 *>(planes + 1 * n_total * words);
    const uint4 *PV = reinterpret_cast<const uint4 *>(planes + 2 * n_total * words);
    const int64_t rowq = words / 4;
    uint4 sA = make_uint4(0, 0, 0, 0), sC = sA, sG = sA, sT = sA, inv = sA;
    auto acc = [&](uint32_t &a, uint32_t &c, uint32_t &g, uint32_t &t, uint32_t &iv, uint32_t x0, uint32_t x1, uint32_t v) {
        a |= v & ~x1 & ~x0;
        c |= v & ~x1 & x0;
        t |= v & x1 & ~x0;
        g |= v & x1 & x0;
        iv |= ~v;
    };
#pragma unroll 4
    for (int64_t s = s0; s < s1; ++s) {
        const uint4 x0 = __ldcs(P0 + s * rowq + q), x1 = __ldcs(P1 + s * rowq + q), v = __ldcs(PV + s * rowq + q);
        acc(sA.x, sC.x, sG.x, sT.x, inv.x, x0.x, x1.x, v.x);
        acc(sA.y, sC.y, sG.y, sT.y, inv.y, x0.y, x1.y, v.y);
        acc(sA.z, sC.z, sG.z, sT.z, inv.z, x0.z, x1.z, v.z);
        acc(sA.w, sC.w, sG.w, sT.w, inv.w, x0.w, x1.w, v.w);
    }
    if (s0 >= s1) return;
    auto flush = [&](int plane, const uint4 &v) {
        uint32_t *d = state + plane * words + q * 4;
        if (v.x) atomicOr(d + 0, v.x);
        if (v.y) atomicOr(d + 1, v.y);
        if (v.z) atomicOr(d + 2, v.z);
        if (v.w) atomicOr(d + 3, v.w);
    };
    flush(0, sA); flush(1, sC); flush(2, sG); flush(3, sT); flush(4, inv);
}

// ------------------------------------------------------------------------------------------ k3a
__device__ __forceinline__ uint32_t keep_word(const uint32_t *state, int64_t words, int64_t w, int64_t G) {
    const uint32_t a = state[w], c = state[words + w], g = state[2 * words + w], t = state[3 * words + w];
    const uint32_t inv = state[4 * words + w];
    uint32_t two = (a & c) | (a & g) | (a & t) | (c & g) | (c & t) | (g & t);
    const uint32_t k = two & ~inv;                        // plane bit order
    uint32_t kc = 0;                                      // column order: bit cl = column 32 w + cl
#pragma unroll
    for (int cl = 0; cl < 32; ++cl) kc |= ((k >> plane_bit(cl)) & 1u) << cl;
    const int64_t base = w * 32;
    if (base + 32 > G) kc &= base >= G ? 0u : (0xFFFFFFFFu >> (32 - int(G - base)));
    return kc;
}

// pass 1: keep mask + per-block kept counts
__global__ void __launch_bounds__(256)
select_count_kernel(const uint32_t *__restrict__ state, int64_t words, int64_t G, uint32_t *__restrict__ keep,
                    uint32_t *__restrict__ block_counts) {
    const int64_t w = int64_t(blockIdx.x) * 256 + threadIdx.x;
    uint32_t k = 0;
    if (w < words) { k = keep_word(state, words, w, G); keep[w] = k; }
    uint32_t cnt = __popc(k);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    __shared__ uint32_t s[8];
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int i = 0; i < 8; ++i) t += s[i];
        block_counts[blockIdx.x] = t;
    }
}

// pass 2: exclusive scan of the block counts (one CTA; the list is a few hundred entries)
__global__ void __launch_bounds__(1024)
select_scan_kernel(uint32_t *__restrict__ block_counts, int64_t nblocks, uint32_t *__restrict__ total) {
    __shared__ uint32_t warp_sum[32];
    __shared__ uint32_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int64_t base = 0; base < nblocks; base += 1024) {
        const int64_t i = base + threadIdx.x;
        uint32_t v = i < nblocks ? block_counts[i] : 0, x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
            if ((threadIdx.x & 31) >= o) x += y;
        }
        if ((threadIdx.x & 31) == 31) warp_sum[threadIdx.x >> 5] = x;
        __syncthreads();
        if (threadIdx.x < 32) {
            uint32_t ws = warp_sum[threadIdx.x], wx = ws;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t y = __shfl_up_sync(0xffffffffu, wx, o);
                if (threadIdx.x >= o) wx += y;
            }
            warp_sum[threadIdx.x] = wx - ws;                 // exclusive prefix of warp sums
        }
        __syncthreads();
        const uint32_t excl = carry + warp_sum[threadIdx.x >> 5] + x - v;
        if (i < nblocks) block_counts[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}

// pass 3: every thread scatters the column indices of its word, in ascending order
__global__ void __launch_bounds__(256)
select_scatter_kernel(const uint32_t *__restrict__ keep, int64_t words, const uint32_t *__restrict__ block_off,
                      uint32_t *__restrict__ idx) {
    const int64_t w = int64_t(blockIdx.x) * 256 + threadIdx.x;
    uint32_t k = w < words ? keep[w] : 0;
    const uint32_t cnt = __popc(k);
    uint32_t x = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
        if ((threadIdx.x & 31) >= o) x += y;
    }
    __shared__ uint32_t s[8];
    if ((threadIdx.x & 31) == 31) s[threadIdx.x >> 5] = x;
    __syncthreads();
    uint32_t before = 0;
    for (int i = 0; i < int(threadIdx.x >> 5); ++i) before += s[i];
    uint32_t pos = block_off[blockIdx.x] + before + x - cnt;
    while (k) {
        const int b = __ffs(k) - 1;
        idx[pos++] = uint32_t(w * 32 + b);
        k &= k - 1;
    }
}

// ------------------------------------------------------------------------------------------ k3b
__global__ void __launch_bounds__(256)
compact_kernel(const uint32_t *__restrict__ planes, int64_t first, int64_t n_total, int64_t words,
               const uint32_t *__restrict__ idx, int64_t L, uint8_t *__restrict__ out, int64_t ld) {
    const int64_t i = blockIdx.y;
    const uint32_t *P0 = planes + (0 * n_total + first + i) * words;
    const uint32_t *P1 = planes + (1 * n_total + first + i) * words;
    for (int64_t l = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; l < L; l += int64_t(gridDim.x) * blockDim.x) {
        const uint32_t c = idx[l];
        const uint32_t w = c >> 5, b = uint32_t(plane_bit(int(c & 31)));
        const uint32_t code = ((P0[w] >> b) & 1u) | (((P1[w] >> b) & 1u) << 1);
        out[i * ld + l] = uint8_t((0x47544341u >> (8 * code)) & 0xFFu);
    }
}

}  // namespace btb

using namespace btb;

// =============================================================================== C-ABI, level 3
extern "C" int64_t btb_plane_words(int64_t G) { return plane_words(G); }

extern "C" int btb_pack_planes(const uint8_t *d_text, int64_t text_bytes, const int64_t *d_rec_off,
                               const int32_t *d_line_w, const int32_t *d_eol, int64_t n_chunk, int64_t first,
                               int64_t n_total, int64_t G, uint32_t *d_planes, int32_t *d_flags, void *stream) {
    if (!d_text || !d_rec_off || !d_line_w || !d_eol || !d_planes || !d_flags || n_chunk <= 0 || G <= 0)
        return fail(BTB_EINVAL, "btb_pack_planes: bad argument");
    const int64_t words = plane_words(G);
    const int wpt = g_pack_wpt == 4 ? 4 : 1;
    const int64_t cols = wpt == 4 ? PackCfg<4>::kCols : PackCfg<1>::kCols;
    const int64_t gx = (words * 32 + cols - 1) / cols;
    for (int64_t y0 = 0; y0 < n_chunk; y0 += 65535) {
        const int64_t ny = std::min<int64_t>(65535, n_chunk - y0);
        dim3 grid((unsigned)gx, (unsigned)ny);
        if (wpt == 4)
            pack_planes_kernel<4><<<grid, kPackThreads, 0, static_cast<cudaStream_t>(stream)>>>(
                d_text, text_bytes, d_rec_off + y0, d_line_w + y0, d_eol + y0, first + y0, n_total, G, words, d_planes, d_flags + y0);
        else
            pack_planes_kernel<1><<<grid, kPackThreads, 0, static_cast<cudaStream_t>(stream)>>>(
                d_text, text_bytes, d_rec_off + y0, d_line_w + y0, d_eol + y0, first + y0, n_total, G, words, d_planes, d_flags + y0);
    }
    BTB_CUDA(cudaGetLastError());
    return BTB_OK;
}

extern "C" int btb_column_mask(const uint32_t *d_planes, int64_t n, int64_t n_total, int64_t G,
                               uint32_t *d_colstate, int accumulate, void *stream_) {
    if (!d_planes || !d_colstate || n < 0 || G <= 0) return fail(BTB_EINVAL, "btb_column_mask: bad argument");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const int64_t words = plane_words(G);
    if (!accumulate) BTB_CUDA(cudaMemsetAsync(d_colstate, 0, size_t(5 * words * 4), stream));
    if (n == 0) return BTB_OK;
    const int64_t gx = (words / 4 + 255) / 256;
    int64_t gy = std::max<int64_t>(1, std::min<int64_t>(n / 32 + 1, (148 * 16 + gx - 1) / gx));
    gy = std::min<int64_t>(gy, 65535);
    column_mask_kernel<<<dim3((unsigned)gx, (unsigned)gy), 256, 0, stream>>>(d_planes, n, n_total, words, d_colstate);
    BTB_CUDA(cudaGetLastError());
    return BTB_OK;
}

extern "C" int btb_select_columns(const uint32_t *d_colstate, int64_t G, int pure_mode, uint32_t *d_keep,
                                  uint32_t *d_idx, uint32_t *d_count, void *stream_) {
    if (!d_colstate || !d_keep || !d_idx || !d_count || G <= 0) return fail(BTB_EINVAL, "btb_select_columns: bad argument");
    if (!pure_mode) return fail(BTB_EINVAL, "only snp-sites -c (pure mode) is implemented on the GPU");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const int64_t words = plane_words(G);
    const int64_t nb = (words + 255) / 256;
    uint32_t *d_blocks = nullptr;
    BTB_CUDA(cudaMallocAsync(reinterpret_cast<void **>(&d_blocks), size_t(nb * 4), stream));
    select_count_kernel<<<(unsigned)nb, 256, 0, stream>>>(d_colstate, words, G, d_keep, d_blocks);
    select_scan_kernel<<<1, 1024, 0, stream>>>(d_blocks, nb, d_count);
    select_scatter_kernel<<<(unsigned)nb, 256, 0, stream>>>(d_keep, words, d_blocks, d_idx);
    BTB_CUDA(cudaGetLastError());
    BTB_CUDA(cudaFreeAsync(d_blocks, stream));
    return BTB_OK;
}

extern "C" int btb_compact_columns(const uint32_t *d_planes, int64_t first, int64_t n_chunk, int64_t n_total,
                                   int64_t G, const uint32_t *d_idx, int64_t L, uint8_t *d_snps, int64_t snps_ld,
                                   void *stream) {
    if (!d_planes || !d_idx || !d_snps || n_chunk <= 0 || L < 0 || snps_ld < L) return fail(BTB_EINVAL, "btb_compact_columns: bad argument");
    if (L == 0) return BTB_OK;
    const int64_t words = plane_words(G);
    const int64_t gx = std::min<int64_t>((L + 255) / 256, 64);
    for (int64_t y0 = 0; y0 < n_chunk; y0 += 65535) {
        const int64_t ny = std::min<int64_t>(65535, n_chunk - y0);
        compact_kernel<<<dim3((unsigned)gx, (unsigned)ny), 256, 0, static_cast<cudaStream_t>(stream)>>>(
            d_planes, first + y0, n_total, words, d_idx, L, d_snps + y0 * snps_ld, snps_ld);
    }
    BTB_CUDA(cudaGetLastError());
    return BTB_OK;
}
