// k1-k3: variable-column extraction (the snp-sites stage) on the GPU.
//
// Replaces `snp-sites <multi_fasta> -c -o <snps.fas>` (btbphylo/phylogeny.py:133; behaviour:
// SURVEY.md App. A.2): a column is kept iff every sample's byte is one of ACGTacgt and at least
// two distinct bases occur; kept columns are emitted upper-cased in genome order.
//
//   k1 pack_planes   : FASTA text -> 3 bit-planes per sample (code bit 0, code bit 1, validity),
//                      32 columns per word, column 32 w + cl at bit cl.  code = (byte >> 1) & 3
//                      (A=0 C=1 T=2 G=3).  4 bytes are classified at a time with SWAR logic.
//                      3 bits per base instead of 8 is what lets 50,000 x 4.35 Mb stay in 180 GB.
//   k2 column_mask   : OR/AND reduction over samples -> seenA, seenC, seenG, seenT, anyInvalid.  The
//                      level-1/2 paths get this from k1 itself (btb_pack_planes_state: the state is
//                      accumulated while the planes are written); the kernel remains for planes that
//                      are already there (level 3, and the rare re-staged records).
//   k3 select/compact: keep-mask, popc + shuffle-scan stream compaction of the kept column
//                      indices, then a gather of those columns for every sample.
// All three are HBM-bound streaming kernels: bytes = N*G (text) + 3/8 N*G (planes written) + 2 N*L.
#include <algorithm>
#include <vector>

#include "common.cuh"

namespace btb {

constexpr int kPackThreads = 256;
constexpr int kMinLineW = 32;                            // shorter lines go through the host walk (<= 1 line end per 32 columns)
// One CTA packs kPackWords plane words (32 columns each, one per thread) of every sample it is given.
// 250 words = 8000 columns: with 60..80-column lines the CTA's text span is just under 2 x 4096 bytes, so the
// 256 threads classify it in two rounds of 16-byte chunks with next to no idle lanes.
constexpr int kPackWords = 250;
constexpr int kPackCols = kPackWords * 32;
constexpr int kSpanBytes = kPackCols + kPackCols / kMinLineW * 2 + 64;          // text bytes a CTA can touch per sample
constexpr int kBitWords = (kSpanBytes + 31) / 32 + 4;                           // flag bits of the span, + the words a funnel shift may touch

__host__ __device__ inline int64_t plane_words(int64_t G) { return ((G + 31) / 32 + 3) / 4 * 4; }

__device__ __forceinline__ uint4 ld_nc_v4(const uint8_t *p) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

template <int LUT>
__device__ __forceinline__ uint32_t lop3(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, %3, %4;" : "=r"(r) : "r"(a), "r"(b), "r"(c), "n"(LUT));
    return r;
}

// ------------------------------------------------------------------------------------------ k1
// Bit order inside a plane word: column c = 32 w + cl is bit cl.
__host__ __device__ inline int plane_bit(int cl) { return cl; }

// SWAR classification of 4 text bytes: one flag per byte lane, at bit 7 of the lane.
//   v   <=> byte is one of ACGTacgt   (exact; checked exhaustively in tests/test_capi_host.py)
//   b0, b1 = bits 1, 2 of a valid byte: code = (byte >> 1) & 3  (A 0, C 1, T 2, G 3)
//   ctl <=> byte < 0x20 (a control character: the line terminators, and nothing else in a well-formed file)
// Bit k of every byte is lined up at bit 7 with a LEFT shift (IMAD.SHL on the FMA pipe, leaving the ALU pipe to
// the LOP3s).  With b_k those aligned copies:  valid = ~b7 & b6 & ~b3 & R,  R = b4 ? (b2 & ~b1 & ~b0) : (b0 & (b1 | ~b2))
// = (u = b2 & ~b1;  b4 ? (u & ~b0) : (b0 & ~u)) -- four LOP3 in all (the decomposition was found by exhaustive search).
struct Class4 { uint32_t v, b0, b1, ctl; };
__device__ __forceinline__ Class4 classify4(uint32_t w) {
    const uint32_t b0 = w << 7, b1 = w << 6, b2 = w << 5, b3 = w << 4, b4 = w << 3, b5 = w << 2, b6 = w << 1;
    constexpr int A = 0xF0, B = 0xCC, C = 0xAA;
    const uint32_t u = lop3<(B & ~A) & 0xFF>(b1, b2, 0u);                          // b2 & ~b1
    const uint32_t R = lop3<((C & A & ~B) | (~C & B & ~A)) & 0xFF>(u, b0, b4);     // b4 ? (u & ~b0) : (b0 & ~u)
    const uint32_t p = lop3<(~A & B & ~C) & 0xFF>(w, b6, b3);                      // ~b7 & b6 & ~b3
    Class4 c;
    c.v = lop3<(A & B & C) & 0xFF>(p, R, 0x80808080u);
    c.b0 = c.v & b1;
    c.b1 = c.v & b2;
    c.ctl = lop3<(~(A | B | C)) & 0xFF>(w, b6, b5) & 0x80808080u;                  // none of bits 7, 6, 5
    return c;
}
// the four lane flags (bits 7, 15, 23, 31) as a nibble, lane 0 (the first text byte) lowest: one multiply gathers
// them into bits 28..31 (all partial products land on distinct bits, so nothing carries)
__device__ __forceinline__ uint32_t gather4(uint32_t flags7) { return (flags7 * 0x00204081u) >> 28; }

// k1 with k2 fused in.  grid.x = blocks of kPackCols columns, grid.y = groups of samples; a CTA walks through its
// samples one after the other:
//   phase A  the CTA's text span (bases AND line terminators, as they lie in the file) is read once, 16 bytes per
//            thread and step, coalesced, classified in registers, and the four flags of every text byte go to
//            four bit arrays in shared memory, in TEXT order -- no staging of the text, no realignment of bytes;
//   phase B  every thread builds its own plane word (32 columns) from those bit arrays at the BIT level: the
//            columns are 32 consecutive flag bits, except that after the `rem` bases left on the current line the
//            text skips e terminator bytes (at most one line end per word since W >= 32) -- three words, three
//            funnel shifts and one masked merge per plane.  The control-character plane must be empty at the
//            columns and the terminator bytes must be the expected ones (checked where a line ends: two byte
//            loads), else the record is flagged and the host re-stages it through its exact parser.
// The column state (seenA, seenC, seenG, seenT, anyInvalid: what k2 used to recompute from the planes) is
// accumulated in registers over the CTA's samples and merged into `state` once per launch: plain read-modify-
// write when the CTA has every sample of the launch (each state word then belongs to exactly one thread; launches
// on one stream are ordered), atomicOr otherwise.
// Bytes: N*G text read once + 3/8 N*G planes written; nothing is read twice.
__global__ void __launch_bounds__(kPackThreads)
pack_planes_kernel(const uint8_t *__restrict__ text, int64_t text_bytes, const int64_t *__restrict__ rec_off,
                   const int32_t *__restrict__ line_w, const int32_t *__restrict__ eol_len, int64_t n_chunk, int64_t first,
                   int64_t n_total, int64_t G, int64_t words, uint32_t *__restrict__ planes,
                   int32_t *__restrict__ flags, uint32_t *__restrict__ state) {
    __shared__ __align__(16) uint32_t bits[2][4][kBitWords];       // [buffer][v, b0, b1, ctl][text bit]
    const int64_t c0 = int64_t(blockIdx.x) * kPackCols;
    const int64_t w = (c0 >> 5) + threadIdx.x;                      // this thread's plane word
    const bool owns = threadIdx.x < kPackWords && w < words;
    const uint32_t G32 = uint32_t(G), cb = uint32_t(c0);           // genome lengths fit 32 bits (the tools use int)
    const uint32_t ce = (c0 >= G) ? cb : (G32 - cb < uint32_t(kPackCols) ? G32 : cb + uint32_t(kPackCols));   // end of this CTA's columns
    const uint32_t c = cb + 32u * threadIdx.x;                      // first column of this thread's word
    const int64_t per = (n_chunk + gridDim.y - 1) / gridDim.y;
    const int64_t s_lo = int64_t(blockIdx.y) * per, s_hi = s_lo + per < n_chunk ? s_lo + per : n_chunk;
    uint32_t sA = 0, sC = 0, sG = 0, sT = 0, sInv = 0;
    uint32_t parity = 0;                                           // bit-array buffer of the next sample that gets that far
    auto is_marker = [](uint32_t b) { return b == '>' || b == '@' || b == '+'; };
    auto byte_at = [&](int64_t a) -> uint32_t { return a < text_bytes ? uint32_t(text[a]) : 0x20u; };
    for (int64_t i = s_lo; i < s_hi; ++i) {
        const uint32_t W = uint32_t(line_w[i]), e = uint32_t(eol_len[i]);
        const int64_t rec = rec_off[i];
        uint32_t *p0 = planes + (0 * n_total + first + i) * words + w;
        uint32_t *p1 = planes + (1 * n_total + first + i) * words + w;
        uint32_t *pv = planes + (2 * n_total + first + i) * words + w;
        if (cb >= ce || (W && W < kMinLineW)) {            // padding words / a layout the host must re-stage
            if (owns) { *p0 = 0; *p1 = 0; *pv = 0; }
            if (cb < ce && threadIdx.x == 0) atomicOr(&flags[i], 1);
            continue;                                       // (uniform for the CTA: no barrier is skipped by a part of it)
        }
        auto off = [&](uint32_t col) -> int64_t { return W ? int64_t(col) + int64_t(col / W) * e : int64_t(col); };
        const int64_t lo = rec + off(cb);
        int64_t hi = rec + off(ce - 1) + 1 + e;           // incl. the terminator after the last column
        if (hi > text_bytes) hi = text_bytes;
        const int64_t alo = lo & ~int64_t(15);
        const int nchunk = int((hi - alo + 15) >> 4) + 1;  // + 1: phase B may look one word beyond
        uint32_t(*buf)[kBitWords] = bits[parity];
        parity ^= 1u;
        // ---- phase A: classify the text, 16 bytes per thread and step
        for (int j = threadIdx.x; j < nchunk; j += kPackThreads) {
            uint4 t4 = make_uint4(0x20202020u, 0x20202020u, 0x20202020u, 0x20202020u);
            if (alo + 16 * int64_t(j) < text_bytes) t4 = ld_nc_v4(text + alo + 16 * int64_t(j));
            const uint32_t tw[4] = {t4.x, t4.y, t4.z, t4.w};
            uint32_t hv = 0, h0 = 0, h1 = 0, hc = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const Class4 f = classify4(tw[k]);
                hv += gather4(f.v) << (4 * k); h0 += gather4(f.b0) << (4 * k);
                h1 += gather4(f.b1) << (4 * k); hc += gather4(f.ctl) << (4 * k);
            }
            reinterpret_cast<uint16_t *>(buf[0])[j] = uint16_t(hv);
            reinterpret_cast<uint16_t *>(buf[1])[j] = uint16_t(h0);
            reinterpret_cast<uint16_t *>(buf[2])[j] = uint16_t(h1);
            reinterpret_cast<uint16_t *>(buf[3])[j] = uint16_t(hc);
        }
        __syncthreads();                                   // (the other buffer is free again: everyone is past its phase B)
        // ---- phase B: this thread's 32 columns, taken from the bit arrays
        if (c < ce) {
            const uint32_t ncols = ce - c < 32u ? ce - c : 32u;
            const uint32_t line = W ? c / W : 0u;
            const uint32_t pos = W ? c - line * W : 0u;    // column inside the current line
            const uint32_t rem = W ? W - pos : 0xFFFFu;    // bases left on that line
            const int64_t at = rec + int64_t(c) + int64_t(line) * e;       // text offset of column c
            const uint32_t o = uint32_t(at - alo);         // = bit index in the arrays
            const uint32_t q = o >> 5, sh = o & 31u;
            const uint32_t M = rem >= 32u ? 0xFFFFFFFFu : (1u << rem) - 1u;          // columns before the line end
            const uint32_t colmask = ncols >= 32u ? 0xFFFFFFFFu : (1u << ncols) - 1u;  // last word of a row: columns < G
            auto take = [&](const uint32_t *a) {
                const uint32_t a0 = a[q], a1 = a[q + 1], a2 = a[q + 2];
                const uint32_t x0 = __funnelshift_r(a0, a1, sh), x1 = __funnelshift_r(a1, a2, sh);
                return ((x0 & M) | (__funnelshift_r(x0, x1, e) & ~M)) & colmask;
            };
            const uint32_t fv = take(buf[0]), f0 = take(buf[1]), f1 = take(buf[2]), fc = take(buf[3]);
            bool irregular = fc != 0;                      // a control character among the bases
            bool marker = false;
            // a line of the sequence region that starts with '>', '@' or '+' is not sequence at all for the parser
            // the tools share (it opens the next record / a quality block): the host must re-index
            if (pos == 0 && (W || c == 0)) marker = is_marker(byte_at(at));
            if (rem <= ncols && c + rem < G32) {           // a line ends inside (or at the end of) this word and more bases follow
                const int64_t t = at + rem;
                if (e == 1) irregular |= byte_at(t) != '\n';
                else irregular |= (byte_at(t) != '\r') | (byte_at(t + 1) != '\n');
                marker |= is_marker(byte_at(t + e));       // first byte of the next line
            }
            if (owns) { *p0 = f0; *p1 = f1; *pv = fv; }
            sA |= fv & ~f1 & ~f0; sC |= fv & ~f1 & f0; sT |= fv & f1 & ~f0; sG |= fv & f1 & f0;
            sInv |= ~fv & colmask;
            if (irregular) atomicOr(&flags[i], 1);
            if (marker) atomicOr(&flags[i], 2);
        } else if (owns) {
            *p0 = 0; *p1 = 0; *pv = 0;                     // padding words of the last column block
        }
    }
    if (state && owns && c < ce && s_hi > s_lo) {
        uint32_t *d = state + w;
        if (gridDim.y == 1) {
            d[0] |= sA; d[words] |= sC; d[2 * words] |= sG; d[3 * words] |= sT; d[4 * words] |= sInv;
        } else {
            if (sA) atomicOr(d, sA);
            if (sC) atomicOr(d + words, sC);
            if (sG) atomicOr(d + 2 * words, sG);
            if (sT) atomicOr(d + 3 * words, sT);
            if (sInv) atomicOr(d + 4 * words, sInv);
        }
    }
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

namespace btb {
// k1 (+ k2 when d_colstate != nullptr: the column state of these samples is OR-ed into it)
int pack_planes(const uint8_t *d_text, int64_t text_bytes, const int64_t *d_rec_off, const int32_t *d_line_w,
                const int32_t *d_eol, int64_t n_chunk, int64_t first, int64_t n_total, int64_t G, uint32_t *d_planes,
                int32_t *d_flags, uint32_t *d_colstate, cudaStream_t stream) {
    const int64_t words = plane_words(G);
    const int64_t gx = (words + kPackWords - 1) / kPackWords;
    // sample groups: one CTA per column block takes ALL samples of the launch (plain state merge) unless that
    // leaves the GPU short of CTAs and every group still gets enough samples to make its atomics cheap
    int sms = 148;
    { int dev = 0; if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
    int64_t gy = 1;
    if (gx < int64_t(sms) * 4) gy = std::max<int64_t>(1, std::min<int64_t>((int64_t(sms) * 6 + gx - 1) / gx, n_chunk / (d_colstate ? 24 : 1)));
    gy = std::max<int64_t>(1, std::min<int64_t>(gy, std::min<int64_t>(n_chunk, 65535)));
    pack_planes_kernel<<<dim3((unsigned)gx, (unsigned)gy), kPackThreads, 0, stream>>>(
        d_text, text_bytes, d_rec_off, d_line_w, d_eol, n_chunk, first, n_total, G, words, d_planes, d_flags, d_colstate);
    BTB_CUDA(cudaGetLastError());
    return BTB_OK;
}
}  // namespace btb

extern "C" int btb_pack_planes(const uint8_t *d_text, int64_t text_bytes, const int64_t *d_rec_off,
                               const int32_t *d_line_w, const int32_t *d_eol, int64_t n_chunk, int64_t first,
                               int64_t n_total, int64_t G, uint32_t *d_planes, int32_t *d_flags, void *stream) {
    if (!d_text || !d_rec_off || !d_line_w || !d_eol || !d_planes || !d_flags || n_chunk <= 0 || G <= 0)
        return fail(BTB_EINVAL, "btb_pack_planes: bad argument");
    return pack_planes(d_text, text_bytes, d_rec_off, d_line_w, d_eol, n_chunk, first, n_total, G, d_planes, d_flags, nullptr,
                       static_cast<cudaStream_t>(stream));
}

extern "C" int btb_pack_planes_state(const uint8_t *d_text, int64_t text_bytes, const int64_t *d_rec_off,
                                     const int32_t *d_line_w, const int32_t *d_eol, int64_t n_chunk, int64_t first,
                                     int64_t n_total, int64_t G, uint32_t *d_planes, int32_t *d_flags,
                                     uint32_t *d_colstate, void *stream) {
    if (!d_text || !d_rec_off || !d_line_w || !d_eol || !d_planes || !d_flags || !d_colstate || n_chunk <= 0 || G <= 0)
        return fail(BTB_EINVAL, "btb_pack_planes_state: bad argument");
    return pack_planes(d_text, text_bytes, d_rec_off, d_line_w, d_eol, n_chunk, first, n_total, G, d_planes, d_flags, d_colstate,
                       static_cast<cudaStream_t>(stream));
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
