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
// One CTA packs kPackThreadsB x WPT plane words (32 columns each, WPT consecutive words per thread) of every sample it
// is given.  250 x WPT words = 8000 x WPT columns: with 60..80-column lines the CTA's text span is just under
// 2 x WPT x 4096 bytes, so the 256 threads classify it in 2 x WPT rounds of 16-byte chunks with next to no idle lanes.
constexpr int kPackThreadsB = 250;                       // threads that own plane words
constexpr uint32_t kNone = 0xFFFFFFFFu;
template <int WPT> struct PackCfg {
    static constexpr int kWords = kPackThreadsB * WPT;
    static constexpr int kCols = kWords * 32;
    static constexpr int kSpanBytes = kCols + kCols / kMinLineW * 2 + 64;       // text bytes a CTA can touch per sample
    static constexpr int kStageBytes = (kSpanBytes + 16 + 15) / 16 * 16;
    static constexpr int kBitWords = (kSpanBytes + 31) / 32 + 4;                // flag bits of the span + what a funnel shift may touch
    static constexpr int kRing = 2;                                             // stage buffers: the text of this sample and of the next
    static constexpr int kSmemBytes = kRing * kStageBytes + 2 * 3 * kBitWords * 4 + 64;
};
int g_pack_wpt = 2;                                      // "pack_wpt": plane words per thread of k1 (1, 2 or 4)
int g_pack_gy = 0;                                       // "pack_gy": sample groups per launch (0 = automatic)

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
//   v  <=> byte is one of ACGTacgt   (exact; checked exhaustively in tests/test_capi_host.py)
//   f1 = bit 2 of a valid byte, f0 = bit 1 of a valid byte: code = (byte >> 1) & 3  (A 0, C 1, T 2, G 3)
//   f0 is ALSO set for a control character (byte < 0x20: the line terminators, and nothing else in a well-formed
//   file).  A control character is never valid, so "f0 without v" marks it: the planes need no fourth flag.
// Bit k of every byte is lined up at bit 7 with a LEFT shift (IMAD.SHL: the FMA pipe, which has twice the ALU
// pipe's rate and little else to do here).  With b_k those aligned copies and m7 = 0x80 where bit 7 is clear:
//   valid = m7 & b6 & ~b3 & R,   R = b4 ? (b2 & ~b1 & ~b0) : (b0 & (b1 | ~b2)) = (u = b2 & ~b1;  b4 ? (u & ~b0) : (b0 & ~u))
// (that two-LOP3 form of R was found by exhaustive search) -- eight LOP3 per four bytes in all.
struct Class4 { uint32_t v, f0, f1; };
__device__ __forceinline__ Class4 classify4(uint32_t w) {
    const uint32_t b0 = w << 7, b1 = w << 6, b2 = w << 5, b3 = w << 4, b4 = w << 3, b5 = w << 2, b6 = w << 1;
    constexpr int A = 0xF0, B = 0xCC, C = 0xAA;
    const uint32_t m7 = lop3<(~A & B) & 0xFF>(w, 0x80808080u, 0u);                 // bit 7 of the lane where the byte's is clear
    const uint32_t u = lop3<(B & ~A) & 0xFF>(b1, b2, 0u);                          // b2 & ~b1
    const uint32_t R = lop3<((C & A & ~B) | (~C & B & ~A)) & 0xFF>(u, b0, b4);     // b4 ? (u & ~b0) : (b0 & ~u)
    const uint32_t p = lop3<(A & B & ~C) & 0xFF>(m7, b6, b3);                      // ~b7 & b6 & ~b3
    const uint32_t ctl = lop3<(A & ~B & ~C) & 0xFF>(m7, b6, b5);                   // none of bits 7, 6, 5
    Class4 c;
    c.v = p & R;
    c.f0 = lop3<((A & B) | C) & 0xFF>(c.v, b1, ctl);
    c.f1 = c.v & b2;
    return c;
}
// The flags of 16 text bytes (four registers, flag = 0x80 in its byte lane) as 16 bits in text order: byte-wise dot
// products with the weights 1, 2, 4, 8 (16, 32, 64, 128 for the second register of a pair) -- IDP4A, FMA pipe --
// give 128 x the flag byte of a register pair; the two pairs are joined and the factor 128 dropped with IMAD / IMAD.HI.
__device__ __forceinline__ uint32_t gather16(uint32_t g0, uint32_t g1, uint32_t g2, uint32_t g3) {
    const uint32_t lo = __dp4a(g1, 0x80402010u, __dp4a(g0, 0x08040201u, 0u));
    const uint32_t hi = __dp4a(g3, 0x80402010u, __dp4a(g2, 0x08040201u, 0u));
    return __umulhi(hi * 256u + lo, 1u << 25);                                      // >> 7
}

// ---- async bulk copies (TMA) into shared memory, completion on an mbarrier
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "PACK_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra PACK_DONE;\n\t"
        "bra PACK_WAIT;\n\t"
        "PACK_DONE:\n\t}"
        ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// k1 with k2 fused in.  grid.x = blocks of PackCfg::kCols columns, grid.y = groups of samples; a CTA walks through
// its samples one after the other, two samples' text in flight:
//   stage    the CTA's text span of a sample (bases AND line terminators, as they lie in the file: 8.2 KB per
//            8000 columns of 60-column lines) comes in by ONE async bulk copy (TMA, cp.async.bulk) into a ring of
//            shared-memory buffers, issued ahead by an elected thread as soon as the buffer of the previous
//            sample is free; its bytes complete an mbarrier;
//   phase A  the span is classified in place, 16 bytes per thread and step, and the three flags of every text
//            byte go to three bit arrays in shared memory, in TEXT order -- no realignment of bytes;
//   phase B  every thread builds its own plane words (32 columns each) from those bit arrays at the BIT level: the
//            columns are 32 consecutive flag bits, except that after the `rem` bases left on the current line the
//            text skips e terminator bytes (at most one line end per word since W >= 32) -- three words, three
//            funnel shifts and one masked merge per plane.  No column may hold a control character and the
//            terminator bytes must be the expected ones (checked where a line ends, in the staged text), else the
//            record is flagged and the host re-stages it through its exact parser.
// Everything of phase B that depends only on the line layout (W, e) -- the bit offset of every word, the merge
// masks, which words hold a line end -- is computed once and kept in registers while the layout stays the same
// (in a consensus multi-FASTA it never changes); per sample only the record's 16-byte misalignment is new.
// The column state (seenA, seenC, seenG, seenT, anyInvalid: what k2 used to recompute from the planes) is
// accumulated in registers over the CTA's samples and merged into `state` once per launch: plain read-modify-
// write when the CTA has every sample of the launch (each state word then belongs to exactly one thread; launches
// on one stream are ordered), atomicOr otherwise.
// Bytes: N*G text read once + 3/8 N*G planes written; nothing is read twice.
template <int WPT>
__global__ void __launch_bounds__(kPackThreads)
pack_planes_kernel(const uint8_t *__restrict__ text, int64_t text_bytes, const int64_t *__restrict__ rec_off,
                   const int32_t *__restrict__ line_w, const int32_t *__restrict__ eol_len, int64_t n_chunk, int64_t first,
                   int64_t n_total, int64_t G, int64_t words, uint32_t *__restrict__ planes,
                   int32_t *__restrict__ flags, uint32_t *__restrict__ state) {
    using Cfg = PackCfg<WPT>;
    constexpr int kBitWords = Cfg::kBitWords;
    extern __shared__ __align__(128) uint8_t pack_smem[];
    constexpr int kRing = Cfg::kRing;
    uint8_t *stage = pack_smem;                                                            // [kRing][kStageBytes]
    uint32_t *bits = reinterpret_cast<uint32_t *>(pack_smem + kRing * Cfg::kStageBytes);   // [2][v, f0, f1][kBitWords]
    uint64_t *full = reinterpret_cast<uint64_t *>(pack_smem + kRing * Cfg::kStageBytes + 2 * 3 * kBitWords * 4);   // [kRing]
    const int64_t c0 = int64_t(blockIdx.x) * Cfg::kCols;
    const int64_t w = (c0 >> 5) + int64_t(threadIdx.x) * WPT;      // this thread's first plane word
    const bool owns = threadIdx.x < kPackThreadsB && w < words;    // (words is a multiple of 4: all WPT words or none)
    const uint32_t G32 = uint32_t(G), cb = uint32_t(c0);          // genome lengths fit 32 bits (the tools use int)
    const int64_t per = (n_chunk + gridDim.y - 1) / gridDim.y;
    const int64_t s_lo = int64_t(blockIdx.y) * per, s_hi = s_lo + per < n_chunk ? s_lo + per : n_chunk;
    // this thread's words of the CTA's first sample in the three planes; the pointers move on by one sample per iteration
    uint32_t *p0 = planes + (first + s_lo) * words + w;
    uint32_t *p1 = p0 + n_total * words;
    uint32_t *pv = p1 + n_total * words;
    auto store_words = [&](const uint32_t (&x0)[WPT], const uint32_t (&x1)[WPT], const uint32_t (&xv)[WPT]) {
        if constexpr (WPT == 4) {
            *reinterpret_cast<uint4 *>(p0) = make_uint4(x0[0], x0[1], x0[2], x0[3]);
            *reinterpret_cast<uint4 *>(p1) = make_uint4(x1[0], x1[1], x1[2], x1[3]);
            *reinterpret_cast<uint4 *>(pv) = make_uint4(xv[0], xv[1], xv[2], xv[3]);
        } else if constexpr (WPT == 2) {
            *reinterpret_cast<uint2 *>(p0) = make_uint2(x0[0], x0[1]);
            *reinterpret_cast<uint2 *>(p1) = make_uint2(x1[0], x1[1]);
            *reinterpret_cast<uint2 *>(pv) = make_uint2(xv[0], xv[1]);
        } else {
            *p0 = x0[0]; *p1 = x1[0]; *pv = xv[0];
        }
    };
    auto next_sample = [&]() { p0 += words; p1 += words; pv += words; };
    const uint32_t zeros[WPT] = {};
    if (c0 >= G) {                                                  // a block of padding words only
        if (owns)
            for (int64_t i = s_lo; i < s_hi; ++i, next_sample()) store_words(zeros, zeros, zeros);
        return;
    }
    const uint32_t ce = G32 - cb < uint32_t(Cfg::kCols) ? G32 : cb + uint32_t(Cfg::kCols);   // end of this CTA's columns
    const uint32_t c = cb + 32u * WPT * threadIdx.x;                // first column of this thread's first word
    auto active = [&](int64_t i) { const uint32_t W = uint32_t(line_w[i]); return !(W && W < uint32_t(kMinLineW)); };
    // layout-dependent values, recomputed when (W, e) changes
    uint32_t W_cur = 0xFFFFFFFFu, e_cur = 0;
    uint32_t rel_cb = 0, span_len = 0;                              // text offset of column cb relative to the record; bytes to its end
    uint32_t d_k[WPT], M_k[WPT], cm_k[WPT];                        // per word: bit offset - a15, columns before the line end, columns < G
    uint32_t end_k[WPT], ls_k[WPT], lsoff_k[WPT];                  // text offset of the line end in the word (kNone: none); line-start bit, its text offset
    uint32_t eol_first = '\n', eol_last_at = 0;                    // first terminator byte; offset of the last one ('\n') from it
    auto relayout = [&](uint32_t W, uint32_t e) {
        W_cur = W; e_cur = e;
        eol_first = e == 2u ? uint32_t('\r') : uint32_t('\n');
        eol_last_at = e ? e - 1u : 0u;
        rel_cb = cb + (W ? cb / W : 0u) * e;
        span_len = (ce - 1) + (W ? (ce - 1) / W : 0u) * e + 1 + e - rel_cb;    // incl. the terminator after the last column
#pragma unroll
        for (int k = 0; k < WPT; ++k) {
            const uint32_t ck = c + 32u * k;
            d_k[k] = 0; M_k[k] = 0xFFFFFFFFu; cm_k[k] = 0u; end_k[k] = kNone; ls_k[k] = 0u; lsoff_k[k] = 0u;   // a padding word: all zero
            if (ck < ce) {
                const uint32_t ncols = ce - ck < 32u ? ce - ck : 32u;
                const uint32_t line = W ? ck / W : 0u;
                const uint32_t pos = W ? ck - line * W : 0u;
                const uint32_t rem = W ? W - pos : 0xFFFFu;         // bases left on the line at column ck
                d_k[k] = ck + line * e - rel_cb;
                M_k[k] = rem >= 32u ? 0xFFFFFFFFu : (1u << rem) - 1u;
                cm_k[k] = ncols >= 32u ? 0xFFFFFFFFu : (1u << ncols) - 1u;
                // a line ends inside (or at the end of) this word and more bases follow: its terminator is checked here
                // (inside the staged span, which includes the terminator after the CTA's last column)
                if (rem <= ncols && ck + rem < G32) end_k[k] = d_k[k] + rem;
                // Every line start is a column of exactly one word: its first one, or the one after a line end inside it.
                if ((W && rem == W) || ck == 0u) { ls_k[k] = 1u; lsoff_k[k] = d_k[k]; }
                else if (rem < ncols) { ls_k[k] = 1u << rem; lsoff_k[k] = d_k[k] + rem + e; }
            }
        }
    };
    const uint32_t bar0 = smem_u32(full);
    auto full_bar_at = [&](uint32_t k) { return bar0 + 8u * k; };                  // the text of buffer k has landed
    auto empty_bar_at = [&](uint32_t k) { return bar0 + 8u * (uint32_t(kRing) + k); };   // every thread is done with buffer k
    const int64_t text_padded = (text_bytes + 15) & ~int64_t(15);  // the allocation is padded to a 16-byte boundary
    auto issue_at = [&](uint32_t W, uint32_t e, int64_t rec, int k) {   // one thread: bulk copy of a sample's span into stage[k]
        uint32_t rel = rel_cb, span = span_len;                     // (the layout registers hold them for the current layout)
        if (W != W_cur || e != e_cur) {
            rel = cb + (W ? cb / W : 0u) * e;
            span = (ce - 1) + (W ? (ce - 1) / W : 0u) * e + 1 + e - rel;
        }
        const int64_t lo = rec + int64_t(rel);
        int64_t hi = lo + int64_t(span);
        if (hi > text_bytes) hi = text_bytes;
        const int64_t alo = lo & ~int64_t(15);
        int64_t bytes = (((hi - alo + 15) >> 4) + 1) * 16;          // + 1 chunk: phase B may look one word beyond
        if (alo + bytes > text_padded) bytes = text_padded - alo;   // (chunks beyond the text read as blanks in phase A)
        mbar_expect_tx(full_bar_at(uint32_t(k)), uint32_t(bytes));
        bulk_g2s(smem_u32(stage + k * Cfg::kStageBytes), text + alo, uint32_t(bytes), full_bar_at(uint32_t(k)));
    };
    auto issue = [&](int64_t i, int k) { issue_at(uint32_t(line_w[i]), uint32_t(eol_len[i]), rec_off[i], k); };
    int64_t nxt = s_lo;                                             // next sample whose text has not been requested (thread 0)
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < kRing; ++k) { mbar_init(full_bar_at(uint32_t(k)), 1); mbar_init(empty_bar_at(uint32_t(k)), kPackThreads); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        while (nxt < s_hi && !active(nxt)) ++nxt;           // the first active sample; the following ones are requested in the loop
        if (nxt < s_hi) issue(nxt++, 0);
    }
    __syncthreads();
    uint32_t sA[WPT] = {}, sC[WPT] = {}, sG[WPT] = {}, sT[WPT] = {}, sInv[WPT] = {};
    uint32_t k_act = 0, ring = 0, ring_phase = 0;                   // active samples done; their stage buffer and its mbarrier phase
    auto is_marker = [](uint32_t b) { return b == '>' || b == '@' || b == '+'; };
    // the record table entry of a sample is read one iteration ahead: no thread waits for it, and thread 0 has what the
    // next request needs at hand
    uint32_t W_n = 0, e_n = 0;
    int64_t rec_n = 0;
    if (s_lo < s_hi) { W_n = uint32_t(line_w[s_lo]); e_n = uint32_t(eol_len[s_lo]); rec_n = rec_off[s_lo]; }
    for (int64_t i = s_lo; i < s_hi; ++i) {
        const uint32_t W = W_n, e = e_n;
        const int64_t rec = rec_n;
        if (i + 1 < s_hi) { W_n = uint32_t(line_w[i + 1]); e_n = uint32_t(eol_len[i + 1]); rec_n = rec_off[i + 1]; }
        if (W && W < uint32_t(kMinLineW)) {                 // a layout the host must re-stage
            if (owns) store_words(zeros, zeros, zeros);
            if (threadIdx.x == 0) atomicOr(&flags[i], 1);
            next_sample();
            continue;                                       // (uniform for the CTA: no barrier is skipped by a part of it)
        }
        if (W != W_cur || e != e_cur) relayout(W, e);
        // per sample only the record's start is new; everything below is 32-bit arithmetic relative to it
        const uint32_t a15 = (uint32_t(rec) + rel_cb) & 15u;        // misalignment of the span's first byte
        // The span may run beyond the end of the text in the last record only (a truncated file): then fewer bytes were
        // staged, and the chunks past them count as blanks.
        int nchunk = int((a15 + span_len + 15u) >> 4) + 1;          // + 1: phase B may look one word beyond
        int live = nchunk;                                          // chunks that hold text at all
        if (rec + int64_t(rel_cb) + int64_t(span_len) + 48 > text_bytes) {      // (48: every chunk below lies inside the text)
            const int64_t room64 = text_bytes - (rec + int64_t(rel_cb));
            const uint32_t room = room64 > int64_t(1 << 30) ? (1u << 30) : (room64 > 0 ? uint32_t(room64) : 0u);
            const uint32_t len = span_len < room ? span_len : room;
            nchunk = int((a15 + len + 15u) >> 4) + 1;
            live = int((a15 + room + 15u) >> 4);
        }
        const int nload = nchunk < live ? nchunk : live;
        uint32_t *buf = bits + (k_act & 1u) * 3 * kBitWords;
        const uint8_t *stg = stage + ring * Cfg::kStageBytes;
        if (threadIdx.x == 0) {     // the text of the NEXT active sample is requested a whole sample ahead: into the other
            const bool follows = nxt == i + 1 && nxt < s_hi && !(W_n && W_n < uint32_t(kMinLineW));   // (the usual case)
            if (!follows) while (nxt < s_hi && !active(nxt)) ++nxt;
            if (nxt < s_hi) {       // buffer, once every thread is done with the sample before this one
                const uint32_t nb = ring + 1u == uint32_t(kRing) ? 0u : ring + 1u;
                if (k_act + 1u >= uint32_t(kRing)) mbar_wait(empty_bar_at(uint32_t(nb)), ((k_act + 1u - uint32_t(kRing)) / uint32_t(kRing)) & 1u);
                if (follows) issue_at(W_n, e_n, rec_n, int(nb)); else issue(nxt, int(nb));
                ++nxt;
            }
        }
        mbar_wait(full_bar_at(uint32_t(ring)), ring_phase);
        // ---- phase A: classify the text, 16 bytes per thread and step
        const uint4 *st4 = reinterpret_cast<const uint4 *>(stg);
        uint16_t *hvp = reinterpret_cast<uint16_t *>(buf), *h0p = reinterpret_cast<uint16_t *>(buf + kBitWords);
        uint16_t *h1p = reinterpret_cast<uint16_t *>(buf + 2 * kBitWords);
#pragma unroll 2
        for (int j = threadIdx.x; j < nload; j += kPackThreads) {
            const uint4 t4 = st4[j];
            const Class4 g0 = classify4(t4.x), g1 = classify4(t4.y), g2 = classify4(t4.z), g3 = classify4(t4.w);
            hvp[j] = uint16_t(gather16(g0.v, g1.v, g2.v, g3.v));
            h0p[j] = uint16_t(gather16(g0.f0, g1.f0, g2.f0, g3.f0));
            h1p[j] = uint16_t(gather16(g0.f1, g1.f1, g2.f1, g3.f1));
        }
        if (nload < nchunk)                                          // blanks: not a base, not a control character
            for (int j = nload + threadIdx.x; j < nchunk; j += kPackThreads) hvp[j] = h0p[j] = h1p[j] = 0;
        __syncthreads();            // this sample's bit arrays are complete
        // ---- phase B: this thread's words, taken from the bit arrays
        uint32_t o0[WPT], o1[WPT], ov[WPT];
        uint32_t irregular = 0;                             // != 0: the record is not laid out as promised
        bool marker = false;
#pragma unroll
        for (int k = 0; k < WPT; ++k) {
            const uint32_t o = a15 + d_k[k];               // bit index in the arrays = byte index in the stage buffer
            const uint32_t q = o >> 5, sh = o & 31u, M = M_k[k], cm = cm_k[k];
            auto take = [&](const uint32_t *a) {
                const uint32_t a0 = a[q], a1 = a[q + 1], a2 = a[q + 2];
                const uint32_t x0 = __funnelshift_r(a0, a1, sh), x1 = __funnelshift_r(a1, a2, sh);
                return ((x0 & M) | (__funnelshift_r(x0, x1, e) & ~M)) & cm;
            };
            const uint32_t fv = take(buf), g0 = take(buf + kBitWords), f1 = take(buf + 2 * kBitWords);
            irregular |= g0 & ~fv;                         // a control character among the bases
            // the terminator of the line that ends in this word must be the expected bytes ('\n', or '\r' '\n')
            const bool has_end = end_k[k] != kNone;
            const uint32_t t = has_end ? a15 + end_k[k] : 0u;
            const uint32_t t_first = stg[t], t_last = stg[t + eol_last_at];
            irregular |= has_end ? ((t_first ^ eol_first) | (t_last ^ uint32_t('\n'))) : 0u;
            // A line of the sequence region that starts with '>', '@' or '+' is not sequence at all for the parser the
            // tools share (it opens the next record / a quality block): the host must re-index.  A marker is never a
            // base: the byte is only looked at when the line's first column is not valid.
            if (ls_k[k] & ~fv) marker |= is_marker(stg[a15 + lsoff_k[k]]);
            o0[k] = g0 & fv; o1[k] = f1; ov[k] = fv;
            const uint32_t lo2 = fv & ~f1, hi2 = fv & f1;   // codes 0/1 (A, C) and 2/3 (T, G)
            sA[k] |= lo2 & ~g0; sC[k] |= lo2 & g0; sT[k] |= hi2 & ~g0; sG[k] |= hi2 & g0;
            sInv[k] |= ~fv & cm;
        }
        mbar_arrive(empty_bar_at(uint32_t(ring)));                       // (the checks above were this thread's last look at the staged text)
        if (owns) store_words(o0, o1, ov);
        next_sample();
        if (irregular) atomicOr(&flags[i], 1);
        if (marker) atomicOr(&flags[i], 2);
        ++k_act;
        if (++ring == uint32_t(kRing)) { ring = 0; ring_phase ^= 1u; }
    }
    if (state && owns && s_hi > s_lo) {
#pragma unroll
        for (int k = 0; k < WPT; ++k) {
            if (c + 32u * k >= ce) continue;
            uint32_t *d = state + w + k;
            if (gridDim.y == 1) {
                d[0] |= sA[k]; d[words] |= sC[k]; d[2 * words] |= sG[k]; d[3 * words] |= sT[k]; d[4 * words] |= sInv[k];
            } else {
                if (sA[k]) atomicOr(d, sA[k]);
                if (sC[k]) atomicOr(d + words, sC[k]);
                if (sG[k]) atomicOr(d + 2 * words, sG[k]);
                if (sT[k]) atomicOr(d + 3 * words, sT[k]);
                if (sInv[k]) atomicOr(d + 4 * words, sInv[k]);
            }
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
    const int wpt = g_pack_wpt == 4 ? 4 : (g_pack_wpt == 1 ? 1 : 2);
    const int64_t cta_words = int64_t(kPackThreadsB) * wpt;
    const int64_t gx = (words + cta_words - 1) / cta_words;
    // sample groups: about 14 CTAs per SM in all (four are resident at a time; measured best on 48 records of
    // 4.35 Mb: 3.3 TB/s with 8 groups against 2.4 with 1), but no fewer than four samples per CTA -- its set-up
    // and the merge of its column state (one coalesced atomicOr per state word and group) have to be worth it
    int sms = 148;
    { int dev = 0; if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
    int64_t gy = std::max<int64_t>(1, std::min<int64_t>((int64_t(sms) * 14 + gx - 1) / gx, n_chunk / 4));
    if (g_pack_gy > 0) gy = g_pack_gy;
    gy = std::max<int64_t>(1, std::min<int64_t>(gy, std::min<int64_t>(n_chunk, 65535)));
    const dim3 grid((unsigned)gx, (unsigned)gy);
#define BTB_PACK_LAUNCH(W_)                                                                                                     \
    do {                                                                                                                        \
        BTB_CUDA(cudaFuncSetAttribute(pack_planes_kernel<W_>, cudaFuncAttributeMaxDynamicSharedMemorySize, PackCfg<W_>::kSmemBytes)); \
        pack_planes_kernel<W_><<<grid, kPackThreads, PackCfg<W_>::kSmemBytes, stream>>>(                                        \
            d_text, text_bytes, d_rec_off, d_line_w, d_eol, n_chunk, first, n_total, G, words, d_planes, d_flags, d_colstate);  \
    } while (0)
    if (wpt == 4) BTB_PACK_LAUNCH(4); else if (wpt == 1) BTB_PACK_LAUNCH(1); else BTB_PACK_LAUNCH(2);
#undef BTB_PACK_LAUNCH
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
