// host_mirror_exp: how fast can the host transpose tile-packed 256x256 uint16 tiles into a row-major matrix?
// (decides whether shipping only the upper triangle over PCIe and mirroring on the host can beat the full D2H)
//   host_mirror_exp <n> <threads> <mode 0: 32x8 blocks streamed, 1: whole tile via a buffer, 2: 1 + copy the upper tile too>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>
#include <chrono>
#include <atomic>
#include <immintrin.h>
static inline void tr8x8(const uint16_t *src, size_t sld, uint16_t *dst, size_t dld) {
    __m128i r[8];
    for (int i = 0; i < 8; ++i) r[i] = _mm_loadu_si128((const __m128i *)(src + i * sld));
    __m128i t0 = _mm_unpacklo_epi16(r[0], r[1]), t1 = _mm_unpackhi_epi16(r[0], r[1]);
    __m128i t2 = _mm_unpacklo_epi16(r[2], r[3]), t3 = _mm_unpackhi_epi16(r[2], r[3]);
    __m128i t4 = _mm_unpacklo_epi16(r[4], r[5]), t5 = _mm_unpackhi_epi16(r[4], r[5]);
    __m128i t6 = _mm_unpacklo_epi16(r[6], r[7]), t7 = _mm_unpackhi_epi16(r[6], r[7]);
    __m128i u0 = _mm_unpacklo_epi32(t0, t2), u1 = _mm_unpackhi_epi32(t0, t2);
    __m128i u2 = _mm_unpacklo_epi32(t1, t3), u3 = _mm_unpackhi_epi32(t1, t3);
    __m128i u4 = _mm_unpacklo_epi32(t4, t6), u5 = _mm_unpackhi_epi32(t4, t6);
    __m128i u6 = _mm_unpacklo_epi32(t5, t7), u7 = _mm_unpackhi_epi32(t5, t7);
    _mm_storeu_si128((__m128i *)(dst + 0 * dld), _mm_unpacklo_epi64(u0, u4));
    _mm_storeu_si128((__m128i *)(dst + 1 * dld), _mm_unpackhi_epi64(u0, u4));
    _mm_storeu_si128((__m128i *)(dst + 2 * dld), _mm_unpacklo_epi64(u1, u5));
    _mm_storeu_si128((__m128i *)(dst + 3 * dld), _mm_unpackhi_epi64(u1, u5));
    _mm_storeu_si128((__m128i *)(dst + 4 * dld), _mm_unpacklo_epi64(u2, u6));
    _mm_storeu_si128((__m128i *)(dst + 5 * dld), _mm_unpackhi_epi64(u2, u6));
    _mm_storeu_si128((__m128i *)(dst + 6 * dld), _mm_unpacklo_epi64(u3, u7));
    _mm_storeu_si128((__m128i *)(dst + 7 * dld), _mm_unpackhi_epi64(u3, u7));
}
int main(int argc, char **argv) {
    size_t n = argc > 1 ? atol(argv[1]) : 32768; int T = argc > 2 ? atoi(argv[2]) : 8; int mode = argc > 3 ? atoi(argv[3]) : 0;
    size_t ld = n, nb = n / 256, tiles = nb * (nb - 1) / 2;
    uint16_t *m = (uint16_t *)aligned_alloc(4096, n * ld * 2);
    uint16_t *tp = (uint16_t *)aligned_alloc(4096, tiles * 65536 * 2);
    { std::vector<std::thread> p; for (int t = 0; t < T; ++t) p.emplace_back([&, t] { size_t r0 = n * t / T, r1 = n * (t + 1) / T; for (size_t r = r0; r < r1; ++r) for (size_t c = 0; c < n; ++c) m[r * ld + c] = 0;
        for (size_t k = tiles * t / T; k < tiles * (t + 1) / T; ++k) for (size_t e = 0; e < 65536; ++e) tp[k * 65536 + e] = uint16_t(k * 7 + e); }); for (auto &th : p) th.join(); }
    auto t0 = std::chrono::steady_clock::now();
    std::atomic<size_t> next{0};
    std::vector<std::thread> p;
    for (int t = 0; t < T; ++t) p.emplace_back([&] {
        for (;;) { size_t k = next.fetch_add(1); if (k >= tiles) break;
            // tile k = (I, J), I < J: enumerate
            size_t I = 0, rem = k; while (rem >= nb - 1 - I) { rem -= nb - 1 - I; ++I; } size_t J = I + 1 + rem;
            const uint16_t *src = tp + k * 65536;
            if (mode == 0) {
              for (size_t j = 0; j < 256; j += 8) for (size_t i = 0; i < 256; i += 32) {
                alignas(64) uint16_t buf[8][32];
                for (int q = 0; q < 4; ++q) tr8x8(src + (i + 8 * q) * 256 + j, 256, &buf[0][8 * q], 32);
                for (int r = 0; r < 8; ++r) { uint16_t *d = m + (J * 256 + j + r) * ld + I * 256 + i;
                    _mm256_stream_si256((__m256i *)d, _mm256_load_si256((const __m256i *)&buf[r][0]));
                    _mm256_stream_si256((__m256i *)(d + 16), _mm256_load_si256((const __m256i *)&buf[r][16])); } }
            } else {
              // transpose whole tile into a local buffer, then copy rows out (memcpy 512 B each) + also the upper tile rows
              alignas(64) static thread_local uint16_t buf[256 * 256];
              for (size_t i = 0; i < 256; i += 8) for (size_t j = 0; j < 256; j += 8) tr8x8(src + i * 256 + j, 256, buf + j * 256 + i, 256);
              for (size_t r = 0; r < 256; ++r) { uint16_t *d = m + (J * 256 + r) * ld + I * 256; const uint16_t *s = buf + r * 256;
                  for (int q = 0; q < 256; q += 16) _mm256_stream_si256((__m256i *)(d + q), _mm256_load_si256((const __m256i *)(s + q))); }
              if (mode == 2) for (size_t r = 0; r < 256; ++r) { uint16_t *d = m + (I * 256 + r) * ld + J * 256; const uint16_t *s = src + r * 256;
                  for (int q = 0; q < 256; q += 16) _mm256_stream_si256((__m256i *)(d + q), _mm256_load_si256((const __m256i *)(s + q))); }
            } } });
    for (auto &th : p) th.join();
    double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    printf("n=%zu threads=%d mode=%d: %.3f s -> %.1f GB/s of tile bytes (%.2f GB)\n", n, T, mode, dt, tiles * 131072.0 / dt / 1e9, tiles * 131072.0 / 1e9);
}
