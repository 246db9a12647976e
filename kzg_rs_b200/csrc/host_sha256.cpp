// Host SHA-256 for the batch transcript (see host_sha256.h): SHA-NI when the CPU has it, portable C++ otherwise.
#include "host_sha256.h"
#include <string.h>
#if defined(__x86_64__) || defined(__i386__)
#include <cpuid.h>
#include <immintrin.h>
#define KZGB200_X86 1
#endif

namespace kzgb200 {
namespace {

alignas(16) const uint32_t K256[64] = {
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01, 0x243185be,
    0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa,
    0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967, 0x27b70a85,
    0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85, 0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3,
    0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f,
    0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};

inline uint32_t rotr(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }

void compress_portable(uint32_t h[8], const uint8_t* p, size_t nblk) {
    for (; nblk; nblk--, p += 64) {
        uint32_t w[64];
        for (int i = 0; i < 16; i++) w[i] = ((uint32_t)p[4 * i] << 24) | ((uint32_t)p[4 * i + 1] << 16) | ((uint32_t)p[4 * i + 2] << 8) | p[4 * i + 3];
        for (int i = 16; i < 64; i++) {
            uint32_t s0 = rotr(w[i - 15], 7) ^ rotr(w[i - 15], 18) ^ (w[i - 15] >> 3);
            uint32_t s1 = rotr(w[i - 2], 17) ^ rotr(w[i - 2], 19) ^ (w[i - 2] >> 10);
            w[i] = w[i - 16] + s0 + w[i - 7] + s1;
        }
        uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
        for (int i = 0; i < 64; i++) {
            uint32_t t1 = hh + (rotr(e, 6) ^ rotr(e, 11) ^ rotr(e, 25)) + ((e & f) ^ (~e & g)) + K256[i] + w[i];
            uint32_t t2 = (rotr(a, 2) ^ rotr(a, 13) ^ rotr(a, 22)) + ((a & b) ^ (a & c) ^ (b & c));
            hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
        }
        h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
    }
}

#ifdef KZGB200_X86
// x86 SHA extensions: the state lives in two registers as (A,B,E,F) / (C,D,G,H); sha256rnds2 does two rounds, msg1 / msg2 the
// two halves of the message schedule for four words at a time.
__attribute__((target("sha,sse4.1,ssse3"))) void compress_shani(uint32_t h[8], const uint8_t* p, size_t nblk) {
    const __m128i bswap = _mm_set_epi64x(0x0c0d0e0f08090a0bULL, 0x0405060700010203ULL);
    __m128i t = _mm_shuffle_epi32(_mm_loadu_si128((const __m128i*)&h[0]), 0xB1);     // C D A B  (high .. low: B A D C)
    __m128i s1 = _mm_shuffle_epi32(_mm_loadu_si128((const __m128i*)&h[4]), 0x1B);    // E F G H reversed
    __m128i s0 = _mm_alignr_epi8(t, s1, 8);                                           // ABEF
    s1 = _mm_blend_epi16(s1, t, 0xF0);                                                // CDGH
    for (; nblk; nblk--, p += 64) {
        const __m128i save0 = s0, save1 = s1;
        __m128i m[4];
        for (int i = 0; i < 4; i++) m[i] = _mm_shuffle_epi8(_mm_loadu_si128((const __m128i*)(p + 16 * i)), bswap);
#pragma GCC unroll 16
        for (int r = 0; r < 16; r++) {
            __m128i wk = _mm_add_epi32(m[r & 3], _mm_load_si128((const __m128i*)&K256[4 * r]));
            s1 = _mm_sha256rnds2_epu32(s1, s0, wk);
            s0 = _mm_sha256rnds2_epu32(s0, s1, _mm_shuffle_epi32(wk, 0x0E));
            if (r < 12) {   // slot r&3 (words 4r..4r+3) <- words 4r+16..4r+19
                __m128i x = _mm_sha256msg1_epu32(m[r & 3], m[(r + 1) & 3]);
                x = _mm_add_epi32(x, _mm_alignr_epi8(m[(r + 3) & 3], m[(r + 2) & 3], 4));
                m[r & 3] = _mm_sha256msg2_epu32(x, m[(r + 3) & 3]);
            }
        }
        s0 = _mm_add_epi32(s0, save0);
        s1 = _mm_add_epi32(s1, save1);
    }
    t = _mm_shuffle_epi32(s0, 0x1B);          // F E B A
    s1 = _mm_shuffle_epi32(s1, 0xB1);         // D C H G
    _mm_storeu_si128((__m128i*)&h[0], _mm_blend_epi16(t, s1, 0xF0));      // A B C D
    _mm_storeu_si128((__m128i*)&h[4], _mm_alignr_epi8(s1, t, 8));         // E F G H
}
bool cpu_has_shani() {
    unsigned a = 0, b = 0, c = 0, d = 0;
    if (!__get_cpuid_count(7, 0, &a, &b, &c, &d)) return false;
    bool sha = (b >> 29) & 1;
    if (!__get_cpuid(1, &a, &b, &c, &d)) return false;
    return sha && ((c >> 19) & 1) /* sse4.1 */ && ((c >> 9) & 1) /* ssse3 */;
}
#else
bool cpu_has_shani() { return false; }
#endif

int g_force_portable = 0;
bool use_shani() {
    static const bool has = cpu_has_shani();
    return has && !g_force_portable;
}
void compress(uint32_t h[8], const uint8_t* p, size_t nblk) {
#ifdef KZGB200_X86
    if (use_shani()) return compress_shani(h, p, nblk);
#endif
    compress_portable(h, p, nblk);
}

}  // namespace

void host_sha256_init(HostSha256* s) {
    static const uint32_t iv[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
    memcpy(s->h, iv, sizeof(iv));
    s->fill = 0;
    s->total = 0;
}
void host_sha256_update(HostSha256* s, const uint8_t* data, size_t len) {
    s->total += len;
    if (s->fill) {
        size_t take = 64 - s->fill < len ? 64 - s->fill : len;
        memcpy(s->buf + s->fill, data, take);
        s->fill += (uint32_t)take; data += take; len -= take;
        if (s->fill < 64) return;
        compress(s->h, s->buf, 1);
        s->fill = 0;
    }
    if (len >= 64) { compress(s->h, data, len / 64); data += len & ~(size_t)63; len &= 63; }
    if (len) { memcpy(s->buf, data, len); s->fill = (uint32_t)len; }
}
void host_sha256_final(HostSha256* s, uint8_t out[32]) {
    uint64_t bits = s->total * 8;
    uint8_t pad[72] = {0x80};
    size_t padlen = (s->fill < 56 ? 56 : 120) - s->fill;
    for (int i = 0; i < 8; i++) pad[padlen + i] = (uint8_t)(bits >> (56 - 8 * i));
    host_sha256_update(s, pad, padlen + 8);
    for (int i = 0; i < 8; i++) { out[4 * i] = (uint8_t)(s->h[i] >> 24); out[4 * i + 1] = (uint8_t)(s->h[i] >> 16); out[4 * i + 2] = (uint8_t)(s->h[i] >> 8); out[4 * i + 3] = (uint8_t)s->h[i]; }
}
int host_sha256_uses_shani() { return use_shani() ? 1 : 0; }
void host_sha256_force_portable(int on) { g_force_portable = on; }

}  // namespace kzgb200
