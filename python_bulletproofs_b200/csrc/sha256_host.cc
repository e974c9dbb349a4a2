// sha256_host.cc -- SHA-256 compression function for the host-side Fiat-Shamir code (transcript.h).
// FIPS 180-4; two implementations with a run-time switch: portable C++, and the x86 SHA extensions
// (SHA-NI: sha256rnds2 / sha256msg1 / sha256msg2) when the CPU reports them (CPUID.7:EBX bit 29).
// Replaces hashlib.sha256 as used by mod_hash (/root/reference/src/utils/utils.py:84-97).
#include <cstddef>
#include <cstdint>
#include <cstring>
#if defined(__x86_64__)
#include <cpuid.h>
#include <immintrin.h>
#endif

namespace bp {

static const uint32_t K256[64] = {
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01,
    0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc,
    0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147,
    0x06ca6351, 0x14292967, 0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85,
    0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08,
    0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208,
    0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};

static inline uint32_t rotr(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }

static void blocks_portable(uint32_t h[8], const uint8_t* p, size_t nblocks) {
  for (; nblocks; nblocks--, p += 64) {
    uint32_t w[64];
    for (int i = 0; i < 16; i++) w[i] = (uint32_t)p[4 * i] << 24 | (uint32_t)p[4 * i + 1] << 16 | (uint32_t)p[4 * i + 2] << 8 | p[4 * i + 3];
    for (int i = 16; i < 64; i++) {
      uint32_t s0 = rotr(w[i - 15], 7) ^ rotr(w[i - 15], 18) ^ (w[i - 15] >> 3);
      uint32_t s1 = rotr(w[i - 2], 17) ^ rotr(w[i - 2], 19) ^ (w[i - 2] >> 10);
      w[i] = w[i - 16] + s0 + w[i - 7] + s1;
    }
    uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
    for (int i = 0; i < 64; i++) {
      uint32_t S1 = rotr(e, 6) ^ rotr(e, 11) ^ rotr(e, 25), ch = (e & f) ^ (~e & g);
      uint32_t t1 = hh + S1 + ch + K256[i] + w[i];
      uint32_t S0 = rotr(a, 2) ^ rotr(a, 13) ^ rotr(a, 22), mj = (a & b) ^ (a & c) ^ (b & c);
      uint32_t t2 = S0 + mj;
      hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
    }
    h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
  }
}

#if defined(__x86_64__)
__attribute__((target("sha,sse4.1,ssse3"))) static void blocks_shani(uint32_t h[8], const uint8_t* p, size_t nblocks) {
  const __m128i bswap = _mm_set_epi64x(0x0c0d0e0f08090a0bULL, 0x0405060700010203ULL);
  __m128i tmp = _mm_loadu_si128((const __m128i*)&h[0]);       // DCBA
  __m128i st1 = _mm_loadu_si128((const __m128i*)&h[4]);       // HGFE
  tmp = _mm_shuffle_epi32(tmp, 0xB1);                         // CDAB
  st1 = _mm_shuffle_epi32(st1, 0x1B);                         // EFGH
  __m128i st0 = _mm_alignr_epi8(tmp, st1, 8);                 // ABEF
  st1 = _mm_blend_epi16(st1, tmp, 0xF0);                      // CDGH
  for (; nblocks; nblocks--, p += 64) {
    const __m128i save0 = st0, save1 = st1;
    __m128i m[4];
    for (int g = 0; g < 16; g++) {                            // rounds 4g .. 4g+3
      if (g < 4) m[g] = _mm_shuffle_epi8(_mm_loadu_si128((const __m128i*)(p + 16 * g)), bswap);
      else {
        // W[4g..4g+3] = msg2( msg1(W[4g-16..], W[4g-12..]) + W[4g-7..4g-4], W[4g-4..] )
        __m128i x = _mm_sha256msg1_epu32(m[(g - 4) & 3], m[(g - 3) & 3]);
        x = _mm_add_epi32(x, _mm_alignr_epi8(m[(g - 1) & 3], m[(g - 2) & 3], 4));
        m[g & 3] = _mm_sha256msg2_epu32(x, m[(g - 1) & 3]);
      }
      __m128i msg = _mm_add_epi32(m[g & 3], _mm_loadu_si128((const __m128i*)&K256[4 * g]));
      st1 = _mm_sha256rnds2_epu32(st1, st0, msg);
      msg = _mm_shuffle_epi32(msg, 0x0E);
      st0 = _mm_sha256rnds2_epu32(st0, st1, msg);
    }
    st0 = _mm_add_epi32(st0, save0);
    st1 = _mm_add_epi32(st1, save1);
  }
  tmp = _mm_shuffle_epi32(st0, 0x1B);                         // FEBA
  st1 = _mm_shuffle_epi32(st1, 0xB1);                         // DCHG
  st0 = _mm_blend_epi16(tmp, st1, 0xF0);                      // DCBA
  st1 = _mm_alignr_epi8(st1, tmp, 8);                         // HGFE
  _mm_storeu_si128((__m128i*)&h[0], st0);
  _mm_storeu_si128((__m128i*)&h[4], st1);
}
static bool cpu_has_shani() {
  unsigned a, b, c, d;
  if (!__get_cpuid_count(7, 0, &a, &b, &c, &d)) return false;
  bool sha = (b >> 29) & 1;
  if (!__get_cpuid(1, &a, &b, &c, &d)) return false;
  bool sse41 = (c >> 19) & 1, ssse3 = (c >> 9) & 1;
  return sha && sse41 && ssse3;
}
#endif

static int g_impl = -1;   // -1 unknown, 0 portable, 1 SHA-NI

void sha256_blocks(uint32_t h[8], const uint8_t* p, size_t nblocks) {
#if defined(__x86_64__)
  if (g_impl < 0) g_impl = cpu_has_shani() ? 1 : 0;
  if (g_impl == 1) { blocks_shani(h, p, nblocks); return; }
#endif
  blocks_portable(h, p, nblocks);
}
void sha256_force_portable(int on) { g_impl = on ? 0 : -1; }
int sha256_impl() {
#if defined(__x86_64__)
  if (g_impl < 0) g_impl = cpu_has_shani() ? 1 : 0;
#else
  g_impl = 0;
#endif
  return g_impl;
}

}  // namespace bp
