// transcript.h -- host-side Fiat-Shamir helpers, restated in C++ so the sequential challenge
// derivation can run inside the library between kernel launches (no Python in the round loop).
//
// Replaces (bit-exactly): mod_hash, point_to_bytes, point_to_b64
// (/root/reference/src/utils/utils.py:84-111) and Transcript.add_point / add_number / get_modp
// (/root/reference/src/utils/transcript.py:16-33).  hashlib's SHA-256 is restated from FIPS 180-4.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include "fq.cuh"

namespace bp {

// ---- SHA-256 (FIPS 180-4) ----------------------------------------------------------------------
struct Sha256 {
  uint32_t h[8];
  uint8_t buf[64];
  uint64_t len = 0;
  size_t fill = 0;
  Sha256() {
    static const uint32_t iv[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
    memcpy(h, iv, sizeof h);
  }
  static uint32_t rotr(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }
  void block(const uint8_t* p) {
    static const uint32_t K[64] = {
        0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01,
        0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc,
        0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147,
        0x06ca6351, 0x14292967, 0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85,
        0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08,
        0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208,
        0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};
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
      uint32_t t1 = hh + S1 + ch + K[i] + w[i];
      uint32_t S0 = rotr(a, 2) ^ rotr(a, 13) ^ rotr(a, 22), mj = (a & b) ^ (a & c) ^ (b & c);
      uint32_t t2 = S0 + mj;
      hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
    }
    h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
  }
  void update(const uint8_t* p, size_t n) {
    len += n;
    while (n) {
      if (fill == 0 && n >= 64) { block(p); p += 64; n -= 64; continue; }
      size_t k = 64 - fill < n ? 64 - fill : n;
      memcpy(buf + fill, p, k); fill += k; p += k; n -= k;
      if (fill == 64) { block(buf); fill = 0; }
    }
  }
  void final(uint8_t out[32]) {
    uint64_t bits = len * 8;
    uint8_t pad = 0x80;
    update(&pad, 1);
    uint8_t z = 0;
    while (fill != 56) update(&z, 1);
    uint8_t lb[8];
    for (int i = 0; i < 8; i++) lb[i] = (uint8_t)(bits >> (56 - 8 * i));
    update(lb, 8);
    for (int i = 0; i < 8; i++) { out[4 * i] = h[i] >> 24; out[4 * i + 1] = h[i] >> 16; out[4 * i + 2] = h[i] >> 8; out[4 * i + 3] = h[i]; }
  }
};

// mod_hash(msg, q): utils.py:84-97.  q is 256 bits so the `% 2**bit_length` mask is a no-op.
inline Fq mod_hash_q(const uint8_t* msg, size_t len) {
  for (unsigned long ctr = 1;; ctr++) {
    char pre[24];
    int pl = snprintf(pre, sizeof pre, "%lu", ctr);
    Sha256 s;
    s.update((const uint8_t*)pre, (size_t)pl);
    s.update(msg, len);
    uint8_t d[32];
    s.final(d);
    Fq x;
    for (int i = 0; i < 8; i++)   // digest is a big-endian integer
      x.v[i] = (uint32_t)d[31 - 4 * i] | (uint32_t)d[30 - 4 * i] << 8 | (uint32_t)d[29 - 4 * i] << 16 | (uint32_t)d[28 - 4 * i] << 24;
    if (fq_geq(x, fq_const_q())) continue;      // x >= p: retry
    if (fq_is_zero(x)) continue;                // non_zero=True
    return x;
  }
}

// str(x): decimal of a 256-bit integer (transcript.py:29-31 via ModP.__str__, utils.py:77-78)
inline std::string fq_to_decimal(const Fq& a) {
  uint32_t w[8];
  memcpy(w, a.v, sizeof w);
  char tmp[80];
  int n = 0;
  bool nz = true;
  while (nz) {
    uint64_t rem = 0;
    nz = false;
    for (int i = 7; i >= 0; i--) {
      uint64_t cur = (rem << 32) | w[i];
      w[i] = (uint32_t)(cur / 1000000000u);
      rem = cur % 1000000000u;
      if (w[i]) nz = true;
    }
    for (int k = 0; k < 9; k++) { tmp[n++] = (char)('0' + rem % 10); rem /= 10; }
  }
  while (n > 1 && tmp[n - 1] == '0') n--;
  std::string s(n, '0');
  for (int i = 0; i < n; i++) s[i] = tmp[n - 1 - i];
  return s;
}

// int(slot) for a plain run of ASCII digits (any length), reduced mod q.  Returns false when the
// slot is empty or contains anything else: Python's int() has more accepting syntax (sign,
// underscores, surrounding whitespace), so the caller must then defer to it.
inline bool decimal_to_fq(const uint8_t* s, size_t n, Fq* out) {
  if (n == 0) return false;
  Fq acc = fq_zero();
  const Fq ten = fq_from_u64(10);
  for (size_t i = 0; i < n; i++) {
    if (s[i] < '0' || s[i] > '9') return false;
    acc = fq_add(fq_mul(acc, ten), fq_from_u64((uint64_t)(s[i] - '0')));
  }
  *out = acc;
  return true;
}

// point_to_b64: utils.py:100-111.  pt = 64-byte little-endian affine, identity = zeros.
inline std::string point_to_b64(const uint8_t pt[64]) {
  static const char* A = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/";
  uint8_t raw[33];
  size_t n;
  bool ident = true;
  for (int i = 0; i < 64; i++) if (pt[i]) { ident = false; break; }
  if (ident) { raw[0] = 0; n = 1; }
  else {
    raw[0] = (pt[32] & 1) ? 3 : 2;                         // parity of y (little-endian byte 0)
    for (int i = 0; i < 32; i++) raw[1 + i] = pt[31 - i];  // x big-endian
    n = 33;
  }
  std::string o;
  for (size_t i = 0; i < n; i += 3) {
    uint32_t v = (uint32_t)raw[i] << 16 | (i + 1 < n ? (uint32_t)raw[i + 1] << 8 : 0) | (i + 2 < n ? raw[i + 2] : 0);
    o += A[(v >> 18) & 63];
    o += A[(v >> 12) & 63];
    o += i + 1 < n ? A[(v >> 6) & 63] : '=';
    o += i + 2 < n ? A[v & 63] : '=';
  }
  return o;
}

inline void fq_from_le(Fq* r, const uint8_t* b) { memcpy(r->v, b, 32); }
inline void fq_to_le(uint8_t* b, const Fq& a) { memcpy(b, a.v, 32); }

}  // namespace bp
