// fp_host.h -- F_p (secp256k1 base field) on the HOST, 4 x 64-bit limbs: just enough to turn an XYZZ point into its canonical
// affine form.  The IPA prover's rounds end in the host's Fiat-Shamir hash anyway (inner_product_prover.py:102-106), and one
// modular inversion is ~2 us on a CPU core against ~45 us as a lone GPU thread at the end of a latency chain, so the round
// kernels hand L and R over in XYZZ coordinates and the host finishes them.  Same canonical (x, y) as ec.cuh:xyzz_to_affine.
#pragma once
#include <cstdint>
#include <cstring>

namespace bp {
namespace fph {
typedef unsigned __int128 u128;
static const uint64_t P64[4] = {0xFFFFFFFEFFFFFC2FULL, 0xFFFFFFFFFFFFFFFFULL, 0xFFFFFFFFFFFFFFFFULL, 0xFFFFFFFFFFFFFFFFULL};
static const uint64_t PC = 0x1000003D1ULL;           // 2^256 = PC (mod p)

inline bool geq_p(const uint64_t* x) { for (int i = 3; i >= 0; i--) { if (x[i] != P64[i]) return x[i] > P64[i]; } return true; }
inline void canon(uint64_t* x) {                     // x < 2^256 -> [0, p)
  if (!geq_p(x)) return;
  u128 bw = 0;
  for (int i = 0; i < 4; i++) { u128 d = (u128)x[i] - P64[i] - (uint64_t)bw; x[i] = (uint64_t)d; bw = (d >> 64) & 1; }
}
inline bool is_zero(const uint64_t* x) { return !(x[0] | x[1] | x[2] | x[3]); }
// r = a * b mod p, canonical
inline void mul(uint64_t r[4], const uint64_t a[4], const uint64_t b[4]) {
  uint64_t t[8] = {0};
  for (int i = 0; i < 4; i++) {
    u128 c = 0;
    for (int j = 0; j < 4; j++) { c += (u128)a[i] * b[j] + t[i + j]; t[i + j] = (uint64_t)c; c >>= 64; }
    t[i + 4] = (uint64_t)c;
  }
  // fold the high half: hi * 2^256 = hi * PC
  uint64_t s[5];
  u128 c = 0;
  for (int i = 0; i < 4; i++) { c += (u128)t[4 + i] * PC + t[i]; s[i] = (uint64_t)c; c >>= 64; }
  s[4] = (uint64_t)c;                                // < 2^34
  c = (u128)s[4] * PC;
  for (int i = 0; i < 4; i++) { c += s[i]; r[i] = (uint64_t)c; c >>= 64; }
  if ((uint64_t)c) {                                 // one more wrap (the sum is then tiny)
    c = PC;
    for (int i = 0; i < 4; i++) { c += r[i]; r[i] = (uint64_t)c; c >>= 64; }
  }
  canon(r);
}
// a^-1 mod p (a canonical, non-zero) by the binary extended Euclidean algorithm
inline void inv(uint64_t r[4], const uint64_t a[4]) {
  auto is_one = [](const uint64_t* x) { return x[0] == 1 && !(x[1] | x[2] | x[3]); };
  auto shr1 = [](uint64_t* x, uint64_t top) { x[0] = x[0] >> 1 | x[1] << 63; x[1] = x[1] >> 1 | x[2] << 63; x[2] = x[2] >> 1 | x[3] << 63; x[3] = x[3] >> 1 | top << 63; };
  auto add = [](uint64_t* x, const uint64_t* y) { u128 c = 0; for (int i = 0; i < 4; i++) { c += (u128)x[i] + y[i]; x[i] = (uint64_t)c; c >>= 64; } return (uint64_t)c; };
  auto sub = [](uint64_t* x, const uint64_t* y) { u128 bw = 0; for (int i = 0; i < 4; i++) { u128 d = (u128)x[i] - y[i] - (uint64_t)bw; x[i] = (uint64_t)d; bw = (d >> 64) & 1; } return (uint64_t)bw; };
  auto geq = [](const uint64_t* x, const uint64_t* y) { for (int i = 3; i >= 0; i--) { if (x[i] != y[i]) return x[i] > y[i]; } return true; };
  auto half = [&](uint64_t* x) { uint64_t top = 0; if (x[0] & 1) top = add(x, P64); shr1(x, top); };
  uint64_t u[4] = {a[0], a[1], a[2], a[3]}, v[4] = {P64[0], P64[1], P64[2], P64[3]}, x1[4] = {1, 0, 0, 0}, x2[4] = {0, 0, 0, 0};
  memset(r, 0, 32);
  if (is_zero(u)) return;
  while (!is_one(u) && !is_one(v)) {
    while (!(u[0] & 1)) { shr1(u, 0); half(x1); }
    while (!(v[0] & 1)) { shr1(v, 0); half(x2); }
    if (geq(u, v)) { sub(u, v); if (sub(x1, x2)) add(x1, P64); }
    else { sub(v, u); if (sub(x2, x1)) add(x2, P64); }
  }
  memcpy(r, is_one(u) ? x1 : x2, 32);
}
inline void add(uint64_t r[4], const uint64_t a[4], const uint64_t b[4]) {      // canonical operands and result
  u128 c = 0;
  for (int i = 0; i < 4; i++) { c += (u128)a[i] + b[i]; r[i] = (uint64_t)c; c >>= 64; }
  if ((uint64_t)c) { c = PC; for (int i = 0; i < 4; i++) { c += r[i]; r[i] = (uint64_t)c; c >>= 64; } }      // - 2^256 + PC = - p
  canon(r);
}
inline void sub(uint64_t r[4], const uint64_t a[4], const uint64_t b[4]) {
  u128 bw = 0;
  for (int i = 0; i < 4; i++) { u128 d = (u128)a[i] - b[i] - (uint64_t)bw; r[i] = (uint64_t)d; bw = (d >> 64) & 1; }
  if ((uint64_t)bw) { u128 c = 0; for (int i = 0; i < 4; i++) { c += (u128)r[i] + P64[i]; r[i] = (uint64_t)c; c >>= 64; } }
}
inline void dbl(uint64_t r[4], const uint64_t a[4]) { add(r, a, a); }

// Jacobian points (x = X/Z^2, y = Y/Z^3; Z = 0: the identity) for the serial tails the host takes over from the device
struct JacH { uint64_t X[4], Y[4], Z[4]; };
inline void jac_from_xyzz(JacH& r, const uint8_t* xyzz) {      // (X, Y, ZZ, ZZZ) -> (X*ZZ, Y*ZZZ, ZZ)   [ZZ^3 = ZZZ^2]
  uint64_t X[4], Y[4], ZZ[4], ZZZ[4];
  memcpy(X, xyzz, 32); canon(X); memcpy(Y, xyzz + 32, 32); canon(Y);
  memcpy(ZZ, xyzz + 64, 32); canon(ZZ); memcpy(ZZZ, xyzz + 96, 32); canon(ZZZ);
  mul(r.X, X, ZZ); mul(r.Y, Y, ZZZ); memcpy(r.Z, ZZ, 32);
  if (is_zero(ZZZ)) memset(r.Z, 0, 32);
}
inline void jac_dbl(JacH& p) {                                  // dbl-2009-l (a = 0): 2M + 5S; the identity stays the identity
  uint64_t A[4], B[4], C[4], D[4], E[4], F[4], t[4];
  mul(A, p.X, p.X); mul(B, p.Y, p.Y); mul(C, B, B);
  add(t, p.X, B); mul(t, t, t); sub(t, t, A); sub(t, t, C); dbl(D, t);
  dbl(E, A); add(E, E, A);
  mul(F, E, E);
  mul(p.Z, p.Y, p.Z); dbl(p.Z, p.Z);
  dbl(t, D); sub(p.X, F, t);
  sub(t, D, p.X); mul(t, E, t);
  dbl(C, C); dbl(C, C); dbl(C, C);
  sub(p.Y, t, C);
}
inline void jac_add(JacH& p, const JacH& q) {                    // complete: identities, P + P, P - P
  if (is_zero(q.Z)) return;
  if (is_zero(p.Z)) { p = q; return; }
  uint64_t Z1Z1[4], Z2Z2[4], U1[4], U2[4], S1[4], S2[4], H[4], R[4], HH[4], HHH[4], V[4], t[4];
  mul(Z1Z1, p.Z, p.Z); mul(Z2Z2, q.Z, q.Z);
  mul(U1, p.X, Z2Z2); mul(U2, q.X, Z1Z1);
  mul(S1, p.Y, q.Z); mul(S1, S1, Z2Z2);
  mul(S2, q.Y, p.Z); mul(S2, S2, Z1Z1);
  sub(H, U2, U1); sub(R, S2, S1);
  if (is_zero(H)) {
    if (is_zero(R)) { jac_dbl(p); return; }
    memset(&p, 0, sizeof(p)); return;
  }
  mul(HH, H, H); mul(HHH, H, HH); mul(V, U1, HH);
  mul(t, p.Z, q.Z); mul(p.Z, t, H);
  mul(p.X, R, R); sub(p.X, p.X, HHH); dbl(t, V); sub(p.X, p.X, t);
  sub(t, V, p.X); mul(t, R, t); mul(S1, S1, HHH); sub(p.Y, t, S1);
}
inline void jac_to_affine(const JacH& p, uint8_t out64[64]) {   // canonical (x, y); the identity becomes 64 zero bytes
  if (is_zero(p.Z)) { memset(out64, 0, 64); return; }
  uint64_t zi[4], zi2[4], x[4], y[4];
  inv(zi, p.Z); mul(zi2, zi, zi);
  mul(x, p.X, zi2); mul(zi2, zi2, zi); mul(y, p.Y, zi2);
  memcpy(out64, x, 32); memcpy(out64 + 32, y, 32);
}
}  // namespace fph

// sum of `count` canonical affine points (64 bytes each, 64 zero bytes = the identity) -> canonical affine: the partial results of
// the ranks of a sharded host-result MSM, added on every rank in the same order
inline void affine_sum_host(const uint8_t* pts64, size_t count, uint8_t out64[64]) {
  using namespace fph;
  JacH acc; memset(&acc, 0, sizeof(acc));
  for (size_t i = 0; i < count; i++) {
    JacH q; memset(&q, 0, sizeof(q));
    memcpy(q.X, pts64 + 64 * i, 32); memcpy(q.Y, pts64 + 64 * i + 32, 32);
    if (is_zero(q.X) && is_zero(q.Y)) continue;
    q.Z[0] = 1;
    jac_add(acc, q);
  }
  jac_to_affine(acc, out64);
}

// Horner over the window sums of ONE MSM, on the host (k_combine's job: /root/reference/src/pippenger/pippenger.py:56-60, the loop
// that squares c times and multiplies the next window in): winsum = U XYZZ points (128 bytes each) as the bucket reduction left
// them, unit U-1 (and U-2 when `dbl`) the top window.  112 dependent doublings at c = 16: ~0.23 ms as a 4-lane chain on the
// device (1.4 us per doubling at 2 GHz with 32-bit multipliers), ~30 us on a host core -- which the result is headed for anyway.
inline void horner_host(const uint8_t* winsum, int c, int W, int U, int dblunits, uint8_t out64[64]) {
  using namespace fph;
  JacH acc, v;
  jac_from_xyzz(acc, winsum + 128 * (size_t)(U - 1));
  if (dblunits) { jac_from_xyzz(v, winsum + 128 * (size_t)(U - 2)); jac_add(acc, v); }
  for (int w = W - 2; w >= 0; w--) {
    for (int d = 0; d < c; d++) jac_dbl(acc);
    jac_from_xyzz(v, winsum + 128 * (size_t)w);
    jac_add(acc, v);
  }
  jac_to_affine(acc, out64);
}

// count XYZZ points (128 little-endian bytes each: X, Y, ZZ, ZZZ, lazy residues < 2^256) -> canonical affine (64 bytes each);
// the identity (ZZ = 0 mod p) becomes 64 zero bytes.  One shared inversion (Montgomery's trick over the ZZZ values).
inline void xyzz_to_affine_host(const uint8_t* xyzz, size_t count, uint8_t* out64) {
  using namespace fph;
  uint64_t pre[8][4], zzz[8][4];
  bool id[8];
  uint64_t acc[4] = {1, 0, 0, 0};
  for (size_t i = 0; i < count && i < 8; i++) {
    uint64_t zz[4];
    memcpy(zz, xyzz + 128 * i + 64, 32); canon(zz);
    memcpy(zzz[i], xyzz + 128 * i + 96, 32); canon(zzz[i]);
    id[i] = is_zero(zz) || is_zero(zzz[i]);
    memcpy(pre[i], acc, 32);
    if (!id[i]) mul(acc, acc, zzz[i]);
  }
  uint64_t inv_all[4];
  inv(inv_all, acc);
  for (size_t k = count > 8 ? 8 : count; k-- > 0;) {
    uint8_t* o = out64 + 64 * k;
    if (id[k]) { memset(o, 0, 64); continue; }
    uint64_t zi3[4], zi2[4], X[4], Y[4], ZZ[4], t[4];
    mul(zi3, inv_all, pre[k]);                       // ZZZ_k^-1
    mul(inv_all, inv_all, zzz[k]);
    memcpy(X, xyzz + 128 * k, 32); canon(X);
    memcpy(Y, xyzz + 128 * k + 32, 32); canon(Y);
    memcpy(ZZ, xyzz + 128 * k + 64, 32); canon(ZZ);
    mul(t, ZZ, ZZ); mul(zi2, zi3, zi3); mul(zi2, zi2, t);      // ZZ^-1 = ZZ^2 * ZZZ^-2   (ZZ^3 = ZZZ^2)
    mul(X, X, zi2); mul(Y, Y, zi3);
    memcpy(o, X, 32); memcpy(o + 32, Y, 32);
  }
}
}  // namespace bp
