// ec.cuh -- secp256k1 (y^2 = x^3 + 7) group law on the device, complete (exceptional cases
// handled), in extended Jacobian "XYZZ" coordinates: x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2,
// identity <=> ZZ == 0.
//
// Replaces: fastecdsa Point.__add__ / Point.__mul__ as used by EC.mult
// (/root/reference/src/pippenger/group.py:27-32) and ModP.__mul__(Point)
// (/root/reference/src/utils/utils.py:43-44).
//
// Formulas: EFD "xyzz" set for short Weierstrass curves (madd-2008-s 8M+2S, add-2008-s 12M+2S,
// dbl-2008-s-1 6M+3S with a = 0, mdbl-2008-s 3M+3S... counted with S = M in the roofline).
#pragma once
#include "fp.cuh"

namespace bp {

struct __align__(16) Affine { Fp x, y; };          // identity = all-zero (not on the curve)
struct __align__(16) XYZZ { Fp X, Y, ZZ, ZZZ; };

BP_DI bool affine_is_identity(const Affine& p) {
  u32 o = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) o |= p.x.v[i] | p.y.v[i];
  return o == 0;
}
BP_DI bool xyzz_is_identity(const XYZZ& p) { return fp_is_zero(p.ZZ); }
BP_DI XYZZ xyzz_identity() { XYZZ r; r.X = fp_zero(); r.Y = fp_zero(); r.ZZ = fp_zero(); r.ZZZ = fp_zero(); return r; }
BP_DI XYZZ xyzz_from_affine(const Affine& p) {
  XYZZ r;
  if (affine_is_identity(p)) return xyzz_identity();
  r.X = p.x; r.Y = p.y; r.ZZ = fp_one(); r.ZZZ = fp_one();
  return r;
}
BP_DI Affine affine_neg(const Affine& p) {   // (0,0) stays (0,0): -0 = 0 after canon
  Affine r; r.x = p.x; r.y = fp_canon(fp_neg(p.y)); return r;
}
BP_DI XYZZ xyzz_neg(const XYZZ& p) { XYZZ r = p; r.Y = fp_neg(p.Y); return r; }

// 2 * (affine, non-identity)
BP_DI XYZZ xyzz_mdbl(const Affine& p) {
  XYZZ r;
  Fp U = fp_dbl(p.y);
  Fp V = fp_sqr(U);
  Fp W = fp_mul(U, V);
  Fp S = fp_mul(p.x, V);
  Fp xx = fp_sqr(p.x);
  Fp M = fp_add(fp_dbl(xx), xx);
  r.X = fp_sub(fp_sqr(M), fp_dbl(S));
  r.Y = fp_sub(fp_mul(M, fp_sub(S, r.X)), fp_mul(W, p.y));
  r.ZZ = V; r.ZZZ = W;
  return r;      // y != 0 on a prime-order curve, so never the identity
}
BP_DI XYZZ xyzz_dbl(const XYZZ& p) {
  if (xyzz_is_identity(p)) return p;
  XYZZ r;
  Fp U = fp_dbl(p.Y);
  Fp V = fp_sqr(U);
  Fp W = fp_mul(U, V);
  Fp S = fp_mul(p.X, V);
  Fp xx = fp_sqr(p.X);
  Fp M = fp_add(fp_dbl(xx), xx);
  r.X = fp_sub(fp_sqr(M), fp_dbl(S));
  r.Y = fp_sub(fp_mul(M, fp_sub(S, r.X)), fp_mul(W, p.Y));
  r.ZZ = fp_mul(V, p.ZZ);
  r.ZZZ = fp_mul(W, p.ZZZ);
  return r;
}
// acc += p (p affine).  Complete: handles acc = O, p = O, p = acc, p = -acc.
BP_DI void xyzz_madd(XYZZ& a, const Affine& p) {
  if (affine_is_identity(p)) return;
  if (xyzz_is_identity(a)) { a.X = p.x; a.Y = p.y; a.ZZ = fp_one(); a.ZZZ = fp_one(); return; }
  Fp U2 = fp_mul(p.x, a.ZZ);
  Fp S2 = fp_mul(p.y, a.ZZZ);
  Fp Pd = fp_sub(U2, a.X);
  Fp R = fp_sub(S2, a.Y);
  if (fp_is_zero(Pd)) {
    if (fp_is_zero(R)) a = xyzz_mdbl(p); else a = xyzz_identity();
    return;
  }
  Fp PP = fp_sqr(Pd);
  Fp PPP = fp_mul(Pd, PP);
  Fp Qv = fp_mul(a.X, PP);
  Fp X3 = fp_sub(fp_sub(fp_sqr(R), PPP), fp_dbl(Qv));
  Fp Y3 = fp_sub(fp_mul(R, fp_sub(Qv, X3)), fp_mul(a.Y, PPP));
  a.X = X3; a.Y = Y3;
  a.ZZ = fp_mul(a.ZZ, PP);
  a.ZZZ = fp_mul(a.ZZZ, PPP);
}
// a += b (both XYZZ).  Complete.
BP_DI void xyzz_add(XYZZ& a, const XYZZ& b) {
  if (xyzz_is_identity(b)) return;
  if (xyzz_is_identity(a)) { a = b; return; }
  Fp U1 = fp_mul(a.X, b.ZZ);
  Fp U2 = fp_mul(b.X, a.ZZ);
  Fp S1 = fp_mul(a.Y, b.ZZZ);
  Fp S2 = fp_mul(b.Y, a.ZZZ);
  Fp Pd = fp_sub(U2, U1);
  Fp R = fp_sub(S2, S1);
  if (fp_is_zero(Pd)) {
    if (fp_is_zero(R)) a = xyzz_dbl(a); else a = xyzz_identity();
    return;
  }
  Fp PP = fp_sqr(Pd);
  Fp PPP = fp_mul(Pd, PP);
  Fp Qv = fp_mul(U1, PP);
  Fp X3 = fp_sub(fp_sub(fp_sqr(R), PPP), fp_dbl(Qv));
  Fp Y3 = fp_sub(fp_mul(R, fp_sub(Qv, X3)), fp_mul(S1, PPP));
  a.X = X3; a.Y = Y3;
  a.ZZ = fp_mul(fp_mul(a.ZZ, b.ZZ), PP);
  a.ZZZ = fp_mul(fp_mul(a.ZZZ, b.ZZZ), PPP);
}
// out-of-line forms for kernels that are not throughput bound (keeps their code inside the instruction caches)
__device__ __noinline__ void xyzz_add_ni(XYZZ& a, const XYZZ& b) { xyzz_add(a, b); }
__device__ __noinline__ void xyzz_madd_ni(XYZZ& a, const Affine& p) { xyzz_madd(a, p); }
__device__ __noinline__ XYZZ xyzz_dbl_ni(const XYZZ& p) { return xyzz_dbl(p); }

// canonical affine output (the parity surface): x = X*ZZ^-1, y = Y*ZZZ^-1, one inversion:
// ZZ^-1 = ZZ^2 * ZZZ^-2 because ZZ^3 = ZZZ^2.
// lone = the calling thread is (nearly) alone in its warp: take the binary-GCD inversion (2.4x quicker for one thread,
// but its data-dependent loops serialise when many lanes of a warp invert at once)
BP_DI Affine xyzz_to_affine(const XYZZ& p, bool lone = false) {
  Affine r;
  if (xyzz_is_identity(p)) { r.x = fp_zero(); r.y = fp_zero(); return r; }
  Fp zi3 = lone ? fp_inv_gcd(p.ZZZ) : fp_inv(p.ZZZ);
  Fp zi2 = fp_mul(fp_sqr(p.ZZ), fp_sqr(zi3));
  r.x = fp_canon(fp_mul(p.X, zi2));
  r.y = fp_canon(fp_mul(p.Y, zi3));
  return r;
}

// ---- vectorised global-memory access: 128-bit loads/stores ------------------------------------
BP_DI Fp ld_fp(const Fp* p) {
  const uint4* q = reinterpret_cast<const uint4*>(p);
  uint4 a = __ldg(q), b = __ldg(q + 1);
  Fp r; r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w; r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}
BP_DI Fp ld_fp_plain(const Fp* p) {       // coherent load: for kernels that re-read what they wrote
  const uint4* q = reinterpret_cast<const uint4*>(p);
  uint4 a = q[0], b = q[1];
  Fp r; r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w; r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}
BP_DI Affine ld_affine(const Affine* p) { Affine r; r.x = ld_fp(&p->x); r.y = ld_fp(&p->y); return r; }
BP_DI void st_fp(Fp* p, const Fp& a) {
  uint4* q = reinterpret_cast<uint4*>(p);
  q[0] = make_uint4(a.v[0], a.v[1], a.v[2], a.v[3]);
  q[1] = make_uint4(a.v[4], a.v[5], a.v[6], a.v[7]);
}
BP_DI void st_xyzz(XYZZ* p, const XYZZ& a) { st_fp(&p->X, a.X); st_fp(&p->Y, a.Y); st_fp(&p->ZZ, a.ZZ); st_fp(&p->ZZZ, a.ZZZ); }
BP_DI XYZZ ld_xyzz(const XYZZ* p) {   // plain (coherent) loads: buckets are written by earlier kernels/threads
  XYZZ r;
  const uint4* q = reinterpret_cast<const uint4*>(p);
  uint4 w[8];
#pragma unroll
  for (int i = 0; i < 8; i++) w[i] = q[i];
  Fp* f = &r.X;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    f[i].v[0] = w[2 * i].x; f[i].v[1] = w[2 * i].y; f[i].v[2] = w[2 * i].z; f[i].v[3] = w[2 * i].w;
    f[i].v[4] = w[2 * i + 1].x; f[i].v[5] = w[2 * i + 1].y; f[i].v[6] = w[2 * i + 1].z; f[i].v[7] = w[2 * i + 1].w;
  }
  return r;
}
BP_DI void st_affine(Affine* p, const Affine& a) { st_fp(&p->x, a.x); st_fp(&p->y, a.y); }

}  // namespace bp
