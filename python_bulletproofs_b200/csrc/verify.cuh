// verify.cuh -- device-side scalar algebra + term construction for batch verification of
// single-value range proofs over one shared generator set.
//
// Replaces, per proof: RangeVerifier.verify (/root/reference/src/rangeproofs/rangeproof_verifier.py:55-97),
// Verifier1.verify (/root/reference/src/innerproduct/inner_product_verifier.py:44-58) and
// Verifier2.get_ss / verify (:91-102,127-147).  Every point equation of the reference is kept as
// its own exact check (no random linear combination), rewritten as "MSM == identity":
//   E1  t_hat*g + taux*h == z^2*V + delta*g + x*T1 + x^2*T2                 rangeproof_verifier.py:73-76
//   E2  P_new == A + x*S + sum(-z*gs_i) + sum((z*y^i + z^2*2^i)*hsp_i) - mu*h + (x1*t_hat)*u
//       evaluated as  A + x*S - z*Gsum + z*Hsum + sum(z^2*2^i*y^-i * hs_i) - mu*h + (x1*t_hat)*u - P_new == O  with the
//       per-generator-set constants Gsum = sum gs_i, Hsum = sum hs_i (the n equal scalars -z collapse into one term)
//                                                              rangeproof_verifier.py:78-84,88-97; inner_product_verifier.py:51
//   E3  u_new == x1*u                                                         inner_product_verifier.py:52
//   E4  MSM(gs||hsp||u_new ; a*s || b*s^-1 || a*b) == P_new + MSM(Ls||Rs ; x_j^2 || x_j^-2)   :134-145
// with hsp_i = y^-i * hs_i folded into the scalars (rangeproof_verifier.py:72).
#pragma once
#include "ec.cuh"
#include "fq.cuh"
#include "fqdev.cuh"
#include "ipa.cuh"
#include "fixedbase.cuh"

namespace bp {

// per-proof scalar slots (standard form)
enum { RS_Y = 0, RS_Z, RS_X, RS_X1, RS_THAT, RS_TAUX, RS_MU, RS_A, RS_B, RS_XS };   // then xs[0..L)
// per-proof point slots
enum { RP_V = 0, RP_A, RP_S, RP_T1, RP_T2, RP_UNEW, RP_PNEW, RP_LS };               // then Ls[0..L), Rs[0..L)

struct RpLayout {
  u32 n, L;            // vector length (power of two), log2 n
  u32 nsc, npt;        // scalars / points per proof
  u32 tpp;             // terms per proof = 3n + 16 + 2L
  u32 fixed;           // fixed table size = 2n + 5  : [gs | hs | g | h | u | sum(gs) | sum(hs)]
};
inline RpLayout rp_layout(u32 n) {
  RpLayout l; l.n = n; l.L = 0; while ((1u << l.L) < n) l.L++;
  l.nsc = RS_XS + l.L; l.npt = RP_LS + 2 * l.L; l.tpp = 3 * n + 16 + 2 * l.L; l.fixed = 2 * n + 5;
  return l;
}

// One thread per proof, ONE field inversion per WARP: y^-1 and the x_j^-1 of 32 proofs by Montgomery's trick -- per-thread
// prefix products of the L+1 values, a product scan across the warp (shuffles), lane 31 inverts the warp's total with the binary
// extended GCD (a lone lane: ~40 us against ~245 us for the 334-product Fermat chain every lane ran before), and the inverse
// is unrolled back through the scan and the local prefixes.  inv[p] = [y^-1, x_0^-1, ..., x_{L-1}^-1] in standard form.
// A zero among the inputs (such proofs are already rejected by the host's transcript checks) is replaced by 1 inside the
// products so that it cannot wipe out its neighbours' inverses; its own "inverse" comes out as 0.
BP_DI Fq shfl_fq(const Fq& a, int src) { Fq r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = __shfl_sync(0xFFFFFFFFu, a.v[i], src); return r; }
BP_DI Fq shfl_up_fq(const Fq& a, int d) { Fq r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = __shfl_up_sync(0xFFFFFFFFu, a.v[i], d); return r; }
__global__ void __launch_bounds__(64) k_rp_invert(const Fq* __restrict__ psc, RpLayout lay, u32 nproofs, Fq* __restrict__ inv) {
  const u32 p = blockIdx.x * blockDim.x + threadIdx.x, lane = threadIdx.x & 31;
  const bool live = p < nproofs;
  const Fq* S = psc + (size_t)(live ? p : 0) * lay.nsc;
  const u32 cnt = lay.L + 1;
  Fq* out = inv + (size_t)(live ? p : 0) * cnt;
  // local prefix products (standard form, parked in out[]): out[i] = v_0 * ... * v_i with zeros read as 1
  Fq acc = fq_one();
  if (live) {
    for (u32 i = 0; i < cnt; i++) {
      Fq v = ld_fq(S + (i == 0 ? RS_Y : RS_XS + i - 1));
      if (!fq_is_zero(v)) acc = fq_mul_ni(acc, v);
      st_fq(out + i, acc);
    }
  }
  // inclusive product scan over the lanes: incl = t_0 * ... * t_lane
  Fq incl = acc;
#pragma unroll 1
  for (int d = 1; d < 32; d <<= 1) {
    Fq up = shfl_up_fq(incl, d);
    if ((int)lane >= d) incl = fq_mul_ni(incl, up);
  }
  Fq tot_inv = fq_zero();
  if (lane == 31) tot_inv = fq_inv_gcd(incl);            // (t_0 ... t_31)^-1: one lane, the others wait at the shuffle
  tot_inv = shfl_fq(tot_inv, 31);
  // exclusive prefix e_lane = t_0 ... t_{lane-1}; suffix products by a second scan from the top:  t_lane^-1 = tot_inv * e_lane * s_lane,
  // s_lane = t_{lane+1} ... t_31
  Fq excl = shfl_up_fq(incl, 1);
  if (lane == 0) excl = fq_one();
  Fq suf = acc;                                          // inclusive suffix product t_lane ... t_31
#pragma unroll 1
  for (int d = 1; d < 32; d <<= 1) {
    Fq dn; 
#pragma unroll
    for (int i = 0; i < 8; i++) dn.v[i] = __shfl_down_sync(0xFFFFFFFFu, suf.v[i], d);
    if ((int)lane + d < 32) suf = fq_mul_ni(suf, dn);
  }
  Fq sx;
#pragma unroll
  for (int i = 0; i < 8; i++) sx.v[i] = __shfl_down_sync(0xFFFFFFFFu, suf.v[i], 1);
  if (lane == 31) sx = fq_one();
  Fq t = fq_mul_ni(fq_mul_ni(tot_inv, excl), sx);        // (this thread's product)^-1
  if (!live) return;
  for (int i = (int)cnt - 1; i >= 0; i--) {
    Fq v = ld_fq(S + (i == 0 ? RS_Y : RS_XS + i - 1));
    Fq prev = i > 0 ? ld_fq(out + i - 1) : fq_one();
    const bool z = fq_is_zero(v);
    st_fq(out + i, z ? fq_zero() : fq_mul_ni(t, prev));  // v_i^-1 = t * (v_0 ... v_{i-1})
    if (!z) t = fq_mul_ni(t, v);
  }
}

// One block per proof, 2n threads (at least 64): thread i < n is the "g side" of position i, thread n + i its "h side".
// Writes the proof's tpp term scalars / point indices, its 4 MSM offsets and (optionally) the variable scalars of svar.cuh.
// All arithmetic in standard form (fqdev.cuh).  Shared work instead of per-thread exponentiations:
//   y^i and y^-i      one Hillis-Steele product scan each (log2 n multiplications per thread; g side y, h side y^-1)
//   s_i               = A[i >> lo] * B[i & (2^lo - 1)] with the two half-products tabulated by the first threads;
//                     s_i^-1 = s_{n-1-i} (complemented bits)
__global__ void __launch_bounds__(256) k_rp_expand(const Fq* __restrict__ psc, const Fq* __restrict__ inv, RpLayout lay, u32 nproofs, u32 pt_base, Fq* __restrict__ tsc,
                                                    u32* __restrict__ tidx, u32* __restrict__ offsets, Fq* __restrict__ vsc) {
  // vsc (optional): the 5 + 2L full-width scalars on proof-specific points, in the order of svar.cuh (V, T1, T2, S, u_new, L_j, R_j)
  extern __shared__ Fq sm[];      // scan[2][2n] | sv[n] | AB[32] | xs[2L]
  const u32 p = blockIdx.x, t = threadIdx.x, n = lay.n, L = lay.L;
  if (p >= nproofs) return;
  const Fq* S = psc + (size_t)p * lay.nsc;
  const Fq* IV = inv + (size_t)p * (L + 1);
  Fq* scan0 = sm; Fq* scan1 = sm + 2 * n; Fq* sv = sm + 4 * n; Fq* AB = sm + 5 * n; Fq* xs = sm + 5 * n + 32;
  const bool hside = t >= n && t < 2 * n, gside = t < n;
  const u32 i = hside ? t - n : t;
  const Fq y = ld_fq(S + RS_Y), z = ld_fq(S + RS_Z);
  if (t < L) { xs[t] = ld_fq(S + RS_XS + t); xs[L + t] = ld_fq(IV + 1 + t); }
  // ---- product scans: scan[t] = base for position >= 1, 1 for position 0
  {
    const Fq base = hside ? ld_fq(IV) : y;
    if (t < 2 * n) scan0[t] = i == 0 ? fq_one() : base;
  }
  __syncthreads();
  Fq* src = scan0; Fq* dst = scan1;
  for (u32 off = 1; off < n; off <<= 1) {
    if (t < 2 * n) {
      Fq v = src[t];
      if (i >= off) v = fq_mul_ni(v, src[t - off]);
      dst[t] = v;
    }
    __syncthreads();
    Fq* tmp = src; src = dst; dst = tmp;
  }
  const Fq ypw = t < 2 * n ? src[t] : fq_zero();          // g side: y^i, h side: y^-i
  // ---- s-vector half products: A over the top `hi` challenge bits, B over the low `lo` bits (bit j counted from the MSB)
  const u32 lo = L / 2, hi = L - lo;
  if (t < (1u << hi) + (1u << lo)) {
    const bool isB = t >= (1u << hi);
    const u32 idx = isB ? t - (1u << hi) : t, nb = isB ? lo : hi, j0 = isB ? hi : 0u;
    Fq acc = fq_one();
    for (u32 j = 0; j < nb; j++) {
      const bool bit = (idx >> (nb - 1 - j)) & 1u;
      acc = fq_mul_ni(acc, bit ? xs[j0 + j] : xs[L + j0 + j]);
    }
    AB[(isB ? 16u : 0u) + idx] = acc;
  }
  __syncthreads();
  if (gside) sv[i] = fq_mul_ni(AB[i >> lo], AB[16 + (i & ((1u << lo) - 1u))]);
  // block sum of y^i for delta(y, z): reuse the scan buffer that is not `src`
  Fq* red = dst;
  if (t < 2 * n) red[t] = gside ? ypw : fq_zero();
  __syncthreads();
  for (u32 off = n >> 1; off > 0; off >>= 1) {
    if (t < off) red[t] = fq_add(red[t], red[t + off]);
    __syncthreads();
  }
  const Fq sum_y = red[0];
  // Terms and MSMs are laid out KIND-MAJOR: all E1 equations of the chunk first, then all E2, E3, E4 (MSM index
  // m = e * nproofs + p).  Warps of the thread-per-unit reduction / Horner kernels then hold equations of one kind, so the
  // light ones (E3: one term, E2, E1) no longer wait for a heavy E4 neighbour in the same warp.
  const size_t len1 = 5, len2 = n + 7, len3 = 2, len4 = 2 * n + 2 + 2 * L;
  const size_t tb1 = (size_t)p * len1, tb2 = (size_t)nproofs * len1 + (size_t)p * len2,
               tb3 = (size_t)nproofs * (len1 + len2) + (size_t)p * len3, tb4 = (size_t)nproofs * (len1 + len2 + len3) + (size_t)p * len4;
  const u32 pb = pt_base + p * lay.npt;                  // point base of this proof (pt_base >= lay.fixed: after the generators)
  const u32 iG = 2 * n, iH = 2 * n + 1, iU = 2 * n + 2, iGsum = 2 * n + 3, iHsum = 2 * n + 4;
  const Fq a = ld_fq(S + RS_A), b = ld_fq(S + RS_B);
  if (gside) {
    st_fq(tsc + tb4 + i, fq_mul_ni(a, sv[i]));                   tidx[tb4 + i] = i;          // a * s_i
  }
  if (hside) {
    const Fq z2 = fq_mul_ni(z, z);
    Fq two_i = fq_zero(); if (i < 256) two_i.v[i >> 5] = 1u << (i & 31);   // 2^i, standard form (host enforces n <= 128)
    two_i = fq_reduce(two_i);
    // ---- E2: z^2 * 2^i * y^-i on hs_i  (z*hs_i is in Hsum)
    st_fq(tsc + tb2 + 2 + i, fq_mul_ni(fq_mul_ni(z2, two_i), ypw));   tidx[tb2 + 2 + i] = n + i;
    // ---- E4: b * s_i^-1 * y^-i on hs_i
    st_fq(tsc + tb4 + n + i, fq_mul_ni(fq_mul_ni(b, sv[n - 1 - i]), ypw));  tidx[tb4 + n + i] = n + i;
  }
  if (t < L) {
    const Fq x2 = fq_mul_ni(xs[t], xs[t]), xi2 = fq_mul_ni(xs[L + t], xs[L + t]);
    st_fq(tsc + tb4 + 2 * n + 2 + t, fq_neg(x2));       tidx[tb4 + 2 * n + 2 + t] = pb + RP_LS + t;        // L_j : -x_j^2
    st_fq(tsc + tb4 + 2 * n + 2 + L + t, fq_neg(xi2));  tidx[tb4 + 2 * n + 2 + L + t] = pb + RP_LS + L + t; // R_j : -x_j^-2
    if (vsc) { Fq* V = vsc + (size_t)p * (5 + 2 * L); st_fq(V + 5 + t, fq_neg(x2)); st_fq(V + 5 + L + t, fq_neg(xi2)); }
  }
  if (t == 32 || (blockDim.x <= 32 && t == 0)) {         // a thread of the second warp: the first one carries the table work above
    const Fq x = ld_fq(S + RS_X), x1 = ld_fq(S + RS_X1);
    const Fq that = ld_fq(S + RS_THAT), taux = ld_fq(S + RS_TAUX), mu = ld_fq(S + RS_MU);
    const Fq one = fq_one(), m1 = fq_neg(one);
    const Fq z2 = fq_mul_ni(z, z), z3 = fq_mul_ni(z2, z);
    // delta = (z - z^2) * sum y^i - z^3 * (2^n - 1)
    Fq two_n = fq_zero();
    if (n < 256) two_n.v[n >> 5] = 1u << (n & 31);
    two_n = fq_reduce(two_n);
    Fq delta = fq_sub(fq_mul_ni(fq_sub(z, z2), sum_y), fq_mul_ni(z3, fq_sub(two_n, one)));
    Fq x2 = fq_mul_ni(x, x);
    // E1
    st_fq(tsc + tb1 + 0, fq_sub(that, delta));  tidx[tb1 + 0] = iG;
    st_fq(tsc + tb1 + 1, taux);                 tidx[tb1 + 1] = iH;
    st_fq(tsc + tb1 + 2, fq_neg(z2));           tidx[tb1 + 2] = pb + RP_V;
    st_fq(tsc + tb1 + 3, fq_neg(x));            tidx[tb1 + 3] = pb + RP_T1;
    st_fq(tsc + tb1 + 4, fq_neg(x2));           tidx[tb1 + 4] = pb + RP_T2;
    // E2 non-generator terms
    st_fq(tsc + tb2 + 0, one);                  tidx[tb2 + 0] = pb + RP_A;
    st_fq(tsc + tb2 + 1, x);                    tidx[tb2 + 1] = pb + RP_S;
    st_fq(tsc + tb2 + 2 + n + 0, fq_neg(mu));              tidx[tb2 + 2 + n + 0] = iH;
    st_fq(tsc + tb2 + 2 + n + 1, fq_mul_ni(x1, that));     tidx[tb2 + 2 + n + 1] = iU;
    st_fq(tsc + tb2 + 2 + n + 2, m1);                      tidx[tb2 + 2 + n + 2] = pb + RP_PNEW;
    st_fq(tsc + tb2 + 2 + n + 3, fq_neg(z));               tidx[tb2 + 2 + n + 3] = iGsum;        // sum_i (-z) * gs_i
    st_fq(tsc + tb2 + 2 + n + 4, z);                       tidx[tb2 + 2 + n + 4] = iHsum;        // sum_i z * hs_i
    // E3
    st_fq(tsc + tb3 + 0, x1);                   tidx[tb3 + 0] = iU;
    st_fq(tsc + tb3 + 1, m1);                   tidx[tb3 + 1] = pb + RP_UNEW;
    // E4 non-generator terms
    const Fq ab = fq_mul_ni(a, b);
    st_fq(tsc + tb4 + 2 * n + 0, ab);           tidx[tb4 + 2 * n + 0] = pb + RP_UNEW;
    st_fq(tsc + tb4 + 2 * n + 1, m1);           tidx[tb4 + 2 * n + 1] = pb + RP_PNEW;
    if (vsc) {
      Fq* V = vsc + (size_t)p * (5 + 2 * L);
      st_fq(V + 0, fq_neg(z2)); st_fq(V + 1, fq_neg(x)); st_fq(V + 2, fq_neg(x2)); st_fq(V + 3, x); st_fq(V + 4, ab);
    }
    offsets[p] = (u32)tb1; offsets[nproofs + p] = (u32)tb2; offsets[2 * nproofs + p] = (u32)tb3; offsets[3 * nproofs + p] = (u32)tb4;
    if (p == nproofs - 1) offsets[4 * nproofs] = nproofs * lay.tpp;
  }
}

// out = sum of n affine points (n <= 1024), one block of 128 threads: the per-generator-set constants Gsum, Hsum
__global__ void __launch_bounds__(128) k_sum_points(const Affine* __restrict__ pts, u32 n, Affine* __restrict__ out) {
  __shared__ XYZZ sm[128];
  XYZZ acc = xyzz_identity();
  for (u32 i = threadIdx.x; i < n; i += 128) { Affine p = ld_affine(pts + i); xyzz_madd_ni(acc, p); }
  sm[threadIdx.x] = acc;
  __syncthreads();
  for (int off = 64; off > 0; off >>= 1) {
    if (threadIdx.x < off) { XYZZ v = sm[threadIdx.x + off]; xyzz_add_ni(acc, v); sm[threadIdx.x] = acc; }
    __syncthreads();
  }
  if (threadIdx.x == 0) st_affine(out, xyzz_to_affine(acc));
}

// accept[p] = all four MSM results of proof p are the identity
// hok[p] = the host's transcript verdict (1 = passed; 0 reject / 2 defer override whatever the equations say)
__global__ void k_rp_accept(const Affine* __restrict__ res, u32 nproofs, const uint8_t* __restrict__ bad, const uint8_t* __restrict__ hok,
                            uint8_t* __restrict__ accept) {
  u32 p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= nproofs) return;
  if (hok && hok[p] != 1) { accept[p] = hok[p]; return; }
  bool ok = bad[p] == 0;                                 // a proof with a point off the curve is rejected (svar.cuh)
  for (u32 e = 0; e < 4; e++) ok = ok && affine_is_identity(ld_affine(res + e * nproofs + p));
  accept[p] = ok ? 1 : 0;
}

// ---- table path of the batch verifier: balanced lookups -------------------------------------------------------------
// One block of 128 threads per proof.  The four equations have very different numbers of generator terms (2, n+4, 1,
// 2n+1), so the 4 warps are dealt out as E4: 2 (interleaved term slices), E2: 1, E1 and E3: 1 (one after the other).  Table
// entries carry their window weight, so any lane may add any (term, window) entry: the number of lanes per equation only
// sets how many partial sums the fold kernels must add afterwards (64 / 32 / 32 / 32 here; 128 / 64 / 32 / 32 with the 256-
// thread layout of round 1, whose folds cost a third of the lookups).  Terms on proof-specific points (idx >= nfixed) are
// skipped: they take the window-parallel pass of svar.cuh.  Lane sums land in part[(e*nproofs + p)*64 + 32*s + lane].
#define BP_RP_SLOTS 64
// Blocks of 96 threads (the default): the E1 / E3 warp of the 128-thread layout is idle after 4 of ~34 steps, a quarter of the
// resident warps of a kernel whose occupancy the registers cap at 16 warps per SM; with three warps per proof -- warp 2 takes
// E2, then E1, then E3: 38 steps against 33 of the E4 warps -- six blocks are resident (18 warps, all busy).
BP_DI u32 rp_warp_passes(u32 warp) { return blockDim.x == 96 ? (warp == 2 ? 3u : 1u) : (warp == 3 ? 2u : 1u); }
BP_DI void rp_warp_role(u32 warp, u32 pass, u32& e, u32& nw, u32& s) {
  if (warp < 2) { e = 3u; nw = 2u; s = warp; }
  else if (blockDim.x == 96) { e = pass == 0 ? 1u : (pass == 1 ? 0u : 2u); nw = 1u; s = 0u; }
  else if (warp == 2) { e = 1u; nw = 1u; s = 0u; }
  else { e = pass == 0 ? 0u : 2u; nw = 1u; s = 0u; }
}
// byte tables: lane = byte window of one term per step
__global__ void __launch_bounds__(128) k_rp_lookup(const Affine* __restrict__ tab, const u32* __restrict__ idx, const Fq* __restrict__ sc,
                                                   const u32* __restrict__ offsets, u32 nproofs, u32 nfixed, XYZZ* __restrict__ part) {
  const u32 p = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (p >= nproofs) return;
#pragma unroll 1
  for (u32 pass = 0; pass < rp_warp_passes(warp); pass++) {
    u32 e, nw, s;
    rp_warp_role(warp, pass, e, nw, s);
    const u32 lo = __ldg(offsets + e * nproofs + p), hi = __ldg(offsets + e * nproofs + p + 1);
    XYZZ acc = xyzz_identity();
    for (u32 t = lo + s; t < hi; t += nw) {
      const u32 gi = __ldg(idx + t);
      if (gi >= nfixed) continue;
      const u32* kw = reinterpret_cast<const u32*>(sc + t);
      Fq k;
#pragma unroll
      for (int i = 0; i < 8; i++) k.v[i] = __ldg(kw + i);
      k = fq_reduce(k);
      const u32 d = (k.v[lane >> 2] >> (8 * (lane & 3))) & 0xFFu;
      if (d) { Affine q = ld_affine(tab + fb_index(gi, lane, d)); xyzz_madd_ni(acc, q); }
    }
    st_xyzz(part + ((size_t)(e * nproofs + p) * BP_RP_SLOTS + 32 * s + lane), acc);
  }
}
// (Measured and not kept, round 2: cutting the block's 101 steps into four equal ranges, segment sums merged through shared memory --
// the E1/E3 warp is idle after 2 steps while the others take 33-34 -- left the kernel at 3.29 ms per 8192 proofs: it is not bound by
// the longest warp of a block but by the chip-wide rate of dependent table-load + addition chains.)
// Same with the 16-bit table: lane l owns window l & 15 of the (l >> 4)-th of two terms taken per step, so a term costs 16
// lookups; the table lives in HBM (8.9 GB), hence the next entry is loaded into registers under the current addition.
template <int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB) k_rp_lookup16(const Affine* __restrict__ tab16, const u32* __restrict__ idx, const Fq* __restrict__ sc,
                                                     const u32* __restrict__ offsets, u32 nproofs, u32 nfixed, XYZZ* __restrict__ part) {
  const u32 p = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (p >= nproofs) return;
  const u32 win = lane & 15u, sub = lane >> 4;
#pragma unroll 1
  for (u32 pass = 0; pass < rp_warp_passes(warp); pass++) {
    u32 e, nw, s;
    rp_warp_role(warp, pass, e, nw, s);
    const u32 lo = __ldg(offsets + e * nproofs + p), hi = __ldg(offsets + e * nproofs + p + 1);
    XYZZ acc = xyzz_identity();
    Affine cur; cur.x = fp_zero(); cur.y = fp_zero();
    bool have = false;
    for (u32 t = lo + 2 * s + sub; ; t += 2 * nw) {
      // fetch the entry of term t (if any) while the previous one is being added
      Affine nxt; nxt.x = fp_zero(); nxt.y = fp_zero();
      bool nhave = false;
      const bool in = t < hi;
      if (in) {
        const u32 gi = __ldg(idx + t);
        if (gi < nfixed) {
          const u32* kw = reinterpret_cast<const u32*>(sc + t);
          Fq k;
#pragma unroll
          for (int i = 0; i < 8; i++) k.v[i] = __ldg(kw + i);
          k = fq_reduce(k);
          const u32 d = (k.v[win >> 1] >> (16 * (win & 1))) & 0xFFFFu;
          if (d) { nxt = ld_affine(tab16 + fb_index16(gi, win, d)); nhave = true; }
        }
      }
      if (have) xyzz_madd(acc, cur);                     // inlined: 128 registers, no stack traffic for the accumulator (7.40 -> 7.33 ms per 8192 proofs)
      cur = nxt; have = nhave;
      if (!__any_sync(BP_FULL_MASK, in)) break;          // both term slots of the warp are past the end
    }
    st_xyzz(part + ((size_t)(e * nproofs + p) * BP_RP_SLOTS + 32 * s + lane), acc);
  }
}
// Fold, step 1 (throughput form): one thread adds 8 consecutive lane sums; 8 (E4) or 4 group sums per equation slot
__global__ void __launch_bounds__(128) k_rp_fold8(const XYZZ* __restrict__ part, u32 nmsm, u32 nproofs, XYZZ* __restrict__ grp) {
  const u32 t = blockIdx.x * blockDim.x + threadIdx.x;          // (m, group of 8)
  const u32 m = t >> 3, gq = t & 7;
  if (m >= nmsm) return;
  const u32 e = m / nproofs, ng = e == 3 ? 8u : 4u;
  if (gq >= ng) return;
  const XYZZ* src = part + (size_t)m * BP_RP_SLOTS + 8 * gq;
  XYZZ v = ld_xyzz(src);
#pragma unroll 1
  for (int j = 1; j < 8; j++) { XYZZ x = ld_xyzz(src + j); xyzz_add_ni(v, x); }
  st_xyzz(grp + (size_t)m * 8 + gq, v);
}
// Fold, step 2: one warp per equation: its <= 8 group sums + the variable part `other[m]` -> out[m]
__global__ void __launch_bounds__(128) k_rp_fold(const XYZZ* __restrict__ part, const XYZZ* __restrict__ other, u32 nmsm, u32 nproofs,
                                                 XYZZ* __restrict__ out) {
  __shared__ XYZZ sm[4][8];
  const u32 warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  u32 m = blockIdx.x * 4 + warp;
  const bool live = m < nmsm;
  if (!live) m = nmsm - 1;
  const u32 e = m / nproofs, nl = e == 3 ? 8u : 4u;
  const int role = lane & 3, base = lane & ~3;
  const u32 q = lane >> 2;
  const XYZZ* src = part + (size_t)m * 8;
  XYZZ v = q < nl ? ld_xyzz(src + q) : xyzz_identity();
  if (role == 0) st_xyzz(&sm[warp][q], v);
  __syncwarp();
#pragma unroll 1
  for (u32 off = 4; off > 0; off >>= 1) {
    XYZZ x = (q < off) ? ld_xyzz(&sm[warp][q + off]) : xyzz_identity();
    v = coop_add(v, x, role, base);
    __syncwarp();
    if (q < off && role == 0) st_xyzz(&sm[warp][q], v);
    __syncwarp();
  }
  if (other) { XYZZ x = ld_xyzz(other + m); v = coop_add(v, x, role, base); }
  if (live && lane == 0) st_xyzz(out + m, v);
}

// same decision from XYZZ sums (table path: fixed-generator part + other terms already added): identity <=> ZZ == 0
__global__ void k_rp_accept_xyzz(const XYZZ* __restrict__ res, u32 nproofs, const uint8_t* __restrict__ bad, const uint8_t* __restrict__ hok,
                                 uint8_t* __restrict__ accept) {
  u32 p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= nproofs) return;
  if (hok && hok[p] != 1) { accept[p] = hok[p]; return; }
  bool ok = bad[p] == 0;
  for (u32 e = 0; e < 4; e++) ok = ok && xyzz_is_identity(ld_xyzz(res + e * nproofs + p));
  accept[p] = ok ? 1 : 0;
}

// host verdicts (transcript checks: 1 = passed, 0 = reject, 2 = the reference would raise) override the device's decisions
__global__ void k_rp_merge_host(const uint8_t* __restrict__ hok, u32 nproofs, uint8_t* __restrict__ accept) {
  const u32 p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p < nproofs && hok[p] != 1) accept[p] = hok[p];
}

}  // namespace bp
