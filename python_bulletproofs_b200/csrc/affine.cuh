// affine.cuh -- batched-affine pair sums: the first levels of the bucket accumulation of a large MSM.
//
// Replaces (same result): the bucket sums of Pippenger.multiexp (/root/reference/src/pippenger/pippenger.py:22-61), as msm.cuh.
//
// A mixed XYZZ addition costs 8M + 2S (666 IMAD.WIDE).  An affine addition costs 2M + 1S once 1/(x2 - x1) is known, and
// Montgomery's trick turns B inversions into one inversion and 3(B-1) multiplications: 5M + 1S (405 IMAD.WIDE) per addition
// when the inversion is shared widely enough.  The bucket-sorted entry list is therefore summed pairwise first:
//   * the scan pads every bucket to a multiple of A = 2^P entries (padding entries = identity), so that the pairs of every one
//     of the P passes are simply the list positions (2j, 2j + 1): no pair straddles a bucket, no per-pass search or scan;
//   * pass 1 gathers the two points of a pair from the point table (sign applied), later passes read the previous pass's
//     output list; each pass writes a list of half the length;
//   * after P passes bucket b owns the contiguous slots [start[b] >> P, start[b+1] >> P) of the last list: k_accumulate (msm.cuh)
//     adds those in XYZZ coordinates as before (1/A of the original additions).
// One warp owns a tile of 32 * B pairs: lane l handles pairs base + k*32 + l (coalesced), multiplies its B denominators into a
// running product (prefix products parked in an L2-resident scratch ring), the 32 lane products are inverted together -- two
// shuffle scans (prefix and suffix products) and ONE binary-GCD inversion by lane 0, which runs on the ALU pipe while the other
// warps of the SM keep the multiplier pipe busy -- and the lanes walk back through their pairs.  Complete: identity operands,
// P + P (tangent slope, denominator 2y) and P + (-P) are handled per pair without poisoning the shared product.
#pragma once
#include "ec.cuh"
#include "coop4.cuh"

namespace bp {

#define BP_AFF_PAD 0xFFFFFFFFu       // entries[].x of a padding slot
#ifndef BP_AFF_B
#define BP_AFF_B 16                   // pairs per lane and tile
#endif

BP_DI Fp shfl_up_fp(const Fp& m, int delta) {
  Fp r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = __shfl_up_sync(BP_FULL_MASK, m.v[i], delta);
  return r;
}

// every lane receives the inverse of its own `acc` (all non-zero mod p)
BP_DI Fp warp_batch_inverse(const Fp& acc, u32 lane) {
  Fp S = acc, T = acc;                       // inclusive prefix / suffix products over the lanes
#pragma unroll 1
  for (int off = 1; off < 32; off <<= 1) {
    const Fp a = shfl_up_fp(S, off), b = shfl_down_fp(T, off);
    const bool up = lane >= (u32)off;        // S <- S * S[lane - off]   else   T <- T * T[lane + off]: both scans share one multiplier site
    const Fp x = fp_mul(S, a);
    const Fp y = fp_mul(T, b);
    if (up) S = x;
    if (lane + (u32)off < 32u) T = y;
  }
  Fp inv = fp_zero();
  if (lane == 31) inv = fp_inv_gcd(S);       // 1 / (product of all 32)
  inv = shfl_fp(inv, 31);
  Fp Sm = shfl_up_fp(S, 1), Tp = shfl_down_fp(T, 1);
  if (lane == 0) Sm = fp_one();
  if (lane == 31) Tp = fp_one();
  return fp_mul(fp_mul(Sm, Tp), inv);        // 1/acc_l = (acc_0 .. acc_{l-1}) * (acc_{l+1} .. acc_31) / total
}

template <bool FIRST>
BP_DI void aff_load_x(const Affine* __restrict__ points, const u32* __restrict__ point_idx, const Affine* __restrict__ phi,
                      const uint2* __restrict__ entries, const Affine* __restrict__ in, u32 j, Fp& x1, Fp& x2, bool& id1, bool& id2,
                      const Affine*& a1, const Affine*& a2, u32& sg) {
  if (FIRST) {
    const uint4 e = __ldg(reinterpret_cast<const uint4*>(entries + 2 * (size_t)j));      // entries (2j, 2j+1): {x, bucket, x, bucket}
    id1 = e.x == BP_AFF_PAD; id2 = e.z == BP_AFF_PAD;
    a1 = id1 ? points : entry_point_ptr(points, point_idx, phi, e.x);
    a2 = id2 ? points : entry_point_ptr(points, point_idx, phi, e.z);
    sg = (id1 ? 0u : e.x >> 31) | (id2 ? 0u : (e.z >> 31) << 1);
  } else {
    a1 = in + 2 * (size_t)j; a2 = a1 + 1; id1 = id2 = false; sg = 0;
  }
  x1 = id1 ? fp_zero() : ld_fp(&a1->x);
  x2 = id2 ? fp_zero() : ld_fp(&a2->x);
}

// B pairs per lane and tile.  n_in_ptr: device word holding the (padded) length of the input list; the pass handles n_in >> 1 pairs.
// tile_ctr: zeroed word, tiles are handed out dynamically.  scratch: gridDim.x * blockDim.x * B field elements.
template <int B, bool FIRST>
__global__ void __launch_bounds__(128, 4) k_aff_pass(const Affine* __restrict__ points, const u32* __restrict__ point_idx, const Affine* __restrict__ phi,
                                                     const uint2* __restrict__ entries, const Affine* __restrict__ in,
                                                     const u32* __restrict__ n_in_ptr, u32 in_shift, Affine* __restrict__ out,
                                                     Fp* __restrict__ scratch, u32* __restrict__ tile_ctr) {
  const u32 lane = threadIdx.x & 31;
  const u32 n_out = (__ldg(n_in_ptr) >> in_shift) >> 1;
  const u32 ntiles = (n_out + 32 * B - 1) / (32 * B);
  Fp* sc = scratch + ((size_t)(blockIdx.x * blockDim.x + threadIdx.x) >> 5) * (32 * B) + lane;      // [k][lane]
  for (;;) {
    u32 tile = 0;
    if (lane == 0) tile = atomicAdd(tile_ctr, 1u);
    tile = __shfl_sync(BP_FULL_MASK, tile, 0);
    if (tile >= ntiles) return;
    const u32 base = tile * (32 * B) + lane;
    // ---- forward: denominators and their running product
    Fp acc = fp_one();
#pragma unroll 1
    for (int k = 0; k < B; k++) {
      const u32 j = base + (u32)k * 32;
      Fp den = fp_one();
      if (j < n_out) {
        Fp x1, x2; bool id1, id2; const Affine *a1, *a2; u32 sg;
        aff_load_x<FIRST>(points, point_idx, phi, entries, in, j, x1, x2, id1, id2, a1, a2, sg);
        {                                               // identity = all-zero point; x = 0 alone cannot tell (curve points with x = 0 exist)
          u32 o1 = 0, o2 = 0;
#pragma unroll
          for (int i = 0; i < 8; i++) { o1 |= x1.v[i]; o2 |= x2.v[i]; }
          if (o1 == 0 && !id1) id1 = affine_is_identity(ld_affine(a1));
          if (o2 == 0 && !id2) id2 = affine_is_identity(ld_affine(a2));
        }
        if (!id1 && !id2) {
          den = fp_sub(x2, x1);
          if (fp_is_zero(den)) {                        // same x: P + P or P - P
            Fp y1 = ld_fp(&a1->y), y2 = ld_fp(&a2->y);
            if (sg & 1u) y1 = fp_neg(y1);
            if (sg & 2u) y2 = fp_neg(y2);
            den = fp_is_zero(fp_sub(y2, y1)) ? fp_dbl(y1) : fp_one();
          }
        }
      }
      acc = fp_mul(acc, den);
      st_fp(sc + (size_t)k * 32, acc);
    }
    Fp inv = warp_batch_inverse(acc, lane);
    // ---- backward: 1/den_k = inv * (den_0 .. den_{k-1}),  inv <- inv * den_k
#pragma unroll 1
    for (int k = B - 1; k >= 0; k--) {
      const u32 j = base + (u32)k * 32;
      if (j >= n_out) continue;                         // (den = 1: inv unchanged)
      Fp x1, x2; bool id1, id2; const Affine *a1, *a2; u32 sg;
      aff_load_x<FIRST>(points, point_idx, phi, entries, in, j, x1, x2, id1, id2, a1, a2, sg);
      Fp y1 = id1 ? fp_zero() : ld_fp(&a1->y), y2 = id2 ? fp_zero() : ld_fp(&a2->y);
      if (!id1) { u32 o = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) o |= x1.v[i] | y1.v[i];
        id1 = o == 0; }
      if (!id2) { u32 o = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) o |= x2.v[i] | y2.v[i];
        id2 = o == 0; }
      if (sg & 1u) y1 = fp_neg(y1);
      if (sg & 2u) y2 = fp_neg(y2);
      Affine r;
      if (id1 || id2) {                                 // den was 1
        r.x = id1 ? x2 : x1; r.y = id1 ? y2 : y1;       // (both identity: zeros)
      } else {
        Fp den = fp_sub(x2, x1), num = fp_sub(y2, y1);
        bool dead = false;
        if (fp_is_zero(den)) {
          if (fp_is_zero(num)) { den = fp_dbl(y1); const Fp xx = fp_sqr(x1); num = fp_add(fp_dbl(xx), xx); }
          else { dead = true; den = fp_one(); }
        }
        Fp dinv = inv;
        if (k > 0) { dinv = fp_mul(inv, ld_fp_plain(sc + (size_t)(k - 1) * 32)); inv = fp_mul(inv, den); }
        const Fp lam = fp_mul(num, dinv);
        r.x = fp_sub(fp_sub(fp_sqr(lam), x1), x2);
        r.y = fp_sub(fp_mul(lam, fp_sub(x1, r.x)), y1);
        if (dead) { r.x = fp_zero(); r.y = fp_zero(); }
      }
      st_affine(out + j, r);
    }
  }
}

// padding entries behind the real ones of every bucket: slots [start[b] + count[b], start[b+1])
__global__ void __launch_bounds__(256) k_aff_pad(const u32* __restrict__ start, const u32* __restrict__ count, u32 nb, uint2* __restrict__ entries) {
  const u32 b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nb) return;
  const u32 e = __ldg(start + b + 1);
  for (u32 i = __ldg(start + b) + __ldg(count + b); i < e; i++) entries[i] = make_uint2(BP_AFF_PAD, b);
}

// after P passes: slot j of the last list belongs to the bucket of entry j << P (a bucket's real entries come first);
// red_entries[j] = {j, bucket}, red_start[b] = start[b] >> P  (b <= nb)
__global__ void __launch_bounds__(256) k_aff_index(const u32* __restrict__ start, u32 nb, const uint2* __restrict__ entries, int P,
                                                   u32* __restrict__ red_start, uint2* __restrict__ red_entries) {
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  const u32 n_red = __ldg(start + nb) >> P;
  if (i <= nb) red_start[i] = __ldg(start + i) >> P;
  if (i < n_red) red_entries[i] = make_uint2(i, __ldg(&entries[(size_t)i << P].y));
}

}  // namespace bp
