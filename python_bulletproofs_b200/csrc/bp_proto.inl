#include <chrono>
#include <atomic>
#include <cuda.h>
// bp_proto.inl -- host drivers of the protocol-level entry points (included by bp_gpu.cu):
// IPA folding rounds with the Fiat-Shamir transcript on the host, the Verifier2 equation,
// batch range-proof verification, and the NCCL plumbing for sharded MSM / batches.

namespace bp {

static ncclComm_t g_comm = nullptr;
static int g_rank = 0, g_nranks = 1;
void nccl_shutdown() { if (g_comm) { ncclCommDestroy(g_comm); g_comm = nullptr; } g_rank = 0; g_nranks = 1; }

#define BP_CU(call)                                                                           \
  do {                                                                                        \
    CUresult r__ = (call);                                                                    \
    if (r__ != CUDA_SUCCESS) return ::bp::fail("%s failed: CUresult %d", #call, (int)r__);    \
  } while (0)
#define BP_NCCL(call)                                                                         \
  do {                                                                                        \
    ncclResult_t r__ = (call);                                                                \
    if (r__ != ncclSuccess) return ::bp::fail("%s failed: %s", #call, ncclGetErrorString(r__)); \
  } while (0)

struct ChallengeForms { Fq x, xinv, xm, xim; };
static ChallengeForms challenge_forms(const Fq& x);
// context of the IPA prover's host nodes (see bp_ipa_prove_hs)
struct IpaHostCtx {
  std::string digest; RunningModHash rh; size_t round = 0, n = 0;
  IpaRound* h_rp = nullptr; uint8_t* h_lr = nullptr;
  bool lr_xyzz = false;            // h_lr holds L, R in XYZZ coordinates (2 x 128 bytes): finished here (fp_host.h)
  std::vector<uint8_t> Ls, Rs, xs;
};
static IpaHostCtx g_ipa_host;
// transcript.add_list_points([L, R]); x = get_modp(q); add_number(x)      inner_product_prover.py:102-106
static void CUDART_CB ipa_host_round(void* p) {
  IpaHostCtx& c = *(IpaHostCtx*)p;
  const size_t r = c.round;
  if (c.lr_xyzz) { uint8_t aff[128]; xyzz_to_affine_host(c.h_lr, 2, aff); memcpy(c.h_lr, aff, 128); }
  memcpy(c.Ls.data() + 64 * r, c.h_lr, 64);
  memcpy(c.Rs.data() + 64 * r, c.h_lr + 64, 64);
  c.digest += point_to_b64(c.h_lr); c.digest += '&';
  c.digest += point_to_b64(c.h_lr + 64); c.digest += '&';
  Fq x = c.rh.challenge((const uint8_t*)c.digest.data(), c.digest.size());
  fq_to_le(c.xs.data() + 32 * r, x);
  c.digest += fq_to_decimal(x); c.digest += '&';
  ChallengeForms f = challenge_forms(x);
  c.h_rp->fold = 1; c.h_rp->xm = f.xm; c.h_rp->xim = f.xim;      // applied at the start of the next round / the last fold
  c.h_rp->m = (u32)(c.n >> (r + 1));
  c.round = r + 1;
}
static ChallengeForms challenge_forms(const Fq& x) {
  ChallengeForms c; c.x = x; c.xinv = fq_inv_host(x); c.xm = fq_to_mont(x); c.xim = fq_to_mont(c.xinv);
  return c;
}

// one folding step on device buffers: P=[u|g|h] (2k each) -> P2=[u|g'|h'] ; a,b (2k) -> a2,b2 (k)
static int fold_step(const Affine* P, Affine* P2, const Fq* a, const Fq* b, Fq* a2, Fq* b2, u32 k, const ChallengeForms& c) {
  ++g.nlaunch, k_fold_points<<<(2 * k + 1 + 63) / 64, 64, 0, g.stream>>>(P, P2, k, c.x, c.xinv);
  ++g.nlaunch, k_fold_scalars<<<(k + 127) / 128, 128, 0, g.stream>>>(a, b, k, c.xm, c.xim, a2, b2);
  BP_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace bp

extern "C" {

int bp_ipa_set_fast_rounds(int on) { g.ipa_fast = on != 0; return 0; }
int bp_ipa_set_graphs(int mode) {
  if (mode < 0 || mode > 2) return fail("bp_ipa_set_graphs: 0 = synchronise per round, 1 = mapped-flag handshake (default), 2 = one CUDA graph with host nodes");
  g.ipa_mode = mode;
  return 0;
}

int bp_sha256(const uint8_t* msg, size_t len, uint8_t out32[32]) {
  Sha256 s; s.update(msg, len); s.final(out32);
  return 0;
}
int bp_sha256_set_portable(int on) { sha256_force_portable(on); return sha256_impl(); }

int bp_mod_hash(const uint8_t* msg, size_t len, uint8_t out32[32]) {
  Fq x = mod_hash_q(msg, len);
  fq_to_le(out32, x);
  return 0;
}

int bp_mod_hash_indexed(const uint8_t* suffix, size_t len, uint32_t first, uint32_t count, uint8_t* out32) {
  // out[i] = mod_hash(str(first + i).encode() + suffix, q): the blinding vectors sL, sR of a range proof
  // (rangeproof_prover.py:48-55, rangeproof_aggreg_prover.py:53-60) -- 2nm hashes per proof, 2.2 us each in Python
  std::string msg;
  msg.reserve(len + 12);
  for (uint32_t i = 0; i < count; i++) {
    msg = std::to_string(first + i);
    msg.append((const char*)suffix, len);
    Fq x = mod_hash_q((const uint8_t*)msg.data(), msg.size());
    fq_to_le(out32 + 32 * (size_t)i, x);
  }
  return 0;
}

int bp_rp_prover_poly1(const uint8_t* aL_bits, const uint8_t* sL32, const uint8_t* sR32, size_t n, size_t m, const uint8_t y32[32],
                       const uint8_t z32[32], uint8_t t1_out[32], uint8_t t2_out[32]) {
  using namespace rpa;
  const size_t nm = n * m;
  if (nm == 0) return fail("bp_rp_prover_poly1: empty vectors");
  const H4 y = reduce(ld(y32)), z = reduce(ld(z32));
  Consts c = position_constants(y, z, n, m);
  const H4 zneg = sub(zero(), z), zm1 = sub(z, one());
  H4 t1 = zero(), t2 = zero();
  for (size_t i = 0; i < nm; i++) {
    const H4 sL = reduce(ld(sL32 + 32 * i)), sR = reduce(ld(sR32 + 32 * i));
    const bool bit = aL_bits[i] != 0;
    const H4 aRz = bit ? z : zm1;                          // aR_i + z  with aR = aL - 1
    const H4 aLz = bit ? add(one(), zneg) : zneg;          // aL_i - z
    const H4 ysR = mul_sm(sR, c.ym[i]);                    // y^i * sR_i
    const H4 inner = add(mul_sm(aRz, c.ym[i]), c.zz[i]);   // y^i (aR_i + z) + zz_i
    const H4 sLm = to_m(sL);
    t1 = add(t1, add(mul_sm(inner, sLm), mul_sm(aLz, to_m(ysR))));
    t2 = add(t2, mul_sm(ysR, sLm));
  }
  st(t1_out, t1); st(t2_out, t2);
  return 0;
}

int bp_rp_prover_poly2(const uint8_t* aL_bits, const uint8_t* sL32, const uint8_t* sR32, size_t n, size_t m, const uint8_t y32[32],
                       const uint8_t z32[32], const uint8_t x32[32], uint8_t* ls32, uint8_t* rs32, uint8_t* yinv32, uint8_t* hsc32,
                       uint8_t* rsy32, uint8_t that_out[32]) {
  using namespace rpa;
  const size_t nm = n * m;
  if (nm == 0) return fail("bp_rp_prover_poly2: empty vectors");
  const H4 y = reduce(ld(y32)), z = reduce(ld(z32)), x = reduce(ld(x32));
  Fq yq; memcpy(yq.v, y.v, 32);
  if (fq_is_zero(yq)) return fail("modular inverse does not exist");
  Fq yiq = fq_inv_host(yq);
  H4 yinv; memcpy(yinv.v, yiq.v, 32);
  Consts c = position_constants(y, z, n, m);
  const H4 xM = to_m(x), yinvM = to_m(yinv);
  const H4 zneg = sub(zero(), z), zm1 = sub(z, one());
  H4 that = zero(), yi = one();                             // y^-i, standard
  for (size_t i = 0; i < nm; i++) {
    const H4 sL = reduce(ld(sL32 + 32 * i)), sR = reduce(ld(sR32 + 32 * i));
    const bool bit = aL_bits[i] != 0;
    const H4 l = add(bit ? add(one(), zneg) : zneg, mul_sm(sL, xM));                       // aL_i - z + sL_i x
    const H4 r = add(mul_sm(add(bit ? z : zm1, mul_sm(sR, xM)), c.ym[i]), c.zz[i]);        // y^i (aR_i + z + sR_i x) + zz_i
    that = add(that, mul_sm(l, to_m(r)));
    st(ls32 + 32 * i, l); st(rs32 + 32 * i, r);
    st(yinv32 + 32 * i, yi);
    st(hsc32 + 32 * i, add(z, mul_sm(c.zz[i], to_m(yi))));                                 // z + zz_i y^-i
    if (rsy32) st(rsy32 + 32 * i, mul_sm(r, to_m(yi)));                                    // r_i y^-i: coefficient of h_i in P
    yi = mul_sm(yi, yinvM);
  }
  st(that_out, that);
  return 0;
}

int bp_rp_verifier_scalars(size_t n, size_t m, const uint8_t y32[32], const uint8_t z32[32], uint8_t* yinv32, uint8_t* hsc32,
                           uint8_t delta_out[32]) {
  using namespace rpa;
  const size_t nm = n * m;
  if (nm == 0) return fail("bp_rp_verifier_scalars: empty vectors");
  const H4 y = reduce(ld(y32)), z = reduce(ld(z32));
  Fq yq; memcpy(yq.v, y.v, 32);
  if (fq_is_zero(yq)) return fail("modular inverse does not exist");                 // y.inv(), utils.py:69-70
  Fq yiq = fq_inv_host(yq);
  H4 yinv; memcpy(yinv.v, yiq.v, 32);
  Consts c = position_constants(y, z, n, m);
  const H4 yinvM = to_m(yinv), zM = to_m(z);
  H4 yi = one(), sum_y = zero();
  for (size_t i = 0; i < nm; i++) {
    sum_y = add(sum_y, from_m(c.ym[i]));
    st(yinv32 + 32 * i, yi);
    st(hsc32 + 32 * i, add(z, mul_sm(c.zz[i], to_m(yi))));                             // z + zz_i y^-i
    yi = mul_sm(yi, yinvM);
  }
  // delta = (z - z^2) * sum y^i - sum_{j=1..m} z^(j+2) * (2^n - 1)      rangeproof_verifier.py:69-71, aggreg :72-79
  const H4 z2 = mul_sm(z, zM);
  H4 two_n = one();
  for (size_t i = 0; i < n; i++) two_n = add(two_n, two_n);
  const H4 tn1M = to_m(sub(two_n, one()));
  H4 delta = mul_sm(sub(z, z2), to_m(sum_y)), zj = mul_sm(z2, zM);                     // z^3
  for (size_t j = 1; j <= m; j++) { delta = sub(delta, mul_sm(zj, tn1M)); zj = mul_sm(zj, zM); }
  st(delta_out, delta);
  return 0;
}

int bp_point_to_b64(const uint8_t pt64[64], char out[45], size_t* out_len) {
  std::string s = point_to_b64(pt64);
  memcpy(out, s.data(), s.size());
  out[s.size()] = 0;
  if (out_len) *out_len = s.size();
  return 0;
}

int bp_ipa_fold_round(const uint8_t* g64, const uint8_t* h64, const uint8_t* a32, const uint8_t* b32, size_t n,
                      const uint8_t x32[32], uint8_t* g_out64, uint8_t* h_out64, uint8_t* a_out32, uint8_t* b_out32) {
  BP_NEED_INIT();
  if (n < 2 || (n & 1)) return fail("bp_ipa_fold_round: n must be even and >= 2");
  u32 k = (u32)(n / 2);
  Fq x; fq_from_le(&x, x32); x = fq_reduce(x);
  if (fq_is_zero(x)) return fail("modular inverse does not exist");          // utils.py:69-70
  ChallengeForms c = challenge_forms(x);
  Affine* P = (Affine*)g.ws_g.ensure((2 * n + 1) * sizeof(Affine));
  Affine* P2 = (Affine*)g.ws_g2.ensure((n + 1) * sizeof(Affine));
  Fq* a = (Fq*)g.ws_a.ensure(n * sizeof(Fq)); Fq* b = (Fq*)g.ws_b.ensure(n * sizeof(Fq));
  Fq* a2 = (Fq*)g.ws_a2.ensure(k * sizeof(Fq)); Fq* b2 = (Fq*)g.ws_b2.ensure(k * sizeof(Fq));
  if (!P || !P2 || !a || !b || !a2 || !b2) return fail("device allocation failed");
  BP_CUDA(cudaMemsetAsync(P, 0, sizeof(Affine), g.stream));
  BP_CUDA(cudaMemcpyAsync(P + 1, g64, n * 64, cudaMemcpyHostToDevice, g.stream));
  BP_CUDA(cudaMemcpyAsync(P + 1 + n, h64, n * 64, cudaMemcpyHostToDevice, g.stream));
  BP_CUDA(cudaMemcpyAsync(a, a32, n * 32, cudaMemcpyHostToDevice, g.stream));
  BP_CUDA(cudaMemcpyAsync(b, b32, n * 32, cudaMemcpyHostToDevice, g.stream));
  if (fold_step(P, P2, a, b, a2, b2, k, c)) return 1;
  BP_CUDA(cudaMemcpyAsync(g_out64, P2 + 1, k * 64, cudaMemcpyDeviceToHost, g.stream));
  BP_CUDA(cudaMemcpyAsync(h_out64, P2 + 1 + k, k * 64, cudaMemcpyDeviceToHost, g.stream));
  BP_CUDA(cudaMemcpyAsync(a_out32, a2, k * 32, cudaMemcpyDeviceToHost, g.stream));
  BP_CUDA(cudaMemcpyAsync(b_out32, b2, k * 32, cudaMemcpyDeviceToHost, g.stream));
  BP_CUDA(cudaStreamSynchronize(g.stream));
  return 0;
}

int bp_ipa_prove_hs(const uint8_t* g64, const uint8_t* h64, const uint8_t* hscale32, const uint8_t u64_[64], const uint8_t* a32,
                    const uint8_t* b32, size_t n, const uint8_t* transcript, size_t transcript_len, uint8_t* Ls64, uint8_t* Rs64,
                    uint8_t* xs32, uint8_t a_out32[32], uint8_t b_out32[32], uint8_t* transcript_out, size_t tout_cap,
                    size_t* tout_len) {
  BP_NEED_INIT();
  if (n == 0 || (n & (n - 1))) return fail("bp_ipa_prove: n must be a power of two");   // inner_product_prover.py:52
  // ping-pong buffers
  Affine* PA = (Affine*)g.ws_g.ensure((2 * n + 1) * sizeof(Affine));
  Affine* PB = (Affine*)g.ws_g2.ensure((n + 1) * sizeof(Affine));
  Fq* aA = (Fq*)g.ws_a.ensure(n * sizeof(Fq)); Fq* bA = (Fq*)g.ws_b.ensure(n * sizeof(Fq));
  Fq* aB = (Fq*)g.ws_a2.ensure(n * sizeof(Fq)); Fq* bB = (Fq*)g.ws_b2.ensure(n * sizeof(Fq));
  Fq* tsc = (Fq*)g.ws_terms_sc.ensure((2 * n + 4) * sizeof(Fq));
  u32* tidx = (u32*)g.ws_idx.ensure((2 * n + 4) * sizeof(u32));
  u32* d_off = (u32*)g.ws_off.ensure(3 * sizeof(u32));
  Affine* d_lr = (Affine*)g.ws_lr.ensure(2 * sizeof(Affine));
  if (!PA || !PB || !aA || !bA || !aB || !bB || !tsc || !tidx || !d_off || !d_lr) return fail("device allocation failed");
  BP_CUDA(cudaMemcpyAsync(PA, u64_, 64, cudaMemcpyHostToDevice, g.stream));
  BP_CUDA(cudaMemcpyAsync(PA + 1, g64, n * 64, cudaMemcpyHostToDevice, g.stream));
  BP_CUDA(cudaMemcpyAsync(PA + 1 + n, h64, n * 64, cudaMemcpyHostToDevice, g.stream));
  BP_CUDA(cudaMemcpyAsync(aA, a32, n * 32, cudaMemcpyHostToDevice, g.stream));
  BP_CUDA(cudaMemcpyAsync(bA, b32, n * 32, cudaMemcpyHostToDevice, g.stream));
  // scalars may arrive unreduced (ModP.x): a -> a mod q happens in fq_reduce inside k_digits for
  // MSM terms, but the folds need reduced inputs, so reduce once here.
  ++g.nlaunch, k_reduce_scalars<<<(unsigned)((n + 127) / 128), 128, 0, g.stream>>>(aA, (u32)n);
  ++g.nlaunch, k_reduce_scalars<<<(unsigned)((n + 127) / 128), 128, 0, g.stream>>>(bA, (u32)n);
  Fq *a = aA, *b = bA;
  // coefficient vectors of the folded generators over the original ones (see k_build_lr_sv); P = PA is never rewritten
  Fq* cg = (Fq*)g.ws_h.ensure(2 * n * sizeof(Fq));
  IpaRound* d_rp = (IpaRound*)g.ws_ipa_rp.ensure(sizeof(IpaRound));
  uint8_t* pin = g.pinned_bytes(1024);
  if (!cg || !d_rp || !pin) return fail("device allocation failed");
  IpaRound* h_rp = (IpaRound*)pin;                 // pinned mirror of the round parameters
  uint8_t* h_lr = pin + 256;                       // pinned landing zone of L, R (affine 2 x 64 bytes, or XYZZ 2 x 128 bytes)
  Fq* ch = cg + n;
  ++g.nlaunch, k_fill_one_mont<<<(unsigned)((2 * n + 127) / 128), 128, 0, g.stream>>>(cg, (u32)(2 * n));
  if (hscale32) {   // effective generators h_i' = hscale_i * h_i (e.g. y^-i, rangeproof_prover.py:77): start ch there
    BP_CUDA(cudaMemcpyAsync(ch, hscale32, n * 32, cudaMemcpyHostToDevice, g.stream));
    ++g.nlaunch, k_to_mont<<<(unsigned)((n + 127) / 128), 128, 0, g.stream>>>(ch, (u32)n);
  }
  // repeated generator set: L and R come from table lookups over the ORIGINAL generators (the s-vector form never
  // rewrites PA, which is what makes a per-set table usable in every round)
  const Affine* tab = nullptr;
  if (fb_enabled() && 2 * n + 1 <= fb.max_points && n > 1) {
    const FbSrc src = {{u64_, g64, h64}, {64, n * 64, n * 64}, 3};
    tab = fb_get(src.hash(0x69706131ull), src, PA, 2 * n + 1);
  }
  const u32 n1 = (u32)n + 1;
  u32 h_off[3] = {0, n1, 2 * n1};
  if (n > 1) BP_CUDA(cudaMemcpyAsync(d_off, h_off, sizeof h_off, cudaMemcpyHostToDevice, g.stream));
  u32 L = 0; while (((size_t)1 << L) < n) L++;
  // Host side of a round (Fiat-Shamir, inner_product_prover.py:102-106): runs as a HOST NODE of the stream / graph
  // (cudaLaunchHostFunc), between the copy-out of L, R and the copy-in of the next round's parameters -- the stream is never
  // synchronised inside the proof.  The context lives in one static object (the library is single threaded by contract).
  IpaHostCtx& hc = g_ipa_host;
  hc.digest.assign((const char*)transcript, transcript_len);
  hc.rh = RunningModHash();
  hc.round = 0; hc.n = n; hc.h_rp = h_rp; hc.h_lr = h_lr; hc.lr_xyzz = false;
  hc.Ls.assign(64 * (size_t)(L ? L : 1), 0); hc.Rs.assign(64 * (size_t)(L ? L : 1), 0); hc.xs.assign(32 * (size_t)(L ? L : 1), 0);
  uint8_t* h_ab = pin + 512;                        // pinned landing zone of the final a, b
  volatile u32* f_g2c = (volatile u32*)(pin + 576);  // device -> host: rounds whose L, R have landed
  volatile u32* f_c2g = (volatile u32*)(pin + 580);  // host -> device: rounds whose challenge is in h_rp
  // Three ways to run the same sequence (bp_ipa_set_graphs): per round [parameters up | fold with the previous challenge |
  // L/R terms | batched MSM | L, R down | Fiat-Shamir on the host], then the last fold and a, b down.
  //   1 (default) the whole proof is enqueued ONCE; the stream itself waits on a mapped host word for each challenge
  //     (cuStreamWaitValue32) and signals each L, R pair through another (cuStreamWriteValue32) while this thread polls:
  //     no stream synchronisation, no launch and no callback thread between two rounds;
  //   2 the sequence with the host side as HOST NODES (cudaLaunchHostFunc), captured into ONE CUDA graph per (n, table)
  //     and replayed: one launch per proof (the callback thread's wake-up costs more than the handshake of mode 1);
  //   0 stream launches with one cudaStreamSynchronize per round (the baseline the other two are measured against).
  int mode = g.ipa_mode;
  if (mode == 1 && !g.memops_ready()) mode = 2;
  // Table rounds of a small proof (n <= 4096) are latency chains: TWO launches per round and no copy operation -- the round's
  // scalar work in one block that reads the parameters straight from the mapped host block (k_ipa_round_prep), then the table
  // MSM whose last block also reduces and writes L, R into mapped host memory (k_fb_msm with a ticket).
  const bool fast = tab != nullptr && n <= 4096 && g.ipa_fast;
  u32* d_ticket = nullptr; XYZZ* d_part = nullptr; Fq* cg2 = nullptr;
  const u32 nbx = (u32)((n1 + 31) / 32);
  if (fast) {
    const bool fresh = g.ws_ipa_ticket.p == nullptr;
    d_ticket = (u32*)g.ws_ipa_ticket.ensure(4 * sizeof(u32));
    d_part = (XYZZ*)fb.blockpart.ensure((size_t)2 * nbx * sizeof(XYZZ));
    cg2 = (Fq*)g.ws_h2.ensure(2 * n * sizeof(Fq));
    if (!d_ticket || !d_part || !cg2) return fail("workspace allocation failed");
    if (fresh) BP_CUDA(cudaMemsetAsync(d_ticket, 0, 4 * sizeof(u32), g.stream));      // the last block of every launch leaves it at zero
  }
  hc.lr_xyzz = fast;
  const size_t nkey = n | (fast ? (size_t)1 << 40 : 0);      // captured proof graphs are specific to the round form
  const unsigned prep_threads = n >= 1024 ? 1024u : (n <= 32 ? 32u : (unsigned)n);
  IpaState st[2] = {{a, b, cg, ch}, {aB, bB, cg2, cg2 ? cg2 + n : nullptr}};      // fast rounds ping-pong the scalar state (k_ipa_round)
  auto enqueue_proof = [&](int md) -> int {
    u32 r = 0;
    int cur = 0;
    for (size_t m = n; m > 1; m >>= 1, r++) {
      if (md == 1 && r > 0) BP_CU(g.cuWaitValue32(g.stream, (CUdeviceptr)(uintptr_t)f_c2g, r, CU_STREAM_WAIT_VALUE_GEQ));
      if (fast) {
        // (the parameter block goes down by a copy: 66 blocks reading the same mapped host words cost ~50 us of serialised PCIe reads)
        BP_CUDA(cudaMemcpyAsync(d_rp, h_rp, sizeof(IpaRound), cudaMemcpyHostToDevice, g.stream));
        ++g.nlaunch, k_ipa_round<<<dim3(nbx, 2), 256, 0, g.stream>>>(tab, st[cur], st[cur ^ 1], (u32)n, d_rp, d_part, d_ticket, (XYZZ*)h_lr);
        cur ^= 1;
      } else {
        BP_CUDA(cudaMemcpyAsync(d_rp, h_rp, sizeof(IpaRound), cudaMemcpyHostToDevice, g.stream));
        ++g.nlaunch, k_round_fold<<<(unsigned)((n + 127) / 128), 128, 0, g.stream>>>(a, b, cg, ch, (u32)n, d_rp);
        ++g.nlaunch, k_build_lr_sv<<<1, 256, 0, g.stream>>>(a, b, cg, ch, (u32)n, &d_rp->m, tsc, tidx);
        if (tab ? fb_msm_run(tab, tidx, tsc, d_off, 2, n1, 0, d_lr, nullptr) : msm_run(PA, tidx, tsc, 2 * n1, d_off, 2, n1, d_lr, nullptr)) return 1;
        BP_CUDA(cudaMemcpyAsync(h_lr, d_lr, 128, cudaMemcpyDeviceToHost, g.stream));
      }
      if (md == 2) BP_CUDA(cudaLaunchHostFunc(g.stream, ipa_host_round, &hc));
      else if (md == 1) BP_CU(g.cuWriteValue32(g.stream, (CUdeviceptr)(uintptr_t)f_g2c, r + 1, 0));
      else { BP_CUDA(cudaStreamSynchronize(g.stream)); ipa_host_round(&hc); }
    }
    if (n > 1) {   // last fold: a, b of length 1   (inner_product_prover.py:109-110)
      if (md == 1) BP_CU(g.cuWaitValue32(g.stream, (CUdeviceptr)(uintptr_t)f_c2g, r, CU_STREAM_WAIT_VALUE_GEQ));
      if (fast) {
        ++g.nlaunch, k_ipa_round_prep<<<1, prep_threads, 0, g.stream>>>(st[cur].a, st[cur].b, st[cur].cg, st[cur].ch, (u32)n, h_rp, tsc, tidx, 0, (Fq*)h_ab);
        BP_CUDA(cudaGetLastError());
        return 0;
      }
      BP_CUDA(cudaMemcpyAsync(d_rp, h_rp, sizeof(IpaRound), cudaMemcpyHostToDevice, g.stream));
      ++g.nlaunch, k_round_fold<<<(unsigned)((n + 127) / 128), 128, 0, g.stream>>>(a, b, cg, ch, (u32)n, d_rp);
    }
    BP_CUDA(cudaMemcpyAsync(h_ab, a, 32, cudaMemcpyDeviceToHost, g.stream));
    BP_CUDA(cudaMemcpyAsync(h_ab + 32, b, 32, cudaMemcpyDeviceToHost, g.stream));
    return 0;
  };
  memset(h_rp, 0, sizeof(IpaRound));
  h_rp->m = (u32)n;
  *f_g2c = 0; *f_c2g = 0;
  bool done = false;
  if (mode == 2 && n > 1 && g.ipa_sized_ok(nkey, tab)) {
    // The first proof of a vector length runs eagerly (it sizes every workspace); from the second one on the same sequence
    // is ONE CUDA-graph launch, captured once per (n, table) and valid while no workspace has moved.
    cudaGraphExec_t gexec = g.ipa_graph_lookup(nkey, tab);
    if (!gexec) {
      cudaGraph_t graph = nullptr;
      const unsigned long long l0 = g.nlaunch, gen0 = alloc_generation();
      BP_CUDA(cudaStreamBeginCapture(g.stream, cudaStreamCaptureModeRelaxed));
      int rc = enqueue_proof(2);
      cudaError_t ce = cudaStreamEndCapture(g.stream, &graph);
      const unsigned nk = (unsigned)(g.nlaunch - l0);
      g.nlaunch = l0;                                  // captured, not launched: counted at every cudaGraphLaunch below
      if (rc || ce != cudaSuccess || !graph || gen0 != alloc_generation()) {
        cudaGetLastError();
        if (graph) cudaGraphDestroy(graph);            // (a workspace grew during the capture: run this proof eagerly, retry next time)
      } else {
        ce = cudaGraphInstantiate(&gexec, graph, 0);
        cudaGraphDestroy(graph);
        if (ce != cudaSuccess) { cudaGetLastError(); gexec = nullptr; }
        else g.ipa_graph_store(nkey, tab, gexec, nk);
      }
    }
    if (gexec) {
      BP_CUDA(cudaGraphLaunch(gexec, g.stream));
      g.nlaunch += g.ipa_graph_kernels(nkey, tab);
      done = true;
    }
  }
  if (!done && mode == 1 && n > 1) {
    const auto t_enq0 = std::chrono::steady_clock::now();
    if (enqueue_proof(1)) { *f_c2g = 0x7FFFFFFFu; cudaStreamSynchronize(g.stream); return 1; }
    // this thread's side of the handshake: wait for L, R of round r, hash, publish the challenge
    const auto t_start = std::chrono::steady_clock::now();
    static const bool timing = getenv("BP_IPA_TIMING") != nullptr;      // per-round wait / host times on stderr (development aid)
    double t_wait[40], t_host[40];
    for (u32 r = 0; r < L; r++) {
      const auto t_r0 = std::chrono::steady_clock::now();
      unsigned spins = 0;
      while (*f_g2c < r + 1) {
        if ((++spins & 0x3FFFu) == 0) {
          if (cudaStreamQuery(g.stream) != cudaErrorNotReady ||
              std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count() > 10.0) {
            *f_c2g = 0x7FFFFFFFu;                      // release every pending wait so that the stream drains
            cudaError_t e = cudaStreamSynchronize(g.stream);
            return fail("bp_ipa_prove: device side of round %u did not complete (%s)", r, cudaGetErrorString(e));
          }
        }
      }
      std::atomic_thread_fence(std::memory_order_acquire);
      const auto t_r1 = std::chrono::steady_clock::now();
      ipa_host_round(&hc);
      std::atomic_thread_fence(std::memory_order_release);
      *f_c2g = r + 1;
      if (timing && r < 40) {
        t_wait[r] = std::chrono::duration<double, std::micro>(t_r1 - t_r0).count();
        t_host[r] = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t_r1).count();
      }
    }
    if (timing) {
      fprintf(stderr, "ipa n=%zu fast=%d enqueue..first wait start %.1f us; per round wait/host us:", n, (int)fast,
              std::chrono::duration<double, std::micro>(t_start - t_enq0).count());
      for (u32 r = 0; r < L && r < 40; r++) fprintf(stderr, " %.0f/%.0f", t_wait[r], t_host[r]);
      fprintf(stderr, "\n");
    }
    done = true;
  }
  if (!done) {
    if (enqueue_proof(mode == 1 ? 0 : mode)) return 1;
    g.ipa_mark_sized(nkey, tab);
  }
  BP_CUDA(cudaStreamSynchronize(g.stream));
  if (L) { memcpy(Ls64, hc.Ls.data(), 64 * (size_t)L); memcpy(Rs64, hc.Rs.data(), 64 * (size_t)L); memcpy(xs32, hc.xs.data(), 32 * (size_t)L); }
  memcpy(a_out32, h_ab, 32); memcpy(b_out32, h_ab + 32, 32);
  const std::string& digest = hc.digest;
  if (tout_len) *tout_len = digest.size();
  if (digest.size() > tout_cap) return fail("bp_ipa_prove: transcript buffer too small (%zu needed)", digest.size());
  memcpy(transcript_out, digest.data(), digest.size());
  return 0;
}

int bp_ipa_prove(const uint8_t* g64, const uint8_t* h64, const uint8_t u64_[64], const uint8_t* a32, const uint8_t* b32,
                 size_t n, const uint8_t* transcript, size_t transcript_len, uint8_t* Ls64, uint8_t* Rs64, uint8_t* xs32,
                 uint8_t a_out32[32], uint8_t b_out32[32], uint8_t* transcript_out, size_t tout_cap, size_t* tout_len) {
  return bp_ipa_prove_hs(g64, h64, nullptr, u64_, a32, b32, n, transcript, transcript_len, Ls64, Rs64, xs32, a_out32, b_out32,
                         transcript_out, tout_cap, tout_len);
}

// p1 (optional) = Protocol 1's own two checks, evaluated in the same device pass: {P, u, x*c, x} of Verifier1 --
// then u64_ / P64 are the PROOF's u_new / P_new and must equal x*u and P + (x*c)*u   (inner_product_verifier.py:51-52)
struct IpaP1 { const uint8_t* P; const uint8_t* u; const uint8_t* xc; const uint8_t* x; };
static int ipa_verify_eq_impl(const uint8_t* g64, const uint8_t* h64, const uint8_t* hscale32, const uint8_t u64_[64], const uint8_t P64[64],
                              size_t n, const uint8_t a32[32], const uint8_t b32[32], const uint8_t* xs32, const uint8_t* Ls64,
                              const uint8_t* Rs64, const IpaP1* pr1, int* accept) {
  BP_NEED_INIT();
  if (n == 0 || (n & (n - 1))) return fail("bp_ipa_verify_eq: n must be a power of two");
  u32 L = 0; while (((size_t)1 << L) < n) L++;
  // ---- scalars on the host: s_i (get_ss, inner_product_verifier.py:91-102) built in O(n) ----
  Fq a, b; fq_from_le(&a, a32); fq_from_le(&b, b32); a = fq_reduce(a); b = fq_reduce(b);
  std::vector<Fq> xm(L), xim(L), x2(L), xi2(L);
  for (u32 j = 0; j < L; j++) {
    Fq x; fq_from_le(&x, xs32 + 32 * j); x = fq_reduce(x);
    if (fq_is_zero(x)) return fail("modular inverse does not exist");
    ChallengeForms c = challenge_forms(x);
    xm[j] = c.xm; xim[j] = c.xim;
    x2[j] = fq_mul(c.x, c.x); xi2[j] = fq_mul(c.xinv, c.xinv);
  }
  std::vector<Fq> s(n);
  Fq s0 = fq_const_r();
  for (u32 j = 0; j < L; j++) s0 = fq_mont(s0, xim[j]);
  s[0] = fq_from_mont(s0);                                   // all bits clear: product of inverses
  for (size_t i = 1; i < n; i++) {
    u32 t = 63 - __builtin_clzll((unsigned long long)i);     // highest set bit of i (counted from the LSB)
    u32 j = L - 1 - t;                                       // its challenge index (bit j is counted from the MSB)
    s[i] = fq_mont(s[i - ((size_t)1 << t)], fq_to_mont(x2[j]));   // x_j^-1 -> x_j : multiply by x_j^2
  }
  size_t T0 = 2 * n + 1, T1 = 2 * L + 1, T = T0 + T1;
  std::vector<uint8_t> hp((T + 3) * 64), hs((T + 3) * 32);
  memcpy(hp.data(), g64, n * 64); memcpy(hp.data() + n * 64, h64, n * 64); memcpy(hp.data() + 2 * n * 64, u64_, 64);
  for (size_t i = 0; i < n; i++) {
    fq_to_le(hs.data() + 32 * i, fq_mul(a, s[i]));                       // a * s_i
    Fq bs = fq_mul(b, s[n - 1 - i]);                                     // b * s_i^-1  (s_i^-1 = s_{n-1-i})
    if (hscale32) { Fq sc; fq_from_le(&sc, hscale32 + 32 * i); bs = fq_mul(bs, fq_reduce(sc)); }   // h_i' = hscale_i * h_i
    fq_to_le(hs.data() + 32 * (n + i), bs);
  }
  fq_to_le(hs.data() + 32 * 2 * n, fq_mul(a, b));
  uint8_t* p1 = hp.data() + T0 * 64; uint8_t* s1 = hs.data() + T0 * 32;
  if (L) { memcpy(p1, Ls64, L * 64); memcpy(p1 + L * 64, Rs64, L * 64); }
  memcpy(p1 + 2 * L * 64, P64, 64);
  for (u32 j = 0; j < L; j++) { fq_to_le(s1 + 32 * j, x2[j]); fq_to_le(s1 + 32 * (L + j), xi2[j]); }
  fq_to_le(s1 + 32 * 2 * L, fq_one());
  uint32_t off[5] = {0, (uint32_t)T0, (uint32_t)T, (uint32_t)T + 2, (uint32_t)T + 3};
  uint8_t out[256];
  if (pr1) {
    uint8_t* pp = hp.data() + T * 64; uint8_t* ss = hs.data() + T * 32;
    memcpy(pp, pr1->P, 64); memcpy(pp + 64, pr1->u, 64); memcpy(pp + 128, pr1->u, 64);
    fq_to_le(ss, fq_one()); memcpy(ss + 32, pr1->xc, 32); memcpy(ss + 64, pr1->x, 32);
  }
  if (bp_msm_batch(hp.data(), hs.data(), off, pr1 ? 4 : 2, out)) return 1;
  bool ok = memcmp(out, out + 64, 64) == 0;
  if (pr1) ok = ok && memcmp(out + 128, P64, 64) == 0 && memcmp(out + 192, u64_, 64) == 0;
  *accept = ok ? 1 : 0;
  return 0;
}

int bp_ipa_verify_eq_hs(const uint8_t* g64, const uint8_t* h64, const uint8_t* hscale32, const uint8_t u64_[64], const uint8_t P64[64],
                        size_t n, const uint8_t a32[32], const uint8_t b32[32], const uint8_t* xs32, const uint8_t* Ls64,
                        const uint8_t* Rs64, int* accept) {
  return ipa_verify_eq_impl(g64, h64, hscale32, u64_, P64, n, a32, b32, xs32, Ls64, Rs64, nullptr, accept);
}

int bp_ipa_verify1_eq_hs(const uint8_t* g64, const uint8_t* h64, const uint8_t* hscale32, const uint8_t u64_[64], const uint8_t P64[64],
                         const uint8_t xc32[32], const uint8_t x32[32], const uint8_t u_new64[64], const uint8_t P_new64[64], size_t n,
                         const uint8_t a32[32], const uint8_t b32[32], const uint8_t* xs32, const uint8_t* Ls64, const uint8_t* Rs64,
                         int* accept) {
  IpaP1 pr1 = {P64, u64_, xc32, x32};
  return ipa_verify_eq_impl(g64, h64, hscale32, u_new64, P_new64, n, a32, b32, xs32, Ls64, Rs64, &pr1, accept);
}


// P = sum a_i g_i + sum bs_i h_i + c u  -- the statement of the inner-product argument (inner_product_prover.py:25-45 computes
// it as P + (x c) u from a P the range prover had to evaluate first, rangeproof_prover.py:78-86).  Over the same point set
// [u | g | h] and with the same table key as bp_ipa_prove_hs, so a repeated generator set answers from the IPA's own table.
int bp_ipa_statement(const uint8_t* g64, const uint8_t* h64, const uint8_t u64_[64], const uint8_t* a32, const uint8_t* bs32,
                     const uint8_t c32[32], size_t n, uint8_t P_out64[64]) {
  BP_NEED_INIT();
  if (n == 0) return fail("bp_ipa_statement: empty vectors");
  const size_t T = 2 * n + 1;
  Affine* PA = (Affine*)g.ws_g.ensure(T * sizeof(Affine));
  Fq* d_sc = (Fq*)g.ws_terms_sc.ensure(T * sizeof(Fq));
  Affine* d_out = (Affine*)g.ws_lr.ensure(2 * sizeof(Affine));
  if (!PA || !d_sc || !d_out) return fail("device allocation failed");
  BP_CUDA(cudaMemcpyAsync(PA, u64_, 64, cudaMemcpyHostToDevice, g.stream));
  BP_CUDA(cudaMemcpyAsync(PA + 1, g64, n * 64, cudaMemcpyHostToDevice, g.stream));
  BP_CUDA(cudaMemcpyAsync(PA + 1 + n, h64, n * 64, cudaMemcpyHostToDevice, g.stream));
  BP_CUDA(cudaMemcpyAsync(d_sc, c32, 32, cudaMemcpyHostToDevice, g.stream));
  BP_CUDA(cudaMemcpyAsync(d_sc + 1, a32, n * 32, cudaMemcpyHostToDevice, g.stream));
  BP_CUDA(cudaMemcpyAsync(d_sc + 1 + n, bs32, n * 32, cudaMemcpyHostToDevice, g.stream));
  const Affine* tab = nullptr;
  if (fb_enabled() && T <= fb.max_points && n > 1) {
    const FbSrc src = {{u64_, g64, h64}, {64, n * 64, n * 64}, 3};
    tab = fb_get(src.hash(0x69706131ull), src, PA, T);
  }
  if (tab) return fb_msm_run_host(tab, nullptr, d_sc, nullptr, 1, T, (u32)T, P_out64);
  if (msm_run(PA, nullptr, d_sc, (u32)T, nullptr, 1, T, d_out, nullptr)) return 1;
  BP_CUDA(cudaMemcpyAsync(P_out64, d_out, 64, cudaMemcpyDeviceToHost, g.stream));
  BP_CUDA(cudaStreamSynchronize(g.stream));
  return 0;
}

int bp_ipa_verify_eq(const uint8_t* g64, const uint8_t* h64, const uint8_t u64_[64], const uint8_t P64[64], size_t n,
                     const uint8_t a32[32], const uint8_t b32[32], const uint8_t* xs32, const uint8_t* Ls64,
                     const uint8_t* Rs64, int* accept) {
  return bp_ipa_verify_eq_hs(g64, h64, nullptr, u64_, P64, n, a32, b32, xs32, Ls64, Rs64, accept);
}

size_t bp_rp_proof_stride(size_t n) {
  size_t L = 0; while (((size_t)1 << L) < n) L++;
  return 5 * 64 + 3 * 32 + 2 * 64 + 2 * 32 + L * 32 + 2 * L * 64;
}

}  // extern "C"
namespace bp {
struct RpStats { double wall_ms = 0, host_ms = 0, gpu_ms = 0; unsigned chunks = 0, threads = 0, table_mode = 0; size_t nproofs = 0; };
static RpStats g_rp_stats;
// chunk lengths of a batch: the host checks of the FIRST chunk are the only ones the GPU cannot hide, and every chunk pays
// fixed latency chains (inversions, the 96-doubling tails), so: a short first chunk, then growing ones, at most 4096 proofs
static void rp_chunk_plan(size_t nproofs, bool tables, unsigned nthreads, std::vector<size_t>& lens) {
  lens.clear();
  if (const char* e = getenv("BP_VERIFY_CHUNK")) {
    size_t ch = (size_t)atol(e); if (ch == 0) ch = 1;
    for (size_t lo = 0; lo < nproofs; lo += ch) lens.push_back(nproofs - lo < ch ? nproofs - lo : ch);
    return;
  }
  if (!tables) { const size_t ch = nproofs < 4096 ? nproofs : 2048; for (size_t lo = 0; lo < nproofs; lo += ch) lens.push_back(nproofs - lo < ch ? nproofs - lo : ch); return; }
  // latency regime: every chunk costs ~0.8 ms of dependent kernels.  The host work in front of a chunk is only the staging pass
  // (~1.5 us per proof per core; the transcript verdicts run after the last chunk is enqueued), so a small batch is ONE chunk
  // (measured, 16 threads: 512 / 1024 / 2048 proofs as one chunk 0.96 / 1.48 / 2.31 ms, as 1/4 + 3/4 1.02 / 1.49 / 2.38 ms)
  // -- unless few host threads make that staging pass itself long (one process per GPU sharing the host: 4 threads, 1024 proofs =
  // 0.38 ms before the GPU starts): then a first quarter gets the device going (1.47 against 1.53 ms)
  if (nproofs <= 2048) {
    if (nproofs >= 512 && (double)nproofs * 1.5 / (double)(nthreads ? nthreads : 1) > 200.0) {
      const size_t a = (nproofs / 4 + 63) & ~(size_t)63;
      lens.push_back(a); lens.push_back(nproofs - a);
    } else lens.push_back(nproofs);
    return;
  }
  size_t left = nproofs, next = nproofs / 8 < 256 ? 256 : (nproofs / 8 > 1024 ? 1024 : nproofs / 8);
  while (left) {
    size_t len = next < left ? next : left;
    if (left - len < 128) len = left;                 // no tiny last chunk
    lens.push_back(len); left -= len;
    next = next * 2 > 4096 ? 4096 : next * 2;
  }
}
}  // namespace bp
extern "C" {

// gather_out != nullptr: after the batch every rank's accept bytes (padded to gather_width) are all-gathered on the device
// through NCCL and copied out once (rank-major) -- no host round trip between the verification and the exchange
static int rp_verify_batch_impl(const uint8_t* gs64, const uint8_t* hs64, const uint8_t g64[64], const uint8_t h64[64],
                       const uint8_t u64_[64], size_t n, const uint8_t* proofs, size_t proof_stride, size_t nproofs,
                       const uint8_t* transcripts, const uint64_t* tr_off, const uint32_t* start_transcript,
                       uint8_t* accept, uint8_t* gather_out, size_t gather_width) {
  BP_NEED_INIT();
  if (n == 0 || (n & (n - 1)) || n > 128) return fail("bp_rp_verify_batch: n must be a power of two <= 128");
  if (gather_out && gather_width < nproofs) return fail("bp_rp_verify_batch_gather: width smaller than this rank's block");
  if (nproofs == 0 && !gather_out) return 0;
  const auto t_call0 = std::chrono::steady_clock::now();
  RpLayout lay = rp_layout((u32)n);
  const u32 L = lay.L;
  if (proof_stride < bp_rp_proof_stride(n)) return fail("bp_rp_verify_batch: proof_stride too small");
  if ((uint64_t)nproofs * lay.tpp >= (1ull << 31)) return fail("bp_rp_verify_batch: batch too large, split it");
  // record offsets inside a packed proof
  const size_t oV = 0, oA = 64, oS = 128, oT1 = 192, oT2 = 256, oTaux = 320, oMu = 352, oThat = 384, oUnew = 416, oPnew = 480,
               oa = 544, ob = 576, oXs = 608, oLs = oXs + 32 * L, oRs = oLs + 64 * L;
  // The batch is cut into chunks: while the GPU evaluates the equations of chunk i, the host threads run the
  // transcript checks of chunk i+1 (pinned, double-buffered staging so the uploads are truly asynchronous).
  unsigned nthreads = std::thread::hardware_concurrency();
  if (nthreads == 0) nthreads = 1;
  if (const char* lws = getenv("LOCAL_WORLD_SIZE")) {      // one process per GPU on a shared host (torchrun): share the cores
    long w = atol(lws);
    // (the calling thread is one of them: OpenMP's master thread works in the parallel region)
    if (w > 1) nthreads = nthreads / (unsigned)w ? nthreads / (unsigned)w : 1;
  }
  if (nthreads > 64) nthreads = 64;
  std::vector<size_t> chunk_len;
  rp_chunk_plan(nproofs, fb_enabled(), nthreads, chunk_len);
  size_t CH = 1;
  for (size_t l : chunk_len) if (l > CH) CH = l;
  const size_t sc_bytes = CH * lay.nsc * 32, pt_bytes = CH * lay.npt * 64, ok_bytes = (CH + 63) & ~(size_t)63;
  const size_t stage_each = sc_bytes + pt_bytes + ok_bytes;
  const size_t hok_bytes = (nproofs + 63) & ~(size_t)63;
  uint8_t* stage = g.pinned_stage(2 * stage_each + hok_bytes + (gather_out ? (size_t)g_nranks * gather_width + 64 : 0));
  if (!stage) return fail("pinned staging allocation failed");
  uint8_t* hok_all = stage + 2 * stage_each;                // host verdicts of the whole batch (verdict pass, after the chunk loop)
  uint8_t* hsc_buf[2] = {stage, stage + stage_each};
  uint8_t* hpt_buf[2] = {stage + sc_bytes, stage + stage_each + sc_bytes};
  size_t chunk_lo = 0;
  int cur = 0;
  // ---- host: transcript checks + challenge extraction, threaded over proofs --------------------
  // Two passes over a proof.  stage_pass: what the DEVICE needs -- proof points and scalars into the pinned staging buffers, y, z, x
  // parsed from their transcript slots, x1 re-derived (one hash) -- ~1.5 us per proof, on the critical path in front of the chunk's
  // upload.  Otherwise: the VERDICT of the reference's three verify_transcript methods (base64 comparisons, the running hash of every
  // IPA round), ~3 us per proof, which nothing on the device waits for: it runs after all chunks are enqueued, underneath the
  // device work, and is merged into the accept bytes at the end (k_rp_merge_host).
  auto work = [&](size_t lo, size_t hi, bool stage_pass) {
    // no heap traffic per proof: slots are (offset, length) pairs in a small stack array, points and scalars are compared in
    // place (b64_point_eq / decimal_slot_eq) instead of through std::string round trips
    struct Slot { size_t first, second; };
    Slot sl_stack[64];
    std::vector<Slot> sl_heap;
    // the first `need` slots of s.split(b"&"); returns how many exist (<= need)
    auto split_first = [&](const uint8_t* s, size_t n, size_t need, Slot*& sl) -> size_t {
      sl = sl_stack;
      if (need > 64) { if (need > n + 1) need = n + 1; sl_heap.resize(need); sl = sl_heap.data(); }
      size_t cnt = 0, st = 0;
      while (cnt < need) {
        const uint8_t* amp = st <= n ? (const uint8_t*)memchr(s + st, '&', n - st) : nullptr;
        const size_t end = amp ? (size_t)(amp - s) : n;
        sl[cnt].first = st; sl[cnt].second = end - st; cnt++;
        if (!amp) break;
        st = end + 1;
      }
      return cnt;
    };
    for (size_t p = lo; p < hi; p++) {
      const uint8_t* pr = proofs + p * proof_stride;
      Slot* sl = nullptr;
      const uint8_t* t0 = transcripts + tr_off[3 * p];
      const size_t t0n = tr_off[3 * p + 1] - tr_off[3 * p];
      const uint8_t* t1 = transcripts + tr_off[3 * p + 1];
      const size_t t1n = tr_off[3 * p + 2] - tr_off[3 * p + 1];
      Fq y = fq_one(), z = fq_one(), x = fq_one(), x1 = fq_one();
      if (stage_pass) {
        uint8_t* sc = hsc_buf[cur] + (p - chunk_lo) * lay.nsc * 32;
        uint8_t* pt = hpt_buf[cur] + (p - chunk_lo) * lay.npt * 64;
        memcpy(pt + 64 * RP_V, pr + oV, 64); memcpy(pt + 64 * RP_A, pr + oA, 64); memcpy(pt + 64 * RP_S, pr + oS, 64);
        memcpy(pt + 64 * RP_T1, pr + oT1, 64); memcpy(pt + 64 * RP_T2, pr + oT2, 64);
        memcpy(pt + 64 * RP_UNEW, pr + oUnew, 64); memcpy(pt + 64 * RP_PNEW, pr + oPnew, 64);
        memcpy(pt + 64 * RP_LS, pr + oLs, 64 * L); memcpy(pt + 64 * (RP_LS + L), pr + oRs, 64 * L);
        memcpy(sc + 32 * RS_THAT, pr + oThat, 32); memcpy(sc + 32 * RS_TAUX, pr + oTaux, 32); memcpy(sc + 32 * RS_MU, pr + oMu, 32);
        memcpy(sc + 32 * RS_A, pr + oa, 32); memcpy(sc + 32 * RS_B, pr + ob, 32);
        memcpy(sc + 32 * RS_XS, pr + oXs, 32 * L);
        // (a proof whose slots do not parse is rejected or deferred by the verdict pass; the device then works on ones)
        if (split_first(t0, t0n, 8, sl) >= 8) {
          if (!decimal_to_fq_fast(t0 + sl[3].first, sl[3].second, &y) || !decimal_to_fq_fast(t0 + sl[4].first, sl[4].second, &z) ||
              !decimal_to_fq_fast(t0 + sl[7].first, sl[7].second, &x) || fq_is_zero(y)) { y = fq_one(); z = fq_one(); x = fq_one(); }
        }
        if (split_first(t1, t1n, 2, sl) >= 2) x1 = mod_hash_q(t1, sl[0].second + 1);
        fq_to_le(sc + 32 * RS_Y, y); fq_to_le(sc + 32 * RS_Z, z); fq_to_le(sc + 32 * RS_X, x); fq_to_le(sc + 32 * RS_X1, x1);
        continue;
      }
      uint8_t verdict = 1;
      // --- RangeVerifier.verify_transcript                         rangeproof_verifier.py:42-53
      if (split_first(t0, t0n, 8, sl) < 8) verdict = 2;               // reference would raise IndexError
      else if (!b64_point_eq(t0 + sl[1].first, sl[1].second, pr + oA) || !b64_point_eq(t0 + sl[2].first, sl[2].second, pr + oS)) verdict = 0;
      else if (!decimal_to_fq_fast(t0 + sl[3].first, sl[3].second, &y) || !decimal_to_fq_fast(t0 + sl[4].first, sl[4].second, &z)) verdict = 2;
      else if (!b64_point_eq(t0 + sl[5].first, sl[5].second, pr + oT1) || !b64_point_eq(t0 + sl[6].first, sl[6].second, pr + oT2)) verdict = 0;
      else if (!decimal_to_fq_fast(t0 + sl[7].first, sl[7].second, &x)) verdict = 2;
      if (verdict == 1 && fq_is_zero(y)) verdict = 2;                 // y.inv() raises in the reference
      // --- Verifier1.verify_transcript                             inner_product_verifier.py:36-42
      if (verdict == 1) {
        if (split_first(t1, t1n, 2, sl) < 2) verdict = 2;
        else {
          // parts[0] + b"&" is the transcript up to and including its first '&'
          x1 = mod_hash_q(t1, sl[0].second + 1);
          if (!decimal_slot_eq(t1 + sl[1].first, sl[1].second, x1)) verdict = 0;
        }
      }
      // --- Verifier2.verify_transcript                             inner_product_verifier.py:104-125
      if (verdict == 1) {
        const uint8_t* t2 = transcripts + tr_off[3 * p + 2];
        size_t t2n = tr_off[3 * p + 3] - tr_off[3 * p + 2];
        const size_t st = start_transcript[p], need = st + 3 * (size_t)L;
        if (L && split_first(t2, t2n, need, sl) < need) verdict = 2;
        RunningModHash rh;
        for (u32 j = 0; j < L && verdict == 1; j++) {
          const Slot& sL = sl[st + 3 * j]; const Slot& sR = sl[st + 3 * j + 1]; const Slot& sX = sl[st + 3 * j + 2];
          if (!b64_point_eq(t2 + sL.first, sL.second, pr + oLs + 64 * j) || !b64_point_eq(t2 + sR.first, sR.second, pr + oRs + 64 * j)) { verdict = 0; break; }
          Fq xj; fq_from_le(&xj, pr + oXs + 32 * j);
          // b"&".join(parts[:idx]) + b"&" is the transcript prefix up to and including the '&' before slot idx;
          // str(xs[j]) == slot == str(expected)  <=>  the slot is the canonical decimal of both
          Fq want = rh.challenge(t2, sX.first);
          if (!fq_eq(xj, want) || !decimal_slot_eq(t2 + sX.first, sX.second, want)) { verdict = 0; break; }
        }
      }
      hok_all[p] = verdict;
    }
  };
  // ---- device buffers and the shared generator table -----------------------------------------------
  // The per-chunk inputs (scalar records, proof points, inverses, expanded terms) are double-buffered: chunk i+1 is
  // uploaded, inverted and expanded on a side stream while chunk i is still in its lookups / bucket pass on the main one.
  const u32 nv = 5 + 2 * L;
  size_t Tc = CH * lay.tpp, npts = lay.fixed + 2 * CH * lay.npt;
  const size_t gwidth = gather_out ? gather_width : nproofs;
  Affine* table = (Affine*)g.ws_pts.ensure(npts * sizeof(Affine));
  Fq* psc2 = (Fq*)g.ws_small.ensure(2 * CH * lay.nsc * sizeof(Fq));
  Fq* tsc2 = (Fq*)g.ws_terms_sc.ensure(2 * Tc * sizeof(Fq));
  u32* tidx2 = (u32*)g.ws_idx.ensure(2 * Tc * sizeof(u32));
  u32* d_off2 = (u32*)g.ws_off.ensure(2 * (4 * CH + 1) * sizeof(u32));
  Affine* d_res = (Affine*)g.ws_out.ensure(4 * CH * sizeof(Affine));
  // accept bytes of this rank [gwidth] | all ranks [R * gwidth] | host verdicts [gwidth] | off-curve flags [2 * CH]
  uint8_t* d_acc = (uint8_t*)g.ws_misc.ensure((size_t)(g_nranks + 2) * gwidth + 2 * CH + 256);
  Fq* d_inv2 = (Fq*)g.ws_a.ensure(2 * CH * (lay.L + 1) * sizeof(Fq));
  if (!table || !psc2 || !tsc2 || !tidx2 || !d_off2 || !d_res || !d_acc || !d_inv2) return fail("device allocation failed");
  uint8_t* d_all = d_acc + gwidth; uint8_t* d_hok = d_all + (size_t)g_nranks * gwidth; uint8_t* d_bad2 = d_hok + gwidth;
  if (g.ensure_aux()) return 1;
  if (g.ensure_stage_events()) return 1;
  if (gather_out) BP_CUDA(cudaMemsetAsync(d_acc, 0, gwidth, g.stream));      // padding bytes of the gathered block
  // Steady state (both generator tables of this set exist and match the caller's bytes): every generator term is a table lookup and
  // nothing reads the generator rows of `table` -- skip their five uploads from pageable memory and the two sums (~0.1 ms per call).
  bool warm = false;
  if (fb_enabled()) {
    const FbSrc src = {{gs64, hs64, g64, h64, u64_}, {n * 64, n * 64, 64, 64, 64}, 5};
    const uint64_t k2 = src.hash(0x72707631ull) ^ ((uint64_t)lay.fixed * 0xD6E8FEB86659FD93ull);
    auto it = fb.tabs.find(k2);
    if (it != fb.tabs.end() && it->second.n == lay.fixed && src.equals(it->second.src)) {
      auto it16 = fb.tabs16.find(k2);
      warm = it16 != fb.tabs16.end() && it16->second.n == lay.fixed && it16->second.parent == it->second.tab;
    }
  }
  if (!warm) {
  BP_CUDA(cudaMemcpyAsync(table, gs64, n * 64, cudaMemcpyHostToDevice, g.stream));
  BP_CUDA(cudaMemcpyAsync(table + n, hs64, n * 64, cudaMemcpyHostToDevice, g.stream));
  BP_CUDA(cudaMemcpyAsync(table + 2 * n, g64, 64, cudaMemcpyHostToDevice, g.stream));
  BP_CUDA(cudaMemcpyAsync(table + 2 * n + 1, h64, 64, cudaMemcpyHostToDevice, g.stream));
  BP_CUDA(cudaMemcpyAsync(table + 2 * n + 2, u64_, 64, cudaMemcpyHostToDevice, g.stream));
  // Gsum = sum gs_i, Hsum = sum hs_i: per-generator-set constants (they are rows of the fixed-base table, so they must be in
  // place before fb_get builds it); kept across calls and confirmed against the generator bytes
  {
    const FbSrc src = {{gs64, hs64, g64, h64, u64_}, {n * 64, n * 64, 64, 64, 64}, 5};
    Affine* sums = (Affine*)g.ws_rp_sums.ensure(2 * sizeof(Affine));
    if (!sums) return fail("device allocation failed");
    if (g.rp_sums_gen == alloc_generation() && !g.rp_sums_src.empty() && src.equals(g.rp_sums_src)) {
      BP_CUDA(cudaMemcpyAsync(table + 2 * n + 3, sums, 2 * sizeof(Affine), cudaMemcpyDeviceToDevice, g.stream));
    } else {
      ++g.nlaunch, k_sum_points<<<1, 128, 0, g.stream>>>(table, (u32)n, table + 2 * n + 3);          // Gsum
      ++g.nlaunch, k_sum_points<<<1, 128, 0, g.stream>>>(table + n, (u32)n, table + 2 * n + 4);      // Hsum
      BP_CUDA(cudaMemcpyAsync(sums, table + 2 * n + 3, 2 * sizeof(Affine), cudaMemcpyDeviceToDevice, g.stream));
      g.rp_sums_src = src.copy(); g.rp_sums_gen = alloc_generation();
    }
  }
  }
  // Repeated generator set: its terms (2n+1 of E4, n+4 of E2, ...: 202 of the 220 terms of a 64-bit proof) are read from
  // the fixed-base table with no doublings and no bucket reduction; the ~21 proof-specific terms (V, A, S, T1, T2, u', P',
  // L_j, R_j) take the window-parallel path of svar.cuh.  Each equation is still checked exactly.
  const Affine *fbtab = nullptr, *fbtab16 = nullptr;
  XYZZ *d_lanes = nullptr, *d_var2 = nullptr, *d_tot = nullptr, *sv_A2 = nullptr, *sv_G2 = nullptr;
  Affine* sv_T2 = nullptr;
  Fq* vsc2 = nullptr; uint4* kd2 = nullptr; u32* kfl2 = nullptr;
  uint64_t fbkey = 0;
  if (fb_enabled()) {
    const FbSrc src = {{gs64, hs64, g64, h64, u64_}, {n * 64, n * 64, 64, 64, 64}, 5};
    fbkey = src.hash(0x72707631ull);
    fbtab = fb_get(fbkey, src, table, lay.fixed);
    fbtab16 = fb_get16(fbkey, fbtab, lay.fixed);
    if (fbtab) {
      d_lanes = (XYZZ*)g.ws_fb_lanes.ensure(4 * CH * (BP_RP_SLOTS + 8) * sizeof(XYZZ));
      d_var2 = (XYZZ*)g.ws_fb_var.ensure(3 * 4 * CH * sizeof(XYZZ));
      // affine tables of the proof points (two parities) + the scratch of k_sv_table: its XYZZ chain and prefix products
      sv_T2 = (Affine*)g.ws_sv_tab.ensure(CH * (size_t)nv * BP_SV_ENT * (2 * sizeof(Affine) + sizeof(XYZZ) + sizeof(Fp)));
      sv_A2 = (XYZZ*)g.ws_sv_acc.ensure(2 * CH * (size_t)(96 + 12) * sizeof(XYZZ));
      vsc2 = (Fq*)g.ws_sv_sc.ensure(2 * CH * (size_t)nv * (sizeof(Fq) + 2 * sizeof(uint4) + sizeof(u32)));
      if (!d_lanes || !d_var2 || !sv_T2 || !sv_A2 || !vsc2) return fail("device allocation failed");
      d_tot = d_var2 + 8 * CH;
      sv_G2 = sv_A2 + 2 * CH * (size_t)96;
      kd2 = (uint4*)(vsc2 + 2 * CH * (size_t)nv);
      kfl2 = (u32*)(kd2 + 2 * CH * (size_t)nv * 2);
      if (g.ensure_var_stream()) return 1;
    }
  }
  const u32 bd = 2 * n < 64 ? 64 : (u32)(2 * n);           // k_rp_expand: g side and h side of every position
  const size_t smem = (5 * n + 32 + 2 * L) * sizeof(Fq);
  const bool timing = getenv("BP_VERIFY_TIMING") != nullptr;
  static const size_t sv_latency_max = [] { const char* e = getenv("BP_SV_LATENCY_MAX"); return e ? (size_t)atol(e) : (size_t)1536; }();
  double host_ms = 0;
  int chunk_no = 0;
  chunk_lo = 0;
  for (size_t ci = 0; ci < chunk_len.size(); chunk_lo += chunk_len[ci], ci++, chunk_no++) {
    const size_t cn = chunk_len[ci], chunk_hi = chunk_lo + cn;
    cur = chunk_no & 1;
    if (chunk_no >= 2) BP_CUDA(cudaEventSynchronize(g.stage_ev[cur]));       // staging buffer free again?
    auto t_h0 = std::chrono::steady_clock::now();
    {   // host: transcript checks of this chunk on all cores (OpenMP keeps its worker pool alive between chunks and calls;
        // creating 16 std::threads per chunk cost ~0.4 ms each time)
      const long nt = (long)(nthreads > cn ? (unsigned)cn : nthreads);
      const size_t per = (cn + (size_t)nt - 1) / (size_t)nt;
#pragma omp parallel for schedule(static, 1) num_threads((int)nt)
      for (long t = 0; t < nt; t++) {
        size_t lo = chunk_lo + (size_t)t * per, hi = lo + per < chunk_hi ? lo + per : chunk_hi;
        if (lo < hi) work(lo, hi, true);
      }
    }
    const double h_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_h0).count();
    host_ms += h_ms;
    if (timing) fprintf(stderr, "chunk %d (%zu proofs): host %.3f ms (%u threads)\n", chunk_no, cn, h_ms, nthreads);
    // side stream: upload, invert, expand (and the variable points' small tables) into the buffers of parity `cur` (free once
    // the main stream is done with chunk i-2); main stream: table lookups, fold, accept; var stream: the window-parallel
    // pass over the proof-specific terms, concurrent with the lookups of the same chunk
    Fq* psc = psc2 + (size_t)cur * CH * lay.nsc; Fq* d_inv = d_inv2 + (size_t)cur * CH * (lay.L + 1);
    Fq* tsc = tsc2 + (size_t)cur * Tc; u32* tidx = tidx2 + (size_t)cur * Tc; u32* d_off = d_off2 + (size_t)cur * (4 * CH + 1);
    uint8_t* d_bad = d_bad2 + (size_t)cur * CH;
    const u32 pt_base = lay.fixed + (u32)(cur * CH * lay.npt);
    if (chunk_no == 0) {   // generator table ready
      BP_CUDA(cudaEventRecord(g.aux_free[0], g.stream)); BP_CUDA(cudaEventRecord(g.aux_free[1], g.stream));
      BP_CUDA(cudaEventRecord(g.ev_rp0, g.stream));
    }
    BP_CUDA(cudaStreamWaitEvent(g.aux_stream, g.aux_free[cur], 0));
    BP_CUDA(cudaMemcpyAsync(table + pt_base, hpt_buf[cur], cn * lay.npt * 64, cudaMemcpyHostToDevice, g.aux_stream));
    BP_CUDA(cudaMemcpyAsync(psc, hsc_buf[cur], cn * lay.nsc * 32, cudaMemcpyHostToDevice, g.aux_stream));
    BP_CUDA(cudaEventRecord(g.stage_ev[cur], g.aux_stream));
    BP_CUDA(cudaMemsetAsync(d_bad, 0, cn, g.aux_stream));
    ++g.nlaunch, k_reduce_scalars<<<(unsigned)((cn * lay.nsc + 127) / 128), 128, 0, g.aux_stream>>>(psc, (u32)(cn * lay.nsc));
    if (cn <= 8) {
      // a handful of proofs (RangeVerifier.verify routes single proofs here): the one field inversion per proof is a
      // 0.38 ms latency chain on the device but 35 us on a host core
      std::vector<Fq> hv(cn * (L + 1));
      for (size_t p = 0; p < cn; p++) {
        const uint8_t* sc = hsc_buf[cur] + p * lay.nsc * 32;
        Fq v[33], pre[33];
        Fq acc = fq_one();
        for (u32 i = 0; i <= L; i++) { fq_from_le(&v[i], sc + 32 * (i == 0 ? (size_t)RS_Y : (size_t)RS_XS + i - 1)); v[i] = fq_reduce(v[i]); acc = fq_mul(acc, v[i]); pre[i] = acc; }
        Fq t = fq_is_zero(acc) ? fq_zero() : fq_inv_host(acc);
        for (int i = (int)L; i >= 0; i--) { hv[p * (L + 1) + i] = fq_mul(t, i > 0 ? pre[i - 1] : fq_one()); t = fq_mul(t, v[i]); }
      }
      BP_CUDA(cudaMemcpyAsync(d_inv, hv.data(), hv.size() * sizeof(Fq), cudaMemcpyHostToDevice, g.aux_stream));
      BP_CUDA(cudaStreamSynchronize(g.aux_stream));        // hv is a stack-scoped pageable buffer
    } else {
      ++g.nlaunch, k_rp_invert<<<(unsigned)((cn + 63) / 64), 64, 0, g.aux_stream>>>(psc, lay, (u32)cn, d_inv);
    }
    Fq* vsc = fbtab ? vsc2 + (size_t)cur * CH * nv : nullptr;
    ++g.nlaunch, k_rp_expand<<<(unsigned)cn, bd, smem, g.aux_stream>>>(psc, d_inv, lay, (u32)cn, pt_base, tsc, tidx, d_off, vsc);
    if (fbtab) {
      const u32 nm = (u32)(4 * cn);
      Affine* sv_T = sv_T2 + (size_t)cur * CH * nv * BP_SV_ENT;
      XYZZ* sv_scr = (XYZZ*)(sv_T2 + 2 * CH * (size_t)nv * BP_SV_ENT); Fp* sv_pre = (Fp*)(sv_scr + CH * (size_t)nv * BP_SV_ENT);
      XYZZ* sv_A = sv_A2 + (size_t)cur * CH * 96; XYZZ* sv_G = sv_G2 + (size_t)cur * CH * 12;
      uint4* kd = kd2 + (size_t)cur * CH * nv * 2; u32* kfl = kfl2 + (size_t)cur * CH * nv;
      XYZZ* d_var = d_var2 + (size_t)cur * 4 * CH;
      const Affine* cpts = table + pt_base;
      ++g.nlaunch, k_sv_table<<<(unsigned)((cn * lay.npt + 127) / 128), 128, 0, g.aux_stream>>>(cpts, lay, (u32)cn, vsc, sv_T, sv_scr, sv_pre, kd, kfl, d_bad, d_var + 2 * cn);
      BP_CUDA(cudaEventRecord(g.aux_ready[cur], g.aux_stream));
      BP_CUDA(cudaStreamWaitEvent(g.stream, g.aux_ready[cur], 0));
      BP_CUDA(cudaStreamWaitEvent(g.var_stream[cur], g.aux_ready[cur], 0));
      ++g.nlaunch, k_sv_main<<<(unsigned)((cn * 32 + 127) / 128), 128, 0, g.var_stream[cur]>>>(cpts, lay, (u32)cn, sv_T, kd, kfl, sv_A);
      ++g.nlaunch, k_sv_comb1<<<(unsigned)((cn * 12 + 127) / 128), 128, 0, g.var_stream[cur]>>>(sv_A, (u32)(cn * 12), sv_G);
      if (cn <= sv_latency_max)        // small chunk: the 96 doublings of the last stage as 4-lane cooperative operations
        ++g.nlaunch, k_sv_comb2q<<<(unsigned)((cn * 12 + 127) / 128), 128, 0, g.var_stream[cur]>>>(sv_G, cpts, lay, (u32)cn, d_var);
      else
        ++g.nlaunch, k_sv_comb2<<<(unsigned)((cn * 3 + 127) / 128), 128, 0, g.var_stream[cur]>>>(sv_G, cpts, lay, (u32)cn, d_var);
      BP_CUDA(cudaEventRecord(g.var_done[cur], g.var_stream[cur]));
      static const unsigned rp_block = [] { const char* e = getenv("BP_RP_WARPS"); return e && atoi(e) == 4 ? 128u : 96u; }();      // verify.cuh: rp_warp_role
      // (8192 proofs: 128-thread blocks 7.43-7.46 ms, 96-thread blocks at 5 per SM 7.30-7.36 ms, at 6 per SM -- 96 registers, 230 bytes
      //  of spills -- 7.50-7.59 ms; gpurun_out/rpwarps2.log)
      if (fbtab16 && rp_block == 96) ++g.nlaunch, k_rp_lookup16<96, 5><<<(unsigned)cn, rp_block, 0, g.stream>>>(fbtab16, tidx, tsc, d_off, (u32)cn, lay.fixed, d_lanes);
      else if (fbtab16) ++g.nlaunch, k_rp_lookup16<128, 4><<<(unsigned)cn, rp_block, 0, g.stream>>>(fbtab16, tidx, tsc, d_off, (u32)cn, lay.fixed, d_lanes);
      else ++g.nlaunch, k_rp_lookup<<<(unsigned)cn, rp_block, 0, g.stream>>>(fbtab, tidx, tsc, d_off, (u32)cn, lay.fixed, d_lanes);
      XYZZ* d_grp = d_lanes + (size_t)4 * CH * BP_RP_SLOTS;
      ++g.nlaunch, k_rp_fold8<<<(nm * 8 + 127) / 128, 128, 0, g.stream>>>(d_lanes, nm, (u32)cn, d_grp);
      BP_CUDA(cudaStreamWaitEvent(g.stream, g.var_done[cur], 0));
      ++g.nlaunch, k_rp_fold<<<(nm + 3) / 4, 128, 0, g.stream>>>(d_grp, d_var, nm, (u32)cn, d_tot);
      ++g.nlaunch, k_rp_accept_xyzz<<<(unsigned)((cn + 127) / 128), 128, 0, g.stream>>>(d_tot, (u32)cn, d_bad, nullptr, d_acc + chunk_lo);
    } else {
      ++g.nlaunch, k_rp_check_points<<<(unsigned)((cn * lay.npt + 127) / 128), 128, 0, g.aux_stream>>>(table + pt_base, lay.npt, (u32)cn, d_bad);
      BP_CUDA(cudaEventRecord(g.aux_ready[cur], g.aux_stream));
      BP_CUDA(cudaStreamWaitEvent(g.stream, g.aux_ready[cur], 0));
      if (msm_run(table, tidx, tsc, (u32)(cn * lay.tpp), d_off, (u32)(4 * cn), lay.tpp / 4, d_res, nullptr)) return 1;
      ++g.nlaunch, k_rp_accept<<<(unsigned)((cn + 127) / 128), 128, 0, g.stream>>>(d_res, (u32)cn, d_bad, nullptr, d_acc + chunk_lo);
    }
    BP_CUDA(cudaEventRecord(g.aux_free[cur], g.stream));
  }
  if (nproofs) {
    // verdict pass over the whole batch, underneath the device work enqueued above
    auto t_v0 = std::chrono::steady_clock::now();
    const long nt = (long)(nthreads > nproofs ? (unsigned)nproofs : nthreads);
    const size_t per = (nproofs + (size_t)nt - 1) / (size_t)nt;
#pragma omp parallel for schedule(static, 1) num_threads((int)nt)
    for (long t = 0; t < nt; t++) {
      size_t lo = (size_t)t * per, hi = lo + per < nproofs ? lo + per : nproofs;
      if (lo < hi) work(lo, hi, false);
    }
    const double v_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_v0).count();
    host_ms += v_ms;
    if (timing) fprintf(stderr, "verdict pass (%zu proofs): host %.3f ms (%u threads)\n", nproofs, v_ms, nthreads);
    BP_CUDA(cudaMemcpyAsync(d_hok, hok_all, nproofs, cudaMemcpyHostToDevice, g.stream));
    ++g.nlaunch, k_rp_merge_host<<<(unsigned)((nproofs + 255) / 256), 256, 0, g.stream>>>(d_hok, (u32)nproofs, d_acc);
    BP_CUDA(cudaEventRecord(g.ev_rp1, g.stream));
  }
  if (gather_out) {
    // the single exchange step of the sharded batch: accept bytes all-gathered device to device over NVLink
    const uint8_t* all = d_acc;
    if (g_comm && g_nranks > 1) { BP_NCCL(ncclAllGather(d_acc, d_all, gwidth, ncclUint8, g_comm, g.stream)); all = d_all; }
    uint8_t* pin = stage + 2 * stage_each + hok_bytes;
    BP_CUDA(cudaMemcpyAsync(pin, all, (size_t)g_nranks * gwidth, cudaMemcpyDeviceToHost, g.stream));
    BP_CUDA(cudaStreamSynchronize(g.stream));
    memcpy(gather_out, pin, (size_t)g_nranks * gwidth);
    if (accept) memcpy(accept, pin + (size_t)g_rank * gwidth, nproofs);
  } else {
    BP_CUDA(cudaMemcpyAsync(accept, d_acc, nproofs, cudaMemcpyDeviceToHost, g.stream));
    BP_CUDA(cudaStreamSynchronize(g.stream));
  }
  RpStats& st = g_rp_stats;
  st.wall_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_call0).count();
  st.host_ms = host_ms; st.chunks = (unsigned)chunk_len.size(); st.threads = nthreads; st.nproofs = nproofs;
  st.table_mode = fbtab16 ? 2u : (fbtab ? 1u : 0u);
  float gms = 0;
  if (nproofs && cudaEventElapsedTime(&gms, g.ev_rp0, g.ev_rp1) != cudaSuccess) { cudaGetLastError(); gms = -1.f; }
  st.gpu_ms = gms;
  if (timing) fprintf(stderr, "batch of %zu: wall %.3f ms, host checks %.3f ms, device span %.3f ms, %u chunks\n", nproofs, st.wall_ms, host_ms, gms, st.chunks);
  return 0;
}

int bp_rp_verify_batch(const uint8_t* gs64, const uint8_t* hs64, const uint8_t g64[64], const uint8_t h64[64],
                       const uint8_t u64_[64], size_t n, const uint8_t* proofs, size_t proof_stride, size_t nproofs,
                       const uint8_t* transcripts, const uint64_t* tr_off, const uint32_t* start_transcript,
                       uint8_t* accept) {
  return rp_verify_batch_impl(gs64, hs64, g64, h64, u64_, n, proofs, proof_stride, nproofs, transcripts, tr_off, start_transcript,
                              accept, nullptr, 0);
}
int bp_rp_verify_batch_gather(const uint8_t* gs64, const uint8_t* hs64, const uint8_t g64[64], const uint8_t h64[64],
                              const uint8_t u64_[64], size_t n, const uint8_t* proofs, size_t proof_stride, size_t nproofs,
                              const uint8_t* transcripts, const uint64_t* tr_off, const uint32_t* start_transcript,
                              size_t width, uint8_t* accept_all) {
  return rp_verify_batch_impl(gs64, hs64, g64, h64, u64_, n, proofs, proof_stride, nproofs, transcripts, tr_off, start_transcript,
                              nullptr, accept_all, width);
}
// [0] wall ms of the last batch call, [1] host transcript-check ms (sum over chunks), [2] device span ms (first chunk's
// first kernel .. last accept kernel, CUDA events on the library stream), [3] chunks, [4] host threads, [5] table mode
// (0 bucket method, 1 byte tables, 2 16-bit tables), [6] proofs
int bp_rp_verify_stats(double out7[7]) {
  const RpStats& st = g_rp_stats;
  out7[0] = st.wall_ms; out7[1] = st.host_ms; out7[2] = st.gpu_ms; out7[3] = st.chunks; out7[4] = st.threads; out7[5] = st.table_mode; out7[6] = (double)st.nproofs;
  return 0;
}

// ---- NCCL ---------------------------------------------------------------------------------------
int bp_nccl_unique_id(uint8_t out128[128]) {
  ncclUniqueId id;
  BP_NCCL(ncclGetUniqueId(&id));
  static_assert(sizeof(id) == 128, "ncclUniqueId is 128 bytes");
  memcpy(out128, &id, 128);
  return 0;
}
int bp_nccl_init(int rank, int nranks, const uint8_t unique_id[128]) {
  BP_NEED_INIT();
  if (g_comm) return 0;
  ncclUniqueId id;
  memcpy(&id, unique_id, 128);
  BP_NCCL(ncclCommInitRank(&g_comm, nranks, id, rank));
  g_rank = rank; g_nranks = nranks;
  return 0;
}
int bp_allgather_bytes(const uint8_t* send, size_t nbytes, uint8_t* recv) {
  BP_NEED_INIT();
  if (g_nranks == 1 || !g_comm) { memcpy(recv, send, nbytes); return 0; }
  uint8_t* d = (uint8_t*)g.ws_misc.ensure(nbytes * (g_nranks + 1));
  if (!d) return fail("device allocation failed");
  BP_CUDA(cudaMemcpyAsync(d, send, nbytes, cudaMemcpyHostToDevice, g.stream));
  BP_NCCL(ncclAllGather(d, d + nbytes, nbytes, ncclUint8, g_comm, g.stream));
  BP_CUDA(cudaMemcpyAsync(recv, d + nbytes, nbytes * g_nranks, cudaMemcpyDeviceToHost, g.stream));
  BP_CUDA(cudaStreamSynchronize(g.stream));
  return 0;
}
// slice MSM on this rank -> ncclAllGather of the 128-byte XYZZ partials over NVLink -> every rank adds
// the R partials and converts to the (identical, canonical) affine result in d_out.  All on g.stream.
// hrec (optional): the slice is points [first, first + n) of that resident handle (which may carry precomputed multiples)
static int msm_sharded_device(const Affine* pts, const Fq* sc, size_t n, Affine* d_out, MsmOpts opt = MsmOpts(), const HandleRec* hrec = nullptr,
                              size_t first = 0) {
  int R = g_comm ? g_nranks : 1;
  XYZZ* d_part = (XYZZ*)g.ws_lr.ensure((size_t)(R + 1) * sizeof(XYZZ));
  if (!d_part) return fail("device allocation failed");
  if (R == 1 && n > 0)      // one rank: nothing to exchange or add, the MSM's own last kernel converts to the canonical affine point
    return hrec ? handle_msm(*hrec, first, sc, n, d_out, nullptr) : msm_run(pts, nullptr, sc, (u32)n, nullptr, 1, n, d_out, nullptr, opt);
  if (n == 0) BP_CUDA(cudaMemsetAsync(d_part, 0, sizeof(XYZZ), g.stream));
  else if (hrec ? handle_msm(*hrec, first, sc, n, nullptr, d_part) : msm_run(pts, nullptr, sc, (u32)n, nullptr, 1, n, nullptr, d_part, opt)) return 1;
  const XYZZ* all = d_part;
  if (R > 1) {   // the single exchange step of the sharded MSM
    BP_NCCL(ncclAllGather(d_part, d_part + 1, sizeof(XYZZ), ncclUint8, g_comm, g.stream));
    all = d_part + 1;
  }
  ++g.nlaunch, k_xyzz_sum<<<1, 32, 0, g.stream>>>(all, (u32)R, d_out);
  BP_CUDA(cudaGetLastError());
  return 0;
}

int bp_msm_sharded(bp_handle points, bp_handle scalars, size_t first, size_t n, uint8_t out64[64]) {
  BP_NEED_INIT();
  HandleRec P, S;
  if (get_handle(points, 0, &P) || get_handle(scalars, 1, &S)) return 1;
  if (first + n > P.n || first + n > S.n) return fail("bp_msm_sharded: slice exceeds the uploaded vectors");
  Affine* d_out = (Affine*)g.ws_out.ensure(sizeof(Affine));
  if (!d_out) return fail("device allocation failed");
  if (msm_sharded_device((const Affine*)P.p + first, (const Fq*)S.p + first, n, d_out, MsmOpts(), &P, first)) return 1;
  BP_CUDA(cudaMemcpyAsync(out64, d_out, 64, cudaMemcpyDeviceToHost, g.stream));
  BP_CUDA(cudaStreamSynchronize(g.stream));
  return 0;
}

// host buffers in, host result out: H2D of this rank's slice + slice MSM + all-gather + sum (the end-to-end form)
int bp_msm_sharded_host(const uint8_t* pts64, const uint8_t* sc32, size_t n, uint8_t out64[64]) {
  BP_NEED_INIT();
  if (n >= (1u << 31)) return fail("bp_msm_sharded_host: n too large");
  Affine* d_pts = (Affine*)g.ws_pts.ensure((n ? n : 1) * sizeof(Affine));
  Fq* d_sc = (Fq*)g.ws_sc.ensure((n ? n : 1) * sizeof(Fq));
  Affine* d_out = (Affine*)g.ws_out.ensure(sizeof(Affine));
  if (!d_pts || !d_sc || !d_out) return fail("device allocation failed");
  MsmOpts opt;
  g.hf.pending = false;
  g.dbg_e2e = getenv("BP_E2E_TIMING") != nullptr;
  if (n && upload_operands(d_pts, pts64, d_sc, sc32, n, &opt)) return 1;
  opt.host_finish = true;
  const int R = g_comm ? g_nranks : 1;
  if (R > 1 && g.hf_enabled && !g.profiling) {
    // several ranks, host result: every rank finishes its own slice on the host (Horner chain + affine conversion, fp_host.h), the
    // 64-byte affine partials are all-gathered (NCCL, through a device staging word) and added on every rank in rank order --
    // instead of k_combine (0.23 ms) + all-gather of XYZZ partials + k_xyzz_sum with its lone-thread inversion
    uint8_t mine[64];
    if (n == 0) memset(mine, 0, 64);
    else {
      if (msm_run(d_pts, nullptr, d_sc, (u32)n, nullptr, 1, n, d_out, nullptr, opt)) { cudaStreamSynchronize(g.copy_stream); return 1; }
      if (msm_finish_to_host(d_out, mine)) return 1;
    }
    std::vector<uint8_t> all((size_t)64 * R);
    if (bp_allgather_bytes(mine, 64, all.data())) return 1;
    affine_sum_host(all.data(), (size_t)R, out64);
    return 0;
  }
  if (msm_sharded_device(d_pts, d_sc, n, d_out, opt)) { cudaStreamSynchronize(g.copy_stream); return 1; }
  g.dbg_rec(8, g.stream);
  if (msm_finish_to_host(d_out, out64)) return 1;
  if (g.dbg_e2e && opt.halves) {
    float t[9] = {0};
    for (int i = 1; i < 9; i++) if (g.dbg_ev[i]) cudaEventElapsedTime(&t[i], g.dbg_ev[0], g.dbg_ev[i]);
    fprintf(stderr, "e2e timeline (ms from copy start): scalars in %.3f | points part 0 in %.3f | part 1 in %.3f || digits start %.3f | sorted %.3f | acc 0 done %.3f | acc 1 done %.3f | result %.3f\n",
            t[1], t[2], t[3], t[4], t[5], t[6], t[7], t[8]);
    g.dbg_e2e = false;
  }
  return 0;
}

int bp_bench_msm_sharded(bp_handle points, bp_handle scalars, size_t first, size_t n, int warmup, int iters, int flush_l2,
                         float* ms_each, uint8_t out64[64]) {
  BP_NEED_INIT();
  HandleRec P, S;
  if (get_handle(points, 0, &P) || get_handle(scalars, 1, &S)) return 1;
  if (first + n > P.n || first + n > S.n) return fail("bp_bench_msm_sharded: slice exceeds the uploaded vectors");
  Affine* d_out = (Affine*)g.ws_out.ensure(sizeof(Affine));
  const size_t flush_bytes = 256u << 20;
  void* d_flush = flush_l2 ? g.ws_flush.ensure(flush_bytes) : nullptr;
  for (int it = 0; it < warmup + iters; it++) {
    if (d_flush) BP_CUDA(cudaMemsetAsync(d_flush, it & 0xff, flush_bytes, g.stream));
    BP_CUDA(cudaEventRecord(g.ev_a, g.stream));
    if (msm_sharded_device((const Affine*)P.p + first, (const Fq*)S.p + first, n, d_out, MsmOpts(), &P, first)) return 1;
    BP_CUDA(cudaEventRecord(g.ev_b, g.stream));
    BP_CUDA(cudaEventSynchronize(g.ev_b));
    float ms = 0;
    BP_CUDA(cudaEventElapsedTime(&ms, g.ev_a, g.ev_b));
    if (it >= warmup) ms_each[it - warmup] = ms;
  }
  BP_CUDA(cudaMemcpy(out64, d_out, 64, cudaMemcpyDeviceToHost));
  return 0;
}

// pinned host staging buffers for callers that want full-rate H2D into bp_msm & co.
void* bp_host_alloc(size_t bytes) {
  void* p = nullptr;
  if (!g.inited && bp_init(-1)) return nullptr;
  if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); fail("cudaHostAlloc(%zu) failed", bytes); return nullptr; }
  return p;
}
int bp_host_free(void* p) { if (p) cudaFreeHost(p); return 0; }

}  // extern "C"

// ---- batched point decompression / lift-x ------------------------------------------
// y = (x^3 + 7)^((p+1)/4); ok = (x < p and y^2 == x^3 + 7).  want: 0 = even y, 1 = odd y (bytes_to_point,
// /root/reference/src/utils/utils.py:119-131), 2 = the root exactly as mod_sqrt(...)[0] returns it, 3 = its negation
// p - y (elliptic_hash, /root/reference/src/utils/elliptic_curve_hash.py:17-23, picks between these two with an md5 bit).
namespace bp {
__global__ void __launch_bounds__(128) k_lift_x(const Fp* __restrict__ xs, const uint8_t* __restrict__ want, u32 n,
                                                Affine* __restrict__ out, uint8_t* __restrict__ ok) {
  u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fp x = ld_fp(xs + i);
  Fp xc = fp_canon(x);
  bool in_range = true;                                   // canonical input required: x < p
#pragma unroll
  for (int k = 0; k < 8; k++) in_range = in_range && (xc.v[k] == x.v[k]);
  Fp seven = fp_zero(); seven.v[0] = 7;
  Fp rhs = fp_add(fp_mul(fp_sqr(x), x), seven);
  Fp y = fp_canon(fp_sqrt(rhs));
  Fp chk = fp_sub(fp_sqr(y), rhs);
  bool good = in_range && fp_is_zero(chk);
  const u32 w = want ? want[i] : 2u;
  bool flip = false;
  if (w == 0u) flip = (y.v[0] & 1u) != 0u;
  else if (w == 1u) flip = (y.v[0] & 1u) == 0u;
  else if (w == 3u) flip = true;
  if (flip) y = fp_canon(fp_neg(y));                      // y != 0: x^3 + 7 = 0 has no root on this curve
  Affine r; r.x = good ? x : fp_zero(); r.y = good ? y : fp_zero();
  st_affine(out + i, r);
  ok[i] = good ? 1 : 0;
}
}  // namespace bp
extern "C" int bp_lift_x_batch(const uint8_t* xs32, const uint8_t* want, size_t n, uint8_t* out64, uint8_t* ok) {
  BP_NEED_INIT();
  if (n == 0) return 0;
  if (!xs32 || !out64 || !ok) return fail("bp_lift_x_batch: null argument");
  Fp* d_x = (Fp*)g.ws_sc.ensure(n * sizeof(Fp));
  Affine* d_out = (Affine*)g.ws_out.ensure(n * sizeof(Affine));
  uint8_t* d_flags = (uint8_t*)g.ws_misc.ensure(2 * n + 256);
  if (!d_x || !d_out || !d_flags) return fail("device allocation failed");
  BP_CUDA(cudaMemcpyAsync(d_x, xs32, n * 32, cudaMemcpyHostToDevice, g.stream));
  if (want) BP_CUDA(cudaMemcpyAsync(d_flags, want, n, cudaMemcpyHostToDevice, g.stream));
  ++g.nlaunch, k_lift_x<<<(unsigned)((n + 127) / 128), 128, 0, g.stream>>>(d_x, want ? d_flags : nullptr, (u32)n, d_out, d_flags + n);
  BP_CUDA(cudaMemcpyAsync(out64, d_out, n * 64, cudaMemcpyDeviceToHost, g.stream));
  BP_CUDA(cudaMemcpyAsync(ok, d_flags + n, n, cudaMemcpyDeviceToHost, g.stream));
  BP_CUDA(cudaStreamSynchronize(g.stream));
  return 0;
}

// ---- a + b for two affine points (fastecdsa Point.__add__ as the reference uses it outside multiexps, e.g.
// /root/reference/src/innerproduct/inner_product_prover.py:33, /root/reference/src/utils/commitments.py:6) ---------------
namespace bp {
__global__ void __launch_bounds__(32) k_point_add(const Affine* __restrict__ ab, Affine* __restrict__ out) {
  if (threadIdx.x != 0) return;
  XYZZ acc = xyzz_from_affine(ld_affine(ab));
  Affine b = ld_affine(ab + 1);
  xyzz_madd_ni(acc, b);                              // complete: identity operands, a = b, a = -b
  st_affine(out, xyzz_to_affine(acc, true));
}
}  // namespace bp
extern "C" int bp_point_add(const uint8_t a64[64], const uint8_t b64[64], uint8_t out64[64]) {
  BP_NEED_INIT();
  Affine* d = (Affine*)g.ws_lr.ensure(4 * sizeof(Affine));
  if (!d) return fail("device allocation failed");
  BP_CUDA(cudaMemcpyAsync(d, a64, 64, cudaMemcpyHostToDevice, g.stream));
  BP_CUDA(cudaMemcpyAsync(d + 1, b64, 64, cudaMemcpyHostToDevice, g.stream));
  ++g.nlaunch, k_point_add<<<1, 32, 0, g.stream>>>(d, d + 2);
  BP_CUDA(cudaMemcpyAsync(out64, d + 2, 64, cudaMemcpyDeviceToHost, g.stream));
  BP_CUDA(cudaStreamSynchronize(g.stream));
  return 0;
}

// ---- arithmetic self-test hooks (used by tests/: known-answer tests vs Python big ints) -----------
namespace bp {
__global__ void k_test_fp(int op, const Fp* a, const Fp* b, u32 n, Fp* out) {
  u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fp x = ld_fp(a + i), y = ld_fp(b + i), r;
  switch (op) {
    case 0: r = fp_mul(x, y); break;
    case 1: r = fp_add(x, y); break;
    case 2: r = fp_sub(x, y); break;
    case 3: r = fp_inv(x); break;
    case 4: r = fp_neg(x); break;
    case 5: r = fp_sqr(x); break;
    case 6: r = fp_inv_gcd(x); break;
    case 7: r = fp_sqrt(fp_sqr(x)); r = fp_mul(r, r); break;      // (sqrt(x^2))^2 == x^2
    default: r = x;
  }
  st_fp(out + i, fp_canon(r));
}
__global__ void k_test_ec(int op, const Affine* a, const Affine* b, u32 n, Affine* out) {
  u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Affine p = ld_affine(a + i), q = ld_affine(b + i);
  XYZZ acc = xyzz_from_affine(p);
  // lift to a non-trivial ZZ so the projective paths are exercised: acc = 2p - p computed projectively
  if (op >= 10) { acc = xyzz_dbl(acc); XYZZ np = xyzz_neg(xyzz_from_affine(p)); xyzz_add(acc, np); op -= 10; }
  switch (op) {
    case 0: xyzz_madd(acc, q); break;
    case 1: { XYZZ t = xyzz_from_affine(q); t = xyzz_dbl(t); XYZZ nq = xyzz_neg(xyzz_from_affine(q)); xyzz_add(t, nq); xyzz_add(acc, t); break; }
    case 2: acc = xyzz_dbl(acc); break;
    case 3: acc = xyzz_neg(acc); break;
  }
  st_affine(out + i, xyzz_to_affine(acc));
}
__global__ void k_test_fq(int op, const Fq* a, const Fq* b, u32 n, Fq* out) {
  u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fq x = fq_reduce(ld_fq(a + i)), y = fq_reduce(ld_fq(b + i)), r;
  switch (op) {
    case 0: r = fq_mul(x, y); break;
    case 1: r = fq_add(x, y); break;
    case 2: r = fq_sub(x, y); break;
    case 3: r = fq_inv(x); break;
    case 4: r = fq_neg(x); break;
    case 5: r = fq_mul_dev(x, y); break;              // standard-form product of fqdev.cuh (pseudo-Mersenne folds)
    case 6: r = fq_inv_dev(x); break;
    case 7: r = fq_sqr_dev(x); break;
    case 8: r = fq_inv_gcd(x); break;
    default: r = x;
  }
  st_fq(out + i, r);
}
template <typename K, typename T>
static int run_test_kernel(K kern, int op, const uint8_t* a, const uint8_t* b, size_t n, uint8_t* out, size_t elt) {
  T* da = (T*)g.ws_pts.ensure(n * elt); T* db = (T*)g.ws_sc.ensure(n * elt); T* dout = (T*)g.ws_out.ensure(n * elt);
  if (!da || !db || !dout) return fail("device allocation failed");
  BP_CUDA(cudaMemcpyAsync(da, a, n * elt, cudaMemcpyHostToDevice, g.stream));
  BP_CUDA(cudaMemcpyAsync(db, b, n * elt, cudaMemcpyHostToDevice, g.stream));
  kern<<<(unsigned)((n + 63) / 64), 64, 0, g.stream>>>(op, da, db, (u32)n, dout);
  BP_CUDA(cudaMemcpyAsync(out, dout, n * elt, cudaMemcpyDeviceToHost, g.stream));
  BP_CUDA(cudaStreamSynchronize(g.stream));
  return 0;
}
}  // namespace bp

extern "C" {
int bp_test_fp(int op, const uint8_t* a32, const uint8_t* b32, size_t n, uint8_t* out32) {
  BP_NEED_INIT();
  if (n == 0) return 0;
  return run_test_kernel<decltype(&k_test_fp), Fp>(k_test_fp, op, a32, b32, n, out32, 32);
}
int bp_test_ec(int op, const uint8_t* a64, const uint8_t* b64, size_t n, uint8_t* out64) {
  BP_NEED_INIT();
  if (n == 0) return 0;
  return run_test_kernel<decltype(&k_test_ec), Affine>(k_test_ec, op, a64, b64, n, out64, 64);
}
int bp_test_xyzz_to_affine_host(const uint8_t* xyzz128, size_t count, uint8_t* out64) {      // host-only: needs no GPU
  if (count > 8) return fail("bp_test_xyzz_to_affine_host: at most 8 points");
  xyzz_to_affine_host(xyzz128, count, out64);
  return 0;
}
int bp_test_affine_sum_host(const uint8_t* pts64, size_t count, uint8_t out64[64]) {      // host-only: needs no GPU
  affine_sum_host(pts64, count, out64);
  return 0;
}
int bp_test_horner_host(const uint8_t* winsum128, int c, int W, int U, int dbl, uint8_t out64[64]) {      // host-only: needs no GPU
  if (c < 1 || W < 1 || U != W + (dbl ? 1 : 0)) return fail("bp_test_horner_host: bad shape");
  horner_host(winsum128, c, W, U, dbl, out64);
  return 0;
}
int bp_test_fq(int op, int on_device, const uint8_t* a32, const uint8_t* b32, size_t n, uint8_t* out32) {
  if (n == 0) return 0;
  if (on_device) { BP_NEED_INIT(); return run_test_kernel<decltype(&k_test_fq), Fq>(k_test_fq, op, a32, b32, n, out32, 32); }
  for (size_t i = 0; i < n; i++) {   // same fq.cuh code on the host
    Fq x, y, r; fq_from_le(&x, a32 + 32 * i); fq_from_le(&y, b32 + 32 * i); x = fq_reduce(x); y = fq_reduce(y);
    switch (op) { case 0: r = fq_mul(x, y); break; case 1: r = fq_add(x, y); break; case 2: r = fq_sub(x, y); break;
                  case 3: case 6: case 8: r = fq_inv(x); break; case 9: r = fq_inv_host(x); break; case 4: r = fq_neg(x); break; case 5: r = fq_mul(x, y); break;
                  case 7: r = fq_mul(x, x); break; default: r = x; }
    fq_to_le(out32 + 32 * i, r);
  }
  return 0;
}
}  // extern "C"
