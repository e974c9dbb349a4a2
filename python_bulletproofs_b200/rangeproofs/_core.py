"""Shared arithmetic of the (aggregated) range proof, written once for m >= 1 values.

For m = 1 every formula of the aggregated protocol (rangeproof_aggreg_prover.py /
rangeproof_aggreg_verifier.py) collapses to the single-value one (rangeproof_prover.py /
rangeproof_verifier.py), including the `rho = H(str(2*n) + t)` derivation, so both public
prover classes and both verifier classes are thin shells around this module.

Division of labour: the transcript stays on the host in Python; the prover's O(nm) Z_q algebra runs in C on the host
(bp_rp_prover_poly1/2, bp_mod_hash_indexed) and hands packed scalar vectors straight to the device
calls; every elliptic-curve operation goes to the GPU as a multi-scalar multiplication or the IPA round loop.  The
pure-Python form of the same algebra (`_prove_python`) is what runs for a group order other than secp256k1's.
"""
from .. import _native as nat
from ..innerproduct.inner_product_prover import NIProver
from ..innerproduct.inner_product_verifier import Verifier1
from ..pippenger import PipSECP256k1
from ..point import Point
from ..utils.transcript import Transcript
from ..utils.utils import ModP, mod_hash, point_to_b64


class Proof:
    """Range proof container (rangeproof_verifier.py:10-22; the aggregated module's copy is identical)."""

    def __init__(self, taux, mu, t_hat, T1, T2, A, S, innerProof, transcript):
        self.taux = taux
        self.mu = mu
        self.t_hat = t_hat
        self.T1 = T1
        self.T2 = T2
        self.A = A
        self.S = S
        self.innerProof = innerProof
        self.transcript = transcript


def _position_constants(y, z, n, nm, q):
    """y^i and z^(2 + i//n) * 2^(i % n) for i < nm."""
    ypow = [1] * nm
    for i in range(1, nm):
        ypow[i] = ypow[i - 1] * y % q
    two = [pow(2, i, q) for i in range(n)]
    zz = []
    zj = z * z % q
    for _ in range(nm // n):
        zz += [zj * t % q for t in two]
        zj = zj * z % q
    return ypow, zz


def inverse_powers(y, n, q):
    """[y^0, y^-1, ..., y^-(n-1)] mod q."""
    yinv = pow(int(y) % q, -1, q)
    out, acc = [], 1
    for _ in range(n):
        out.append(acc)
        acc = acc * yinv % q
    return out


def scale_generators(hs, y, q):
    """hsp[i] = y^-i * hs[i]  (rangeproof_prover.py:77): one batched device call."""
    n = len(hs)
    yinv = pow(int(y % q), -1, q)
    ks, acc = [], 1
    for _ in range(n):
        ks.append(acc)
        acc = acc * yinv % q
    raw = nat.scalar_mul_batch_bytes(nat.pack_points(hs), nat.pack_scalars(ks), n)
    return [Point.from_bytes64(raw, 64 * i) for i in range(n)]


_ONE_B, _ZERO_B = (1).to_bytes(32, "little"), bytes(32)


def prove(vs, n, g, h, gs, hs, gammas, u, group, transcript: Transcript):
    """Same proof, byte for byte, as `_prove_python`; vectors travel as packed bytes between the C algebra and the device calls."""
    q = group.q
    if q != nat.Q:
        return _prove_python(vs, n, g, h, gs, hs, gammas, u, group, transcript)
    m = len(vs)
    nm = n * m
    # shape checks the reference makes in vector_commitment / NIProver (commitments.py:10, inner_product_prover.py:14)
    assert len(gs) == len(hs) == nm and len(gammas) == m
    bits = bytes((v.x >> i) & 1 for v in vs for i in range(n))              # aL, value-major
    t0 = transcript.digest
    alpha = mod_hash(b"alpha" + t0, q).x
    sLR = nat.mod_hash_indexed_raw(t0, 0, 2 * nm)                              # sL | sR
    sL_b, sR_b = sLR[:32 * nm], sLR[32 * nm:]
    rho = mod_hash(str(2 * n).encode() + t0, q).x                              # reference quirk: index 2*n even when m > 1
    gs_b, hs_b, h_b, g_b = nat.pack_points(gs), nat.pack_points(hs), nat.pack_point(h), nat.pack_point(g)
    base_b = gs_b + hs_b + h_b
    qm1_b = (q - 1).to_bytes(32, "little")
    aL_b = b"".join([_ONE_B if b else _ZERO_B for b in bits])
    aR_b = b"".join([_ZERO_B if b else qm1_b for b in bits])                   # aR = aL - 1 mod q
    raw = nat.msm_batch_bytes(base_b + base_b, aL_b + aR_b + nat.pack_scalar(alpha) + sLR + nat.pack_scalar(rho),
                              [0, 2 * nm + 1, 4 * nm + 2])
    A, S = Point.from_bytes64(raw, 0), Point.from_bytes64(raw, 64)
    transcript.add_list_points([A, S])
    y = transcript.get_modp(q)
    transcript.add_number(y)
    z = transcript.get_modp(q)
    transcript.add_number(z)
    yv, zv = y.x, z.x
    t1, t2 = nat.rp_prover_poly1(bits, sL_b, sR_b, n, m, yv, zv)
    t1_that = transcript.digest
    tau1 = mod_hash(b"tau1" + t1_that, q).x
    tau2 = mod_hash(b"tau2" + t1_that, q).x
    gh_b = g_b + h_b
    raw = nat.msm_batch_bytes(gh_b + gh_b, nat.pack_scalars([t1, tau1, t2, tau2]), [0, 2, 4])
    T1, T2 = Point.from_bytes64(raw, 0), Point.from_bytes64(raw, 64)
    transcript.add_list_points([T1, T2])
    x = transcript.get_modp(q)
    transcript.add_number(x)
    xv = x.x
    ls_b, rs_b, yinv_b, hsc_b, t_hat, rsy_b = nat.rp_prover_poly2(bits, sL_b, sR_b, n, m, yv, zv, xv)
    zpow = zv * zv % q
    gsum = 0
    for j in range(m):
        gsum += zpow * gammas[j].x
        zpow = zpow * zv % q
    taux = (tau2 * xv * xv + tau1 * xv + gsum) % q
    mu = (alpha + rho * xv) % q
    # P = A + x*S - mu*h + sum(-z * gs_i) + sum((z + zz_i*y^-i) * hs_i) (rangeproof_prover.py:78-86) is, in the prover's own
    # scalars, sum l_i gs_i + sum (r_i y^-i) hs_i (the h terms cancel: mu = alpha + rho x).  Protocol 1 therefore gets its
    # statement P + (x1 t_hat) u as ONE multiexp over the fixed generators (bp_ipa_statement, served by the IPA's table);
    # A and S, which change with every proof, never enter a multiexp again.
    packed = {"n": nm, "g": gs_b, "h": hs_b, "a": ls_b, "b": rs_b, "h_scale": yinv_b, "bs": rsy_b}
    inner = NIProver(gs, hs, u, None, ModP(t_hat, q), None, None, group, _packed=packed)
    return Proof(ModP(taux, q), ModP(mu, q), ModP(t_hat, q), T1, T2, A, S, inner.prove(), transcript.digest)


def _prove_python(vs, n, g, h, gs, hs, gammas, u, group, transcript: Transcript):
    q = group.q
    m = len(vs)
    nm = n * m
    aL = []
    for v in vs:
        aL += [(v.x >> i) & 1 for i in range(n)]            # reversed(bin(v).zfill(n))[:n]
    aR = [(bit - 1) % q for bit in aL]
    t0 = transcript.digest
    alpha = mod_hash(b"alpha" + t0, q).x
    sL = [mod_hash(str(i).encode() + t0, q).x for i in range(nm)]
    sR = [mod_hash(str(i).encode() + t0, q).x for i in range(nm, 2 * nm)]
    rho = mod_hash(str(2 * n).encode() + t0, q).x           # reference quirk: index 2*n even when m > 1
    # A = <aL,gs> + <aR,hs> + alpha*h ; S likewise: two (2nm+1)-term MSMs in one device pass
    base = gs + hs + [h]
    A, S = PipSECP256k1.multiexp_batch([base, base], [aL + aR + [alpha], sL + sR + [rho]])
    transcript.add_list_points([A, S])
    y = transcript.get_modp(q)
    transcript.add_number(y)
    z = transcript.get_modp(q)
    transcript.add_number(z)
    yv, zv = y.x, z.x
    ypow, zz = _position_constants(yv, zv, n, nm, q)
    ysR = [ypow[i] * sR[i] % q for i in range(nm)]
    t1 = (sum(sL[i] * (ypow[i] * (aR[i] + zv) + zz[i]) for i in range(nm))
          + sum((aL[i] - zv) * ysR[i] for i in range(nm))) % q
    t2 = sum(sL[i] * ysR[i] for i in range(nm)) % q
    t1_that = transcript.digest
    tau1 = mod_hash(b"tau1" + t1_that, q).x
    tau2 = mod_hash(b"tau2" + t1_that, q).x
    T1, T2 = PipSECP256k1.multiexp_batch([[g, h], [g, h]], [[t1, tau1], [t2, tau2]])
    transcript.add_list_points([T1, T2])
    x = transcript.get_modp(q)
    transcript.add_number(x)
    xv = x.x
    ls = [(aL[i] - zv + sL[i] * xv) % q for i in range(nm)]
    rs = [(ypow[i] * (aR[i] + zv + sR[i] * xv) + zz[i]) % q for i in range(nm)]
    t_hat = sum(a * b for a, b in zip(ls, rs)) % q
    zpow = zv * zv % q
    gsum = 0
    for j in range(m):
        gsum += zpow * gammas[j].x
        zpow = zpow * zv % q
    taux = (tau2 * xv * xv + tau1 * xv + gsum) % q
    mu = (alpha + rho * xv) % q
    # hsp_i = y^-i * hs_i (rangeproof_prover.py:77) is never materialised: its scalars are folded into the MSMs
    yinv_pow = inverse_powers(yv, nm, q)
    # P - mu*h = A + x*S + sum(-z * gs_i) + sum((z*y^i + zz_i) * y^-i * hs_i) - mu*h  (rangeproof_prover.py:78-86) is
    # handed to the IPA prover as an unevaluated MSM: Protocol 1 only needs P + (x1*t_hat)*u, one device pass.
    inner = NIProver(gs, hs, u, None, ModP(t_hat, q), [ModP(v, q) for v in ls], [ModP(v, q) for v in rs], group,
                     _h_scale=yinv_pow,
                     _P_msm=([A, S, h] + gs + hs,
                             [1, xv, -mu] + [-zv] * nm + [(zv + zz[i] * yinv_pow[i]) % q for i in range(nm)]))
    return Proof(ModP(taux, q), ModP(mu, q), ModP(t_hat, q), T1, T2, A, S, inner.prove(), transcript.digest)


class VerifierCore:
    """verify_transcript + verify shared by RangeVerifier and AggregRangeVerifier."""

    def assertThat(self, expr: bool):
        if not expr:
            raise Exception("Proof invalid")

    def verify_transcript(self):
        """rangeproof_verifier.py:42-53: A, S, T1, T2 must sit in slots 1, 2, 5, 6; y, z, x are READ
        from slots 3, 4, 7 (never re-hashed)."""
        proof = self.proof
        p = proof.taux.p
        slots = proof.transcript.split(b"&")
        self.assertThat(slots[1] == point_to_b64(proof.A))
        self.assertThat(slots[2] == point_to_b64(proof.S))
        self.y = ModP(int(slots[3]), p)
        self.z = ModP(int(slots[4]), p)
        self.assertThat(slots[5] == point_to_b64(proof.T1))
        self.assertThat(slots[6] == point_to_b64(proof.T2))
        self.x = ModP(int(slots[7]), p)

    def _verify(self, Vs):
        self.verify_transcript()
        q = self.proof.taux.p
        proof, g, h, gs, hs = self.proof, self.g, self.h, self.gs, self.hs
        nm, m = len(gs), len(Vs)
        assert m >= 1 and nm % m == 0 and len(hs) == nm          # vector_commitment / Verifier2 shapes (commitments.py:10)
        n = nm // m
        xv, yv, zv = self.x.x % q, self.y.x % q, self.z.x % q
        zpows = [pow(zv, j + 2, q) for j in range(m + 1)]
        if yv == 0:
            raise Exception("modular inverse does not exist")          # y.inv(), utils.py:69-70
        if q == nat.Q:
            # scalar preparation in C (bp_rp_verifier_scalars), vectors as packed bytes; hsp_i = y^-i * hs_i stays implicit
            yinv_pow, hsc_b, delta = nat.rp_verifier_scalars(n, m, yv, zv)
            gs_b, hs_b, g_b, h_b = nat.pack_points(gs), nat.pack_points(hs), nat.pack_point(g), nat.pack_point(h)
            raw = nat.msm_batch_bytes(
                g_b + h_b + nat.pack_points(list(Vs) + [g, proof.T1, proof.T2]) + nat.pack_points([proof.A, proof.S]) + h_b + gs_b + hs_b,
                nat.pack_scalars([proof.t_hat, proof.taux] + zpows[:m] + [delta, xv, xv * xv % q] + [1, xv, -(proof.mu.x)])
                + nat.pack_scalar(-zv) * nm + hsc_b,
                [0, 2, 2 + m + 3, 2 + m + 3 + 3 + 2 * nm])
            lhs, rhs, P_inner = (Point.from_bytes64(raw, 64 * j) for j in range(3))
        else:
            ypow, zz = _position_constants(yv, zv, n, nm, q)
            delta = ((zv - zv * zv) * sum(ypow) - sum(zpows[j] * (2 ** n - 1) for j in range(1, m + 1))) % q
            yinv_pow = inverse_powers(yv, nm, q)                             # hsp_i = y^-i * hs_i stays implicit
            # t_hat*g + taux*h == sum z^(j+2) V_j + delta*g + x*T1 + x^2*T2      (one device pass, exact compare)
            lhs, rhs, P_inner = PipSECP256k1.multiexp_batch(
                [[g, h], list(Vs) + [g, proof.T1, proof.T2], [proof.A, proof.S, h] + gs + hs],
                [[proof.t_hat, proof.taux], zpows[:m] + [delta, xv, xv * xv % q],
                 [1, xv, -(proof.mu.x)] + [-zv] * nm + [(zv + zz[i] * yinv_pow[i]) % q for i in range(nm)]])
        self.assertThat(lhs == rhs)
        return Verifier1(gs, hs, self.u, P_inner, proof.t_hat, proof.innerProof, _h_scale=yinv_pow).verify()
