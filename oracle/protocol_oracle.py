"""CPU restatement of the reference's protocol layer around the hot path.

TEST INFRASTRUCTURE ONLY (checker for tests/, smoke() and bench.py's CPU-baseline legs).
python_bulletproofs_b200 never imports this module.

Style: plain functions over Python ints (scalars, always reduced mod q) and (x, y) tuples
(points, identity = None); elliptic-curve work goes to the C oracle (oracle/ecc.py).
Every function cites the reference lines (relative to /root/reference/) it follows.
Parity pinning: tests/test_oracle.py replays tests/golden/*.json, which were produced by the
UNMODIFIED reference (oracle/gen_golden.py), through these functions and requires identical
proofs, transcripts and accept/reject decisions.

A "proof" here is a plain dict, the serialisation of SURVEY.md A.6:
  range proof : taux mu t_hat (ints)  T1 T2 A S (points)  transcript (bytes)  ip (dict)
  ip (Proof1) : u_new P_new (points)  transcript (bytes)  p2 (dict)
  p2 (Proof2) : a b (ints)  xs (ints)  Ls Rs (points)  transcript (bytes)  start (int)
"""
import base64
import hashlib

from . import ecc

Q = ecc.Q
P = ecc.P


# ---- hashing / encoding --------------------------------------------------------------------
def mod_hash(msg, q=Q, non_zero=True):
    """src/utils/utils.py:84-97 -- counter-prefixed SHA-256 with rejection sampling."""
    ctr = 0
    mask = (1 << q.bit_length()) - 1
    while True:
        ctr += 1
        x = int.from_bytes(hashlib.sha256(str(ctr).encode() + msg).digest(), "big") & mask
        if x < q and not (non_zero and x == 0):
            return x


def elliptic_hash(msg):
    """src/utils/elliptic_curve_hash.py:7-23 -- try-and-increment, md5 parity bit picks y."""
    ctr = 0
    while True:
        ctr += 1
        pre = str(ctr).encode() + msg
        x = int.from_bytes(hashlib.sha256(pre).digest(), "big")
        if x >= P:
            continue
        y = pow((x * x * x + 7) % P, (P + 1) // 4, P)
        if (y * y - x * x * x - 7) % P == 0:
            keep = int.from_bytes(hashlib.md5(pre).digest(), "big") & 1
            return (x, y) if keep else (x, P - y)


def enc_point(pt):
    """src/utils/utils.py:100-106 -- SEC1 compressed; identity is the single byte 00."""
    if pt is None:
        return b"\x00"
    return (b"\x03" if pt[1] & 1 else b"\x02") + pt[0].to_bytes(32, "big")


def b64_point(pt):
    """src/utils/utils.py:109-111."""
    return base64.b64encode(enc_point(pt))


def dec_point(b):
    """src/utils/utils.py:119-131 (the identity branch there is dead code)."""
    x = int.from_bytes(b[1:], "big")
    return ecc.lift_x(x, 0 if b[0] == 2 else 1)


def transcript_start(seed=b""):
    """src/utils/transcript.py:13-14."""
    return base64.b64encode(seed) + b"&"


def inv(x):
    """src/utils/utils.py:66-72 (egcd); q is prime so this is x^-1 mod q."""
    if x % Q == 0:
        raise Exception("modular inverse does not exist")
    return pow(x, -1, Q)


def dot(a, b):
    """src/utils/utils.py:134-137."""
    return sum(x * y for x, y in zip(a, b)) % Q


def commit(g, h, x, r):
    """src/utils/commitments.py:5-6."""
    return ecc.point_add(ecc.scalar_mul(g, x), ecc.scalar_mul(h, r))


def vec_commit(g, h, a, b, algo="bucket"):
    """src/utils/commitments.py:9-13 -> Pippenger.multiexp result semantics."""
    return ecc.msm(list(g) + list(h), list(a) + list(b), algo)


# ---- inner-product argument ----------------------------------------------------------------
def ipa_prove2(g, h, u, a, b, transcript=None):
    """FastNIProver2: src/innerproduct/inner_product_prover.py:50-110."""
    assert len(g) == len(h) == len(a) == len(b) and len(a) & (len(a) - 1) == 0
    digest = transcript_start()                       # Transcript() with empty seed  (:62)
    start = 1
    if transcript:
        digest += transcript                          # (:63-65)
        start = len(transcript.split(b"&"))
    g, h, a, b = list(g), list(h), [x % Q for x in a], [x % Q for x in b]
    xs, Ls, Rs = [], [], []
    while len(a) > 1:
        k = len(a) // 2
        cl, cr = dot(a[:k], b[k:]), dot(a[k:], b[:k])                                   # :96-97
        L = ecc.point_add(vec_commit(g[k:], h[:k], a[:k], b[k:]), ecc.scalar_mul(u, cl))  # :98
        R = ecc.point_add(vec_commit(g[:k], h[k:], a[k:], b[:k]), ecc.scalar_mul(u, cr))  # :99
        Ls.append(L)
        Rs.append(R)
        digest += b64_point(L) + b"&" + b64_point(R) + b"&"                              # :102
        x = mod_hash(digest)                                                             # :104
        xs.append(x)
        digest += str(x).encode() + b"&"                                                 # :106
        xi = inv(x)
        g = ecc.fold(g[:k], g[k:], xi, x)                                                # :107
        h = ecc.fold(h[:k], h[k:], x, xi)                                                # :108
        a = [(x * lo + xi * hi) % Q for lo, hi in zip(a[:k], a[k:])]                     # :109
        b = [(xi * lo + x * hi) % Q for lo, hi in zip(b[:k], b[k:])]                     # :110
    return {"a": a[0], "b": b[0], "xs": xs, "Ls": Ls, "Rs": Rs, "transcript": digest, "start": start}


def ipa_prove1(g, h, u, Pt, c, a, b, seed=b""):
    """NIProver (Protocol 1): src/innerproduct/inner_product_prover.py:25-45."""
    digest = transcript_start(seed)
    x = mod_hash(digest)
    digest += str(x).encode() + b"&"
    P_new = ecc.point_add(Pt, ecc.scalar_mul(u, x * c % Q))
    u_new = ecc.scalar_mul(u, x)
    return {"u_new": u_new, "P_new": P_new, "transcript": digest,
            "p2": ipa_prove2(g, h, u_new, a, b, digest)}


def ss_vector(xs, n):
    """Verifier2.get_ss: src/innerproduct/inner_product_verifier.py:91-102."""
    logn = n.bit_length() - 1
    xinv = [inv(x) for x in xs]
    out = []
    for i in range(n):
        s = 1
        for j in range(logn):
            s = s * (xs[j] if (i >> (logn - 1 - j)) & 1 else xinv[j]) % Q
        out.append(s)
    return out


def ipa_verify2(g, h, u, Pt, p2):
    """Verifier2.verify: src/innerproduct/inner_product_verifier.py:104-147."""
    n = len(g)
    logn = n.bit_length() - 1
    parts = p2["transcript"].split(b"&")
    st = p2["start"]
    for i in range(logn):                                                     # :104-125
        if parts[st + 3 * i] != b64_point(p2["Ls"][i]):
            return False
        if parts[st + 3 * i + 1] != b64_point(p2["Rs"][i]):
            return False
        want = str(mod_hash(b"&".join(parts[:st + 3 * i + 2]) + b"&")).encode()
        if not (str(p2["xs"][i]).encode() == parts[st + 3 * i + 2] == want):
            return False
    ss = ss_vector(p2["xs"], n)                                               # :133
    lhs = ecc.msm(list(g) + list(h) + [u],
                  [p2["a"] * s % Q for s in ss] + [p2["b"] * inv(s) % Q for s in ss]
                  + [p2["a"] * p2["b"] % Q])                                  # :134-139
    rhs = ecc.point_add(Pt, ecc.msm(list(p2["Ls"]) + list(p2["Rs"]),
                                    [x * x % Q for x in p2["xs"]]
                                    + [inv(x) ** 2 % Q for x in p2["xs"]]))   # :140-143
    return lhs == rhs


def ipa_verify1(g, h, u, Pt, c, ip):
    """Verifier1.verify: src/innerproduct/inner_product_verifier.py:36-58."""
    parts = ip["transcript"].split(b"&")
    if parts[1] != str(mod_hash(parts[0] + b"&")).encode():                   # :36-42
        return False
    x = int(parts[1]) % Q
    if ip["P_new"] != ecc.point_add(Pt, ecc.scalar_mul(u, x * c % Q)):        # :51
        return False
    if ip["u_new"] != ecc.scalar_mul(u, x):                                   # :52
        return False
    return ipa_verify2(g, h, ip["u_new"], ip["P_new"], ip["p2"])


# ---- range proofs (single value = aggregated with m = 1 up to the quirks noted) ------------
def _bits_le(v, n):
    """rangeproof_prover.py:42 -- reversed(bin(v).zfill(n))[:n]: the n low bits, LSB first."""
    return [(v >> i) & 1 for i in range(n)]


def _zpow2(z, i, n):
    """z^(2 + i//n) * 2^(i % n): the per-position constant of r(X) (aggreg_prover.py:98)."""
    return pow(z, 2 + i // n, Q) * pow(2, i % n, Q) % Q


def range_prove(vs, n, g, h, gs, hs, gammas, u, seed=b""):
    """AggregNIRangeProver.prove (rangeproof_aggreg_prover.py:36-146); with one value it is
    NIRangeProver.prove (rangeproof_prover.py:35-112) -- the formulas coincide for m = 1,
    including rho = H(str(2n) + t)."""
    m = len(vs)
    nm = n * m
    assert len(gs) == len(hs) == nm
    t = transcript_start(seed)
    aL = []
    for v in vs:
        aL += _bits_le(v % Q, n)
    aR = [(x - 1) % Q for x in aL]                                            # :42-45
    alpha = mod_hash(b"alpha" + t)
    A = ecc.point_add(vec_commit(gs, hs, aL, aR), ecc.scalar_mul(h, alpha))   # :47
    sL = [mod_hash(str(i).encode() + t) for i in range(nm)]
    sR = [mod_hash(str(i).encode() + t) for i in range(nm, 2 * nm)]
    rho = mod_hash(str(2 * n).encode() + t)                                   # quirk: 2*n, not 2*n*m
    S = ecc.point_add(vec_commit(gs, hs, sL, sR), ecc.scalar_mul(h, rho))
    t += b64_point(A) + b"&" + b64_point(S) + b"&"
    y = mod_hash(t)
    t += str(y).encode() + b"&"
    z = mod_hash(t)
    t += str(z).encode() + b"&"
    ypow = [pow(y, i, Q) for i in range(nm)]
    zz = [_zpow2(z, i, n) for i in range(nm)]
    # t(X) coefficients                                                        (:93-101 / :117-131)
    t1 = (dot(sL, [(ypow[i] * (aR[i] + z) + zz[i]) % Q for i in range(nm)])
          + dot([(aL[i] - z) % Q for i in range(nm)], [ypow[i] * sR[i] % Q for i in range(nm)])) % Q
    t2 = dot(sL, [ypow[i] * sR[i] % Q for i in range(nm)])
    tau1 = mod_hash(b"tau1" + t)
    tau2 = mod_hash(b"tau2" + t)
    T1 = commit(g, h, t1, tau1)
    T2 = commit(g, h, t2, tau2)
    t += b64_point(T1) + b"&" + b64_point(T2) + b"&"
    x = mod_hash(t)
    t += str(x).encode() + b"&"
    ls = [(aL[i] - z + sL[i] * x) % Q for i in range(nm)]                     # :103-112 / :133-146
    rs = [(ypow[i] * (aR[i] + z + sR[i] * x) + zz[i]) % Q for i in range(nm)]
    t_hat = dot(ls, rs)
    taux = (tau2 * x * x + tau1 * x + sum(pow(z, 2 + j, Q) * gammas[j] for j in range(m))) % Q
    mu = (alpha + rho * x) % Q
    yinv = inv(y)
    hsp = ecc.scalar_mul_batch(hs, [pow(yinv, i, Q) for i in range(nm)])      # :77 / :82
    Pt = ecc.point_add(ecc.point_add(A, ecc.scalar_mul(S, x)),
                       ecc.msm(list(gs) + hsp,
                               [(-z) % Q] * nm + [(z * ypow[i] + zz[i]) % Q for i in range(nm)]))
    ip = ipa_prove1(gs, hsp, u, ecc.point_add(Pt, ecc.scalar_mul(h, (-mu) % Q)), t_hat, ls, rs)
    return {"taux": taux, "mu": mu, "t_hat": t_hat, "T1": T1, "T2": T2, "A": A, "S": S,
            "transcript": t, "ip": ip}


def range_verify(Vs, g, h, gs, hs, u, proof):
    """AggregRangeVerifier.verify (rangeproof_aggreg_verifier.py:42-108); with one commitment
    it is RangeVerifier.verify (rangeproof_verifier.py:42-97).  Returns False where the
    reference raises Exception("Proof invalid"); a non-numeric y/z/x slot raises ValueError
    exactly as the reference's int() does."""
    parts = proof["transcript"].split(b"&")
    if parts[1] != b64_point(proof["A"]) or parts[2] != b64_point(proof["S"]):
        return False
    y, z = int(parts[3]) % Q, int(parts[4]) % Q
    if parts[5] != b64_point(proof["T1"]) or parts[6] != b64_point(proof["T2"]):
        return False
    x = int(parts[7]) % Q
    nm, m = len(gs), len(Vs)
    n = nm // m
    ypow = [pow(y, i, Q) for i in range(nm)]
    delta = ((z - z * z) * sum(ypow) - sum(pow(z, j + 2, Q) * (2 ** n - 1) for j in range(1, m + 1))) % Q
    yinv = inv(y)
    hsp = ecc.scalar_mul_batch(hs, [pow(yinv, i, Q) for i in range(nm)])
    lhs = commit(g, h, proof["t_hat"], proof["taux"])
    rhs = ecc.msm(list(Vs) + [g, proof["T1"], proof["T2"]],
                  [pow(z, j + 2, Q) for j in range(m)] + [delta, x, x * x % Q])
    if lhs != rhs:
        return False
    Pt = ecc.point_add(ecc.point_add(proof["A"], ecc.scalar_mul(proof["S"], x)),
                       ecc.msm(list(gs) + hsp,
                               [(-z) % Q] * nm + [(z * ypow[i] + _zpow2(z, i, n)) % Q for i in range(nm)]))
    return ipa_verify1(gs, hsp, u, ecc.point_add(Pt, ecc.scalar_mul(h, (-proof["mu"]) % Q)),
                       proof["t_hat"], proof["ip"])


# ---- JSON (de)serialisation of proofs, shared by gen_golden.py and the tests -----------------
def _pt(pt):
    return enc_point(pt).hex()


def _unpt(s):
    b = bytes.fromhex(s)
    return None if b == b"\x00" else dec_point(b)


def proof2_to_json(p2):
    return {"a": str(p2["a"]), "b": str(p2["b"]), "xs": [str(x) for x in p2["xs"]],
            "Ls": [_pt(p) for p in p2["Ls"]], "Rs": [_pt(p) for p in p2["Rs"]],
            "transcript": p2["transcript"].decode("latin1"), "start": p2["start"]}


def proof2_from_json(j):
    return {"a": int(j["a"]), "b": int(j["b"]), "xs": [int(x) for x in j["xs"]],
            "Ls": [_unpt(p) for p in j["Ls"]], "Rs": [_unpt(p) for p in j["Rs"]],
            "transcript": j["transcript"].encode("latin1"), "start": j["start"]}


def proof1_to_json(ip):
    return {"u_new": _pt(ip["u_new"]), "P_new": _pt(ip["P_new"]),
            "transcript": ip["transcript"].decode("latin1"), "p2": proof2_to_json(ip["p2"])}


def proof1_from_json(j):
    return {"u_new": _unpt(j["u_new"]), "P_new": _unpt(j["P_new"]),
            "transcript": j["transcript"].encode("latin1"), "p2": proof2_from_json(j["p2"])}


def range_to_json(pr):
    return {"taux": str(pr["taux"]), "mu": str(pr["mu"]), "t_hat": str(pr["t_hat"]),
            "T1": _pt(pr["T1"]), "T2": _pt(pr["T2"]), "A": _pt(pr["A"]), "S": _pt(pr["S"]),
            "transcript": pr["transcript"].decode("latin1"), "ip": proof1_to_json(pr["ip"])}


def range_from_json(j):
    return {"taux": int(j["taux"]), "mu": int(j["mu"]), "t_hat": int(j["t_hat"]),
            "T1": _unpt(j["T1"]), "T2": _unpt(j["T2"]), "A": _unpt(j["A"]), "S": _unpt(j["S"]),
            "transcript": j["transcript"].encode("latin1"), "ip": proof1_from_json(j["ip"])}
