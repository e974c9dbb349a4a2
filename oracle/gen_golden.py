"""Generate tests/golden/*.json by running the UNMODIFIED reference.

TEST INFRASTRUCTURE ONLY.  Run in the build container (where /root/reference exists):

    PYTHONDONTWRITEBYTECODE=1 python oracle/gen_golden.py [--only msm,ipa_small,...]

It puts oracle/fastecdsa_standin (a plain-Python re-creation of the third-party `fastecdsa`
API) and /root/reference on sys.path, imports the reference's own
`src` package, and records inputs + outputs of the reference's public entry points:
  * Pippenger.multiexp            (src/pippenger/pippenger.py:22)
  * NIProver / FastNIProver2 / Verifier1 / Verifier2   (src/innerproduct/*)
  * NIRangeProver / RangeVerifier / Aggreg*            (src/rangeproofs/*)
The reference cannot travel to the GPU box, so the vectors are committed.
Inputs follow the construction of the reference's own tests (src/tests/test_*.py) with the
os.urandom seeds replaced by the fixed byte strings of SURVEY.md 8(d).
"""
import argparse
import io
import contextlib
import json
import os
import random
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("BP_REFERENCE", "/root/reference")
sys.dont_write_bytecode = True
sys.path.insert(0, os.path.join(HERE, "fastecdsa_standin"))
sys.path.insert(0, REF)
sys.path.insert(0, ROOT)

from fastecdsa.curve import secp256k1 as CURVE          # noqa: E402  (the stand-in)
from fastecdsa.point import Point                        # noqa: E402
from src.pippenger import PipSECP256k1                   # noqa: E402  (the reference)
from src.utils.utils import mod_hash, ModP, inner_product, point_to_bytes  # noqa: E402
from src.utils.commitments import vector_commitment, commitment            # noqa: E402
from src.utils.elliptic_curve_hash import elliptic_hash                    # noqa: E402
from src.innerproduct.inner_product_prover import NIProver, FastNIProver2  # noqa: E402
from src.innerproduct.inner_product_verifier import Verifier1, Verifier2   # noqa: E402
from src.rangeproofs import (NIRangeProver, RangeVerifier, AggregNIRangeProver,  # noqa: E402
                             AggregRangeVerifier)

q = CURVE.q
p = CURVE.p
OUT = os.path.join(ROOT, "tests", "golden")


def enc(pt):
    return point_to_bytes(pt).hex()


def quiet(fn):
    """Run a verifier; the reference prints "OK" on success and raises on failure."""
    buf = io.StringIO()
    try:
        with contextlib.redirect_stdout(buf):
            return bool(fn())
    except Exception as e:  # noqa: BLE001
        if str(e) == "Proof invalid":
            return False
        raise


def dump(name, obj):
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, name + ".json")
    with open(path, "w") as f:
        json.dump(obj, f, indent=0, separators=(",", ":"))
        f.write("\n")
    print("wrote", path, os.path.getsize(path), "bytes", flush=True)


# ---- serialisation of reference objects --------------------------------------
def p2_json(p2):
    return {"a": str(p2.a.x % q), "b": str(p2.b.x % q), "xs": [str(x.x % q) for x in p2.xs],
            "Ls": [enc(P_) for P_ in p2.Ls], "Rs": [enc(P_) for P_ in p2.Rs],
            "transcript": p2.transcript.decode("latin1"), "start": p2.start_transcript}


def p1_json(p1):
    return {"u_new": enc(p1.u_new), "P_new": enc(p1.P_new),
            "transcript": p1.transcript.decode("latin1"), "p2": p2_json(p1.proof2)}


def range_json(pr):
    return {"taux": str(pr.taux.x % q), "mu": str(pr.mu.x % q), "t_hat": str(pr.t_hat.x % q),
            "T1": enc(pr.T1), "T2": enc(pr.T2), "A": enc(pr.A), "S": enc(pr.S),
            "transcript": pr.transcript.decode("latin1"), "ip": p1_json(pr.innerProof)}


# ---- MSM vectors ----------------------------------------------------------------------------
def c3_inputs(lgn, n=None):
    """SURVEY.md 8(d) 'C3 synthetic inputs' (seeded; same lift as src/utils/utils.py:127-131)."""
    rng = random.Random(0xB2000000 + lgn)
    n = n if n is not None else 1 << lgn
    ks = [rng.getrandbits(256) % q for _ in range(n)]
    pts = []
    while len(pts) < n:
        x = rng.getrandbits(256)
        if x >= p:
            continue
        y = pow((x ** 3 + 7) % p, (p + 1) // 4, p)
        if (y * y - x ** 3 - 7) % p:
            continue
        if rng.getrandbits(1):
            y = p - y
        pts.append(Point(x, y, CURVE))
    return pts, ks


def gen_msm():
    cases = []
    for lgn in range(0, 11):
        pts, ks = c3_inputs(lgn)
        t = time.time()
        r = PipSECP256k1.multiexp(pts, ks)
        cases.append({"kind": "c3", "lgn": lgn, "n": len(pts), "out": enc(r)})
        print("msm 2^%d" % lgn, round(time.time() - t, 2), "s", flush=True)
    for n in (3, 12, 19, 20, 129, 257):
        pts, ks = c3_inputs(40 + n, n)
        cases.append({"kind": "c3", "lgn": 40 + n, "n": n, "out": enc(PipSECP256k1.multiexp(pts, ks))})
    # explicit degenerate inputs: identity, duplicates, P and -P, zero / >= q scalars, ModP scalars
    pts, ks = c3_inputs(99, 24)
    pts[1] = Point.IDENTITY_ELEMENT
    pts[3] = pts[2]
    ks[3] = ks[2]
    pts[5] = -pts[4]
    ks[5] = ks[4]
    pts[7] = pts[6]
    ks[7] = (q - ks[6]) % q
    ks[8] = 0
    ks[9] = q
    ks[10] = q + 5
    ks[11] = 2 ** 256 - 1
    ks[12] = 1
    ks[13] = q - 1
    ks[14] = ModP(ks[14], q)
    ks[15] = 2 ** 300 + 17            # Python ints are unbounded: multiexp reduces mod q (:26)
    for i in range(16, 24):
        pts[i] = pts[16]               # eight copies of one point, eight different scalars
    r = PipSECP256k1.multiexp(pts, ks)
    cases.append({"kind": "explicit", "name": "degenerate24",
                  "pts": [enc(P_) for P_ in pts], "ks": [str(k.x) if isinstance(k, ModP) else str(k) for k in ks],
                  "out": enc(r)})
    # all scalars equal / range-proof shaped scalars (rangeproof_prover.py:42-47,83)
    pts, ks = c3_inputs(98, 128)
    same = [ks[0]] * 128
    cases.append({"kind": "explicit", "name": "same_scalar128", "pts": [enc(P_) for P_ in pts],
                  "ks": [str(k) for k in same], "out": enc(PipSECP256k1.multiexp(pts, same))})
    bits = [(0xDEADBEEFCAFEF00D >> i) & 1 for i in range(64)]
    shaped = bits + [(b - 1) % q for b in bits]
    cases.append({"kind": "explicit", "name": "bits_and_minus_one128", "pts": [enc(P_) for P_ in pts],
                  "ks": [str(k) for k in shaped], "out": enc(PipSECP256k1.multiexp(pts, shaped))})
    zero = [0] * 128
    cases.append({"kind": "explicit", "name": "all_zero128", "pts": [enc(P_) for P_ in pts],
                  "ks": [str(k) for k in zero], "out": enc(PipSECP256k1.multiexp(pts, zero))})
    cases.append({"kind": "explicit", "name": "empty", "pts": [], "ks": [],
                  "out": enc(PipSECP256k1.multiexp([], []))})
    dump("msm", {"source": "Pippenger.multiexp, src/pippenger/pippenger.py:22-61", "cases": cases})


# ---- inner-product argument -----------------------------------------------------------------
def ipa_inputs(N, seeds):
    """Construction of src/tests/test_innerprod.py:104-116."""
    g = [elliptic_hash(str(i).encode() + seeds[0], CURVE) for i in range(N)]
    h = [elliptic_hash(str(i).encode() + seeds[1], CURVE) for i in range(N)]
    u = elliptic_hash(seeds[2], CURVE)
    a = [mod_hash(str(i).encode() + seeds[3], q) for i in range(N)]
    b = [mod_hash(str(i).encode() + seeds[4], q) for i in range(N)]
    return g, h, u, a, b


def gen_ipa(name, sizes, seed_prefix):
    cases = []
    for N in sizes:
        seeds = [seed_prefix + str(i).encode() for i in range(6)]
        g, h, u, a, b = ipa_inputs(N, seeds)
        P_ = vector_commitment(g, h, a, b)
        c = inner_product(a, b)
        t = time.time()
        proof = NIProver(g, h, u, P_, c, a, b, CURVE, seeds[5]).prove()
        tp = time.time() - t
        ok = quiet(Verifier1(g, h, u, P_, c, proof).verify)
        # stand-alone Protocol 2 (src/tests/test_innerprod.py:16-31): P includes c*u
        P2 = P_ + c * u
        proof2 = FastNIProver2(g, h, u, P2, a, b, CURVE).prove()
        ok2 = quiet(Verifier2(g, h, u, P2, proof2).verify)
        # soundness smoke (test_innerprod.py:138-224): wrong c must be rejected
        bad = quiet(Verifier1(g, h, u, P_, c + ModP(1, q), proof).verify)
        cases.append({"N": N, "seeds": [s.decode() for s in seeds], "P": enc(P_), "c": str(c.x % q),
                      "proof1": p1_json(proof), "verify1": ok, "verify1_wrong_c": bad,
                      "P2": enc(P2), "proof2": p2_json(proof2), "verify2": ok2})
        print("ipa N=%d prove %.1fs verify=%s/%s bad=%s" % (N, tp, ok, ok2, bad), flush=True)
    dump(name, {"source": "NIProver/FastNIProver2/Verifier1/Verifier2, src/innerproduct/*", "cases": cases})


# ---- range proofs ---------------------------------------------------------------------------
def range_inputs(nm, seeds):
    """Construction of src/tests/test_rangeproofs.py:20-29."""
    gs = [elliptic_hash(str(i).encode() + seeds[0], CURVE) for i in range(nm)]
    hs = [elliptic_hash(str(i).encode() + seeds[1], CURVE) for i in range(nm)]
    g = elliptic_hash(seeds[2], CURVE)
    h = elliptic_hash(seeds[3], CURVE)
    u = elliptic_hash(seeds[4], CURVE)
    return gs, hs, g, h, u


def flip_digit(proof, attr):
    """test_rangeproofs.py:91-114 style tamper: change one decimal digit of a scalar field."""
    val = getattr(proof, attr)
    s = str(val.x)
    d = "1" if s[-1] != "1" else "2"
    setattr(proof, attr, ModP(int(s[:-1] + d), q))


def gen_range(name, configs):
    cases = []
    for cfg in configs:
        n, seed_prefix, v = cfg["n"], cfg["seed"], cfg["v"]
        seeds = [seed_prefix + str(i).encode() for i in range(7)]
        gs, hs, g, h, u = range_inputs(n, seeds)
        gamma = mod_hash(seeds[5], q)
        vm = ModP(v, q)
        V = commitment(g, h, vm, gamma)
        t = time.time()
        proof = NIRangeProver(vm, n, g, h, gs, hs, gamma, u, CURVE, seeds[6]).prove()
        tp = time.time() - t
        ok = quiet(RangeVerifier(V, g, h, gs, hs, u, proof).verify)
        pj = range_json(proof)
        wrongV = quiet(RangeVerifier(V + g, g, h, gs, hs, u, proof).verify)
        flip_digit(proof, "t_hat")
        bad_that = quiet(RangeVerifier(V, g, h, gs, hs, u, proof).verify)
        cases.append({"n": n, "v": str(v), "seeds": [s.decode() for s in seeds], "V": enc(V),
                      "gamma": str(gamma.x), "proof": pj, "verify": ok,
                      "verify_wrong_V": wrongV, "verify_t_hat_flipped": bad_that})
        print("range n=%d v=%d prove %.1fs verify=%s wrongV=%s flipped=%s" % (n, v, tp, ok, wrongV, bad_that), flush=True)
    dump(name, {"source": "NIRangeProver/RangeVerifier, src/rangeproofs/rangeproof_{prover,verifier}.py", "cases": cases})


def gen_aggreg(name, configs):
    cases = []
    for cfg in configs:
        n, m, seed_prefix = cfg["n"], cfg["m"], cfg["seed"]
        seeds = [seed_prefix + str(i).encode() for i in range(7)]
        gs, hs, g, h, u = range_inputs(n * m, seeds)
        vs_int = [cfg["v"](j) for j in range(m)]
        vs = [ModP(v, q) for v in vs_int]
        gammas = [mod_hash(seeds[5], q) for _ in range(m)]     # test_aggreg_rangeproofs.py:30
        Vs = [commitment(g, h, vs[i], gammas[i]) for i in range(m)]
        t = time.time()
        proof = AggregNIRangeProver(vs, n, g, h, gs, hs, gammas, u, CURVE, seeds[6]).prove()
        tp = time.time() - t
        ok = quiet(AggregRangeVerifier(Vs, g, h, gs, hs, u, proof).verify)
        Vbad = list(Vs)
        Vbad[-1] = Vbad[-1] + h
        wrongV = quiet(AggregRangeVerifier(Vbad, g, h, gs, hs, u, proof).verify)
        cases.append({"n": n, "m": m, "vs": [str(v) for v in vs_int], "seeds": [s.decode() for s in seeds],
                      "Vs": [enc(V) for V in Vs], "gammas": [str(x.x) for x in gammas],
                      "proof": range_json(proof), "verify": ok, "verify_wrong_V": wrongV})
        print("aggreg m=%d n=%d prove %.1fs verify=%s wrongV=%s" % (m, n, tp, ok, wrongV), flush=True)
    dump(name, {"source": "AggregNIRangeProver/AggregRangeVerifier, src/rangeproofs/rangeproof_aggreg_*.py", "cases": cases})


GOLD = 0x9E3779B97F4A7C15

JOBS = {
    "msm": gen_msm,
    "ipa_small": lambda: gen_ipa("ipa_small", [1, 2, 4, 8, 16, 64], b"ipas"),
    "ipa_c2": lambda: gen_ipa("ipa_c2", [1024], b"ipa"),
    "range_small": lambda: gen_range("range_small", [
        {"n": 2, "seed": b"rs2_", "v": 3}, {"n": 8, "seed": b"rs8_", "v": 200},
        {"n": 16, "seed": b"rs16_", "v": 65535}, {"n": 16, "seed": b"rx16_", "v": 65536},   # out of range
        {"n": 32, "seed": b"rs32_", "v": 0}]),
    "range_c1": lambda: gen_range("range_c1", [{"n": 64, "seed": b"seed", "v": 2 ** 63 + 12345}]),
    "aggreg_small": lambda: gen_aggreg("aggreg_small", [
        {"n": 8, "m": 2, "seed": b"ag2_", "v": lambda j: (GOLD * (j + 1)) % 2 ** 8},
        {"n": 16, "m": 4, "seed": b"ag4_", "v": lambda j: (GOLD * (j + 1)) % 2 ** 16},
        {"n": 4, "m": 8, "seed": b"ag8_", "v": lambda j: j if j != 3 else 16}]),              # one value out of range
    "aggreg_c4": lambda: gen_aggreg("aggreg_c4", [
        {"n": 64, "m": 16, "seed": b"agg", "v": lambda j: (GOLD * (j + 1)) % 2 ** 64}]),
}

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    names = [s for s in args.only.split(",") if s] or list(JOBS)
    for nme in names:
        t0 = time.time()
        JOBS[nme]()
        print("== %s done in %.1fs" % (nme, time.time() - t0), flush=True)
