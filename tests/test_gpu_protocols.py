"""Parity of the GPU-backed provers / verifiers (the reference's class API re-hosted on libbpgpu)
with the golden proofs recorded from the unmodified reference and with the CPU oracle:
identical serialised proofs, identical transcripts, identical accept/reject."""
import contextlib
import ctypes
import io
import random

import pytest

from oracle import ecc, protocol_oracle as po
from helpers import gens, ipa_inputs
from python_bulletproofs_b200 import Point, secp256k1, _native as nat
from python_bulletproofs_b200.utils import ModP, commitment, vector_commitment, inner_product, mod_hash, elliptic_hash
from python_bulletproofs_b200.innerproduct import NIProver, FastNIProver2, Verifier1, Verifier2, Proof2
from python_bulletproofs_b200.rangeproofs import (NIRangeProver, RangeVerifier, AggregNIRangeProver,
                                                   AggregRangeVerifier, verify_range_proofs_batch)

pytestmark = pytest.mark.gpu
Q = ecc.Q


def P_(t):
    return Point.IDENTITY_ELEMENT if t is None else Point(t[0], t[1], secp256k1)


def T_(p):
    return None if p.curve is None else (p.x, p.y)


def M_(v):
    return ModP(v, Q)


def quiet(fn):
    with contextlib.redirect_stdout(io.StringIO()) as buf:
        try:
            ok = fn()
        except Exception as e:   # noqa: BLE001
            if str(e) != "Proof invalid":
                raise
            return False
    assert ok is True and buf.getvalue() == "OK\n"     # inner_product_verifier.py:146-147
    return True


def p2_json(p2):
    return po.proof2_to_json({"a": p2.a.x, "b": p2.b.x, "xs": [x.x for x in p2.xs], "Ls": [T_(p) for p in p2.Ls],
                              "Rs": [T_(p) for p in p2.Rs], "transcript": p2.transcript, "start": p2.start_transcript})


def p1_json(p1):
    return {"u_new": po.enc_point(T_(p1.u_new)).hex(), "P_new": po.enc_point(T_(p1.P_new)).hex(),
            "transcript": p1.transcript.decode("latin1"), "p2": p2_json(p1.proof2)}


def range_json(pr):
    return {"taux": str(pr.taux.x), "mu": str(pr.mu.x), "t_hat": str(pr.t_hat.x),
            "T1": po.enc_point(T_(pr.T1)).hex(), "T2": po.enc_point(T_(pr.T2)).hex(),
            "A": po.enc_point(T_(pr.A)).hex(), "S": po.enc_point(T_(pr.S)).hex(),
            "transcript": pr.transcript.decode("latin1"), "ip": p1_json(pr.innerProof)}


def test_host_hashes_match_oracle():
    for msg in (b"", b"abc", b"&", b"x" * 1000):
        assert mod_hash(msg, Q).x == po.mod_hash(msg)
        out = ctypes.create_string_buffer(32)
        nat.check(nat.load().bp_mod_hash(msg, len(msg), out))
        assert int.from_bytes(out.raw, "little") == po.mod_hash(msg)
        assert T_(elliptic_hash(msg, secp256k1)) == po.elliptic_hash(msg)
    for pt in (None, ecc.G, ecc.py_mul(ecc.G, 12345)):
        out = ctypes.create_string_buffer(45)
        ln = ctypes.c_size_t()
        nat.check(nat.load().bp_point_to_b64(ecc.pack_point(pt), out, ctypes.byref(ln)))
        assert out.raw[:ln.value] == po.b64_point(pt)


@pytest.mark.parametrize("n", [2, 4, 16, 64, 256])
def test_fold_round_vs_oracle(n):
    g, h, u, a, b = ipa_inputs(min(n, 16), ["f0", "f1", "f2", "f3", "f4"])
    rng = random.Random(n)
    g = (g * (n // len(g) + 1))[:n]
    h = (h * (n // len(h) + 1))[:n]
    a = [rng.getrandbits(256) % Q for _ in range(n)]
    b = [rng.getrandbits(256) % Q for _ in range(n)]
    x = rng.getrandbits(256) % Q
    k = n // 2
    go, ho = ctypes.create_string_buffer(64 * k), ctypes.create_string_buffer(64 * k)
    ao, bo = ctypes.create_string_buffer(32 * k), ctypes.create_string_buffer(32 * k)
    nat.check(nat.load().bp_ipa_fold_round(ecc.pack_points(g), ecc.pack_points(h), ecc.pack_scalars(a), ecc.pack_scalars(b), n,
                                           x.to_bytes(32, "little"), go, ho, ao, bo))
    xi = pow(x, -1, Q)
    assert ecc.unpack_points(go.raw, k) == ecc.fold(g[:k], g[k:], xi, x)          # inner_product_prover.py:107
    assert ecc.unpack_points(ho.raw, k) == ecc.fold(h[:k], h[k:], x, xi)          # :108
    assert nat.unpack_scalars(ao.raw, k) == [(x * lo + xi * hi) % Q for lo, hi in zip(a[:k], a[k:])]   # :109
    assert nat.unpack_scalars(bo.raw, k) == [(xi * lo + x * hi) % Q for lo, hi in zip(b[:k], b[k:])]   # :110


@pytest.mark.parametrize("name", ["ipa_small", "ipa_c2"])
def test_ipa_golden(golden, name):
    for case in golden(name)["cases"]:
        N = case["N"]
        g, h, u, a, b = ipa_inputs(N, case["seeds"])
        g, h, u = [P_(t) for t in g], [P_(t) for t in h], P_(u)
        a, b = [M_(v) for v in a], [M_(v) for v in b]
        Pt = vector_commitment(g, h, a, b)
        assert po.enc_point(T_(Pt)).hex() == case["P"]
        c = inner_product(a, b)
        assert str(c) == case["c"]
        proof = NIProver(g, h, u, Pt, c, a, b, secp256k1, case["seeds"][5].encode()).prove()
        assert p1_json(proof) == case["proof1"]
        assert quiet(Verifier1(g, h, u, Pt, c, proof).verify) is case["verify1"]
        assert quiet(Verifier1(g, h, u, Pt, c + M_(1), proof).verify) is case["verify1_wrong_c"]
        P2 = Pt + c * u
        assert po.enc_point(T_(P2)).hex() == case["P2"]
        proof2 = FastNIProver2(g, h, u, P2, a, b, secp256k1).prove()
        assert p2_json(proof2) == case["proof2"]
        assert quiet(Verifier2(g, h, u, P2, proof2).verify) is case["verify2"]


@pytest.mark.parametrize("fast", [1, 0])
@pytest.mark.parametrize("mode", [0, 2, 1])
def test_ipa_prover_round_loop_modes(golden, mode, fast):
    """bp_ipa_set_graphs: the three ways the host's Fiat-Shamir step meets the stream (per-round synchronisation, one CUDA
    graph with host nodes, mapped-flag handshake) produce the same bytes -- checked against the golden C2 proof (n = 2^10,
    recorded from the unmodified reference) and the small cases, several proofs in a row so that mode 2 replays its graph."""
    lib = nat.load()
    try:
        nat.check(lib.bp_ipa_set_graphs(mode))
        nat.check(lib.bp_ipa_set_fast_rounds(fast))      # two-launch table rounds (k_ipa_round_prep + ticketed k_fb_msm) or the four-launch form
        for name in ("ipa_c2", "ipa_small"):
            for case in golden(name)["cases"]:
                N = case["N"]
                g, h, u, a, b = ipa_inputs(N, case["seeds"])
                g, h, u = [P_(t) for t in g], [P_(t) for t in h], P_(u)
                a, b = [M_(v) for v in a], [M_(v) for v in b]
                P2 = vector_commitment(g, h, a, b) + inner_product(a, b) * u
                for _ in range(3):
                    assert p2_json(FastNIProver2(g, h, u, P2, a, b, secp256k1).prove()) == case["proof2"]
    finally:
        lib.bp_ipa_set_graphs(1)
        lib.bp_ipa_set_fast_rounds(1)


def test_ipa_soundness_smoke():
    """src/tests/test_innerprod.py:33-98,226-268: wrong P / a / b / u / transcript => Proof invalid."""
    N = 8
    g, h, u, a, b = ipa_inputs(N, ["s0", "s1", "s2", "s3", "s4"])
    g, h, u = [P_(t) for t in g], [P_(t) for t in h], P_(u)
    a, b = [M_(v) for v in a], [M_(v) for v in b]
    c = inner_product(a, b)
    P2 = vector_commitment(g, h, a, b) + c * u
    proof = FastNIProver2(g, h, u, P2, a, b, secp256k1).prove()
    assert quiet(Verifier2(g, h, u, P2, proof).verify)
    assert not quiet(Verifier2(g, h, u, P2 + u, proof).verify)
    assert not quiet(Verifier2(g, h, 2 * u, P2, proof).verify)
    for field in ("a", "b"):
        bad = Proof2(proof.a, proof.b, proof.xs, proof.Ls, proof.Rs, proof.transcript, proof.start_transcript)
        setattr(bad, field, getattr(bad, field) + M_(1))
        assert not quiet(Verifier2(g, h, u, P2, bad).verify)
    t = bytearray(proof.transcript)
    t[len(t) // 2] ^= 1
    bad = Proof2(proof.a, proof.b, proof.xs, proof.Ls, proof.Rs, bytes(t), proof.start_transcript)
    assert not quiet(Verifier2(g, h, u, P2, bad).verify)
    with pytest.raises(AssertionError):
        FastNIProver2(g[:3], h[:3], u, P2, a[:3], b[:3], secp256k1)


@pytest.mark.parametrize("name", ["range_small", "range_c1"])
def test_range_golden(golden, name):
    for case in golden(name)["cases"]:
        n, v = case["n"], int(case["v"])
        gs, hs, g, h, u = gens(n, case["seeds"])
        gs, hs, g, h, u = [P_(t) for t in gs], [P_(t) for t in hs], P_(g), P_(h), P_(u)
        gamma = mod_hash(case["seeds"][5].encode(), Q)
        V = commitment(g, h, M_(v), gamma)
        assert po.enc_point(T_(V)).hex() == case["V"]
        proof = NIRangeProver(M_(v), n, g, h, gs, hs, gamma, u, secp256k1, case["seeds"][6].encode()).prove()
        assert range_json(proof) == case["proof"]
        assert quiet(RangeVerifier(V, g, h, gs, hs, u, proof).verify) is case["verify"]
        assert quiet(RangeVerifier(V + g, g, h, gs, hs, u, proof).verify) is case["verify_wrong_V"]
        s = str(proof.t_hat.x)
        proof.t_hat = M_(int(s[:-1] + ("1" if s[-1] != "1" else "2")))
        assert quiet(RangeVerifier(V, g, h, gs, hs, u, proof).verify) is case["verify_t_hat_flipped"]


@pytest.mark.parametrize("name", ["aggreg_small", "aggreg_c4"])
def test_aggreg_golden(golden, name):
    for case in golden(name)["cases"]:
        n, m = case["n"], case["m"]
        gs, hs, g, h, u = gens(n * m, case["seeds"])
        gs, hs, g, h, u = [P_(t) for t in gs], [P_(t) for t in hs], P_(g), P_(h), P_(u)
        vs = [M_(int(v)) for v in case["vs"]]
        gammas = [M_(int(x)) for x in case["gammas"]]
        Vs = [commitment(g, h, vs[i], gammas[i]) for i in range(m)]
        assert [po.enc_point(T_(V)).hex() for V in Vs] == case["Vs"]
        proof = AggregNIRangeProver(vs, n, g, h, gs, hs, gammas, u, secp256k1, case["seeds"][6].encode()).prove()
        assert range_json(proof) == case["proof"]
        assert quiet(AggregRangeVerifier(Vs, g, h, gs, hs, u, proof).verify) is case["verify"]
        Vbad = Vs[:-1] + [Vs[-1] + h]
        assert quiet(AggregRangeVerifier(Vbad, g, h, gs, hs, u, proof).verify) is case["verify_wrong_V"]


def test_range_transcript_tamper_value_error(golden):
    """A non-numeric y slot raises ValueError from int(), as in rangeproof_verifier.py:49."""
    case = golden("range_small")["cases"][1]
    n = case["n"]
    gs, hs, g, h, u = gens(n, case["seeds"])
    gs, hs, g, h, u = [P_(t) for t in gs], [P_(t) for t in hs], P_(g), P_(h), P_(u)
    gamma = mod_hash(case["seeds"][5].encode(), Q)
    v = M_(int(case["v"]))
    V = commitment(g, h, v, gamma)
    proof = NIRangeProver(v, n, g, h, gs, hs, gamma, u, secp256k1, case["seeds"][6].encode()).prove()
    parts = proof.transcript.split(b"&")
    parts[3] = b"12x4"
    proof.transcript = b"&".join(parts)
    with pytest.raises(ValueError):
        RangeVerifier(V, g, h, gs, hs, u, proof).verify()


@pytest.mark.parametrize("n,count", [(8, 40), (64, 64), (128, 16), (2, 16)])
def test_batch_verify_decisions(n, count):
    """bp_rp_verify_batch vs the oracle verifier, proof by proof, incl. corrupted proofs of every kind (>= 64
    decisions at 64 bits).  The batch runs three times: bucket method (first sight of the generator set), byte tables, 16-bit
    tables -- the decisions must not depend on the path."""
    seeds = ["b0", "b1", "b2", "b3", "b4"]
    ogs, ohs, og, oh, ou = gens(n, seeds)
    gs, hs, g, h, u = [P_(t) for t in ogs], [P_(t) for t in ohs], P_(og), P_(oh), P_(ou)
    rng = random.Random(5)
    Vs, proofs, oVs, oproofs = [], [], [], []
    for i in range(count):
        v = rng.getrandbits(n)
        gamma = mod_hash(b"gamma%d" % i, Q)
        V = commitment(g, h, M_(v), gamma)
        pr = NIRangeProver(M_(v), n, g, h, gs, hs, gamma, u, secp256k1, b"p%d" % i).prove()
        kind = i % 8
        if kind == 1:
            s = str(pr.t_hat.x); pr.t_hat = M_(int(s[:-1] + ("1" if s[-1] != "1" else "2")))
        elif kind == 2:
            V = V + g
        elif kind == 3:
            pr.innerProof.proof2.a = pr.innerProof.proof2.a + M_(1)
        elif kind == 4:
            t = bytearray(pr.innerProof.proof2.transcript); t[-5] = ord("1") if t[-5] != ord("1") else ord("2")
            pr.innerProof.proof2.transcript = bytes(t)
        elif kind == 5:
            pr.mu = pr.mu + M_(1)
        elif kind == 6:
            pr.innerProof.u_new = pr.innerProof.u_new + g
        Vs.append(V); proofs.append(pr)
        oVs.append(T_(V)); oproofs.append(po.range_from_json(range_json(pr)))
    want = [po.range_verify([oV], og, oh, ogs, ohs, ou, opr) for oV, opr in zip(oVs, oproofs)]
    for _ in range(3):
        assert verify_range_proofs_batch(Vs, g, h, gs, hs, u, proofs) == want
    assert want.count(True) == sum(1 for i in range(count) if i % 8 in (0, 7))
    # the class API agrees proof by proof
    for i in (0, 1, 4):
        assert quiet(RangeVerifier(Vs[i], g, h, gs, hs, u, proofs[i]).verify) is want[i]


@pytest.mark.parametrize("n,m,count", [(8, 2, 24), (16, 4, 16), (8, 1, 16), (64, 16, 3)])
def test_aggreg_batch_verify_decisions(n, m, count):
    """bp_rp_verify_aggreg_batch (aggregated proofs, m values x n bits, up to 1024 generators per side) vs the oracle's
    AggregRangeVerifier restatement, proof by proof, incl. corrupted proofs of every kind.  Run three times: bucket method over the
    generator rows (first sight of the set), then the set's byte table -- the decisions must not depend on the path."""
    from python_bulletproofs_b200.rangeproofs import AggregNIRangeProver, AggregRangeVerifier, verify_aggreg_range_proofs_batch
    seeds = ["ab%d_%d_%d" % (n, m, i) for i in range(5)]
    nm = n * m
    ogs, ohs, og, oh, ou = gens(nm, seeds)
    gs, hs, g, h, u = [P_(t) for t in ogs], [P_(t) for t in ohs], P_(og), P_(oh), P_(ou)
    rng = random.Random(1000 * n + m)
    Vs_list, proofs, oVs_list, oproofs = [], [], [], []
    for i in range(count):
        vs = [M_(rng.getrandbits(n)) for _ in range(m)]
        gammas = [mod_hash(b"ga%d_%d" % (i, j), Q) for j in range(m)]
        Vs = [commitment(g, h, vs[j], gammas[j]) for j in range(m)]
        pr = AggregNIRangeProver(vs, n, g, h, gs, hs, gammas, u, secp256k1, b"ap%d" % i).prove()
        kind = i % 8
        if kind == 1:
            s = str(pr.t_hat.x); pr.t_hat = M_(int(s[:-1] + ("1" if s[-1] != "1" else "2")))
        elif kind == 2:
            Vs[m - 1] = Vs[m - 1] + g
        elif kind == 3:
            pr.innerProof.proof2.b = pr.innerProof.proof2.b + M_(1)
        elif kind == 4:
            t = bytearray(pr.innerProof.proof2.transcript); t[-5] = ord("1") if t[-5] != ord("1") else ord("2")
            pr.innerProof.proof2.transcript = bytes(t)
        elif kind == 5:
            pr.taux = pr.taux + M_(1)
        elif kind == 6:
            pr.innerProof.P_new = pr.innerProof.P_new + h
        Vs_list.append(Vs); proofs.append(pr)
        oVs_list.append([T_(V) for V in Vs]); oproofs.append(po.range_from_json(range_json(pr)))
    want = [po.range_verify(oVs, og, oh, ogs, ohs, ou, opr) for oVs, opr in zip(oVs_list, oproofs)]
    for _ in range(3):
        assert verify_aggreg_range_proofs_batch(Vs_list, g, h, gs, hs, u, proofs) == want
    assert want.count(True) == sum(1 for i in range(count) if i % 8 in (0, 7))
    for i in (0, 1):                                      # the class API agrees
        assert quiet(AggregRangeVerifier(Vs_list[i], g, h, gs, hs, u, proofs[i]).verify) is want[i]


def test_aggreg_batch_verify_off_curve_points_and_slot_forms():
    """bp_rp_verify_aggreg_batch on inputs the class API can never build or compares as strings: proof points that are not on
    secp256k1 (fastecdsa's constructor would raise) are rejected on both paths; a leading zero in a challenge slot is a
    different string (reject), leading zeros in the y slot are the same number (accept)."""
    from python_bulletproofs_b200.rangeproofs import AggregNIRangeProver
    from python_bulletproofs_b200.rangeproofs.batch import PackedAggregBatch, verify_aggreg_packed
    n, m, count = 8, 2, 8
    nm, L = n * m, 4
    ogs, ohs, og, oh, ou = gens(nm, ["aoc%d" % i for i in range(5)])
    gs, hs, g, h, u = [P_(t) for t in ogs], [P_(t) for t in ohs], P_(og), P_(oh), P_(ou)
    rng = random.Random(77)
    Vs_list, proofs = [], []
    for i in range(count):
        vs = [M_(rng.getrandbits(n)) for _ in range(m)]
        gammas = [mod_hash(b"gq%d_%d" % (i, j), Q) for j in range(m)]
        Vs_list.append([commitment(g, h, vs[j], gammas[j]) for j in range(m)])
        proofs.append(AggregNIRangeProver(vs, n, g, h, gs, hs, gammas, u, secp256k1, b"aq%d" % i).prove())
    def edit(tr, slot, fn):
        parts = tr.split(b"&"); parts[slot] = fn(parts[slot]); return b"&".join(parts)
    p2 = proofs[5].innerProof.proof2
    p2.transcript = edit(p2.transcript, p2.start_transcript + 2, lambda s_: b"0" + s_)          # reject
    proofs[6].transcript = edit(proofs[6].transcript, 3, lambda s_: b"00" + s_)                # same number: accept
    batch = PackedAggregBatch.from_proofs(Vs_list, proofs, n)
    stride = batch.stride
    assert stride == nat.load().bp_rp_aggreg_proof_stride(n, m) == 64 * m + 4 * 64 + 3 * 32 + 2 * 64 + 2 * 32 + 32 * L + 128 * L
    rec = bytearray(batch.records)
    oV1, oT1, oPnew, oR0 = 64, 64 * m + 128, 64 * m + 256 + 96 + 64, 64 * m + 256 + 96 + 128 + 64 + 32 * L + 64 * L
    for k, o in enumerate((oV1, oT1, oPnew, oR0)):
        base = (k + 1) * stride + o
        if k % 2 == 0:
            rec[base + 32] ^= 1                                  # y altered: not on the curve
        else:
            rec[base:base + 32] = (ecc.P + 5).to_bytes(32, "little")      # x >= p: not canonical
    batch.records = bytes(rec)
    want = bytes([1, 0, 0, 0, 0, 0, 1, 1])
    for _ in range(3):                                           # bucket method over the generator rows, then the byte table
        assert verify_aggreg_packed(batch, g, h, gs, hs, u) == want


def _small_batch(n, count, tag):
    seeds = [tag + "%d" % i for i in range(5)]
    ogs, ohs, og, oh, ou = gens(n, seeds)
    gs, hs, g, h, u = [P_(t) for t in ogs], [P_(t) for t in ohs], P_(og), P_(oh), P_(ou)
    rng = random.Random(11)
    Vs, proofs = [], []
    for i in range(count):
        v = rng.getrandbits(n)
        gamma = mod_hash(b"gm%d" % i, Q)
        Vs.append(commitment(g, h, M_(v), gamma))
        proofs.append(NIRangeProver(M_(v), n, g, h, gs, hs, gamma, u, secp256k1, b"q%d" % i).prove())
    return gs, hs, g, h, u, Vs, proofs


def test_batch_verify_rejects_points_off_the_curve():
    """Proof-supplied coordinates that are not on secp256k1 (fastecdsa's Point constructor would raise ValueError, so the
    reference can never see them) are rejected by the C ABI on every path, and Point() itself raises."""
    from python_bulletproofs_b200.rangeproofs.batch import PackedBatch, verify_packed
    n, count = 16, 12
    gs, hs, g, h, u, Vs, proofs = _small_batch(n, count, "oc")
    batch = PackedBatch.from_proofs(Vs, proofs, n)
    stride, L = batch.stride, 4
    rec = bytearray(batch.records)
    off = {"V": 0, "A": 64, "S": 128, "T1": 192, "T2": 256, "u_new": 416, "P_new": 480, "L0": 608 + 32 * L, "R3": 608 + 32 * L + 64 * L + 64 * 3}
    for k, (name, o) in enumerate(off.items()):
        base = (k + 1) * stride + o
        if k % 2 == 0:
            rec[base + 32] ^= 1                      # y altered: (x, y') is not on the curve
        else:
            rec[base:base + 32] = (ecc.P + 5).to_bytes(32, "little")     # x >= p: not canonical
    batch.records = bytes(rec)
    want = bytes([1] + [0] * len(off) + [1] * (count - 1 - len(off)))
    for _ in range(3):                                # bucket method, byte tables, 16-bit tables
        assert verify_packed(batch, g, h, gs, hs, u) == want
    with pytest.raises(ValueError):
        Point(5, 7, secp256k1)
    with pytest.raises(ValueError):
        Point(ecc.P + 1, 2, secp256k1)
    assert Point(0, 0, None) == Point.IDENTITY_ELEMENT


def test_batch_verify_transcript_slot_forms():
    """Slots are compared as the reference compares them: str(x) == slot for the challenge slots (a leading zero or a
    value + q is a different string), int(slot) for y, z, x (leading zeros are the same number)."""
    n, count = 8, 6
    gs, hs, g, h, u, Vs, proofs = _small_batch(n, count, "ts")
    def edit(tr, slot, fn):
        parts = tr.split(b"&"); parts[slot] = fn(parts[slot]); return b"&".join(parts)
    # 1: leading zero in a Protocol-2 challenge slot -> reject; 2: challenge + q in that slot -> reject
    p2 = proofs[1].innerProof.proof2
    p2.transcript = edit(p2.transcript, p2.start_transcript + 2, lambda s: b"0" + s)
    p2 = proofs[2].innerProof.proof2
    p2.transcript = edit(p2.transcript, p2.start_transcript + 2, lambda s: str(int(s) + Q).encode())
    # 3: leading zeros in the y slot of the range transcript -> same number, still accepted
    proofs[3].transcript = edit(proofs[3].transcript, 3, lambda s: b"000" + s)
    # 4: Protocol-1 challenge slot with a leading zero -> reject
    proofs[4].innerProof.transcript = edit(proofs[4].innerProof.transcript, 1, lambda s: b"0" + s)
    # 5: z slot replaced by z + q -> same residue, accepted (ModP arithmetic reduces)
    proofs[5].transcript = edit(proofs[5].transcript, 4, lambda s: str(int(s) + Q).encode())
    want = [True, False, False, True, False, True]
    ogs, ohs, og, oh, ou = [T_(t) for t in gs], [T_(t) for t in hs], T_(g), T_(h), T_(u)
    assert [po.range_verify([T_(V)], og, oh, ogs, ohs, ou, po.range_from_json(range_json(pr))) for V, pr in zip(Vs, proofs)] == want
    for _ in range(3):
        assert verify_range_proofs_batch(Vs, g, h, gs, hs, u, proofs) == want
    for i in range(count):
        assert quiet(RangeVerifier(Vs[i], g, h, gs, hs, u, proofs[i]).verify) is want[i]


def test_h_scale_equals_materialised_generators():
    """bp_ipa_prove_hs / bp_ipa_verify_eq_hs with (hs, y^-i) must give the same proof and decisions as the reference's
    materialised hsp list (rangeproof_prover.py:77)."""
    N = 32
    g, h, u, a, b = ipa_inputs(N, ["hs0", "hs1", "hs2", "hs3", "hs4"])
    y = po.mod_hash(b"some-y")
    yinv = [pow(y, -i, Q) for i in range(N)]
    hsp = ecc.scalar_mul_batch(h, yinv)
    g_, h_, hsp_, u_ = [P_(t) for t in g], [P_(t) for t in h], [P_(t) for t in hsp], P_(u)
    am, bm = [M_(v) for v in a], [M_(v) for v in b]
    c = inner_product(am, bm)
    P2 = vector_commitment(g_, hsp_, am, bm) + c * u_
    ref = FastNIProver2(g_, hsp_, u_, P2, am, bm, secp256k1).prove()
    got = FastNIProver2(g_, h_, u_, P2, am, bm, secp256k1, _h_scale=yinv).prove()
    assert p2_json(got) == p2_json(ref)
    assert p2_json(ref) == po.proof2_to_json(po.ipa_prove2(g, hsp, u, a, b))
    assert quiet(Verifier2(g_, h_, u_, P2, got, _h_scale=yinv).verify)
    assert quiet(Verifier2(g_, hsp_, u_, P2, got).verify)
    assert not quiet(Verifier2(g_, h_, u_, P2, got).verify)          # wrong generators without the scale


@pytest.mark.parametrize("n,m", [(8, 1), (64, 1), (4, 2), (16, 4)])
def test_c_algebra_prover_equals_python_algebra_prover(n, m):
    """rangeproofs/_core.prove (polynomial algebra in C, packed vectors) against _prove_python (the same algebra in Python
    ints) and against the oracle prover: identical proofs, byte for byte."""
    from python_bulletproofs_b200.rangeproofs import _core
    from python_bulletproofs_b200.utils.transcript import Transcript
    seeds = ["ca%d" % i for i in range(7)]
    ogs, ohs, og, oh, ou = gens(n * m, seeds)
    gs, hs, g, h, u = [P_(t) for t in ogs], [P_(t) for t in ohs], P_(og), P_(oh), P_(ou)
    rng = random.Random(n * 100 + m)
    vs = [M_(rng.getrandbits(n)) for _ in range(m)]
    if m > 1:
        vs[0] = M_(0); vs[-1] = M_(2 ** n - 1)
    gammas = [mod_hash(b"gm%d" % j, Q) for j in range(m)]
    a = _core.prove(vs, n, g, h, gs, hs, gammas, u, secp256k1, Transcript(b"seedX"))
    b = _core._prove_python(vs, n, g, h, gs, hs, gammas, u, secp256k1, Transcript(b"seedX"))
    assert range_json(a) == range_json(b)
    want = po.range_prove([v.x for v in vs], n, og, oh, ogs, ohs, [gm.x for gm in gammas], ou, b"seedX")
    assert range_json(a) == po.range_to_json(want)
