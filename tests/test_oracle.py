"""Pins the CPU oracle (oracle/) against the golden vectors produced by the UNMODIFIED
reference (oracle/gen_golden.py) and against published / big-int known answers.
CPU-only; the oracle is test infrastructure, never the product path."""
import random

import pytest

from oracle import ecc, protocol_oracle as po
from helpers import c3_inputs, explicit_case, gens, ipa_inputs

Q, P = ecc.Q, ecc.P


def test_published_multiples_of_G():
    # SEC 2 / widely published secp256k1 test vectors
    assert ecc.scalar_mul(ecc.G, 2)[0] == 0xC6047F9441ED7D6D3045406E95C07CD85C778E4B8CEF3CA7ABAC09B95C709EE5
    assert ecc.scalar_mul(ecc.G, 3)[0] == 0xF9308A019258C31049344F85F89D5229B531C845836F99B08601F113BCE036F9
    assert ecc.scalar_mul(ecc.G, Q) is None
    assert ecc.scalar_mul(ecc.G, Q - 1) == (ecc.GX, P - ecc.GY)


def test_field_ops_vs_bigint():
    import ctypes
    rng = random.Random(7)
    L = ecc.lib()
    vals = [0, 1, 2, P - 1, P - 2, 2 ** 255, 0x1000003D1, 2 ** 256 - 1 - P] + [rng.getrandbits(256) % P for _ in range(200)]
    out = ctypes.create_string_buffer(32)
    for a in vals:
        for b in (vals[rng.randrange(len(vals))], vals[rng.randrange(len(vals))]):
            L.orc_fe_mul(a.to_bytes(32, "little"), b.to_bytes(32, "little"), out)
            assert int.from_bytes(out.raw, "little") == a * b % P
        if a:
            L.orc_fe_inv(a.to_bytes(32, "little"), out)
            assert int.from_bytes(out.raw, "little") == pow(a, -1, P)


def test_group_law_vs_python_ints():
    rng = random.Random(11)
    pts = [ecc.py_mul(ecc.G, rng.getrandbits(256)) for _ in range(12)]
    for a in pts[:6]:
        for b in pts[6:]:
            assert ecc.point_add(a, b) == ecc.py_add(a, b)
        assert ecc.point_add(a, a) == ecc.py_add(a, a)
        assert ecc.point_add(a, ecc.point_neg(a)) is None
        assert ecc.point_add(a, None) == a and ecc.point_add(None, a) == a
        k = rng.getrandbits(256)
        assert ecc.scalar_mul(a, k) == ecc.py_mul(a, k)
    assert all(ecc.on_curve(a) for a in pts) and not ecc.on_curve((1, 1))


@pytest.mark.parametrize("algo", ["naive", "bucket", "subset"])
def test_msm_golden(golden, algo):
    """Every MSM output of the reference's Pippenger.multiexp (pippenger.py:22-61)."""
    for case in golden("msm")["cases"]:
        if case["kind"] == "c3":
            if algo == "naive" and case["n"] > 300:
                continue
            pts, ks = c3_inputs(case["lgn"], case["n"])
        else:
            pts, ks = explicit_case(case)
        got = ecc.msm(pts, ks, algo)
        assert po.enc_point(got).hex() == case["out"], case.get("name", case.get("lgn"))


def test_msm_bucket_threads_and_length_mismatch():
    pts, ks = c3_inputs(9)
    assert ecc.msm(pts, ks, "bucket", threads=4) == ecc.msm(pts, ks, "subset")
    with pytest.raises(Exception, match="Different number of group elements and exponents"):
        ecc.msm(pts, ks[:-1])


def test_hashes_and_codecs(golden):
    # mod_hash / elliptic_hash outputs are embedded in the goldens: V = v*g + gamma*h
    case = golden("range_small")["cases"][0]
    gs, hs, g, h, u = gens(case["n"], case["seeds"])
    gamma = po.mod_hash(case["seeds"][5].encode())
    assert str(gamma) == case["gamma"]
    V = po.commit(g, h, int(case["v"]), gamma)
    assert po.enc_point(V).hex() == case["V"]
    assert po.dec_point(po.enc_point(V)) == V
    assert po.enc_point(None) == b"\x00"


@pytest.mark.parametrize("name", ["ipa_small", "ipa_c2"])
def test_ipa_golden(golden, name):
    for case in golden(name)["cases"]:
        N = case["N"]
        g, h, u, a, b = ipa_inputs(N, case["seeds"])
        Pt = po.vec_commit(g, h, a, b)
        assert po.enc_point(Pt).hex() == case["P"]
        c = po.dot(a, b)
        assert str(c) == case["c"]
        proof = po.ipa_prove1(g, h, u, Pt, c, a, b, case["seeds"][5].encode())
        assert po.proof1_to_json(proof) == case["proof1"]
        assert po.ipa_verify1(g, h, u, Pt, c, proof) is case["verify1"] is True
        assert po.ipa_verify1(g, h, u, Pt, (c + 1) % Q, proof) is case["verify1_wrong_c"] is False
        P2 = ecc.point_add(Pt, ecc.scalar_mul(u, c))
        assert po.enc_point(P2).hex() == case["P2"]
        p2 = po.ipa_prove2(g, h, u, a, b)
        assert po.proof2_to_json(p2) == case["proof2"]
        assert po.ipa_verify2(g, h, u, P2, p2) is case["verify2"] is True
        # from-json round trip feeds the verifier the reference's own bytes
        assert po.ipa_verify1(g, h, u, Pt, c, po.proof1_from_json(case["proof1"]))


@pytest.mark.parametrize("name", ["range_small", "range_c1"])
def test_range_golden(golden, name):
    for case in golden(name)["cases"]:
        n, v = case["n"], int(case["v"])
        gs, hs, g, h, u = gens(n, case["seeds"])
        gamma = int(case["gamma"])
        V = po.dec_point(bytes.fromhex(case["V"]))
        proof = po.range_prove([v], n, g, h, gs, hs, [gamma], u, case["seeds"][6].encode())
        assert po.range_to_json(proof) == case["proof"]
        assert po.range_verify([V], g, h, gs, hs, u, proof) is case["verify"]
        assert po.range_verify([ecc.point_add(V, g)], g, h, gs, hs, u, proof) is case["verify_wrong_V"] is False
        s = str(proof["t_hat"])
        bad = dict(proof, t_hat=int(s[:-1] + ("1" if s[-1] != "1" else "2")))
        assert po.range_verify([V], g, h, gs, hs, u, bad) is case["verify_t_hat_flipped"] is False


@pytest.mark.parametrize("name", ["aggreg_small", "aggreg_c4"])
def test_aggreg_golden(golden, name):
    for case in golden(name)["cases"]:
        n, m = case["n"], case["m"]
        gs, hs, g, h, u = gens(n * m, case["seeds"])
        vs = [int(v) for v in case["vs"]]
        gammas = [int(x) for x in case["gammas"]]
        Vs = [po.dec_point(bytes.fromhex(s)) for s in case["Vs"]]
        proof = po.range_prove(vs, n, g, h, gs, hs, gammas, u, case["seeds"][6].encode())
        assert po.range_to_json(proof) == case["proof"]
        assert po.range_verify(Vs, g, h, gs, hs, u, proof) is case["verify"]
        Vbad = Vs[:-1] + [ecc.point_add(Vs[-1], h)]
        assert po.range_verify(Vbad, g, h, gs, hs, u, proof) is case["verify_wrong_V"] is False


def test_survey_appendix_d_known_answers(golden):
    """SURVEY.md Appendix D values recorded from the reference in the survey session."""
    c1 = golden("range_c1")["cases"][0]
    A = po.dec_point(bytes.fromhex(c1["proof"]["A"]))
    assert A[0] == 0x093d5b9ef7e996a2030fa797b9e8ea26b6151ae5aaee35c5b909f8c435d3c550
    assert c1["proof"]["transcript"].startswith("c2VlZDY=&Agk9W5736ZaiAw+nl7no6ia2FRrlqu41xbkJ+MQ108VQ&")
    assert len(c1["proof"]["transcript"]) == 423
    assert len(c1["proof"]["ip"]["transcript"]) == 79 and c1["proof"]["ip"]["p2"]["start"] == 3
    c2 = golden("ipa_c2")["cases"][0]
    import hashlib
    inner = c2["proof1"]["p2"]          # the Proof2 nested in NIProver's Proof1 (1767-byte transcript)
    assert len(inner["transcript"]) == 1767
    assert hashlib.sha256(inner["transcript"].encode("latin1")).hexdigest() == \
        "ba7e23f1170a86897644dd05e8cc55bf59fd0e51add64036cb8ccdf3892f5579"
    assert inner["a"] == "45772048434343676761746858681413549692749859860001827780065551035206530309759"
    assert inner["b"] == "89557398269276448073630310818547859234577444433827909274445100190913033905793"
