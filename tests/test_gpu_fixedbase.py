"""Fixed-base table path (fixedbase.cuh / bp_fixed.inl): MSMs over a repeated generator set answered from precomputed
tables must equal the bucket method and the oracle bit for bit; the IPA / range-proof provers must produce the same
golden proofs whether or not their generators have tables."""
import ctypes
import random

import pytest

from oracle import ecc
from oracle import protocol_oracle as po
from python_bulletproofs_b200 import _native as nat
from gpu_util import gpu_msm_raw
from helpers import fast_points

pytestmark = pytest.mark.gpu
Q, P = ecc.Q, ecc.P


def _stats():
    v = [ctypes.c_uint64() for _ in range(4)]
    nat.check(nat.load().bp_fb_stats(*[ctypes.byref(x) for x in v]))
    return dict(zip(("tables", "bytes", "hits", "builds"), (x.value for x in v)))


@pytest.fixture
def fb_mode():
    lib = nat.load()
    nat.init(0)
    nat.check(lib.bp_fb_clear())

    def set_mode(m):
        nat.check(lib.bp_fb_set_mode(m))
    yield set_mode
    nat.check(lib.bp_fb_clear())
    nat.check(lib.bp_fb_set_mode(1))


def _sc(ks):
    return b"".join(int(k).to_bytes(32, "little") for k in ks)


EDGE = [0, 1, 2, 255, 256, 257, 0xFF00, 2 ** 248, 2 ** 255, Q - 1, Q, Q + 1, 2 ** 256 - 1, (Q - 1) // 2, 0x0101010101010101010101010101010101010101010101010101010101010101]


@pytest.mark.parametrize("n", [1, 2, 3, 31, 32, 33, 65, 129, 300, 1000])
def test_table_msm_equals_oracle(fb_mode, n):
    fb_mode(2)
    rng = random.Random(1000 + n)
    pts = fast_points(n, 500 + n)
    if n >= 3:
        pts[1] = pts[0]                                   # repeated point
        pts[2] = (pts[0][0], P - pts[0][1])               # and its negation
    if n >= 33:
        pts[17] = None                                    # identity among the generators
    pb = ecc.pack_points(pts)
    for trial in range(3):
        ks = [rng.getrandbits(256) for _ in range(n)]
        for i in range(min(n, len(EDGE))):
            if trial != 1:
                ks[(i * 7) % n] = EDGE[(i + trial) % len(EDGE)]
        if trial == 2:
            ks = [0] * n                                   # all-zero scalars -> identity
        want = ecc.msm(pts, [k % Q for k in ks])
        assert gpu_msm_raw(pb, _sc(ks), n) == want, (n, trial)
    st = _stats()
    assert st["builds"] >= 1 and st["hits"] >= 2


def test_table_is_built_at_second_use_and_results_never_change(fb_mode):
    fb_mode(1)
    n = 130
    rng = random.Random(7)
    pts = fast_points(n, 4242)
    pb = ecc.pack_points(pts)
    b0 = _stats()["builds"]
    outs = []
    for call in range(4):
        ks = [rng.getrandbits(256) % Q for _ in range(n)]
        got = gpu_msm_raw(pb, _sc(ks), n)
        assert got == ecc.msm(pts, ks), call
        outs.append(_stats())
    assert outs[0]["builds"] == b0                        # first sighting: bucket method, no table
    assert outs[1]["builds"] == b0 + 1                    # second: table built and used
    assert outs[3]["hits"] >= outs[1]["hits"] + 2
    # switching the tables off gives the same answers
    fb_mode(0)
    ks = [rng.getrandbits(256) % Q for _ in range(n)]
    a = gpu_msm_raw(pb, _sc(ks), n)
    fb_mode(2)
    assert gpu_msm_raw(pb, _sc(ks), n) == a == ecc.msm(pts, ks)


def test_golden_msm_cases_through_tables(fb_mode, golden):
    from helpers import c3_inputs, explicit_case
    from gpu_util import gpu_msm
    fb_mode(2)
    for case in golden("msm")["cases"]:
        pts, ks = c3_inputs(case["lgn"], case["n"]) if case["kind"] == "c3" else explicit_case(case)
        if not 0 < len(pts) <= 4000:
            continue
        assert po.enc_point(gpu_msm(pts, ks)).hex() == case["out"], case.get("name", case.get("lgn"))


@pytest.mark.parametrize("mode", [0, 2])
def test_provers_emit_the_golden_proofs_with_and_without_tables(fb_mode, golden, mode):
    fb_mode(mode)
    import test_gpu_protocols as tp
    # the golden-proof tests of the protocol suite, re-run under this table mode (twice, so that mode 2 also replays
    # captured IPA-round graphs that read a table)
    for _ in range(2):
        tp.test_ipa_golden(golden, "ipa_small")
        tp.test_range_golden(golden, "range_small")
        tp.test_range_golden(golden, "range_c1")
    tp.test_ipa_golden(golden, "ipa_c2")
    tp.test_aggreg_golden(golden, "aggreg_small")
    if mode == 2:
        assert _stats()["hits"] > 10


def test_batch_of_msms_over_one_point_list_uses_one_table(fb_mode):
    fb_mode(2)
    rng = random.Random(99)
    for n, nmsm in ((2, 2), (129, 2), (65, 5)):
        pts = fast_points(n, 9000 + n)
        pb = ecc.pack_points(pts)
        kss = [[rng.getrandbits(256) for _ in range(n)] for _ in range(nmsm)]
        kss[-1][0] = 0
        raw = nat.msm_batch_bytes(pb * nmsm, b"".join(_sc(ks) for ks in kss), [n * j for j in range(nmsm + 1)])
        for j, ks in enumerate(kss):
            assert ecc.unpack_point(raw[64 * j:64 * j + 64]) == ecc.msm(pts, [k % Q for k in ks]), (n, j)
    assert _stats()["hits"] >= 2


@pytest.mark.parametrize("mode", [0, 2])
def test_batch_verifier_decisions_with_and_without_tables(fb_mode, mode):
    """Every corruption kind of the protocol suite's batch-verifier test, oracle-checked, under both table modes."""
    fb_mode(mode)
    import test_gpu_protocols as tp
    tp.test_batch_verify_decisions(8, 40)
    tp.test_batch_verify_decisions(64, 24)
    if mode == 2:
        assert _stats()["tables"] >= 2
