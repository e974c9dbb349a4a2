"""CPU-only checks of the host side: the C-ABI library loads and exports every symbol that
include/bp_gpu.h declares (no compute calls), host-side Fiat-Shamir code restated in C matches
the oracle, the drop-in classes keep the reference's shapes/exceptions, and the sharding logic
works across 2 gloo ranks."""
import ctypes
import os
import re
import subprocess
import sys

import pytest

from oracle import ecc, protocol_oracle as po
from python_bulletproofs_b200 import _native as nat, sharding

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "bp_gpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bp_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(nat.LIB_PATH)
    names = header_symbols()
    assert len(names) >= 30
    for name in names:
        assert hasattr(lib, name), name
    assert set(names) == set(nat.PROTOTYPES), set(names) ^ set(nat.PROTOTYPES)


def test_no_cpu_fallback_without_gpu():
    lib = nat.load()
    if lib.bp_device_count() > 0:
        pytest.skip("a GPU is visible")
    out = ctypes.create_string_buffer(64)
    assert lib.bp_msm(ecc.pack_point(ecc.G), (5).to_bytes(32, "little"), 1, out) != 0
    assert b"no CPU fallback" in lib.bp_last_error()
    from python_bulletproofs_b200.pippenger import PipSECP256k1
    from python_bulletproofs_b200 import secp256k1
    with pytest.raises(nat.BpGpuError):
        PipSECP256k1.multiexp([secp256k1.G], [5])


def test_host_c_transcript_helpers_match_oracle():
    """bp_mod_hash / bp_point_to_b64 are pure host code: utils.py:84-111 restated in C."""
    lib = nat.load()
    out = ctypes.create_string_buffer(32)
    for msg in (b"", b"&", b"alpha" + b"x" * 70, bytes(range(256)) * 5):
        assert lib.bp_mod_hash(msg, len(msg), out) == 0
        assert int.from_bytes(out.raw, "little") == po.mod_hash(msg)
    b64 = ctypes.create_string_buffer(45)
    ln = ctypes.c_size_t()
    for k in (1, 2, 3, 0xDEADBEEF, ecc.Q - 1):
        pt = ecc.py_mul(ecc.G, k)
        assert lib.bp_point_to_b64(ecc.pack_point(pt), b64, ctypes.byref(ln)) == 0
        assert b64.raw[:ln.value] == po.b64_point(pt)
    lib.bp_point_to_b64(bytes(64), b64, ctypes.byref(ln))
    assert b64.raw[:ln.value] == b"AA=="
    # the indexed form used for the sL / sR blinding vectors (rangeproof_prover.py:48-55)
    for suffix, first, count in ((b"", 0, 3), (b"abc&def&", 0, 130), (b"t" * 200, 120, 140), (b"x", 99990, 25), (b"y", 0, 0)):
        assert nat.mod_hash_indexed(suffix, first, count) == [po.mod_hash(str(i).encode() + suffix) for i in range(first, first + count)]


def test_sha256_both_implementations_match_hashlib():
    """csrc/sha256_host.cc: portable and (when the CPU has them) SHA-NI compression functions."""
    import hashlib
    import random
    lib = nat.load()
    out = ctypes.create_string_buffer(32)
    rng = random.Random(1)
    try:
        for force in (1, 0):
            lib.bp_sha256_set_portable(force)
            for n in list(range(0, 130)) + [255, 256, 1000, 4096, 5555]:
                m = rng.randbytes(n)
                assert lib.bp_sha256(m, n, out) == 0
                assert out.raw == hashlib.sha256(m).digest(), (force, n)
    finally:
        lib.bp_sha256_set_portable(0)


def test_host_fq_arithmetic_matches_bigint():
    import random
    rng = random.Random(4)
    Q = ecc.Q
    a = [0, 1, Q - 1, Q, 2 ** 256 - 1] + [rng.getrandbits(256) for _ in range(300)]
    b = [rng.getrandbits(256) for _ in a]
    le = lambda v: b"".join(x.to_bytes(32, "little") for x in v)   # noqa: E731
    out = ctypes.create_string_buffer(32 * len(a))
    for op, fn in ((0, lambda x, y: x * y % Q), (1, lambda x, y: (x + y) % Q), (2, lambda x, y: (x - y) % Q), (4, lambda x, y: -x % Q)):
        assert nat.load().bp_test_fq(op, 0, le(a), le(b), len(a), out) == 0
        got = [int.from_bytes(out.raw[32 * i:32 * i + 32], "little") for i in range(len(a))]
        assert got == [fn(x % Q, y % Q) for x, y in zip(a, b)], op
    nz = [x % Q or 1 for x in a] + [2, 3, Q - 2, (Q + 1) // 2, 2 ** 255 % Q] + [1 << k for k in range(1, 256, 9)]
    out = ctypes.create_string_buffer(32 * len(nz))
    for op in (3, 9):           # 3: portable Fermat form (shared with the device), 9: fq_inv_host, the binary extended GCD of the IPA round loop
        nat.load().bp_test_fq(op, 0, le(nz), le(nz), len(nz), out)
        assert [int.from_bytes(out.raw[32 * i:32 * i + 32], "little") for i in range(len(nz))] == [pow(x, -1, Q) for x in nz], op


def test_modp_quirks_and_api_shapes():
    """SURVEY.md A.4 reduction rules + constructor assertions, all host-side."""
    from python_bulletproofs_b200.utils import ModP, mod_hash, inner_product, point_to_bytes, bytes_to_point, Transcript
    from python_bulletproofs_b200 import Point, secp256k1
    from python_bulletproofs_b200.innerproduct import FastNIProver2, NIProver
    from python_bulletproofs_b200.pippenger import PipSECP256k1
    p = secp256k1.q
    a, b = ModP(p - 1, p), ModP(5, p)
    assert (a + b).x == 4 and (a + 7).x == p + 6 and (a * 3).x == 3 * (p - 1) and (a - 3).x == p - 4
    assert (3 - b).x == p - 2 and (-ModP(0, p)).x == p and (b ** 3).x == 125 and b % 3 == 2
    assert (b.inv() * b) == ModP(1, p) and str(b) == "5"
    with pytest.raises(Exception, match="modular inverse does not exist"):
        ModP(0, p).inv()
    assert mod_hash(b"abc", p).x == po.mod_hash(b"abc")
    assert inner_product([a, b], [b, b]).x == ((p - 1) * 5 + 25) % p
    G = secp256k1.G
    assert bytes_to_point(point_to_bytes(G)) == G and point_to_bytes(Point.IDENTITY_ELEMENT) == b"\x00"
    t = Transcript(b"seed6")
    assert t.digest == b"c2VlZDY=&"
    with pytest.raises(Exception, match="Different number of group elements and exponents"):
        PipSECP256k1.multiexp([G], [])
    assert PipSECP256k1.multiexp([], []) == Point.IDENTITY_ELEMENT       # no device call for the empty product
    with pytest.raises(AssertionError):
        FastNIProver2([G, G, G], [G, G, G], G, G, [b] * 3, [b] * 3, secp256k1)
    with pytest.raises(AssertionError):
        NIProver([G, G], [G], G, G, b, [b, b], [b, b], secp256k1)


def test_slice_bounds_partition():
    for n in (0, 1, 7, 8192, 2 ** 20 + 3):
        for world in (1, 2, 3, 4, 8):
            sl = sharding.all_slices(n, world)
            assert sl[0][0] == 0 and sl[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(sl, sl[1:]))
            sizes = [hi - lo for lo, hi in sl]
            assert max(sizes) - min(sizes) <= 1


_GLOO_WORKER = r'''
import os, sys, ctypes
sys.path.insert(0, sys.argv[1])
import torch.distributed as dist
from python_bulletproofs_b200 import sharding, _native as nat
from oracle import ecc
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % sys.argv[2], rank=int(sys.argv[3]), world_size=2)
rank, world = dist.get_rank(), dist.get_world_size()
# 1. the id-sharing path used by init_nccl (without touching NCCL/GPU): rank 0's bytes reach rank 1
blob = sharding.share_bytes(bytes(range(128)) if rank == 0 else None, 0)
assert blob == bytes(range(128))
# 2. sharded MSM host logic: every rank owns a slice, partial results are combined in rank order;
#    emulate the device partial with the oracle so the combination logic is what is tested
import random
rng = random.Random(1)
n = 101
pts = [ecc.py_mul(ecc.G, rng.getrandbits(64)) for _ in range(n)]
ks = [rng.getrandbits(256) for _ in range(n)]
lo, hi = sharding.slice_bounds(n, rank, world)
part = ecc.msm(pts[lo:hi], ks[lo:hi])
parts = [None] * world
dist.all_gather_object(parts, part)
total = None
for p in parts:
    total = ecc.py_add(total, p)
assert total == ecc.msm(pts, ks), "sharded sum differs"
# 3. accept-bitmap stitching for a proof batch split in contiguous blocks
counts = [h - l for l, h in sharding.all_slices(11, world)]
mine = bytes([(i * 7 + 1) % 2 for i in range(*sharding.slice_bounds(11, rank, world))])
got = [None] * world
dist.all_gather_object(got, mine)
assert b"".join(got) == bytes([(i * 7 + 1) % 2 for i in range(11)])
dist.barrier()
dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_two_rank_gloo_sharding(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(_GLOO_WORKER)
    port = str(29500 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, port, str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
             for r in range(2)]
    outs = [p.communicate(timeout=240)[0].decode() for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and "rank %d ok" % r in o, o


def test_host_c_range_proof_algebra_matches_python_formulas():
    """bp_rp_prover_poly1/2 and bp_rp_verifier_scalars (csrc/rp_algebra.h, host only) against the formulas of
    rangeproof_prover.py:93-112 / rangeproof_aggreg_prover.py:117-146 / rangeproof_verifier.py:69-72 in Python ints."""
    import random
    from python_bulletproofs_b200.rangeproofs._core import _position_constants, inverse_powers
    q = ecc.Q
    rng = random.Random(11)
    for n, m in ((1, 1), (8, 1), (64, 1), (4, 3), (64, 16)):
        nm = n * m
        bits = bytes(rng.getrandbits(1) for _ in range(nm))
        sL = [rng.randrange(q) for _ in range(nm)]
        sR = [rng.randrange(q) for _ in range(nm)]
        sL[0], sR[-1] = 0, q - 1
        y, z, x = rng.randrange(1, q), rng.randrange(q), rng.randrange(q)
        if n == 8:
            y, z, x = q - 1, 0, 1
        aL = list(bits)
        aR = [(b - 1) % q for b in aL]
        ypow, zz = _position_constants(y, z, n, nm, q)
        ysR = [ypow[i] * sR[i] % q for i in range(nm)]
        t1 = (sum(sL[i] * (ypow[i] * (aR[i] + z) + zz[i]) for i in range(nm)) + sum((aL[i] - z) * ysR[i] for i in range(nm))) % q
        t2 = sum(sL[i] * ysR[i] for i in range(nm)) % q
        sLb, sRb = nat.pack_scalars(sL), nat.pack_scalars(sR)
        assert nat.rp_prover_poly1(bits, sLb, sRb, n, m, y, z) == (t1, t2)
        ls = [(aL[i] - z + sL[i] * x) % q for i in range(nm)]
        rs = [(ypow[i] * (aR[i] + z + sR[i] * x) + zz[i]) % q for i in range(nm)]
        yinv = inverse_powers(y, nm, q)
        hsc = [(z + zz[i] * yinv[i]) % q for i in range(nm)]
        got = nat.rp_prover_poly2(bits, sLb, sRb, n, m, y, z, x)
        assert [nat.unpack_scalars(got[k], nm) for k in range(4)] == [ls, rs, yinv, hsc]
        assert got[4] == sum(a * b for a, b in zip(ls, rs)) % q
        assert nat.unpack_scalars(got[5], nm) == [r * yi % q for r, yi in zip(rs, yinv)]
        zp = [pow(z, j + 2, q) for j in range(m + 1)]
        delta = ((z - z * z) * sum(ypow) - sum(zp[j] * (2 ** n - 1) for j in range(1, m + 1))) % q
        vy, vh, vd = nat.rp_verifier_scalars(n, m, y, z)
        assert (nat.unpack_scalars(vy, nm), nat.unpack_scalars(vh, nm), vd) == (yinv, hsc, delta)
    with pytest.raises(nat.BpGpuError):
        nat.rp_verifier_scalars(4, 1, 0, 5)                        # y = 0: "modular inverse does not exist"


def test_host_xyzz_to_affine_matches_bigint():
    """csrc/fp_host.h: the host finish of the IPA prover's L, R (XYZZ -> canonical affine, one shared inversion) against
    big-int arithmetic: random re-projections of curve points (lazy residues up to 2^256 - 1 included) and the identity."""
    import random
    from oracle import ecc
    P = ecc.P if hasattr(ecc, "P") else 2 ** 256 - 2 ** 32 - 977
    rng = random.Random(99)
    pts = [ecc.py_mul(ecc.G, rng.getrandbits(200) + 1) for _ in range(6)]
    le = lambda v: int(v).to_bytes(32, "little")      # noqa: E731
    for trial in range(20):
        recs, want = [], []
        cnt = rng.randrange(1, 9)
        for i in range(cnt):
            if rng.random() < 0.2:
                recs.append(le(rng.getrandbits(256)) + le(rng.getrandbits(256)) + le(0 if rng.random() < 0.5 else P) + le(rng.getrandbits(256)))
                want.append(bytes(64))
                continue
            x, y = pts[rng.randrange(6)]
            z = rng.randrange(1, P)
            zz, zzz = z * z % P, z * z * z % P
            vals = [x * zz % P, y * zzz % P, zz, zzz]
            vals = [v + P if rng.random() < 0.3 and v + P < 2 ** 256 else v for v in vals]       # non-canonical representatives
            recs.append(b"".join(le(v) for v in vals))
            want.append(le(x) + le(y))
        out = ctypes.create_string_buffer(64 * cnt)
        assert nat.load().bp_test_xyzz_to_affine_host(b"".join(recs), cnt, out) == 0
        assert out.raw == b"".join(want), trial


def test_host_horner_over_window_sums_matches_oracle():
    """fp_host.h: horner_host -- the Horner chain over the window sums that host-result MSMs (bp_msm) hand to the host instead of
    running k_combine: sum_w 2^(c*w) * S_w with complete Jacobian arithmetic (identity sums, equal sums: P + P, opposite sums: P - P,
    non-canonical coordinate representatives), for shapes with and without a second unit of the top window -- against the oracle."""
    import random
    P = 2 ** 256 - 2 ** 32 - 977
    rng = random.Random(1234)
    le = lambda v: int(v).to_bytes(32, "little")      # noqa: E731
    base = [ecc.py_mul(ecc.G, rng.getrandbits(250) + 1) for _ in range(5)]

    def rec(pt):
        if pt is None:
            return le(rng.getrandbits(256)) + le(rng.getrandbits(256)) + le(0 if rng.random() < 0.5 else P) + le(0)
        x, y = pt
        z = rng.randrange(1, P)
        zz, zzz = z * z % P, z * z * z % P
        vals = [x * zz % P, y * zzz % P, zz, zzz]
        vals = [v + P if rng.random() < 0.3 and v + P < 2 ** 256 else v for v in vals]
        return b"".join(le(v) for v in vals)

    lib = nat.load()
    for c, W, dbl in ((16, 8, 1), (13, 10, 0), (5, 26, 0), (1, 128, 1), (8, 16, 1), (3, 43, 0)):
        for trial in range(4):
            U = W + dbl
            sums = []
            for u in range(U):
                r = rng.random()
                sums.append(None if r < 0.2 else base[rng.randrange(5)])
            if trial == 1:                       # top window: both units the same point (P + P), next window its negation after the doublings' input
                sums[U - 1] = base[0]
                if dbl:
                    sums[U - 2] = base[0]
            if trial == 2 and dbl:               # P - P in the top window
                sums[U - 1] = base[1]; sums[U - 2] = ecc.point_neg(base[1])
            if trial == 3:                       # everything empty
                sums = [None] * U
            want = None
            for u in range(U):
                w = u if u < W else W - 1        # the second unit of the top window carries the top window's weight
                if sums[u] is not None:
                    want = ecc.point_add(want, ecc.py_mul(sums[u], 1 << (c * w)))
            out = ctypes.create_string_buffer(64)
            assert lib.bp_test_horner_host(b"".join(rec(s_) for s_ in sums), c, W, U, dbl, out) == 0
            assert out.raw == ecc.pack_point(want), (c, W, dbl, trial)


def test_host_affine_sum_matches_oracle():
    """fp_host.h: affine_sum_host -- the rank partials of a sharded host-result MSM added on the host: identities, repeated points
    (P + P), opposite points, a single point, nothing at all."""
    import random
    rng = random.Random(77)
    base = [ecc.py_mul(ecc.G, rng.getrandbits(250) + 1) for _ in range(4)]
    lib = nat.load()
    cases = [[], [None], [base[0]], [base[0], base[0]], [base[1], ecc.point_neg(base[1])], [None, base[2], None, base[3], base[2], ecc.point_neg(base[3])],
             [base[i % 4] for i in range(8)]]
    for pts in cases:
        want = None
        for p_ in pts:
            want = ecc.point_add(want, p_)
        out = ctypes.create_string_buffer(64)
        assert lib.bp_test_affine_sum_host(b"".join(ecc.pack_point(p_) for p_ in pts), len(pts), out) == 0
        assert out.raw == ecc.pack_point(want), pts
