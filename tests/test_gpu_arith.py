"""Known-answer tests of the device arithmetic (fp.cuh, fq.cuh, ec.cuh) against Python big ints
and the CPU oracle.  Bit-exact."""
import random

import pytest

from oracle import ecc
from gpu_util import call_test

pytestmark = pytest.mark.gpu
P, Q = ecc.P, ecc.Q


def _le(vals):
    return b"".join(v.to_bytes(32, "little") for v in vals)


def _un(b):
    return [int.from_bytes(b[i:i + 32], "little") for i in range(0, len(b), 32)]


EDGE = [0, 1, 2, 977, 2 ** 32, 2 ** 32 + 977, P - 1, P - 2, P, P + 1, 2 ** 256 - 1, 2 ** 256 - 2, 2 ** 255, 2 ** 255 - 1,
        2 ** 256 - 2 ** 32 - 978, (1 << 224) - 1, 0xFFFFFFFF, 0xFFFFFFFF00000000, Q, Q - 1]


def _pairs(rng, n_rand, mod=None):
    a = list(EDGE) + [rng.getrandbits(256) for _ in range(n_rand)]
    pairs = [(x, y) for x in EDGE for y in EDGE] + [(a[rng.randrange(len(a))], rng.getrandbits(256)) for _ in range(n_rand)]
    return [p[0] for p in pairs], [p[1] for p in pairs]


@pytest.mark.parametrize("op,fn", [(0, lambda a, b: a * b % P), (1, lambda a, b: (a + b) % P), (2, lambda a, b: (a - b) % P),
                                   (4, lambda a, b: (-a) % P), (5, lambda a, b: a * a % P)])
def test_fp_ops_lazy_inputs(op, fn):
    """Inputs are arbitrary 256-bit residues (the device keeps lazy values in [0, 2^256))."""
    rng = random.Random(100 + op)
    a, b = _pairs(rng, 4000)
    got = _un(call_test("bp_test_fp", op, _le(a), _le(b), len(a), 32))
    assert got == [fn(x, y) for x, y in zip(a, b)]


def test_fp_inv():
    rng = random.Random(5)
    a = [1, 2, P - 1, P - 2, 977, 2 ** 255] + [rng.getrandbits(256) % P or 1 for _ in range(300)]
    got = _un(call_test("bp_test_fp", 3, _le(a), _le(a), len(a), 32))
    assert got == [pow(x, -1, P) for x in a]


@pytest.mark.parametrize("dev", [0, 1])
@pytest.mark.parametrize("op,fn", [(0, lambda a, b: a * b % Q), (1, lambda a, b: (a + b) % Q), (2, lambda a, b: (a - b) % Q),
                                   (4, lambda a, b: (-a) % Q)])
def test_fq_ops(op, fn, dev):
    rng = random.Random(200 + op)
    a, b = _pairs(rng, 1500)
    got = _un(call_test("bp_test_fq", op, _le(a), _le(b), len(a), 32, dev))
    assert got == [fn(x % Q, y % Q) for x, y in zip(a, b)]


@pytest.mark.parametrize("dev", [0, 1])
def test_fq_inv(dev):
    rng = random.Random(6)
    a = [1, 2, Q - 1] + [rng.getrandbits(256) % Q or 1 for _ in range(100)]
    got = _un(call_test("bp_test_fq", 3, _le(a), _le(a), len(a), 32, dev))
    assert got == [pow(x, -1, Q) for x in a]


def test_ec_ops_including_exceptional_cases():
    rng = random.Random(9)
    base = [ecc.py_mul(ecc.G, rng.getrandbits(256)) for _ in range(24)]
    A, B = [], []
    for i, p in enumerate(base):
        q = base[(i + 7) % len(base)]
        A += [p, p, p, None, p, None]
        B += [q, p, ecc.point_neg(p), q, None, None]      # generic, doubling, inverse, identities
    pa, pb = ecc.pack_points(A), ecc.pack_points(B)
    n = len(A)
    want_add = [ecc.py_add(a, b) for a, b in zip(A, B)]
    for op in (0, 1, 10, 11):           # mixed add, full add, each also from a re-projected accumulator
        got = ecc.unpack_points(call_test("bp_test_ec", op, pa, pb, n, 64), n)
        assert got == want_add, op
    for op in (2, 12):
        got = ecc.unpack_points(call_test("bp_test_ec", op, pa, pb, n, 64), n)
        assert got == [ecc.py_add(a, a) for a in A], op
    got = ecc.unpack_points(call_test("bp_test_ec", 3, pa, pb, n, 64), n)
    assert got == [ecc.point_neg(a) for a in A]
