"""Host-side scalar type, hash-to-scalar and point codecs (reference: src/utils/utils.py).

These stay on the host by decree of the north star (Fiat-Shamir and challenge derivation are
sequential string work); their observable behaviour -- including ModP's reduction quirks
 -- matches the reference so transcripts are byte-identical.
"""
import base64
from hashlib import sha256
from typing import List

from ..curve import secp256k1
from ..point import Point

CURVE = secp256k1
BYTE_LENGTH = CURVE.q.bit_length() // 8


def egcd(a, b):
    """Extended Euclid, iterative: returns (g, s, t) with s*a + t*b = g  (utils.py:15-21)."""
    s0, s1, t0, t1 = 1, 0, 0, 1
    while b:
        k, r = divmod(a, b)
        a, b = b, r
        s0, s1 = s1, s0 - k * s1
        t0, t1 = t1, t0 - k * t1
    return a, s0, t0


class ModP:
    """Integer mod p with the reference's exact mixed-type behaviour (utils.py:24-81):
    ModP (+,-,*) ModP reduces; ModP + int and ModP * int do NOT reduce; ModP - int reduces;
    int - ModP is -(ModP - int); -ModP(0) has x == p; `% m` yields a plain int;
    ModP * Point multiplies the point by .x."""

    __slots__ = ("x", "p")

    def __init__(self, x, p):
        self.x = x
        self.p = p

    def _same(self, other):
        assert self.p == other.p
        return other.x

    def __add__(self, y):
        if isinstance(y, int):
            return ModP(self.x + y, self.p)
        return ModP((self.x + self._same(y)) % self.p, self.p)

    __radd__ = __add__

    def __mul__(self, y):
        if isinstance(y, int):
            return ModP(self.x * y, self.p)
        if isinstance(y, Point) or (hasattr(y, "curve") and hasattr(y, "x")):
            return self.x * y
        return ModP((self.x * self._same(y)) % self.p, self.p)

    def __sub__(self, y):
        if isinstance(y, int):
            return ModP((self.x - y) % self.p, self.p)
        return ModP((self.x - self._same(y)) % self.p, self.p)

    def __rsub__(self, y):
        return -(self - y)

    def __pow__(self, n):
        return ModP(pow(self.x, n, self.p), self.p)

    def __mod__(self, other):
        return self.x % other

    def __neg__(self):
        return ModP(self.p - self.x, self.p)

    def inv(self):
        g, s, _ = egcd(self.x, self.p)
        if g != 1:
            raise Exception("modular inverse does not exist")
        return ModP(s % self.p, self.p)

    def __eq__(self, y):
        return (self.p == y.p) and (self.x % self.p == y.x % self.p)

    def __hash__(self):
        return hash((self.x % self.p, self.p))

    def __str__(self):
        return str(self.x)

    __repr__ = __str__


def mod_hash(msg: bytes, p: int, non_zero: bool = True) -> ModP:
    """Counter-prefixed SHA-256 with rejection sampling (utils.py:84-97)."""
    bits = p.bit_length()
    counter = 0
    while True:
        counter += 1
        x = int.from_bytes(sha256(str(counter).encode() + msg).digest(), "big") % (1 << bits)
        if x >= p or (non_zero and x == 0):
            continue
        return ModP(x, p)


def point_to_bytes(g) -> bytes:
    """SEC1-compressed encoding; the identity is b"\\x00" (utils.py:100-106)."""
    if g == Point.IDENTITY_ELEMENT:
        return b"\x00"
    return (b"\x03" if g.y % 2 else b"\x02") + g.x.to_bytes(BYTE_LENGTH, "big")


def point_to_b64(g) -> bytes:
    return base64.b64encode(point_to_bytes(g))


def b64_to_point(s: bytes):
    return bytes_to_point(base64.b64decode(s))


def bytes_to_point(b: bytes):
    """Decompress (utils.py:119-131).  Like the reference, the identity encoding is not decoded
    (its `b == 0` test compares bytes with an int and never fires)."""
    p = CURVE.p
    want_odd = 0 if b[0] == 2 else 1
    x = int.from_bytes(b[1:], "big")
    y = pow((x ** 3 + CURVE.a * x + CURVE.b) % p, (p + 1) // 4, p)
    return Point(x, y if y % 2 == want_odd else p - y, CURVE)


def bytes_to_points(bs) -> list:
    """[bytes_to_point(b) for b in bs] with the square roots taken on the GPU in one launch (bp_lift_x_batch).
    An x that is not on the curve raises ValueError (the reference would return a point off the curve)."""
    from .. import _native as nat
    xs = [int.from_bytes(b[1:], "big") for b in bs]
    want = bytes(0 if b[0] == 2 else 1 for b in bs)
    pts = nat.lift_x_batch(xs, want)
    if any(xy is None for xy in pts):
        raise ValueError("bytes_to_points: x coordinate not on the curve")
    return [Point(x, y, CURVE) for x, y in pts]


def b64_to_points(ss) -> list:
    return bytes_to_points([base64.b64decode(s) for s in ss])


def inner_product(a: List[ModP], b: List[ModP]) -> ModP:
    """<a, b> in Z_p (utils.py:134-137)."""
    assert len(a) == len(b)
    p = a[0].p
    return ModP(sum(int(x.x) * int(y.x) for x, y in zip(a, b)) % p, p)
