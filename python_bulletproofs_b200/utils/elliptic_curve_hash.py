"""Try-and-increment hash-to-curve used to derive generators
(reference: src/utils/elliptic_curve_hash.py:7-23).  Host-side fixture code, not on the hot path."""
from hashlib import md5, sha256

from ..point import Point


def elliptic_hash(msg: bytes, CURVE):
    p = CURVE.p
    attempt = 0
    while True:
        attempt += 1
        pre = str(attempt).encode() + msg
        x = int.from_bytes(sha256(pre).digest(), "big")
        if x >= p:
            continue
        y = pow((x ** 3 + CURVE.a * x + CURVE.b) % p, (p + 1) // 4, p)   # p = 3 (mod 4)
        if not CURVE.is_point_on_curve((x, y)):
            continue
        flip = int.from_bytes(md5(pre).digest(), "big") % 2 == 0
        return Point(x, p - y, CURVE) if flip else Point(x, y, CURVE)
