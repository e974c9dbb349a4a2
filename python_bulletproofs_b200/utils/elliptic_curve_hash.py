"""Try-and-increment hash-to-curve used to derive generators
(reference: src/utils/elliptic_curve_hash.py:7-23).

`elliptic_hash` is the reference's one-message function (host big-int arithmetic, a single square root).
`elliptic_hash_batch` derives many generators at once: the SHA-256 / MD5 try-and-increment stays on the host, the
modular square roots and curve checks of every pending candidate run in one `bp_lift_x_batch` launch per attempt
round (deriving 2^20 generators costs ~30 s of Python square roots otherwise)."""
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


def elliptic_hash_batch(msgs, CURVE):
    """[elliptic_hash(m, CURVE) for m in msgs], square roots on the GPU (secp256k1 only)."""
    from .. import _native as nat
    from ..curve import secp256k1
    if CURVE.p != secp256k1.p or CURVE.a != 0 or CURVE.b != 7:
        return [elliptic_hash(m, CURVE) for m in msgs]
    p = CURVE.p
    out = [None] * len(msgs)
    pending = list(range(len(msgs)))
    attempt = 0
    while pending:
        attempt += 1
        idx, xs, want = [], [], bytearray()
        for i in pending:
            pre = str(attempt).encode() + msgs[i]
            x = int.from_bytes(sha256(pre).digest(), "big")
            if x >= p:
                continue                                   # retried with the next counter, like the reference's loop
            idx.append(i)
            xs.append(x)
            want.append(3 if int.from_bytes(md5(pre).digest(), "big") % 2 == 0 else 2)
        lifted = nat.lift_x_batch(xs, want)
        done = set()
        for i, xy in zip(idx, lifted):
            if xy is not None:
                out[i] = Point(xy[0], xy[1], CURVE)
                done.add(i)
        pending = [i for i in pending if i not in done]
    return out
