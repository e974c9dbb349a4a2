"""ctypes front-end of the C oracle (oracle/ecc_oracle.c) plus byte packers.

TEST INFRASTRUCTURE ONLY -- see the header of ecc_oracle.c.  Imported by tests/,
__graft_entry__.smoke() and bench.py's CPU-baseline legs; never by python_bulletproofs_b200.

Points cross this boundary as (x, y) integer tuples, identity = None.
"""
import ctypes
import os
import subprocess

P = 0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFEFFFFFC2F
Q = 0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFEBAAEDCE6AF48A03BBFD25E8CD0364141
GX = 0x79BE667EF9DCBBAC55A06295CE870B07029BFCDB2DCE28D959F2815B16F81798
GY = 0x483ADA7726A3C4655DA4FBFC0E1108A8FD17B448A68554199C47D08FFB10D4B8
G = (GX, GY)

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libecc_oracle.so")
_lib = None


def build(force=False):
    """Compile oracle/ecc_oracle.c with the recipe in oracle/Makefile."""
    src = os.path.join(_HERE, "ecc_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE], stdout=subprocess.DEVNULL)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
    return _lib


# ---- packers -------------------------------------------------------------------------------
def pack_point(pt):
    if pt is None:
        return bytes(64)
    return pt[0].to_bytes(32, "little") + pt[1].to_bytes(32, "little")


def unpack_point(b):
    b = bytes(b)
    if b == bytes(64):
        return None
    return (int.from_bytes(b[:32], "little"), int.from_bytes(b[32:64], "little"))


def pack_points(pts):
    return b"".join(pack_point(p) for p in pts)


def unpack_points(b, n):
    b = bytes(b)
    return [unpack_point(b[64 * i:64 * i + 64]) for i in range(n)]


def pack_scalars(ks, reduce=False):
    """32-byte little-endian each.  The C side reduces mod q once (values must be < 2^256)."""
    if reduce:
        ks = [int(k) % Q for k in ks]
    return b"".join(int(k).to_bytes(32, "little") for k in ks)


# ---- C calls ------------------------------------------------------------------------------
def point_add(a, b):
    out = ctypes.create_string_buffer(64)
    lib().orc_point_add(pack_point(a), pack_point(b), out)
    return unpack_point(out.raw)


def point_neg(a):
    return None if a is None else (a[0], (-a[1]) % P)


def on_curve(a):
    return bool(lib().orc_on_curve(pack_point(a)))


def scalar_mul_batch(pts, ks):
    n = len(pts)
    out = ctypes.create_string_buffer(64 * max(n, 1))
    lib().orc_scalar_mul_batch(pack_points(pts), pack_scalars(ks, True), ctypes.c_size_t(n), out)
    return unpack_points(out.raw, n)


def scalar_mul(pt, k):
    return scalar_mul_batch([pt], [k])[0]


def fold(lo, hi, x_lo, x_hi):
    n = len(lo)
    out = ctypes.create_string_buffer(64 * max(n, 1))
    lib().orc_fold(pack_points(lo), pack_points(hi), ctypes.c_size_t(n),
                   pack_scalars([x_lo], True), pack_scalars([x_hi], True), out)
    return unpack_points(out.raw, n)


def msm(pts, ks, algo="bucket", threads=1):
    """Result semantics of Pippenger.multiexp (/root/reference/src/pippenger/pippenger.py:22-29)."""
    if len(pts) != len(ks):
        raise Exception("Different number of group elements and exponents")
    return msm_bytes(pack_points(pts), pack_scalars(ks, True), len(pts), algo, threads)


def msm_bytes(pts_b, sc_b, n, algo="bucket", threads=1):
    out = ctypes.create_string_buffer(64)
    L = lib()
    if algo == "naive":
        L.orc_msm_naive(pts_b, sc_b, ctypes.c_size_t(n), out)
    elif algo == "subset":
        L.orc_msm_subset(pts_b, sc_b, ctypes.c_size_t(n), out)
    else:
        L.orc_msm_bucket(pts_b, sc_b, ctypes.c_size_t(n), out, ctypes.c_int(threads))
    return unpack_point(out.raw)


def max_threads():
    return int(lib().orc_max_threads())


# ---- pure-Python affine group law (independent cross-check of the C code) ---------------
def py_add(a, b):
    if a is None:
        return b
    if b is None:
        return a
    if a[0] == b[0]:
        if (a[1] + b[1]) % P == 0:
            return None
        lam = 3 * a[0] * a[0] * pow(2 * a[1], -1, P) % P
    else:
        lam = (b[1] - a[1]) * pow(b[0] - a[0], -1, P) % P
    x3 = (lam * lam - a[0] - b[0]) % P
    return (x3, (lam * (a[0] - x3) - a[1]) % P)


def py_mul(a, k):
    k %= Q
    acc = None
    while k:
        if k & 1:
            acc = py_add(acc, a)
        a = py_add(a, a)
        k >>= 1
    return acc


def lift_x(x, odd):
    """Point with given x and y parity, or None (same lift as reference utils.py:127-131)."""
    y = pow((x * x * x + 7) % P, (P + 1) // 4, P)
    if (y * y - x * x * x - 7) % P:
        return None
    return (x, y if (y & 1) == odd else P - y)
