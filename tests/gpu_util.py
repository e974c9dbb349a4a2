"""Helpers shared by the GPU parity tests: the product library through its C ABI (ctypes)."""
import ctypes

from python_bulletproofs_b200 import _native as nat
from oracle import ecc


def gpu_msm(pts, ks):
    """(x, y)-tuple points + int scalars -> tuple / None via bp_msm."""
    raw = nat.msm_bytes(ecc.pack_points(pts), b"".join(int(k % ecc.Q).to_bytes(32, "little") for k in ks), len(pts))
    return ecc.unpack_point(raw)


def gpu_msm_raw(pts_b, sc_b, n):
    return ecc.unpack_point(nat.msm_bytes(pts_b, sc_b, n))


def call_test(name, op, a_b, b_b, n, elt, *extra):
    out = ctypes.create_string_buffer(elt * max(n, 1))
    fn = getattr(nat.load(), name)
    nat.check(fn(op, *extra, a_b, b_b, n, out))
    return out.raw[:elt * n]
