"""Point type with the surface the reference uses from fastecdsa.point.Point
(`.x .y .curve`, `IDENTITY_ELEMENT`, `+`, `int * Point`, `==`).

The group law itself runs on the GPU: `P + Q` is one launch of the complete addition (bp_point_add), `k * P` a 1-term
call into the CUDA MSM (libbpgpu, no CPU fallback).  They exist for API compatibility (e.g. `commitment`,
utils/commitments.py); the provers/verifiers of this package batch their point work into
larger device calls instead of using these operators in loops.
"""
import ctypes

from . import _native as nat
from .curve import secp256k1


class Point:
    IDENTITY_ELEMENT = None   # assigned below; also visible on instances (group.py:29 uses G.IDENTITY_ELEMENT)

    __slots__ = ("x", "y", "curve", "_pk")

    def __init__(self, x, y, curve=secp256k1):
        # fastecdsa's Point raises ValueError for coordinates that are not on the curve; so does this one -- every point a
        # verifier receives (V, A, S, T1, T2, L_j, R_j, u_new, P_new) is thereby on secp256k1 before it reaches the device
        # formulas (which assume curve points; the C ABI's batch verifier checks its packed inputs itself)
        if curve is not None and not (0 <= x < curve.p and 0 <= y < curve.p and curve.is_point_on_curve((x, y))):
            raise ValueError("coordinates are not on curve <%s>\n\tx=%x\n\ty=%x" % (curve, x, y))
        self.x, self.y, self.curve = x, y, curve
        self._pk = None        # (x, y, 64-byte wire form) cached by _native.pack_point; revalidated against x, y on use

    # -- boundary helpers
    @classmethod
    def from_bytes64(cls, b, off=0):
        xy = nat.unpack_xy(b, off)
        if xy is None:
            return cls.IDENTITY_ELEMENT
        pt = cls.__new__(cls)              # device output: canonical curve point, no need to re-check
        pt.x, pt.y, pt.curve = xy[0], xy[1], secp256k1
        pt._pk = (pt.x, pt.y, bytes(b[off:off + 64]))
        return pt

    def _is_identity(self):
        return self.curve is None

    # -- group law on the device
    def __add__(self, other):
        if not hasattr(other, "x") or not hasattr(other, "curve"):
            return NotImplemented
        out = ctypes.create_string_buffer(64)
        nat.check(nat.load().bp_point_add(nat.pack_point(self), nat.pack_point(other), out))
        return Point.from_bytes64(out.raw, 0)

    def __radd__(self, other):
        return self.__add__(other)

    def __neg__(self):
        if self._is_identity():
            return self
        pt = Point.__new__(Point)
        pt.x, pt.y, pt.curve, pt._pk = self.x, (-self.y) % self.curve.p, self.curve, None
        return pt

    def __sub__(self, other):
        return self + (-other)

    def __mul__(self, k):
        if not isinstance(k, int):
            k = k % secp256k1.q          # ModP and friends
        return Point.from_bytes64(nat.msm_bytes(nat.pack_point(self), nat.pack_scalar(k), 1))

    __rmul__ = __mul__

    def __eq__(self, other):
        if not hasattr(other, "x") or not hasattr(other, "curve"):
            return NotImplemented
        a, b = self.curve is None, other.curve is None
        if a or b:
            return a and b
        return self.x == other.x and self.y == other.y

    def __hash__(self):
        return hash((self.x, self.y, self.curve is None))

    def __repr__(self):
        return "<POINT AT INFINITY>" if self._is_identity() else "X: 0x%x\nY: 0x%x\n(On curve <%s>)" % (self.x, self.y, self.curve)


Point.IDENTITY_ELEMENT = Point(0, 0, None)
