"""`Pippenger(group).multiexp(gs, es)` -- drop-in for src/pippenger/pippenger.py:8-61.

Same contract as the reference: returns prod g_i^e_i (= sum e_i * g_i on the curve) as a group
element, exponents reduced mod the group order, empty input -> the unit, length mismatch ->
Exception('Different number of group elements and exponents').  The work is done by the
signed-digit bucket MSM of libbpgpu on the B200 (csrc/msm.cuh); the reference's subset-table
algorithm (pippenger.py:31-94) is not reproduced, only its result.

Besides lists of Points, `gs` may be a `DevicePoints` handle (device-resident generators,
see device.py) so that large vectors are marshalled once.
"""
import ctypes

from .. import _native as nat
from ..device import DevicePoints, DeviceScalars
from ..point import Point


class Pippenger:
    def __init__(self, group):
        self.G = group
        self.order = group.order
        self.lamb = group.order.bit_length()

    def multiexp(self, gs, es):
        if len(gs) != len(es):
            raise Exception('Different number of group elements and exponents')
        if self.order != nat.Q:
            raise NotImplementedError("libbpgpu implements secp256k1 only")
        n = len(gs)
        if n == 0:
            return self.G.unit
        out = ctypes.create_string_buffer(64)
        lib = nat.load()
        if isinstance(gs, DevicePoints):
            if isinstance(es, DeviceScalars):
                nat.check(lib.bp_msm_hh(gs.handle, es.handle, n, out))
            else:
                nat.check(lib.bp_msm_h(gs.handle, nat.pack_scalars(es), n, out))
        else:
            nat.check(lib.bp_msm(nat.pack_points(gs), nat.pack_scalars(es), n, out))
        return Point.from_bytes64(out.raw)

    def multiexp_batch(self, gss, ess):
        """Several independent multiexps in one device pass (not in the reference; used by the
        provers/verifiers here for the per-round L/R pairs and the verifier equations)."""
        offsets = [0]
        for gs, es in zip(gss, ess):
            if len(gs) != len(es):
                raise Exception('Different number of group elements and exponents')
            offsets.append(offsets[-1] + len(gs))
        raw = nat.msm_batch_bytes(b"".join(nat.pack_points(g) for g in gss),
                                  b"".join(nat.pack_scalars(e) for e in ess), offsets)
        return [Point.from_bytes64(raw, 64 * j) for j in range(len(gss))]
