"""Inner-product argument, prover side (reference: src/innerproduct/inner_product_prover.py).

`FastNIProver2.prove` hands the whole round loop to libbpgpu (bp_ipa_prove): per round the two
(2k+1)-term commitments L, R, the shared-scalar generator folding and the a/b folding run in
CUDA, while the Fiat-Shamir challenge is hashed on the host between the launches.
"""
import ctypes
from typing import Optional

from .. import _native as nat
from ..curve import secp256k1
from ..pippenger import PipSECP256k1
from ..point import Point
from ..utils.transcript import Transcript
from ..utils.utils import ModP
from .inner_product_verifier import Proof1, Proof2


_U_NEW = {}      # (u.x, u.y, x) -> x*u: with the default seed Protocol 1's challenge is a constant, so is u_new per generator u


def _scaled_u(u, x):
    key = (u.x, u.y, int(x % nat.Q))
    hit = _U_NEW.get(key)
    if hit is None:
        if len(_U_NEW) > 256:
            _U_NEW.clear()
        hit = _U_NEW[key] = PipSECP256k1.multiexp([u], [x])
    return hit


class NIProver:
    """Protocol 1 (inner_product_prover.py:11-45)."""

    def __init__(self, g, h, u, P, c, a, b, group, seed=b"", _h_scale=None, _P_msm=None, _packed=None):
        assert _packed is not None or len(g) == len(h) == len(a) == len(b)
        self.g, self.h, self.u, self.P, self.c, self.a, self.b = g, h, u, P, c, a, b
        self.group = group
        self.transcript = Transcript(seed)
        self._h_scale = _h_scale      # private: effective generators are _h_scale[i] * h[i] (never materialised)
        self._P_msm = _P_msm          # private: P given as (points, scalars) of an MSM not yet evaluated (P may be None)
        # private: wire-form operands from the range prover's C algebra -- dict(n, g, h, a, b, h_scale, bs = b*h_scale)
        # of packed bytes; a, b, _h_scale and _P_msm are then ignored
        self._packed = _packed

    def prove(self) -> Proof1:
        x = self.transcript.get_modp(self.group.q)
        self.transcript.add_number(x)
        # P_new = P + (x*c)*u ; u_new = x*u.  When the caller passed P as an unevaluated MSM (range-proof prover), P_new
        # is that MSM with one more term: one device pass instead of two (P itself is not part of the proof).
        if self._packed is not None:
            pk = self._packed
            u_new = _scaled_u(self.u, x)
            out = ctypes.create_string_buffer(64)
            nat.check(nat.load().bp_ipa_statement(pk["g"], pk["h"], nat.pack_point(u_new), pk["a"], pk["bs"],
                                                  nat.pack_scalar(self.c), pk["n"], out))
            P_new = Point.from_bytes64(out.raw, 0)          # = P + (x*c)*u with u_new = x*u:  sum a_i g_i + sum bs_i h_i + c*u_new
        elif self._P_msm is not None:
            pts, scs = self._P_msm
            P_new, u_new = PipSECP256k1.multiexp_batch([list(pts) + [self.u], [self.u]], [list(scs) + [x * self.c], [x]])
        else:
            # u_new = x*u (cached per generator), P_new = P + (x*c)*u: a one-term multiexp on u (table) and one point addition
            u_new = _scaled_u(self.u, x)
            P_new = self.P + PipSECP256k1.multiexp([self.u], [x * self.c])
        inner = FastNIProver2(self.g, self.h, u_new, P_new, self.a, self.b, self.group, self.transcript.digest,
                              _h_scale=self._h_scale, _packed=self._packed)
        return Proof1(u_new, P_new, inner.prove(), self.transcript.digest)


class FastNIProver2:
    """Protocol 2 (inner_product_prover.py:48-110)."""

    def __init__(self, g, h, u, P, a, b, group, transcript: Optional[bytes] = None, _h_scale=None, _packed=None):
        self._packed = _packed
        n_ = _packed["n"] if _packed is not None else len(a)
        assert _packed is not None or len(g) == len(h) == len(a) == len(b)
        self._h_scale = _h_scale
        assert n_ & (n_ - 1) == 0
        self.log_n = n_.bit_length() - 1
        self.n = n_
        self.g, self.h, self.u, self.P, self.a, self.b = g, h, u, P, a, b
        self.group = group
        self.transcript = Transcript()
        if transcript:
            self.transcript.digest += transcript
            self.init_transcript_length = len(transcript.split(b"&"))
        else:
            self.init_transcript_length = 1

    def prove(self) -> Proof2:
        n, rounds, q = self.n, self.log_n, self.group.q
        if q != nat.Q:
            raise NotImplementedError("libbpgpu implements secp256k1 only")
        start = self.transcript.digest
        cap = len(start) + rounds * (2 * 45 + 80) + 16
        Ls = ctypes.create_string_buffer(64 * max(rounds, 1))
        Rs = ctypes.create_string_buffer(64 * max(rounds, 1))
        xs = ctypes.create_string_buffer(32 * max(rounds, 1))
        a_out, b_out = ctypes.create_string_buffer(32), ctypes.create_string_buffer(32)
        t_out = ctypes.create_string_buffer(cap)
        t_len = ctypes.c_size_t(0)
        if self._packed is not None:
            pk = self._packed
            gb, hb, hscale, ab, bb = pk["g"], pk["h"], pk["h_scale"], pk["a"], pk["b"]
        else:
            hscale = nat.pack_scalars(self._h_scale) if self._h_scale is not None else None
            gb, hb, ab, bb = nat.pack_points(self.g), nat.pack_points(self.h), nat.pack_scalars(self.a), nat.pack_scalars(self.b)
        nat.check(nat.load().bp_ipa_prove_hs(
            gb, hb, hscale, nat.pack_point(self.u), ab, bb, n, start, len(start),
            Ls, Rs, xs, a_out, b_out, t_out, cap, ctypes.byref(t_len)))
        self.transcript.digest = t_out.raw[:t_len.value]
        return Proof2(
            ModP(int.from_bytes(a_out.raw, "little"), q),
            ModP(int.from_bytes(b_out.raw, "little"), q),
            [ModP(v, q) for v in nat.unpack_scalars(xs.raw, rounds)],
            [Point.from_bytes64(Ls.raw, 64 * i) for i in range(rounds)],
            [Point.from_bytes64(Rs.raw, 64 * i) for i in range(rounds)],
            self.transcript.digest,
            self.init_transcript_length,
        )
