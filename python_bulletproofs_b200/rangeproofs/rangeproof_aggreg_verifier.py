"""Aggregated range proof verifier (reference: src/rangeproofs/rangeproof_aggreg_verifier.py:10-108)."""
from ..curve import secp256k1
from ._core import Proof, VerifierCore

CURVE = secp256k1

__all__ = ["Proof", "AggregRangeVerifier"]


class AggregRangeVerifier(VerifierCore):
    def __init__(self, Vs, g, h, gs, hs, u, proof: Proof):
        self.Vs, self.g, self.h, self.gs, self.hs, self.u, self.proof = Vs, g, h, gs, hs, u, proof

    def verify(self):
        return self._verify(list(self.Vs))
