"""Single-value range proof verifier (reference: src/rangeproofs/rangeproof_verifier.py:10-97)."""
from ..curve import secp256k1
from ._core import Proof, VerifierCore

CURVE = secp256k1

__all__ = ["Proof", "RangeVerifier"]


class RangeVerifier(VerifierCore):
    def __init__(self, V, g, h, gs, hs, u, proof: Proof):
        self.V, self.g, self.h, self.gs, self.hs, self.u, self.proof = V, g, h, gs, hs, u, proof

    def verify(self):
        """True (after printing OK) or Exception("Proof invalid")."""
        return self._verify([self.V])
