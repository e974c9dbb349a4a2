"""Single-value range proof verifier (reference: src/rangeproofs/rangeproof_verifier.py:10-97)."""
from ..curve import secp256k1
from ._core import Proof, VerifierCore

CURVE = secp256k1

__all__ = ["Proof", "RangeVerifier"]


class RangeVerifier(VerifierCore):
    def __init__(self, V, g, h, gs, hs, u, proof: Proof):
        self.V, self.g, self.h, self.gs, self.hs, self.u, self.proof = V, g, h, gs, hs, u, proof

    def verify(self):
        """True (after printing OK) or Exception("Proof invalid").

        A proof over 2..128 generator pairs first goes through the batch verifier's C entry point as a batch of one
        (every check of the reference in one call: transcript checks on the host, the four point equations on the
        device).  Only when that does not accept is the step-by-step path below replayed, so that the reference's
        exception (and its type: "Proof invalid", ValueError, IndexError, "modular inverse does not exist") surfaces."""
        n = len(self.gs)
        if 2 <= n <= 128 and n & (n - 1) == 0 and len(self.hs) == n:
            try:
                from .batch import PackedBatch, verify_packed
                acc = verify_packed(PackedBatch.from_proofs([self.V], [self.proof], n), self.g, self.h, self.gs, self.hs, self.u)
            except Exception:          # malformed proof object: let the step-by-step path raise what the reference raises
                acc = b"\x00"
            if acc == b"\x01":
                print("OK")
                return True
        return self._verify([self.V])
