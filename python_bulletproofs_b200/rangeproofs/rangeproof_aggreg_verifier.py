"""Aggregated range proof verifier (reference: src/rangeproofs/rangeproof_aggreg_verifier.py:10-108)."""
from ..curve import secp256k1
from ._core import Proof, VerifierCore

CURVE = secp256k1

__all__ = ["Proof", "AggregRangeVerifier"]


class AggregRangeVerifier(VerifierCore):
    def __init__(self, Vs, g, h, gs, hs, u, proof: Proof):
        self.Vs, self.g, self.h, self.gs, self.hs, self.u, self.proof = Vs, g, h, gs, hs, u, proof

    def verify(self):
        """True (after printing OK) or Exception("Proof invalid").

        Like RangeVerifier.verify: the proof first goes through the batch entry point for aggregated proofs as a batch of one
        (bp_rp_verify_aggreg_batch: transcript checks and all scalars on the host, the four point equations on the device in one
        pass); only when that does not accept is the step-by-step path replayed, so that the reference's exception and its
        type surface unchanged."""
        Vs = list(self.Vs)
        nm, m = len(self.gs), len(Vs)
        if m >= 1 and nm % m == 0 and 2 <= nm <= 2048 and nm & (nm - 1) == 0 and len(self.hs) == nm:
            try:
                from .batch import PackedAggregBatch, verify_aggreg_packed
                acc = verify_aggreg_packed(PackedAggregBatch.from_proofs([Vs], [self.proof], nm // m), self.g, self.h, self.gs, self.hs, self.u)
            except Exception:          # malformed proof object: let the step-by-step path raise what the reference raises
                acc = b"\x00"
            if acc == b"\x01":
                print("OK")
                return True
        return self._verify(Vs)
