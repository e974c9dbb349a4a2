"""Aggregated range proof prover (reference: src/rangeproofs/rangeproof_aggreg_prover.py:10-146)."""
from typing import List

from ..utils.transcript import Transcript
from . import _core


class AggregNIRangeProver:
    def __init__(self, vs: List, n: int, g, h, gs: List, hs: List, gammas: List, u, group, seed: bytes = b""):
        self.vs, self.n, self.g, self.h, self.gs, self.hs = vs, n, g, h, gs, hs
        self.gammas, self.u, self.group = gammas, u, group
        self.transcript = Transcript(seed)
        self.m = len(vs)

    def prove(self):
        return _core.prove(self.vs, self.n, self.g, self.h, self.gs, self.hs, self.gammas, self.u,
                           self.group, self.transcript)
