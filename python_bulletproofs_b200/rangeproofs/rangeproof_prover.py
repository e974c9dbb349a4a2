"""Single-value range proof prover (reference: src/rangeproofs/rangeproof_prover.py:10-112)."""
from typing import List

from ..utils.transcript import Transcript
from . import _core


class NIRangeProver:
    def __init__(self, v, n: int, g, h, gs: List, hs: List, gamma, u, group, seed: bytes = b""):
        self.v, self.n, self.g, self.h, self.gs, self.hs = v, n, g, h, gs, hs
        self.gamma, self.u, self.group = gamma, u, group
        self.transcript = Transcript(seed)

    def prove(self):
        return _core.prove([self.v], self.n, self.g, self.h, self.gs, self.hs, [self.gamma], self.u,
                           self.group, self.transcript)
