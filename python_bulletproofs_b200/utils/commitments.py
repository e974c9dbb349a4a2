"""Pedersen / vector commitments (reference: src/utils/commitments.py:5-13)."""
from ..pippenger import PipSECP256k1


def commitment(g, h, x, r):
    """x*g + r*h, evaluated as one 2-term device MSM."""
    return PipSECP256k1.multiexp([g, h], [x, r])


def vector_commitment(g, h, a, b):
    assert len(g) == len(h) == len(a) == len(b)
    return PipSECP256k1.multiexp(g + h, a + b)
