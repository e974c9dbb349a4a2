"""Group plug-in of the multi-exponentiation (reference: src/pippenger/group.py:6-16,27-32).

`EC` is the only group on the hot path.  The reference's `MultIntModP` toy group (an
operation counter over Z_p*, group.py:19-24, modp.py) has no caller in the proofs and is out
of scope.
"""
from abc import ABC, abstractmethod


class Group(ABC):
    def __init__(self, unit, order):
        self.unit = unit
        self.order = order

    @abstractmethod
    def mult(self, x, y):
        """the group operation"""

    def square(self, x):
        return self.mult(x, x)


class EC(Group):
    """Points of an elliptic curve written multiplicatively: mult = point addition."""

    def __init__(self, curve):
        super().__init__(curve.G.IDENTITY_ELEMENT, curve.q)
        self.curve = curve

    def mult(self, x, y):
        return x + y
