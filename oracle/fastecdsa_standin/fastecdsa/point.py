"""`fastecdsa.point` stand-in: affine short-Weierstrass arithmetic on Python ints.

The identity is the point (0, 0) with curve None, reachable as `Point.IDENTITY_ELEMENT`
on the class and on instances (the reference uses `curve.G.IDENTITY_ELEMENT`,
/root/reference/src/pippenger/group.py:29).
"""


class Point:
    IDENTITY_ELEMENT = None  # set below

    def __init__(self, x, y, curve=None):
        self.x, self.y, self.curve = x, y, curve

    def _is_identity(self):
        return self.curve is None

    def __eq__(self, other):
        if not isinstance(other, Point):
            return NotImplemented
        if self._is_identity() or other._is_identity():
            return self._is_identity() and other._is_identity()
        return self.x == other.x and self.y == other.y

    def __hash__(self):
        return hash((self.x, self.y))

    def __neg__(self):
        if self._is_identity():
            return self
        return Point(self.x, (-self.y) % self.curve.p, self.curve)

    def __add__(self, other):
        if self._is_identity():
            return other
        if other._is_identity():
            return self
        c = self.curve
        p = c.p
        if self.x == other.x:
            if (self.y + other.y) % p == 0:
                return Point.IDENTITY_ELEMENT
            lam = (3 * self.x * self.x + c.a) * pow(2 * self.y, -1, p) % p
        else:
            lam = (other.y - self.y) * pow(other.x - self.x, -1, p) % p
        x3 = (lam * lam - self.x - other.x) % p
        return Point(x3, (lam * (self.x - x3) - self.y) % p, c)

    def __sub__(self, other):
        return self + (-other)

    def __mul__(self, k):
        if self._is_identity():
            return self
        k = int(k) % self.curve.q
        acc, base = Point.IDENTITY_ELEMENT, self
        while k:
            if k & 1:
                acc = acc + base
            base = base + base
            k >>= 1
        return acc

    __rmul__ = __mul__

    def __repr__(self):
        return "<identity>" if self._is_identity() else "X: 0x%x\nY: 0x%x" % (self.x, self.y)


Point.IDENTITY_ELEMENT = Point(0, 0, None)
