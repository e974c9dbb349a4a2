"""`fastecdsa.curve` stand-in: Curve container + secp256k1 constants (SEC 2, v2.0 §2.4.1)."""


class Curve:
    def __init__(self, name, p, a, b, q, gx, gy):
        self.name, self.p, self.a, self.b, self.q, self.gx, self.gy = name, p, a, b, q, gx, gy

    def is_point_on_curve(self, point):
        x, y = point
        return (y * y - (x * x * x + self.a * x + self.b)) % self.p == 0

    @property
    def G(self):
        from .point import Point
        return Point(self.gx, self.gy, self)

    def __repr__(self):
        return self.name


secp256k1 = Curve(
    "secp256k1",
    0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFEFFFFFC2F,
    0,
    7,
    0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFEBAAEDCE6AF48A03BBFD25E8CD0364141,
    0x79BE667EF9DCBBAC55A06295CE870B07029BFCDB2DCE28D959F2815B16F81798,
    0x483ADA7726A3C4655DA4FBFC0E1108A8FD17B448A68554199C47D08FFB10D4B8,
)
