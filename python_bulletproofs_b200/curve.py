"""secp256k1 parameters behind the attribute names the reference reads from fastecdsa's Curve
(`.p .a .b .q .G`, `is_point_on_curve`)."""


class Curve:
    def __init__(self, name, p, a, b, q, gx, gy):
        self.name, self.p, self.a, self.b, self.q, self.gx, self.gy = name, p, a, b, q, gx, gy

    def is_point_on_curve(self, xy):
        x, y = xy
        return (y * y - (x * x * x + self.a * x + self.b)) % self.p == 0

    @property
    def G(self):
        from .point import Point
        return Point(self.gx, self.gy, self)

    def __repr__(self):
        return self.name


secp256k1 = Curve(
    "secp256k1",
    p=2 ** 256 - 2 ** 32 - 977,
    a=0,
    b=7,
    q=0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFEBAAEDCE6AF48A03BBFD25E8CD0364141,
    gx=0x79BE667EF9DCBBAC55A06295CE870B07029BFCDB2DCE28D959F2815B16F81798,
    gy=0x483ADA7726A3C4655DA4FBFC0E1108A8FD17B448A68554199C47D08FFB10D4B8,
)
