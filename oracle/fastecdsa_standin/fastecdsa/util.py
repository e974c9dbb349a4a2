"""`fastecdsa.util` stand-in: mod_sqrt for p = 3 (mod 4) returning both roots."""


def mod_sqrt(a, p):
    r = pow(a, (p + 1) // 4, p)
    return r, p - r
