"""Shared input builders for the parity tests (oracle side only; test infrastructure)."""
import random

from oracle import ecc, protocol_oracle as po

Q, P = ecc.Q, ecc.P


def c3_inputs(lgn, n=None):
    """SURVEY.md 8(d) C3 synthetic inputs -- same stream as oracle/gen_golden.py:c3_inputs."""
    rng = random.Random(0xB2000000 + lgn)
    n = n if n is not None else 1 << lgn
    ks = [rng.getrandbits(256) % Q for _ in range(n)]
    pts = []
    while len(pts) < n:
        x = rng.getrandbits(256)
        if x >= P:
            continue
        y = pow((x ** 3 + 7) % P, (P + 1) // 4, P)
        if (y * y - x ** 3 - 7) % P:
            continue
        if rng.getrandbits(1):
            y = P - y
        pts.append((x, y))
    return pts, ks


def fast_points(n, seed):
    """n pseudo-random curve points, cheap to make at large n: small multiples walk
    P_{i+1} = P_i + D (affine adds on Python ints would be slow, so use x-lifts)."""
    rng = random.Random(seed)
    pts = []
    while len(pts) < n:
        pt = ecc.lift_x(rng.getrandbits(256) % P, rng.getrandbits(1))
        if pt is not None:
            pts.append(pt)
    return pts


def explicit_case(case):
    pts = [None if s == "00" else po.dec_point(bytes.fromhex(s)) for s in case["pts"]]
    ks = [int(k) for k in case["ks"]]
    return pts, ks


def gens(nm, seeds):
    """Generator construction of src/tests/test_rangeproofs.py:20-29 via the oracle's elliptic_hash."""
    seeds = [s.encode() if isinstance(s, str) else s for s in seeds]
    gs = [po.elliptic_hash(str(i).encode() + seeds[0]) for i in range(nm)]
    hs = [po.elliptic_hash(str(i).encode() + seeds[1]) for i in range(nm)]
    return gs, hs, po.elliptic_hash(seeds[2]), po.elliptic_hash(seeds[3]), po.elliptic_hash(seeds[4])


def ipa_inputs(N, seeds):
    """src/tests/test_innerprod.py:104-116."""
    seeds = [s.encode() if isinstance(s, str) else s for s in seeds]
    g = [po.elliptic_hash(str(i).encode() + seeds[0]) for i in range(N)]
    h = [po.elliptic_hash(str(i).encode() + seeds[1]) for i in range(N)]
    u = po.elliptic_hash(seeds[2])
    a = [po.mod_hash(str(i).encode() + seeds[3]) for i in range(N)]
    b = [po.mod_hash(str(i).encode() + seeds[4]) for i in range(N)]
    return g, h, u, a, b
