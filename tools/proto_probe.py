"""Wall-clock probe of the protocol-level entry points (configs C1, C2, C4, C5). Development aid."""
import contextlib, io, os, random, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from python_bulletproofs_b200 import secp256k1, _native as nat
from python_bulletproofs_b200.utils import ModP, commitment, mod_hash, elliptic_hash, vector_commitment, inner_product
from python_bulletproofs_b200.innerproduct import NIProver, Verifier1
from python_bulletproofs_b200.rangeproofs import NIRangeProver, RangeVerifier, AggregNIRangeProver, AggregRangeVerifier
from python_bulletproofs_b200.rangeproofs.batch import PackedBatch, verify_packed

nat.init(0)
q = secp256k1.q
def T(label, fn, reps=3):
    best = 1e9
    for _ in range(reps):
        t = time.perf_counter()
        with contextlib.redirect_stdout(io.StringIO()):
            r = fn()
        best = min(best, time.perf_counter() - t)
    print("%-44s %9.2f ms" % (label, best * 1e3), flush=True)
    return r
def gens(n, pre):
    gs = [elliptic_hash(str(i).encode() + pre + b"0", secp256k1) for i in range(n)]
    hs = [elliptic_hash(str(i).encode() + pre + b"1", secp256k1) for i in range(n)]
    return gs, hs, elliptic_hash(pre + b"2", secp256k1), elliptic_hash(pre + b"3", secp256k1), elliptic_hash(pre + b"4", secp256k1)

# C2
N = 1024
g, h, _, _, u = gens(N, b"ipa")
a = [mod_hash(str(i).encode() + b"a", q) for i in range(N)]
b = [mod_hash(str(i).encode() + b"b", q) for i in range(N)]
P = T("C2 vector_commitment 2048 terms", lambda: vector_commitment(g, h, a, b))
c = inner_product(a, b)
pr = T("C2 IPA prove n=1024 (NIProver)", lambda: NIProver(g, h, u, P, c, a, b, secp256k1, b"s").prove())
T("C2 IPA verify n=1024 (Verifier1)", lambda: Verifier1(g, h, u, P, c, pr).verify())
# C1
n = 64
gs, hs, g1, h1, u1 = gens(n, b"seed")
gamma = mod_hash(b"g", q); v = ModP(2 ** 63 + 12345, q)
V = commitment(g1, h1, v, gamma)
pr1 = T("C1 range prove n=64", lambda: NIRangeProver(v, n, g1, h1, gs, hs, gamma, u1, secp256k1, b"x").prove())
T("C1 range verify n=64", lambda: RangeVerifier(V, g1, h1, gs, hs, u1, pr1).verify())
# C4
m = 16
gs4, hs4, g4, h4, u4 = gens(n * m, b"agg")
vs = [ModP((0x9E3779B97F4A7C15 * (j + 1)) % 2 ** 64, q) for j in range(m)]
gammas = [mod_hash(b"g%d" % j, q) for j in range(m)]
Vs = [commitment(g4, h4, vs[j], gammas[j]) for j in range(m)]
pr4 = T("C4 aggregated prove m=16 x 64", lambda: AggregNIRangeProver(vs, n, g4, h4, gs4, hs4, gammas, u4, secp256k1, b"y").prove(), reps=6)
T("C4 aggregated verify", lambda: AggregRangeVerifier(Vs, g4, h4, gs4, hs4, u4, pr4).verify(), reps=6)
# C5
rng = random.Random(5)
Vl, pl = [], []
for i in range(32):
    vv = ModP(rng.getrandbits(64), q); gm = mod_hash(b"gamma%d" % i, q)
    Vl.append(commitment(g1, h1, vv, gm)); pl.append(NIRangeProver(vv, n, g1, h1, gs, hs, gm, u1, secp256k1, b"p%d" % i).prove())
for total in (1024, 8192):
    batch = PackedBatch.from_proofs((Vl * (total // 32)), (pl * (total // 32)), n)
    acc = T("C5 verify_packed %d proofs" % total, lambda: verify_packed(batch, g1, h1, gs, hs, u1))
    assert acc == b"\x01" * total
# effect of the CUDA-graph replay of the IPA rounds
for on in (0, 2, 1):
    nat.load().bp_ipa_set_graphs(on)
    T("C2 IPA prove n=1024, mode=%d" % on, lambda: NIProver(g, h, u, P, c, a, b, secp256k1, b"s").prove(), reps=5)
    T("C1 range prove n=64, mode=%d" % on, lambda: NIRangeProver(v, n, g1, h1, gs, hs, gamma, u1, secp256k1, b"x").prove(), reps=5)
