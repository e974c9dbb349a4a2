"""Batch verification of aggregated range proofs (C4-shaped: m values x n bits). Development aid.
usage: python tools/aggreg_probe.py [m] [n] [total] [distinct]"""
import contextlib, io, os, random, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from python_bulletproofs_b200 import secp256k1, _native as nat
from python_bulletproofs_b200.utils import ModP, commitment, mod_hash, elliptic_hash
from python_bulletproofs_b200.rangeproofs import AggregNIRangeProver, AggregRangeVerifier
from python_bulletproofs_b200.rangeproofs.batch import PackedAggregBatch, verify_aggreg_packed
nat.init(0)
q = secp256k1.q
m = int(sys.argv[1]) if len(sys.argv) > 1 else 16
n = int(sys.argv[2]) if len(sys.argv) > 2 else 64
total = int(sys.argv[3]) if len(sys.argv) > 3 else 256
distinct = int(sys.argv[4]) if len(sys.argv) > 4 else 16
nm = n * m
pre = b"agg"
gs = [elliptic_hash(str(i).encode() + pre + b"0", secp256k1) for i in range(nm)]
hs = [elliptic_hash(str(i).encode() + pre + b"1", secp256k1) for i in range(nm)]
g, h, u = (elliptic_hash(pre + s, secp256k1) for s in (b"2", b"3", b"4"))
rng = random.Random(4)
Vl, pl = [], []
t = time.perf_counter()
for i in range(distinct):
    vs = [ModP(rng.getrandbits(n), q) for _ in range(m)]
    gammas = [mod_hash(b"g%d_%d" % (i, j), q) for j in range(m)]
    Vl.append([commitment(g, h, vs[j], gammas[j]) for j in range(m)])
    pl.append(AggregNIRangeProver(vs, n, g, h, gs, hs, gammas, u, secp256k1, b"y%d" % i).prove())
print("proved %d aggregated proofs (m=%d x n=%d) in %.2f s" % (distinct, m, n, time.perf_counter() - t), flush=True)
with contextlib.redirect_stdout(io.StringIO()):
    t = time.perf_counter(); ok = AggregRangeVerifier(Vl[0], g, h, gs, hs, u, pl[0]).verify(); one = time.perf_counter() - t
    t = time.perf_counter(); ok = AggregRangeVerifier(Vl[1], g, h, gs, hs, u, pl[1]).verify(); one = time.perf_counter() - t
print("AggregRangeVerifier.verify (one proof, class API): %.2f ms" % (one * 1e3), flush=True)
batch = PackedAggregBatch.from_proofs(Vl * (total // distinct), pl * (total // distinct), n)
for it in range(6):
    t = time.perf_counter(); acc = verify_aggreg_packed(batch, g, h, gs, hs, u); dt = time.perf_counter() - t
    assert acc == b"\x01" * total, acc[:16]
    print("verify_aggreg_packed %d proofs: %.2f ms -> %.0f proofs/s, %.0f values/s" % (total, dt * 1e3, total / dt, total * m / dt), flush=True)
