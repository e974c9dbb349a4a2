"""Breakdown of bp_rp_verify_batch (config C5). Development aid."""
import contextlib, io, os, random, sys, time, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from python_bulletproofs_b200 import secp256k1, _native as nat
from python_bulletproofs_b200.utils import ModP, commitment, mod_hash, elliptic_hash
from python_bulletproofs_b200.rangeproofs import NIRangeProver
from python_bulletproofs_b200.rangeproofs.batch import PackedBatch, verify_packed
nat.init(0)
q = secp256k1.q; n = 64
pre = b"seed"
gs = [elliptic_hash(str(i).encode() + pre + b"0", secp256k1) for i in range(n)]
hs = [elliptic_hash(str(i).encode() + pre + b"1", secp256k1) for i in range(n)]
g1, h1, u1 = (elliptic_hash(pre + s, secp256k1) for s in (b"2", b"3", b"4"))
rng = random.Random(5)
Vl, pl = [], []
DISTINCT = int(os.environ.get("BP_PROBE_DISTINCT", "32"))      # distinct proofs, tiled up to `total`
for i in range(DISTINCT):
    vv = ModP(rng.getrandbits(64), q); gm = mod_hash(b"gamma%d" % i, q)
    Vl.append(commitment(g1, h1, vv, gm)); pl.append(NIRangeProver(vv, n, g1, h1, gs, hs, gm, u1, secp256k1, b"p%d" % i).prove())
total = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
batch = PackedBatch.from_proofs((Vl * (total // DISTINCT)), (pl * (total // DISTINCT)), n)
print("%d proofs, %d distinct" % (total, DISTINCT), flush=True)
for it in range(3):
    t = time.perf_counter(); acc = verify_packed(batch, g1, h1, gs, hs, u1); dt = time.perf_counter() - t
    print("verify_packed %d: %.2f ms  window c=%d" % (total, dt * 1e3, nat.load().bp_msm_last_window()), flush=True)
lib = nat.load(); lib.bp_msm_set_profiling(1)
acc = verify_packed(batch, g1, h1, gs, hs, u1)
st = (ctypes.c_float * 7)(); lib.bp_msm_stage_ms(st)
print("msm stages ms:", ["%.2f" % x for x in st])
# raw C call timing (no Python packing)
gsb, hsb = nat.pack_points(gs), nat.pack_points(hs)
gb, hb, ub = nat.pack_point(g1), nat.pack_point(h1), nat.pack_point(u1)
accept = ctypes.create_string_buffer(total)
lib.bp_msm_set_profiling(0)
ts = []
for it in range(int(os.environ.get("BP_PROBE_REPS", "3"))):
    t = time.perf_counter()
    nat.check(lib.bp_rp_verify_batch(gsb, hsb, gb, hb, ub, n, batch.records, batch.stride, total, batch.blob, batch.tr_off, batch.starts, accept))
    ts.append((time.perf_counter() - t) * 1e3)
    if it < 3:
        print("raw bp_rp_verify_batch: %.2f ms" % ts[-1])
print("raw bp_rp_verify_batch over %d calls: min %.2f median %.2f ms" % (len(ts), min(ts), sorted(ts)[len(ts) // 2]))
