"""C2 / C1 prover latency probe (development aid): python tools/ipa_probe.py [n] [reps]"""
import contextlib, io, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from python_bulletproofs_b200 import secp256k1, _native as nat
from python_bulletproofs_b200.utils import mod_hash, elliptic_hash, vector_commitment, inner_product
from python_bulletproofs_b200.innerproduct import FastNIProver2
nat.init(0)
q = secp256k1.q
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
g = [elliptic_hash(str(i).encode() + b"ipa0", secp256k1) for i in range(N)]
h = [elliptic_hash(str(i).encode() + b"ipa1", secp256k1) for i in range(N)]
u = elliptic_hash(b"ipa2", secp256k1)
a = [mod_hash(str(i).encode() + b"a", q) for i in range(N)]
b = [mod_hash(str(i).encode() + b"b", q) for i in range(N)]
P2 = vector_commitment(g, h, a, b) + inner_product(a, b) * u
import ctypes
if os.environ.get("BP_WARM"):            # keep the SM clock up: a latency-bound proof alone does not make the GPU leave its idle clocks
    macs, ms = ctypes.c_double(), ctypes.c_float()
    for _ in range(int(os.environ["BP_WARM"])):
        nat.load().bp_imad_peak(4096, ctypes.byref(macs), ctypes.byref(ms))
for i in range(reps):
    t = time.perf_counter()
    FastNIProver2(g, h, u, P2, a, b, secp256k1).prove()
    print("prove n=%d: %.3f ms" % (N, (time.perf_counter() - t) * 1e3), flush=True)
