"""Timing harness for the UNMODIFIED reference (baseline "B1" of BASELINE.md 3).

TEST / BENCH INFRASTRUCTURE ONLY -- python_bulletproofs_b200 never imports this module.

`__graft_entry__.build()` copies /root/reference/src (git-ignored, never committed) to baseline/_ref/src so that the
reference's own Python travels to the GPU box; this module puts it on sys.path together with oracle/fastecdsa_standin (the
plain-Python re-creation of the absent third-party `fastecdsa` C extension) and times

  msm     Pippenger.multiexp (src/pippenger/pippenger.py:22-61) on 2^lgn-term slices of the C3 workload, one slice per
          worker process (the reference is single threaded; a sum over point slices is how it uses more than one core)
  verify  RangeVerifier.verify (src/rangeproofs/rangeproof_verifier.py:55-86) on `count` 64-bit proofs made by the
          reference's own NIRangeProver in the workers (untimed), fanned out over the worker processes by proof

and prints one JSON object.  Used by bench.py's `--impl reference` arm and its `cpu_baseline` legs (as a subprocess, so that a
CUDA context never meets fork()).  Exit code 3 when baseline/_ref is absent.

    python -m oracle.ref_runner msm --lgn 10 --procs 16 --reps 3
    python -m oracle.ref_runner verify --count 64 --procs 16
"""
import argparse
import contextlib
import io
import json
import multiprocessing as mp
import os
import random
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.path.join(ROOT, "baseline", "_ref")


def available():
    return os.path.isdir(os.path.join(REF, "src", "pippenger"))


def _import_reference():
    sys.dont_write_bytecode = True
    for p in (os.path.join(HERE, "fastecdsa_standin"), REF):
        if p not in sys.path:
            sys.path.insert(0, p)


def _c3_slice(lgn, worker):
    """2^lgn seeded C3 terms (SURVEY.md 8d construction: lift-x points, uniform scalars) for one worker."""
    from fastecdsa.curve import secp256k1
    from fastecdsa.point import Point
    p, q = secp256k1.p, secp256k1.q
    rng = random.Random((0xB2000000 + 20) * 1000 + worker)
    n = 1 << lgn
    ks = [rng.getrandbits(256) % q for _ in range(n)]
    pts = []
    while len(pts) < n:
        x = rng.getrandbits(256)
        if x >= p:
            continue
        y = pow((x ** 3 + 7) % p, (p + 1) // 4, p)
        if (y * y - x ** 3 - 7) % p:
            continue
        if rng.getrandbits(1):
            y = p - y
        pts.append(Point(x, y, secp256k1))
    return pts, ks


_SLICE = None
_BARRIER = None
_WORKER = None


def _init_worker(counter, barrier):
    """Pool initializer: a worker id from a shared counter, and the barrier that makes every pool.map below hand exactly one
    task to every worker (a worker waiting at the barrier cannot pick up a second one)."""
    global _BARRIER, _WORKER
    with counter.get_lock():
        _WORKER = counter.value
        counter.value += 1
    _BARRIER = barrier
    _import_reference()


def _msm_prepare(lgn):
    global _SLICE
    _BARRIER.wait()
    _SLICE = _c3_slice(lgn, _WORKER)
    return True


def _msm_run(_):
    from src.pippenger import PipSECP256k1
    _BARRIER.wait()
    t = time.perf_counter()
    r = PipSECP256k1.multiexp(*_SLICE)
    return time.perf_counter() - t, (r.x, r.y)


def run_msm(lgn, procs, reps):
    """-> dict(value pts/s over all workers, per-rep wall seconds)."""
    with mp.Pool(procs, initializer=_init_worker, initargs=(mp.Value("i", 0), mp.Barrier(procs))) as pool:
        pool.map(_msm_prepare, [lgn] * procs, chunksize=1)
        walls, single = [], []
        for _ in range(reps):
            t = time.perf_counter()
            out = pool.map(_msm_run, range(procs), chunksize=1)
            walls.append(time.perf_counter() - t)
            single.append(sum(o[0] for o in out) / len(out))
    n = 1 << lgn
    return {"task": "msm", "lgn": lgn, "procs": procs, "reps": reps, "wall_s": walls, "mean_single_call_s": sum(single) / len(single),
            "pts_per_s": procs * n * len(walls) / sum(walls), "pts_per_s_one_core": n / (sum(single) / len(single))}


_PROOFS = None


def _verify_prepare(shares):
    """Each worker proves its share of the sample with the reference prover (same inputs as bench.py's C5 stream)."""
    global _PROOFS
    _BARRIER.wait()
    first, count = shares[_WORKER]
    from fastecdsa.curve import secp256k1
    from src.utils.utils import mod_hash, ModP
    from src.utils.commitments import commitment
    from src.utils.elliptic_curve_hash import elliptic_hash
    from src.rangeproofs import NIRangeProver
    q, n = secp256k1.q, 64
    seeds = [b"seed%d" % i for i in range(5)]
    gs = [elliptic_hash(str(i).encode() + seeds[0], secp256k1) for i in range(n)]
    hs = [elliptic_hash(str(i).encode() + seeds[1], secp256k1) for i in range(n)]
    g, h, u = (elliptic_hash(s, secp256k1) for s in seeds[2:5])
    rng = random.Random(5)
    vals = [rng.getrandbits(64) for _ in range(first + count)]
    items = []
    for i in range(first, first + count):
        gamma = mod_hash(b"gamma%d" % i, q)
        V = commitment(g, h, ModP(vals[i], q), gamma)
        pr = NIRangeProver(ModP(vals[i], q), n, g, h, gs, hs, gamma, u, secp256k1, b"p%d" % i).prove()
        if i % 16 == 15:
            s = str(pr.t_hat.x)
            pr.t_hat = ModP(int(s[:-1] + ("1" if s[-1] != "1" else "2")), q)
        items.append((V, pr))
    _PROOFS = (g, h, gs, hs, u, items, first)
    return count


def _verify_run(_):
    from src.rangeproofs import RangeVerifier
    _BARRIER.wait()
    g, h, gs, hs, u, items, first = _PROOFS
    out = []
    t = time.perf_counter()
    for V, pr in items:
        try:
            with contextlib.redirect_stdout(io.StringIO()):
                out.append(bool(RangeVerifier(V, g, h, gs, hs, u, pr).verify()))
        except Exception as e:   # noqa: BLE001
            if str(e) != "Proof invalid":
                raise
            out.append(False)
    return time.perf_counter() - t, first, out


def run_verify(count, procs):
    procs = min(procs, count)
    base, extra = divmod(count, procs)
    shares, lo = [], 0
    for w in range(procs):
        c = base + (1 if w < extra else 0)
        shares.append((lo, c))
        lo += c
    with mp.Pool(procs, initializer=_init_worker, initargs=(mp.Value("i", 0), mp.Barrier(procs))) as pool:
        t = time.perf_counter()
        pool.map(_verify_prepare, [shares] * procs, chunksize=1)
        prove_s = time.perf_counter() - t
        t = time.perf_counter()
        out = pool.map(_verify_run, range(procs), chunksize=1)
        wall = time.perf_counter() - t
    decisions = [None] * count
    for _, first, dec in out:
        decisions[first:first + len(dec)] = dec
    busy = sum(o[0] for o in out)
    return {"task": "verify", "count": count, "procs": procs, "wall_s": wall, "prove_wall_s": prove_s,
            "verifies_per_s": count / wall, "verifies_per_s_one_core": count / busy,
            "decisions": "".join("1" if d else "0" for d in decisions)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("task", choices=["msm", "verify", "probe"])
    ap.add_argument("--lgn", type=int, default=10)
    ap.add_argument("--procs", type=int, default=os.cpu_count() or 1)
    ap.add_argument("--reps", type=int, default=1)
    ap.add_argument("--count", type=int, default=64)
    a = ap.parse_args()
    if not available():
        print(json.dumps({"unavailable": "baseline/_ref/src missing (populated by __graft_entry__.build() where /root/reference exists)"}))
        sys.exit(3)
    if a.task == "probe":
        print(json.dumps({"available": True}))
    elif a.task == "msm":
        print(json.dumps(run_msm(a.lgn, a.procs, a.reps)))
    else:
        print(json.dumps(run_verify(a.count, a.procs)))


if __name__ == "__main__":
    main()
