"""bench.py -- headline benchmark of the hot path on B200.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched under torchrun)
    python bench.py --impl reference --gpus N --steps K --warmup W

Metric (BASELINE.json): secp256k1 MSM throughput in Mpts/s at 2^20 points per GPU; a "step" is one
full multi-scalar multiplication over the resident 2^20-term slice (N > 1: slice MSM on every
rank + one ncclAllGather of the 128-byte partials + the N-way sum = one MSM over N * 2^20 terms,
weak scaling).  Secondary, in the same JSON line under "verify": 64-bit range-proof verifies/s
for a batch of 8192 proofs (BASELINE config 5).

No PyTorch on the data path: device work is libbpgpu (ctypes); torch.distributed (gloo) is only the
launcher-side rendezvous for N > 1 (NCCL id exchange, barriers, max-over-ranks of the timings).
The oracle (oracle/) is used here only for the cpu_baseline leg and for the in-bench parity check.
"""
import argparse
import ctypes
import json
import os
import random
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GX = 0x79BE667EF9DCBBAC55A06295CE870B07029BFCDB2DCE28D959F2815B16F81798
GY = 0x483ADA7726A3C4655DA4FBFC0E1108A8FD17B448A68554199C47D08FFB10D4B8
METRIC = "secp256k1 MSM Mpts/s at 2^20"
LIMB_MACS_PER_FIELD_MUL = 136          # 8-limb Montgomery CIOS, the algorithmic unit
FIELD_MULS_PER_MADD = 10               # XYZZ + affine, 8M + 2S
FIELD_MULS_PER_ADD = 14                # XYZZ + XYZZ, 12M + 2S


def synth_inputs(n, seed):
    """Seeded synthetic workload: points k_i * G (made on the GPU by the scalar-mul batch kernel),
    scalars uniform 256-bit (reduced mod q on the device)."""
    from python_bulletproofs_b200 import _native as nat
    rng = random.Random(seed)
    pts = nat.scalar_mul_batch_bytes(nat.pack_xy(GX, GY) * n, rng.randbytes(32 * n), n)
    return pts, rng.randbytes(32 * n)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, t0=None, t1=None):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        rows = [r for t, r in self.rows if (t0 is None or t >= t0 - 0.2) and (t1 is None or t <= t1 + 0.2)] or [r for _, r in self.rows]
        sm = [float(r[0]) for r in rows if r and r[0].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [nm for k, nm in enumerate(names) if any(len(r) > 3 + k and r[3 + k].startswith("Active") for r in rows)]
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": float(rows[0][1]) if rows and rows[0][1].replace(".", "").isdigit() else None,
                "power_w_max": max([float(r[2]) for r in rows if r[2].replace(".", "").isdigit()], default=None),
                "samples": len(rows), "reasons": reasons}


def dist_setup(world):
    if world <= 1:
        return None
    import torch.distributed as dist
    dist.init_process_group("gloo")          # rendezvous only; MASTER_ADDR/PORT, RANK, WORLD_SIZE from torchrun
    return dist


def max_over_ranks(dist, x):
    if dist is None:
        return x
    import torch
    t = torch.tensor([x], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


def pinned_copy(nat, data):
    buf = nat.load().bp_host_alloc(len(data))
    if not buf:
        raise RuntimeError("bp_host_alloc failed")
    ctypes.memmove(buf, data, len(data))
    return buf


# ------------------------------------------------------------------------------------------------
def run_ours(args):
    from python_bulletproofs_b200 import _native as nat, sharding
    rank, local_rank, world = sharding.env_rank()
    if world != args.gpus and world > 1:
        args.gpus = world
    dist = dist_setup(world)
    nat.init(local_rank)
    lib = nat.load()
    if dist is not None:
        sharding.init_nccl()
    n = 1 << args.lgn

    # IMAD.WIDE issue-rate microbenchmark = the measured roofline denominator
    macs, ms_peak = ctypes.c_double(), ctypes.c_float()
    nat.check(lib.bp_imad_peak(1 << 16, ctypes.byref(macs), ctypes.byref(ms_peak)))

    pts, sc = synth_inputs(n, 0xB2000000 + args.lgn + 1000 * rank)
    hp, hs = ctypes.c_uint64(), ctypes.c_uint64()
    nat.check(lib.bp_points_upload(pts, n, ctypes.byref(hp)))
    nat.check(lib.bp_scalars_upload(sc, n, ctypes.byref(hs)))
    out = ctypes.create_string_buffer(64)

    # ---- timed region: K resident MSM steps, CUDA events on the library stream, L2 flushed between steps
    sampler = ClockSampler(local_rank)
    times = (ctypes.c_float * args.steps)()
    if dist is not None:
        dist.barrier()
    t0 = time.time()
    nat.check(lib.bp_bench_msm_sharded(hp, hs, 0, n, args.warmup, args.steps, 1, times, out))
    t1 = time.time()
    if dist is not None:
        dist.barrier()
    total_ms = max_over_ranks(dist, float(sum(times)))
    result_hex = out.raw.hex()
    value = world * n * args.steps / (total_ms * 1e-3) / 1e6

    # ---- per-stage device times of one MSM and of the dominant kernel (k_accumulate) alone: CUDA events recorded on
    #      the library stream around the stages / the kernel while profiling mode is on
    stage = (ctypes.c_float * 7)()
    kms = ctypes.c_float()
    acc_ms, stage_rows = [], []
    lib.bp_msm_set_profiling(1)
    for _ in range(5):
        nat.check(lib.bp_msm_hh(hp, hs, n, out))
        nat.check(lib.bp_msm_stage_ms(stage))
        nat.check(lib.bp_msm_accumulate_kernel_ms(ctypes.byref(kms)))
        acc_ms.append(kms.value)
        stage_rows.append(list(stage))
    lib.bp_msm_set_profiling(0)
    med = [statistics.median(r[i] for r in stage_rows) for i in range(7)]
    stages = {k: round(float(v), 4) for k, v in zip(["digits", "scan", "scatter", "accumulate_stage", "reduce", "combine", "total"], med)}
    c = lib.bp_msm_last_window()
    W = (128 + c - 1) // c                    # GLV: two halves < 2^128 per scalar; the top window absorbs the recoding carry
    ent = ctypes.c_uint64()
    nat.check(lib.bp_msm_last_entries(ctypes.byref(ent)))
    acc = statistics.median(acc_ms)
    # Work of one k_accumulate launch = one mixed add (XYZZ + affine, 8M + 2S) per non-zero signed digit; ent.value is
    # that count, read back from the device (~ 2 halves * n * 128 / c).
    #  * issued: this implementation spends 72 IMAD.WIDE per field multiplication (64 + 8, pseudo-Mersenne fold) and 45
    #    per squaring, i.e. 666 32x32->64 multiply-accumulates per mixed add -- the count the IMAD.WIDE pipe peak bounds;
    #  * algorithmic (SURVEY.md 8d units): 10 field multiplications x 136 limb-MAC (8-limb Montgomery CIOS) = 1360.
    issued_macs = ent.value * (8 * 72 + 2 * 45)
    alg_macs = ent.value * FIELD_MULS_PER_MADD * LIMB_MACS_PER_FIELD_MUL
    achieved = issued_macs / (acc * 1e-3) / 1e12
    peak = macs.value / 1e12
    nominal = 148 * 64 * 1965.0 * 1e6 / 1e12
    roofline = {"bound": "imad", "kernel": "k_accumulate", "achieved": round(achieved, 3), "peak": round(peak, 3),
                "unit": "T IMAD.WIDE (32x32+64 limb-MAC)/s", "frac": round(achieved / peak, 4), "traffic": 1547183120,
                "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum of one k_accumulate launch, ncu --set full with its default cache flush between replay passes, i.e. cold L2 (profiles/r1_accumulate_v3_ncu.txt); algorithmic bytes: 16.78 M gathers x (64 B point + 8 B entry) = 1.21 GB",
                "peak_source": "measured in this process: bp_imad_peak, data-dependent IMAD.WIDE.U32 stream (SASS checked); "
                               "IMAD.WIDE issues at half the 32-bit IMAD rate on B200, with or without the carry predicate",
                "nominal_imad32_peak": round(nominal, 2),
                "ncu_pipe_fmaheavy_active_pct": 80.4,
                "window_bits": c, "windows": W, "glv": True, "mixed_adds_per_launch": ent.value, "kernel_ms": round(acc, 4),
                "share_of_step": round(acc / med[6], 3),
                "issued_limb_macs_per_launch": issued_macs,
                "algorithmic_limb_macs_per_launch_survey_units": alg_macs,
                "algorithmic_rate_survey_units_T_per_s": round(alg_macs / (acc * 1e-3) / 1e12, 3),
                "whole_msm_algorithmic_limb_macs_per_pt": round((alg_macs + (W + (1 if 128 % c == 0 else 0)) * (1 << (c - 1)) * 2 * FIELD_MULS_PER_ADD * LIMB_MACS_PER_FIELD_MUL) / n, 1),
                "hbm": {"bound": "hbm", "achieved": round(ent.value * 72 / (acc * 1e-3) / 1e9, 1), "unit": "GB/s",
                        "peak": measured_hbm(), "note": "gathered 64 B point + 8 B entry per mixed add; not the binding resource"}}

    # ---- end to end through the C ABI with pinned HOST buffers (H2D + MSM + D2H inside the timed region)
    pp, ps = pinned_copy(nat, pts), pinned_copy(nat, sc)
    e2e_t = []
    for it in range(2 + args.steps):
        if dist is not None:
            dist.barrier()
        a = time.perf_counter()
        nat.check(lib.bp_msm_sharded_host(pp, ps, n, out))
        e2e_t.append(max_over_ranks(dist, time.perf_counter() - a))
    e2e_t = e2e_t[2:]
    assert out.raw.hex() == result_hex, "end-to-end result differs from resident result"
    e2e = {"value": round(world * n / (sum(e2e_t) / len(e2e_t)) / 1e6, 3), "unit": "Mpts/s", "h2d_bytes_per_step": 96 * n,
           "d2h_bytes_per_step": 64, "api": "bp_msm_sharded_host (C ABI, pinned host buffers)" if world > 1 else "bp_msm / bp_msm_sharded_host (C ABI, pinned host buffers)"}
    lib.bp_host_free(pp)
    lib.bp_host_free(ps)
    clocks = sampler.stop(t0, time.time())       # sampled from the start of the timed region through the e2e loop

    # multi-GPU parity: every rank's own slice result (no NCCL) is gathered through the launcher's process group and
    # summed with the CPU oracle's group law on rank 0; it must equal the point every rank got from the sharded call
    sharded_ok = None
    if dist is not None:
        nat.check(lib.bp_msm_hh(hp, hs, n, out))
        mine = out.raw
        parts = [None] * world
        dist.all_gather_object(parts, mine)
        if rank == 0:
            from oracle import ecc
            tot = None
            for pb in parts:
                tot = ecc.point_add(tot, ecc.unpack_point(pb))
            sharded_ok = ecc.pack_point(tot).hex() == result_hex

    line = {"metric": METRIC, "value": round(value, 3), "unit": "Mpts/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(total_ms / args.steps, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32x8 (256-bit modular integer)", "data": "synthetic",
            "config": {"workload": "C3: standalone MSM, 2^%d secp256k1 points + 256-bit scalars resident per GPU" % args.lgn,
                       "terms_per_gpu": n, "terms_total": world * n, "window_bits": c,
                       "l2": "flushed between steps (256 MiB memset outside the timed events)",
                       "parallelism": "slice%d" % world, "seed": "0xB2000000+lgn(+1000*rank), points k_i*G"},
            "gpu_launches": 13 * args.steps,
            "clocks": clocks, "e2e": e2e, "roofline": roofline, "stages_ms": stages, "result": result_hex}

    if rank == 0 and world == 1:
        line["cpu_baseline"], bit_exact = cpu_baseline(pts, sc, n, result_hex)
        line["bit_exact_vs_oracle"] = bit_exact
    if sharded_ok is not None:
        line["sharded_result_equals_sum_of_slices"] = sharded_ok
    if not args.no_verify:
        try:
            line["verify"] = bench_verify(args, nat, dist, rank, world)
        except Exception as e:   # noqa: BLE001  -- the headline metric must still print
            line["verify"] = {"error": repr(e)}
    lib.bp_handle_free(hp)
    lib.bp_handle_free(hs)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def measured_hbm():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f)["hbm_gbs"]
    except Exception:   # noqa: BLE001
        return 6650.0


def cpu_baseline(pts, sc, n, gpu_hex):
    """The oracle port (C, bucket method, all host threads) on the SAME 2^lgn inputs, plus the
    reference's own subset-table algorithm restated in C on a 2^12 sample (it is impractical beyond
    ~2^14)."""
    from oracle import ecc
    cores = ecc.max_threads()
    t = time.perf_counter()
    res = ecc.msm_bytes(pts, sc, n, "bucket", cores)
    dt = time.perf_counter() - t
    bit_exact = ecc.pack_point(res).hex() == gpu_hex
    ns = 1 << 12
    t = time.perf_counter()
    ecc.msm_bytes(pts[:64 * ns], sc[:32 * ns], ns, "subset", 1)
    dts = time.perf_counter() - t
    return ({"value": round(n / dt / 1e6, 4), "unit": "Mpts/s", "cores": cores, "kind": "port",
             "sample": "full workload (2^%d terms) once, oracle/ecc_oracle.c bucket MSM on %d OpenMP threads, %.2f s; "
                       "reference's subset-table algorithm restated in C, 1 thread, 2^12 terms: %.4f Mpts/s"
                       % (n.bit_length() - 1, cores, dt, ns / dts / 1e6)}, bit_exact)


# ---- secondary metric: batch verification of 64-bit range proofs (config 5) ----------------------
def bench_verify(args, nat, dist, rank, world):
    import contextlib
    import io
    from python_bulletproofs_b200 import Point, secp256k1
    from python_bulletproofs_b200.rangeproofs import NIRangeProver
    from python_bulletproofs_b200.rangeproofs.batch import PackedBatch, verify_packed
    from python_bulletproofs_b200.utils import ModP, commitment, mod_hash, elliptic_hash
    q, nbits = secp256k1.q, 64
    seeds = [b"seed%d" % i for i in range(5)]
    gs = [elliptic_hash(str(i).encode() + seeds[0], secp256k1) for i in range(nbits)]
    hs = [elliptic_hash(str(i).encode() + seeds[1], secp256k1) for i in range(nbits)]
    g, h, u = (elliptic_hash(s, secp256k1) for s in seeds[2:5])
    rng = random.Random(5)
    distinct = args.verify_distinct
    Vs, proofs = [], []
    t = time.perf_counter()
    for i in range(distinct):
        v = rng.getrandbits(64)
        gamma = mod_hash(b"gamma%d" % i, q)
        Vs.append(commitment(g, h, ModP(v, q), gamma))
        pr = NIRangeProver(ModP(v, q), nbits, g, h, gs, hs, gamma, u, secp256k1, b"p%d" % i).prove()
        if i % 16 == 15:       # every 16th proof corrupted: flip one decimal digit of t_hat
            s = str(pr.t_hat.x)
            pr.t_hat = ModP(int(s[:-1] + ("1" if s[-1] != "1" else "2")), q)
        proofs.append(pr)
    prove_s = (time.perf_counter() - t) / distinct
    total = args.verify_proofs
    reps = (total + distinct - 1) // distinct
    batch = PackedBatch.from_proofs((Vs * reps)[:total], (proofs * reps)[:total], nbits)
    lo, hi = (0, total) if world == 1 else __import__("python_bulletproofs_b200.sharding", fromlist=["x"]).slice_bounds(total, rank, world)
    times = []
    acc = b""
    for it in range(2 + 8):      # the first calls build the generator tables (byte windows, then 16-bit windows)
        if dist is not None:
            dist.barrier()
        a = time.perf_counter()
        acc = verify_packed(batch, g, h, gs, hs, u, lo, hi - lo)
        if world > 1:
            from python_bulletproofs_b200 import sharding
            acc = sharding.gather_accept(acc, [b - a_ for a_, b in sharding.all_slices(total, world)])
        times.append(max_over_ranks(dist, time.perf_counter() - a))
    best = min(times[2:])
    want = bytes([0 if (i % distinct) % 16 == 15 else 1 for i in range(total)])
    return {"metric": "64-bit range-proof verifies/s", "value": round(total / best, 1), "unit": "verifies/s", "n_gpus": world,
            "scaling": "strong", "proofs": total, "distinct_proofs": distinct, "decisions_ok": acc == want,
            "rejected": acc.count(b"\x00"), "ms_per_batch": round(best * 1e3, 2),
            "note": "end to end through bp_rp_verify_batch with HOST buffers (packed proofs + transcripts H2D, accept bytes D2H); "
                    "proofs made by the GPU prover (%.1f ms/proof), %d distinct proofs tiled to %d, every 16th corrupted" % (prove_s * 1e3, distinct, total)}


# ------------------------------------------------------------------------------------------------
def run_reference(args):
    """Reference arm: the CPU implementation of the path (oracle port; the reference itself is pure
    Python + an absent third-party C extension and cannot travel), all host threads, same metric."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import ecc
    cores = ecc.max_threads()
    lgs = min(args.lgn, 18)
    ns = 1 << lgs
    rng = random.Random(0xB2000000 + args.lgn)
    base = [ecc.py_mul(ecc.G, rng.getrandbits(256)) for _ in range(16)]
    pts_l = ecc.scalar_mul_batch([base[i % 16] for i in range(ns)], [rng.getrandbits(256) for _ in range(ns)])
    pts, sc = ecc.pack_points(pts_l), rng.randbytes(32 * ns)
    ts = []
    for it in range(args.warmup + args.steps):
        t = time.perf_counter()
        ecc.msm_bytes(pts, sc, ns, "bucket", cores)
        ts.append(time.perf_counter() - t)
    ts = ts[args.warmup:]
    val = round(ns * len(ts) / sum(ts) / 1e6, 4)
    sample = "2^%d-term sample of the 2^%d workload per step, oracle/ecc_oracle.c bucket MSM, %d OpenMP threads" % (lgs, args.lgn, cores)
    print(json.dumps({"impl": "reference", "metric": METRIC, "value": val, "unit": "Mpts/s", "n_gpus": args.gpus, "steps": args.steps,
                      "warmup": args.warmup, "ms_per_step": round(1e3 * sum(ts) / len(ts), 3), "higher_is_better": True, "scaling": "weak",
                      "vs_baseline": None, "dtype": "u64x4 (256-bit modular integer, CPU)", "data": "synthetic",
                      "config": {"workload": "C3: standalone MSM, 2^%d secp256k1 points (CPU arm: bounded sample)" % args.lgn, "sample_terms": ns},
                      "cpu_baseline": {"value": val, "unit": "Mpts/s", "cores": cores, "kind": "port", "sample": sample},
                      "e2e": {"value": val, "unit": "Mpts/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--lgn", type=int, default=20)
    ap.add_argument("--verify-proofs", type=int, default=8192)
    ap.add_argument("--verify-distinct", type=int, default=64)
    ap.add_argument("--no-verify", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
