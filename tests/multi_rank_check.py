"""Worker of tests/test_gpu_multi.py (one process per GPU under torch.distributed.run): the two sharded paths of SURVEY.md 8(e)
through NCCL -- a point-sliced MSM (ncclAllGather of the XYZZ partials) and a proof-sharded batch verification (ncclAllGather of
the accept bytes) -- against the CPU oracle.  Prints MULTI_OK on rank 0."""
import ctypes
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch.distributed as dist
    from oracle import ecc, protocol_oracle as po
    from python_bulletproofs_b200 import Point, secp256k1, sharding, _native as nat
    from python_bulletproofs_b200.device import DevicePoints, DeviceScalars
    from python_bulletproofs_b200.rangeproofs import NIRangeProver
    from python_bulletproofs_b200.rangeproofs.batch import PackedBatch, verify_packed_sharded
    from python_bulletproofs_b200.utils import ModP, commitment, mod_hash
    rank, local_rank, world = sharding.env_rank()
    dist.init_process_group("gloo", init_method="tcp://%s:%s" % (os.environ.get("MASTER_ADDR", "127.0.0.1"), os.environ["MASTER_PORT"]),
                            rank=rank, world_size=world)
    nat.init(local_rank)
    sharding.init_nccl()
    lib = nat.load()
    q = secp256k1.q
    # ---- one MSM of n terms cut into contiguous slices; every rank holds the same seeded vectors and takes its slice
    n = 6000
    rng = random.Random(2024)
    base = [ecc.py_mul(ecc.G, rng.getrandbits(255) + 1) for _ in range(8)]
    pts = ecc.scalar_mul_batch([base[i % 8] for i in range(n)], [rng.getrandbits(256) for _ in range(n)])
    ks = [rng.getrandbits(256) for _ in range(n)]
    want = ecc.msm(pts, [k % q for k in ks], "bucket", 2)
    dp = DevicePoints(raw=ecc.pack_points(pts))
    ds = DeviceScalars(raw=b"".join(k.to_bytes(32, "little") for k in ks))
    out = ctypes.create_string_buffer(64)
    lo, hi = sharding.slice_bounds(n, rank, world)
    for pre in (False, True):
        if pre:
            dp.precompute(0)
        nat.check(lib.bp_msm_sharded(dp.handle, ds.handle, lo, hi - lo, out))
        assert ecc.unpack_point(out.raw) == want, ("sharded MSM", rank, pre)
    # host-operand form (H2D of the slice inside the call)
    pb = ecc.pack_points(pts[lo:hi]); sb = b"".join(k.to_bytes(32, "little") for k in ks[lo:hi])
    nat.check(lib.bp_msm_sharded_host(pb, sb, hi - lo, out))
    assert ecc.unpack_point(out.raw) == want, ("sharded host MSM", rank)
    # ---- a batch of range proofs sharded by proof, accept bytes all-gathered on the device
    nbits, total = 8, 37
    seeds = [b"mr%d" % i for i in range(5)]
    ogs = [po.elliptic_hash(str(i).encode() + seeds[0]) for i in range(nbits)]
    ohs = [po.elliptic_hash(str(i).encode() + seeds[1]) for i in range(nbits)]
    og, oh, ou = (po.elliptic_hash(s) for s in seeds[2:5])
    mk = lambda t: Point(t[0], t[1], secp256k1)   # noqa: E731
    gs, hs, g, h, u = [mk(t) for t in ogs], [mk(t) for t in ohs], mk(og), mk(oh), mk(ou)
    Vs, proofs, expect = [], [], []
    import contextlib, io
    for i in range(total):
        v = rng.getrandbits(nbits)
        gamma = mod_hash(b"gm%d" % i, q)
        Vs.append(commitment(g, h, ModP(v, q), gamma))
        pr = NIRangeProver(ModP(v, q), nbits, g, h, gs, hs, gamma, u, secp256k1, b"p%d" % i).prove()
        bad = i % 5 == 3
        if bad:
            pr.t_hat = ModP((pr.t_hat.x + 1) % q, q)
        proofs.append(pr); expect.append(0 if bad else 1)
    batch = PackedBatch.from_proofs(Vs, proofs, nbits)
    for _ in range(3):                      # bucket method, byte tables, 16-bit tables
        got = verify_packed_sharded(batch, g, h, gs, hs, u, rank, world)
        assert list(got) == expect, ("sharded verify", rank, list(got))
    dist.barrier()
    if rank == 0:
        print("MULTI_OK world=%d" % world, flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
