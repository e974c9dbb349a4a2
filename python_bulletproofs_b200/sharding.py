"""Host-side plumbing for the two paths that shard across GPUs:

  * a large MSM: contiguous point slice per rank, one all-gather of 128-byte XYZZ partials;
  * a batch of independent range-proof verifications: contiguous block of proofs per rank, one
    all-gather of the accept bytes.

One process per GPU (torchrun).  `torch.distributed` is used ONLY as the launcher's rendezvous
(sharing the ncclUniqueId, barriers, max-over-ranks of timings); the data-path collective is the
ncclAllGather issued by libbpgpu on its own stream (bp_msm_sharded / bp_allgather_bytes).
"""
import ctypes
import os

from . import _native as nat


def slice_bounds(n, rank, world):
    """Contiguous, balanced partition of range(n): the first n % world ranks get one extra unit."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def all_slices(n, world):
    return [slice_bounds(n, r, world) for r in range(world)]


def env_rank():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def share_bytes(payload, src=0):
    """Broadcast a small bytes object from `src` through the launcher's process group."""
    import torch.distributed as dist
    box = [payload if dist.get_rank() == src else None]
    dist.broadcast_object_list(box, src=src)
    return box[0]


def init_nccl():
    """Create libbpgpu's NCCL communicator across the ranks of an initialised torch.distributed job."""
    import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    lib = nat.load()
    uid = None
    if rank == 0:
        buf = ctypes.create_string_buffer(128)
        nat.check(lib.bp_nccl_unique_id(buf))
        uid = buf.raw
    uid = share_bytes(uid, 0)
    nat.check(lib.bp_nccl_init(rank, world, uid))
    return rank, world


def gather_accept(local_accept, counts):
    """All-gather per-rank accept bytes (padded to the largest block) and stitch them in rank order."""
    width = max(counts)
    send = local_accept + bytes(width - len(local_accept))
    recv = ctypes.create_string_buffer(width * len(counts))
    nat.check(nat.load().bp_allgather_bytes(send, width, recv))
    out = b""
    for r, c in enumerate(counts):
        out += recv.raw[r * width:r * width + c]
    return out
