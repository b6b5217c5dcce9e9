"""Multi-GPU plumbing of the ray stream: rays are independent, so every rank holds a full BVH replica and
traces its own contiguous slice; the only exchange is the final gather of compact hit records to rank 0.
Works with any torch.distributed backend (NCCL over NVLink on the GPU box, gloo in the CPU tests)."""
import torch
import torch.distributed as dist

# columns of the [n, 24] float32 RTCRayHit view that make up a compact 32-byte hit record:
# tfar, Ng.x, Ng.y, Ng.z, u, v, primID, geomID
HIT_COLS = (8, 12, 13, 14, 15, 16, 17, 18)


def shard_bounds(total, rank, world):
    """Contiguous slice [begin, end) of a `total`-ray stream owned by `rank` (keeps whatever coherence the stream had)."""
    return total * rank // world, total * (rank + 1) // world


def compact_hits(rayhits, out=None):
    """[n,24] RTCRayHit view -> contiguous [n,8] compact hit records (one gather kernel)."""
    idx = torch.tensor(HIT_COLS, device=rayhits.device)
    if out is None:
        return rayhits.index_select(1, idx)
    torch.index_select(rayhits, 1, idx, out=out)
    return out


def gather_hits(local, dst=0, group=None, async_op=False):
    """Gather every rank's compact hit records on `dst`.  Returns (list of per-rank tensors on dst | None, work)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if world == 1:
        return [local], None
    bufs = [torch.empty_like(local) for _ in range(world)] if rank == dst else None
    work = dist.gather(local, bufs, dst=dst, group=group, async_op=async_op)
    return bufs, work
