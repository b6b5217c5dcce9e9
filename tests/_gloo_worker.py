"""Worker of tests/test_sharding_gloo.py: one rank of a world_size-N gloo job (env: RANK, WORLD_SIZE, MASTER_*)."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from embree_b200 import sharding  # noqa: E402


def fake_traced(n, first):
    """RTCRayHit [n,24] view with recognisable hit fields derived from the global ray index."""
    r = torch.zeros((n, 24), dtype=torch.float32)
    g = torch.arange(first, first + n, dtype=torch.float32)
    r[:, 8] = g * 0.5
    r[:, 12:15] = g[:, None] + torch.tensor([0.1, 0.2, 0.3])
    r[:, 15], r[:, 16] = g * 0.25, g * 0.125
    ri = r.view(torch.int32)
    ri[:, 17] = torch.arange(first, first + n, dtype=torch.int32)
    ri[:, 18] = 0
    return r


def main():
    total = int(sys.argv[1])
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    b, e = sharding.shard_bounds(total, rank, world)
    compact = sharding.compact_hits(fake_traced(e - b, b))
    assert compact.shape == (e - b, 8) and compact.is_contiguous()
    # chunked gather exactly as bench.py does it (chunk c gathered while chunk c+1 would trace)
    n = e - b
    pieces = [[] for _ in range(world)]
    for c in range(4):
        cb, ce = sharding.shard_bounds(n, c, 4)
        bufs, _ = sharding.gather_hits(compact[cb:ce].contiguous(), dst=0)
        if rank == 0:
            for r in range(world):
                pieces[r].append(bufs[r])
    if rank == 0:
        got = torch.cat([torch.cat(p, 0) for p in pieces], 0)
        want = sharding.compact_hits(fake_traced(total, 0))
        assert torch.equal(got, want), "gathered hit buffer differs from the single-process result"
        print("GATHER_OK", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
