"""N>1 path on CPU: world_size-2 gloo processes exercise the ray-stream sharding and the compact hit gather that
bench.py runs over NCCL (embree_b200/sharding.py).  No GPU and no product kernels involved: the traced records are
synthesised so the test checks exactly the host-side plumbing (slice bounds, record compaction, gather order)."""
import os
import socket

import subprocess
import sys

from embree_b200 import sharding


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_shard_bounds_cover_stream():
    for total in (0, 1, 7, 64, 1000003):
        for world in (1, 2, 3, 8):
            segs = [sharding.shard_bounds(total, r, world) for r in range(world)]
            assert segs[0][0] == 0 and segs[-1][1] == total
            assert all(segs[i][1] == segs[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in segs]
            assert max(sizes) - min(sizes) <= 1


def test_gather_world2_gloo():
    """Two gloo ranks; equal slices (the bench's weak-scaling layout) so every rank contributes same-sized chunks."""
    port = _free_port()
    here = os.path.dirname(os.path.abspath(__file__))
    procs = []
    for rank in range(2):
        env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE="2", LOCAL_RANK=str(rank))
        procs.append(subprocess.Popen([sys.executable, os.path.join(here, "_gloo_worker.py"), "4096"], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=180)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert "GATHER_OK" in outs[0]
