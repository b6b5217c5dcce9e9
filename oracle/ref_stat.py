#!/usr/bin/env python3
"""TEST INFRASTRUCTURE ONLY: traversal counters of the UNMODIFIED reference on a given ray set.

    python oracle/ref_stat.py <numPhi> <rays.npy> [threads]

Loads oracle/_ref/libembree4_stat.so.4 (the reference built with EMBREE_STAT_COUNTERS=ON by `oracle/build_ref.py --stat`),
commits createTriangleSphere(numPhi), traces the RTCRayHit[] records of <rays.npy> with rtcIntersect1 and exits; the
reference prints its counters from a static destructor at process exit (kernels/common/stat.cpp:13-17,
`normal.trav_nodes` / `trav_leaves` / `trav_prims`: stat.h:82-86, bvh_intersector1.cpp:87,101,
triangle_intersector.h:23 -- one "prim" is one Triangle4 block).  bench.py runs this as a subprocess and parses the text:
its `reference_nodes_per_ray` is the node-visit count the GPU traversal is compared with on the SAME rays."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
STAT_SO = os.path.join(ROOT, "oracle", "_ref", "libembree4_stat.so.4")


def parse(text):
    """'#nodes = 12.3M' lines of the ABSOLUTE block -> dict of floats (in units of rays / visits, not millions)."""
    out = {}
    block = text.split("--------- ABSOLUTE ---------")[-1]
    for key, name in (("#normal_travs", "rays"), ("#nodes ", "nodes"), ("#leaves", "leaves"), ("#prims ", "blocks")):
        for line in block.splitlines():
            if line.strip().startswith(key):
                out[name] = float(line.split("=")[1].strip().rstrip("M")) * 1e6
                break
    return out


if __name__ == "__main__":
    from embree_b200 import scenes
    from embree_b200.rtc import RAYHIT_DTYPE, RTCLib
    from tests.parity import api_trace_mt
    phi, path = int(sys.argv[1]), sys.argv[2]
    threads = int(sys.argv[3]) if len(sys.argv) > 3 else (len(os.sched_getaffinity(0)) or 1)
    R = RTCLib(STAT_SO)
    dev = R.new_device(None)
    v, t = scenes.triangle_sphere(phi)
    sc = R.rtcNewScene(dev)
    R.rtcSetSceneBuildQuality(sc, 1)
    _, keep = R.add_triangle_mesh(dev, sc, v, t, mask=0xFFFFFFFF)
    R.rtcCommitScene(sc)
    rays = np.load(path).view(RAYHIT_DTYPE).reshape(-1)
    api_trace_mt(R, sc, rays, threads)
    sys.stdout.flush()
