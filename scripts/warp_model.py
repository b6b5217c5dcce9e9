"""Offline model of the trace kernel's warp-synchronous scheduling (CPU only, no GPU needed).

The CPU emulation (tests/emu) logs, per ray, the operation sequence of the production traversal order: 'N' = one node
step, 'T' = one triangle test.  This script replays those sequences through a model of trace.cu's per-warp loop
(32 lanes, rays handed out in 32-ray blocks, batched refill, one node step + at most one batched triangle step per
iteration) and counts warp-level phase executions under alternative scheduling policies.  It answers "how many
warp-instructions per ray would policy X issue" -- the quantity the round-1 experiments showed the kernel is bound by
(next to memory latency) -- before any GPU time is spent.  Cost weights are SASS instruction counts of the phases.

    python scripts/warp_model.py [numPhi] [image_w] [replicate]
"""
import ctypes as C
import os
import subprocess
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from embree_b200 import scenes  # noqa: E402
import bench  # noqa: E402

C_NODE, C_TRI, C_LOOP, C_REFILL, C_POP = 300, 150, 70, 90, 40
C_SPREAD = int(os.environ.get("C_SPREAD", "100"))   # guess: 9 ray values broadcast + segmented min of (t, item) + write-back, by shuffles   # instructions per executed phase (cuobjdump, trace_kernel<1,0,0,0,0>)


def load_emu():
    subprocess.check_call([os.path.join(ROOT, "tests", "emu", "build.sh")])
    e = C.CDLL(os.path.join(ROOT, "tests", "emu", "_build", "libemu.so"))
    e.emu_build.restype = C.c_void_p
    e.emu_build.argtypes = [C.c_void_p, C.c_uint, C.c_void_p, C.c_uint, C.c_uint, C.c_uint, C.c_int]
    e.emu_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_void_p]
    e.emu_trace_ops.restype = C.c_uint64
    e.emu_trace_ops.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64]
    return e


def ray_ops(e, h, rays):
    cap = len(rays) * 256
    buf = np.zeros(cap, np.uint8)
    n = e.emu_trace_ops(h, rays.ctypes.data, len(rays), buf.ctypes.data, cap)
    assert n <= cap
    buf = buf[:n]
    ends = np.nonzero(buf == 0)[0]
    starts = np.concatenate([[0], ends[:-1] + 1])
    return buf, starts, ends


def simulate(buf, starts, ends, policy, n_warps=64):
    """Replays the op sequences.  policy: dict(tri_batch_min, tri_wait_max, refill_min, tris_per_step)."""
    nrays = len(starts)
    nblocks = nrays // 32
    tbm, twm, rmin, tps = policy["tri_batch_min"], policy["tri_wait_max"], policy["refill_min"], policy.get("tris_per_step", 1)
    isN = (buf == ord("N"))
    tot = dict(iter=0, node=0, tri=0, refill=0, node_lanes=0, tri_lanes=0, inst=0)
    for w in range(n_warps):
        blocks = list(range(w, nblocks, n_warps))
        if not blocks:
            continue
        bi, consumed = 0, 0                      # current block and how many of its rays were handed out
        pos = np.zeros(32, np.int64)             # cursor into buf per lane
        end = np.zeros(32, np.int64)
        active = np.zeros(32, bool)
        tri_wait = 0
        while True:
            # ---- refill
            idle = ~active
            n_idle = int(idle.sum())
            have = bi < len(blocks)
            if have and n_idle and (n_idle >= rmin or not active.any()):
                lanes = np.nonzero(idle)[0]
                k = 0
                while k < len(lanes) and bi < len(blocks):
                    take = min(len(lanes) - k, 32 - consumed)
                    ids = blocks[bi] * 32 + consumed + np.arange(take)
                    pos[lanes[k:k + take]] = starts[ids]
                    end[lanes[k:k + take]] = ends[ids]
                    active[lanes[k:k + take]] = True
                    consumed += take
                    k += take
                    if consumed == 32:
                        bi += 1
                        consumed = 0
                tot["refill"] += 1
                tot["inst"] += C_REFILL
            if not active.any():
                break
            tot["iter"] += 1
            tot["inst"] += C_LOOP
            done = active & (pos >= end)
            cur = np.where(active & ~done, isN[np.minimum(pos, len(buf) - 1)], False)
            wantN = active & ~done & cur
            wantT = active & ~done & ~cur
            # ---- node step
            if wantN.any():
                tot["node"] += 1
                tot["node_lanes"] += int(wantN.sum())
                tot["inst"] += C_NODE
                pos[wantN] += 1
            # ---- triangle step (batched)
            nT = int(wantT.sum())
            if nT and (nT >= tbm or not wantN.any() or tri_wait + 1 >= twm):
                tri_wait = 0
                tot["tri"] += 1
                tot["tri_lanes"] += nT
                if tps == "spread":
                    # warp-wide redistribution: every pending triangle of every T-lane's current run is one work item,
                    # up to 32 items are tested in one phase by whichever lanes are free (ray broadcast by shuffles,
                    # segmented min back to the owner): +C_SPREAD instructions
                    tot["inst"] += C_TRI + C_SPREAD
                    budget = 32
                    for l in np.nonzero(wantT)[0]:
                        run = 0
                        while pos[l] + run < end[l] and not isN[pos[l] + run]:
                            run += 1
                        take = min(run, budget)
                        pos[l] += take
                        budget -= take
                        tot["tri_items"] = tot.get("tri_items", 0) + take
                        if budget == 0:
                            break
                else:
                    tot["inst"] += C_TRI + (tps - 1) * int(C_TRI * 0.7)
                    for _ in range(tps):
                        still = wantT & (pos < end) & ~isN[np.minimum(pos, len(buf) - 1)]
                        pos[still] += 1
            elif nT:
                tri_wait += 1
            # ---- finish
            fin = active & (pos >= end)
            if fin.any():
                tot["inst"] += C_POP
                active[fin] = False
    return tot


def main():
    phi = int(sys.argv[1]) if len(sys.argv) > 1 else 400
    w = int(sys.argv[2]) if len(sys.argv) > 2 else 160
    rep = int(sys.argv[3]) if len(sys.argv) > 3 else 8
    e = load_emu()
    v, t = scenes.triangle_sphere(phi)
    t0 = time.time()
    h = e.emu_build(v.ctypes.data, len(v), t.ctypes.data, len(t), 0, 0xFFFFFFFF, 3)
    print(f"scene: {len(t)} triangles, emu build {time.time() - t0:.1f}s")
    prim = scenes.as_numpy_rayhits(scenes.primary_rays(w, w * 9 // 16, eye=bench.EYE, look=bench.LOOK))
    e.emu_trace(h, prim.ctypes.data, len(prim), 0, None)
    pt = torch.from_numpy(prim.view(np.float32).reshape(-1, 24).copy())
    n = len(prim) * rep
    ids = torch.arange(n, dtype=torch.int64)
    rays = scenes.as_numpy_rayhits(bench.bounce_rays(pt, ids, replicate=rep))
    buf, starts, ends = ray_ops(e, h, rays)
    lens = ends - starts
    nN = int((buf == ord("N")).sum()); nT = int((buf == ord("T")).sum())
    print(f"{len(rays)} diffuse-bounce rays: nodes/ray {nN / len(rays):.2f}, tris/ray {nT / len(rays):.2f}, ops/ray {lens.mean():.1f} (max {lens.max()})")
    base = dict(tri_batch_min=6, tri_wait_max=3, refill_min=4)
    policies = [("production (6,3,4)", base),
                ("tri_batch_min=1 (eager triangles)", dict(base, tri_batch_min=1, tri_wait_max=1)),
                ("tri_batch_min=12, wait 6", dict(base, tri_batch_min=12, tri_wait_max=6)),
                ("refill_min=1", dict(base, refill_min=1)),
                ("refill_min=12", dict(base, refill_min=12)),
                ("2 triangles per triangle step", dict(base, tris_per_step=2)),
                ("3 triangles per triangle step", dict(base, tris_per_step=3)),
                ("2 triangles per step, batch 8 wait 4", dict(base, tris_per_step=2, tri_batch_min=8, tri_wait_max=4)),
                ("warp-spread triangles (6,3)", dict(base, tris_per_step="spread")),
                ("warp-spread triangles, eager (1,1)", dict(base, tris_per_step="spread", tri_batch_min=1, tri_wait_max=1)),
                ("warp-spread triangles (4,2)", dict(base, tris_per_step="spread", tri_batch_min=4, tri_wait_max=2)),
                ("warp-spread triangles (10,4)", dict(base, tris_per_step="spread", tri_batch_min=10, tri_wait_max=4))]
    print(f"{'policy':42s} {'iter/ray':>8s} {'nodeph/ray':>10s} {'triph/ray':>9s} {'N lanes':>7s} {'T lanes':>7s} {'warp-inst/ray':>13s}")
    ref = None
    for name, pol in policies:
        r = simulate(buf, starts, ends, pol)
        nr = (len(starts) // 32) * 32
        ipr = r["inst"] / nr
        ref = ref or ipr
        print(f"{name:42s} {r['iter'] / nr:8.3f} {r['node'] / nr:10.3f} {r['tri'] / nr:9.3f} {r['node_lanes'] / max(r['node'], 1):7.1f} "
              f"{r['tri_lanes'] / max(r['tri'], 1):7.1f} {ipr:13.1f}  ({ipr / ref * 100:5.1f} %)")


if __name__ == "__main__":
    main()
