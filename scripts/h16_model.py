"""Offline experiment (CPU emulation): cost in extra node / triangle visits, and safety, of evaluating the BVH8 slab test
in packed HALF precision relative to the ray's entry time into the node (tests/emu/emu.cpp node_hitmask_h16).

    python scripts/h16_model.py [numPhi] [image_w] [replicate]
"""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))
from embree_b200 import scenes  # noqa: E402
import bench  # noqa: E402
from warp_model import load_emu  # noqa: E402
from tests.parity import compare_hits  # noqa: E402


def main():
    phi = int(sys.argv[1]) if len(sys.argv) > 1 else 400
    w = int(sys.argv[2]) if len(sys.argv) > 2 else 160
    rep = int(sys.argv[3]) if len(sys.argv) > 3 else 8
    e = load_emu()
    e.emu_trace_h16.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_float, C.c_int, C.c_void_p]
    e.emu_trace_h2.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
    for name, (v, t) in (("sphere", scenes.triangle_sphere(phi)), ("terrain", scenes.terrain(int(phi * 1.0)))):
        h = e.emu_build(v.ctypes.data, len(v), t.ctypes.data, len(t), 0, 0xFFFFFFFF, 3)
        if name == "sphere":
            prim = scenes.as_numpy_rayhits(scenes.primary_rays(w, w * 9 // 16, eye=bench.EYE, look=bench.LOOK))
        else:
            prim = scenes.as_numpy_rayhits(scenes.primary_rays(w, w * 9 // 16, eye=(0.0, 0.6, -1.2), look=(0.0, -0.5, 1.0)))
        e.emu_trace(h, prim.ctypes.data, len(prim), 0, None)
        pt = torch.from_numpy(prim.view(np.float32).reshape(-1, 24).copy())
        n = len(prim) * rep
        rays = scenes.as_numpy_rayhits(bench.bounce_rays(pt, torch.arange(n, dtype=torch.int64), replicate=rep))
        st = np.zeros(2, np.uint64)
        base = rays.copy()
        e.emu_trace(h, base.ctypes.data, len(base), 0, st.ctypes.data)
        print(f"{name}: {len(t)} triangles, {len(rays)} diffuse-bounce rays, fp32 test: nodes/ray {st[0] / len(rays):.2f} tris/ray {st[1] / len(rays):.2f}, "
              f"hit rate {(base['geomID'] != 0xFFFFFFFF).mean():.3f}")
        for magic in (0,):   # magic=1 (conversion folded into the FMA) needs a 4x smaller scale and ~1 cell of pad: not pursued
            for pad in (0.0, 0.25, 0.5, 1.0, 1.5):
                s2 = np.zeros(2, np.uint64)
                got = rays.copy()
                e.emu_trace_h16(h, got.ctypes.data, len(got), pad, magic, s2.ctypes.data)
                rep_ = compare_hits(base, got, 1e-6)
                lost = int(((base["geomID"] != 0xFFFFFFFF) & (got["geomID"] == 0xFFFFFFFF)).sum())
                farther = int(((base["geomID"] != 0xFFFFFFFF) & (got["geomID"] != 0xFFFFFFFF) & (got["tfar"] > base["tfar"])).sum())
                print(f"  half, magic={magic} pad={pad:4.2f} cells: nodes/ray {s2[0] / len(rays):6.2f} ({s2[0] / st[0] * 100 - 100:+5.1f} %)  tris/ray {s2[1] / len(rays):5.2f} "
                      f"({s2[1] / st[1] * 100 - 100:+5.1f} %)  lost hits {lost}  farther hits {farther}  id mismatches {rep_['id_mismatch']} ties {rep_['tie']}")
        s2 = np.zeros(2, np.uint64)
        got = rays.copy()
        e.emu_trace_h2(h, got.ctypes.data, len(got), s2.ctypes.data)
        same = all((base[f].view(np.uint32) == got[f].view(np.uint32)).all() for f in ("tfar", "u", "v", "Ng_x", "Ng_y", "Ng_z", "primID", "geomID"))
        print(f"  rt_core_h2.cuh node_hitmask_h2 (the device source, _Float16 emulation): nodes/ray {s2[0] / len(rays):6.2f} ({s2[0] / st[0] * 100 - 100:+5.1f} %)  "
              f"tris/ray {s2[1] / len(rays):5.2f} ({s2[1] / st[1] * 100 - 100:+5.1f} %)  results bit-identical to the fp32 traversal: {same}")
        e.emu_free(h)


if __name__ == "__main__":
    main()
