import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import embree_b200
from embree_b200 import scenes
lib = embree_b200.load(); dev = lib.new_device(None)
v, t = scenes.triangle_sphere(1581)
rays = scenes.as_numpy_rayhits(scenes.incoherent_rays_reference(1 << 20, org=(0.3, 0.1, -0.2)))
outs = []
for it in range(3):
    sc = lib.rtcNewScene(dev); _, k = lib.add_triangle_mesh(dev, sc, v, t, mask=0xFFFFFFFF); lib.rtcCommitScene(sc); lib.check(dev)
    st = lib.scene_stats(sc); out = lib.intersect(sc, rays.copy(), "1M"); outs.append(out)
    print("build", it, "nodes", st.num_nodes, "sah", st.sah_cost, "build_ms", round(st.build_ms, 2), "identical to first:", bool((out.view(np.uint8) == outs[0].view(np.uint8)).all()), flush=True)
    lib.rtcReleaseScene(sc)
