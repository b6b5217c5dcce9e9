import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import embree_b200
from embree_b200 import scenes
lib = embree_b200.load(); dev = lib.new_device(None)
for phi in [int(x) for x in sys.argv[1:]] or [501]:
    v, t = scenes.triangle_sphere(phi)
    for q in (0, 0, 0, 1, 1, 1):
        sc = lib.rtcNewScene(dev); lib.rtcSetSceneBuildQuality(sc, q)
        _, k = lib.add_triangle_mesh(dev, sc, v, t, mask=0xFFFFFFFF)
        t0 = time.perf_counter(); lib.rtcCommitScene(sc); wall = time.perf_counter() - t0; lib.check(dev)
        st = lib.scene_stats(sc); print(f"tris {len(t):9d} quality {q} device build {st.build_ms:8.3f} ms  commit wall {wall*1e3:8.3f} ms  nodes {st.num_nodes} depth {st.max_depth}", flush=True)
        lib.rtcReleaseScene(sc)
