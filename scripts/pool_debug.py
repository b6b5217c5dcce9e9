import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import embree_b200
from embree_b200 import scenes
torch.cuda.set_device(0)
lib = embree_b200.load(); dev = lib.new_device(os.environ.get("DEVCFG", "gpu=0"))
v, t = scenes.triangle_sphere(1581)
def commit(tag):
    sc = lib.rtcNewScene(dev)
    _, k = lib.add_triangle_mesh(dev, sc, v, t, mask=0xFFFFFFFF)
    t0 = time.perf_counter(); lib.rtcCommitScene(sc); wall = time.perf_counter() - t0; lib.check(dev)
    st = lib.scene_stats(sc); print(f"{tag}: device build {st.build_ms:8.2f} ms wall {wall*1e3:8.2f} ms", flush=True)
    return sc
s1 = commit("first (cold)")
s2 = commit("second, first still alive")
lib.rtcReleaseScene(s1); lib.rtcReleaseScene(s2)
s3 = commit("third after release")
x = torch.empty((1 << 26, 24), device="cuda"); x.fill_(1.0); torch.cuda.synchronize()
s4 = commit("fourth after a 6.4 GB torch alloc + sync, third alive")
lib.rtcReleaseScene(s3)
torch.cuda.empty_cache(); torch.cuda.synchronize()
s5 = commit("fifth after empty_cache")
s6 = commit("sixth")
