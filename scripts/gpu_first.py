"""First GPU shake-out: build + trace + parity on a sphere, timings with CUDA events."""
import ctypes as C, sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import embree_b200
from embree_b200 import scenes
from embree_b200.rtc import *
from tests.parity import load_oracle, load_reference, compare_hits, api_trace_mt

lib = embree_b200.load()
dev = lib.new_device("verbose=2")
num_phi = int(sys.argv[1]) if len(sys.argv) > 1 else 501
nrays = int(sys.argv[2]) if len(sys.argv) > 2 else (1 << 24)
v, t = scenes.triangle_sphere(num_phi)
print("tris", len(t))
for q in (0,):
    sc = lib.rtcNewScene(dev)
    lib.rtcSetSceneBuildQuality(sc, q)
    gid, keep = lib.add_triangle_mesh(dev, sc, v, t, mask=0xFFFFFFFF)
    t0 = time.time(); lib.rtcCommitScene(sc); print("commit wall", time.time() - t0); lib.check(dev)
    st = lib.scene_stats(sc)
    print("stats: nodes", st.num_nodes, "tris", st.num_triangles, "build_ms", st.build_ms, "sah", st.sah_cost, "depth", st.max_depth)
    rays = scenes.incoherent_rays_reference(nrays, device="cuda")
    torch.cuda.synchronize()
    a = lib.args()
    stream = torch.cuda.current_stream().cuda_stream
    for it in range(3):
        r = rays.clone()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        lib.rtcb200Intersect1MDevice(sc, C.c_void_p(r.data_ptr()), nrays, C.byref(a), C.c_void_p(stream))
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print(f"trace {nrays} rays: {ms:.3f} ms = {nrays/ms*1e-3:.1f} Mrays/s")
    lib.check(dev)
    # stats pass
    lib.rtcb200SetSceneStatCounters(sc, 1); lib.rtcb200ResetSceneStatCounters(sc)
    r2 = rays[:1<<20].clone()
    lib.rtcb200Intersect1MDevice(sc, C.c_void_p(r2.data_ptr()), 1<<20, C.byref(a), C.c_void_p(stream)); torch.cuda.synchronize()
    st = lib.scene_stats(sc); print("nodes/ray", st.trav_nodes/st.trav_rays, "tris/ray", st.trav_tris/st.trav_rays)
    lib.rtcb200SetSceneStatCounters(sc, 0)
    # parity on a subset vs reference + oracle
    nsub = 200000
    sub = scenes.as_numpy_rayhits(rays[:nsub].cpu())
    got = scenes.as_numpy_rayhits(r[:nsub].cpu())
    want = load_oracle().trace(v, t, sub.copy(), nthreads=8)
    print("vs oracle:", compare_hits(want, got))
    R = load_reference()
    if R:
        rd = R.new_device(None); rs = R.rtcNewScene(rd); _, k2 = R.add_triangle_mesh(rd, rs, v, t, mask=0xFFFFFFFF); R.rtcCommitScene(rs)
        ncpu = os.cpu_count()
        big = scenes.as_numpy_rayhits(rays[:2000000].cpu())
        t0 = time.time(); api_trace_mt(R, rs, big, ncpu); dt = time.time() - t0
        print(f"reference rtcIntersect1 x{ncpu} threads: {len(big)/dt*1e-6:.1f} Mrays/s")
        print("vs reference:", compare_hits(big[:nsub], got))
    # host path
    h = scenes.as_numpy_rayhits(rays[:1<<20].cpu())
    t0 = time.time(); lib.intersect(sc, h, "1M"); print("host 1M path: %.1f Mrays/s" % ((1<<20)/(time.time()-t0)*1e-6))
    print("host vs device:", compare_hits(scenes.as_numpy_rayhits(r[:1<<20].cpu()), h))
    one = lib.intersect(sc, scenes.as_numpy_rayhits(rays[:64].cpu()), "1")
    print("single-ray path ids equal:", (one["primID"] == got["primID"][:64]).all())
    p16 = lib.intersect(sc, scenes.as_numpy_rayhits(rays[:100].cpu()), "16")
    print("packet16 path ids equal:", (p16["primID"] == got["primID"][:100]).all())
    lib.check(dev)
    lib.rtcReleaseScene(sc)
print("launches", lib.rtcb200GetLaunchCount())
