"""Parameter sweep of the trace kernel / collapse policy on the headline workload (one scene build per policy)."""
import ctypes as C, os, sys, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import embree_b200
from embree_b200 import scenes
import bench

lib = embree_b200.load()
dev = lib.new_device(None)
phi = int(sys.argv[1]) if len(sys.argv) > 1 else 1581
n = int(sys.argv[2]) if len(sys.argv) > 2 else (1 << 24)
v, t = scenes.triangle_sphere(phi)
a = lib.args()
stream = torch.cuda.current_stream().cuda_stream
devt = torch.device("cuda", 0)

def trace(sc, tensor, count):
    lib.rtcb200Intersect1MDevice(sc, C.c_void_p(tensor.data_ptr()), count, C.byref(a), C.c_void_p(stream))

def timeit(sc, A, B, reps=3):
    best = 1e9
    for _ in range(reps):
        B.copy_(A)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); trace(sc, B, A.shape[0]); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best

A = None
policies = [x for x in os.environ.get("POLICIES", "0,3,3:20,3:50,3:80").split(",")]
for pol_s in policies:
    pol = int(pol_s.split(":")[0])
    lib.rtcb200SetTuning(b"sah_small", int(pol_s.split(":")[2]) if pol_s.count(":") > 1 else 8)
    lib.rtcb200SetTuning(b"collapse_policy", pol)
    lib.rtcb200SetTuning(b"c_tri", int(pol_s.split(":")[1]) if ":" in pol_s else 30)
    sc = lib.rtcNewScene(dev)
    _, keep = lib.add_triangle_mesh(dev, sc, v, t, mask=0xFFFFFFFF)
    lib.rtcCommitScene(sc); lib.check(dev)
    st = lib.scene_stats(sc)
    if A is None:
        prim = scenes.primary_rays(bench.PRIMARY_W, bench.PRIMARY_H, eye=bench.EYE, look=bench.LOOK, device=devt)
        trace(sc, prim, prim.shape[0]); torch.cuda.synchronize()
        stride = (1 << 26) // n
        A = torch.empty((n, 24), dtype=torch.float32, device=devt)
        CH = 1 << 22
        for c0 in range(0, n, CH):
            ids = (torch.arange(c0, min(c0 + CH, n), device=devt, dtype=torch.int64)) * stride
            A[c0:c0 + len(ids)] = bench.bounce_rays(prim, ids)
        B = A.clone()
    lib.rtcb200SetSceneStatCounters(sc, 1); lib.rtcb200ResetSceneStatCounters(sc)
    S = A[::16].contiguous(); trace(sc, S, S.shape[0]); torch.cuda.synchronize()
    s2 = lib.scene_stats(sc); lib.rtcb200SetSceneStatCounters(sc, 0)
    print(f"policy {pol_s}: nodes {st.num_nodes} sah {st.sah_cost:.2f} depth {st.max_depth} build {st.build_ms:.1f} ms | nodes/ray {s2.trav_nodes/s2.trav_rays:.2f} tris/ray {s2.trav_tris/s2.trav_rays:.2f}", flush=True)
    base = dict(tri_batch_min=6, tri_wait_max=3, blocks_per_sm=8, use_tma=1, refill_min=4)
    combos = [dict()]
    if pol_s == policies[-1] or os.environ.get("FULL"):
        combos = [dict(), dict(refill_min=1), dict(refill_min=2), dict(refill_min=6), dict(refill_min=8), dict(refill_min=12),
                  dict(tri_batch_min=1, tri_wait_max=1), dict(tri_batch_min=4, tri_wait_max=2), dict(tri_batch_min=8, tri_wait_max=4),
                  dict(use_tma=0)]
    for c in combos:
        cfg = dict(base, **c)
        for k, val in cfg.items():
            lib.rtcb200SetTuning(k.encode(), val)
        ms = timeit(sc, A, B)
        print(f"   {c}: {ms:7.3f} ms  {n/ms*1e-3:8.1f} Mrays/s", flush=True)
    for k, val in base.items():
        lib.rtcb200SetTuning(k.encode(), val)
    lib.rtcReleaseScene(sc)
