"""A/B of the host pipeline's copy back (rtcb200SetTuning "host_d2h_partial"): full 96-byte records vs the 64 bytes a query can change.
Same scene / rays as bench.py's e2e leg; also checks that both deliver identical records and leave the untouched bytes alone."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import embree_b200
from embree_b200 import scenes
import bench

lib = embree_b200.load()
dev = lib.new_device(None)
n = int(sys.argv[1]) if len(sys.argv) > 1 else (1 << 26)
v, t = scenes.triangle_sphere(1581)
sc = lib.rtcNewScene(dev)
_, keep = lib.add_triangle_mesh(dev, sc, v, t, mask=0xFFFFFFFF)
lib.rtcCommitScene(sc); lib.check(dev)
devt = torch.device("cuda", 0)
a = lib.args()
stream = torch.cuda.current_stream().cuda_stream
prim = scenes.primary_rays(bench.PRIMARY_W, bench.PRIMARY_H, eye=bench.EYE, look=bench.LOOK, device=devt)
lib.rtcb200Intersect1MDevice(sc, C.c_void_p(prim.data_ptr()), prim.shape[0], C.byref(a), C.c_void_p(stream)); torch.cuda.synchronize()
stride = (1 << 26) // n
host = torch.empty((n, 24), dtype=torch.float32, pin_memory=True)
CH = 1 << 22
for c0 in range(0, n, CH):
    ids = (torch.arange(c0, min(c0 + CH, n), device=devt, dtype=torch.int64)) * stride
    host[c0:c0 + len(ids)] = bench.bounce_rays(prim, ids).cpu()
src = host.clone().pin_memory()
res = {}
for mode in (1, 0, 1, 0):
    lib.rtcb200SetTuning(b"host_d2h_partial", mode)
    best = 1e9
    for rep in range(3):
        host.copy_(src)
        t0 = time.perf_counter()
        lib.rtcb200Intersect1M(sc, C.c_void_p(host.data_ptr()), n, C.byref(a))
        best = min(best, time.perf_counter() - t0)
    res[mode] = host.clone()
    print(f"host_d2h_partial {mode}: {best*1e3:8.2f} ms  {n/best*1e-6:7.1f} Mrays/s", flush=True)
print("identical records:", bool(torch.equal(res[0].view(torch.int32), res[1].view(torch.int32))))
# packets: RTCRayHit16 through rtcb200IntersectNM, both modes
from embree_b200.rtc import to_packets, from_packets
small = scenes.as_numpy_rayhits(src[:1 << 20].clone())
out = {}
for mode in (1, 0):
    lib.rtcb200SetTuning(b"host_d2h_partial", mode)
    out[mode] = lib.intersect(sc, small.copy(), "16M")
print("packets identical:", bool((out[0].view(np.uint8) == out[1].view(np.uint8)).all()), "hits", int((out[1]["geomID"] != 0xFFFFFFFF).sum()))
