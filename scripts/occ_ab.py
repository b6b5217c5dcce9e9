"""A/B of the warp-wide triangle redistribution in the ANY-HIT kernels (rtcb200SetTuning "tri_spread_occluded") on the headline stream
(every 2nd ray of bench.py's diffuse-bounce rays on the 10 M-triangle sphere, RTCRay records, device-resident, CUDA events)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import embree_b200
from embree_b200 import scenes
import bench

lib = embree_b200.load()
dev = lib.new_device(None)
v, t = scenes.triangle_sphere(1581)
sc, keep, _ = bench.commit(lib, dev, v, t)
devt = torch.device("cuda", 0)
a = lib.args()
stream = torch.cuda.current_stream().cuda_stream
prim = scenes.primary_rays(bench.PRIMARY_W, bench.PRIMARY_H, eye=bench.EYE, look=bench.LOOK, device=devt)
lib.rtcb200Intersect1MDevice(sc, C.c_void_p(prim.data_ptr()), prim.shape[0], C.byref(a), C.c_void_p(stream)); torch.cuda.synchronize()
n = 1 << 25
R = torch.empty((n, 12), dtype=torch.float32, device=devt)
CH = 1 << 22
for c0 in range(0, n, CH):
    ids = torch.arange(c0, min(c0 + CH, n), device=devt, dtype=torch.int64) * 2
    R[c0:c0 + len(ids)] = bench.bounce_rays(prim, ids)[:, :12]
R[::3, 8] = 0.8          # a third of the rays end before the surface: not occluded
out = {}
for mode in (1, 0, 1, 0):
    lib.rtcb200SetTuning(b"tri_spread_occluded", mode)
    W = R.clone()
    best = 1e9
    for it in range(6):
        W.copy_(R)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        lib.rtcb200Occluded1MDevice(sc, C.c_void_p(W.data_ptr()), n, C.byref(a), C.c_void_p(stream))
        e1.record()
        torch.cuda.synchronize()
        if it >= 2:
            best = min(best, e0.elapsed_time(e1))
    out[mode] = W[:, 8].clone()
    print(f"tri_spread_occluded {mode}: {best:7.3f} ms  {n / best * 1e-3:7.1f} Mrays/s  occluded {float((W[:, 8] == float('-inf')).float().mean()):.4f}", flush=True)
print("identical:", bool(torch.equal(out[0].view(torch.int32), out[1].view(torch.int32))))
