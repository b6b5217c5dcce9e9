"""Single-GPU cost of the fused hit gather's delivery modes (records go to a LOCAL buffer): plain trace vs
rtcb200Intersect1MGatherDevice with gather_mode 0 (one store per record) and 1 (32-ray blocks re-sent as 1 KB)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import embree_b200
from embree_b200 import scenes
import bench

lib = embree_b200.load()
dev = lib.new_device(None)
n = 1 << 25
v, t = scenes.triangle_sphere(1581)
sc, keep, _ = bench.commit(lib, dev, v, t)
devt = torch.device("cuda", 0)
a = lib.args()
st = torch.cuda.current_stream().cuda_stream
prim = scenes.primary_rays(bench.PRIMARY_W, bench.PRIMARY_H, eye=bench.EYE, look=bench.LOOK, device=devt)
lib.rtcb200Intersect1MDevice(sc, C.c_void_p(prim.data_ptr()), prim.shape[0], C.byref(a), C.c_void_p(st))
torch.cuda.synchronize()
A = torch.empty((n, 24), dtype=torch.float32, device=devt)
for c0 in range(0, n, 1 << 22):
    ids = torch.arange(c0, min(c0 + (1 << 22), n), device=devt, dtype=torch.int64) * 2
    A[c0:c0 + len(ids)] = bench.bounce_rays(prim, ids)
B = A.clone()
out = torch.empty((n, 8), dtype=torch.float32, device=devt)
ref = None
for name, mode in (("plain", None), ("gather_mode0", 0), ("gather_mode1", 1), ("gather_mode0", 0), ("gather_mode1", 1)):
    if mode is not None:
        lib.rtcb200SetTuning(b"gather_mode", mode)
    ts = []
    for it in range(7):
        B.copy_(A)
        out.fill_(7.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if mode is None:
            lib.rtcb200Intersect1MDevice(sc, C.c_void_p(B.data_ptr()), n, C.byref(a), C.c_void_p(st))
        else:
            lib.rtcb200Intersect1MGatherDevice(sc, C.c_void_p(B.data_ptr()), n, C.byref(a), C.c_void_p(st), C.c_void_p(out.data_ptr()))
        e1.record()
        torch.cuda.synchronize()
        if it >= 2:
            ts.append(e0.elapsed_time(e1))
    lib.check(dev)
    if mode is not None:
        if ref is None:
            ref = out.clone()
        assert torch.equal(ref.view(torch.int32), out.view(torch.int32)), name
    print(f"{name:14s} {sum(ts)/len(ts):7.3f} ms  {n/(sum(ts)/len(ts))*1e-3:8.1f} Mrays/s", flush=True)
