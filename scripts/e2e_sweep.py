"""Sweep of the host-buffer pipeline (chunk size x streams in flight) of rtcb200Intersect1M on the headline scene."""
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
for streams in (3, 4):
    for lg in (22, 21, 20, 19):
        lib.rtcb200SetTuning(b"host_chunk_log2", lg); lib.rtcb200SetTuning(b"host_streams", streams)
        best = 1e9
        for rep in range(3):
            host.copy_(src)
            t0 = time.perf_counter()
            lib.rtcb200Intersect1M(sc, C.c_void_p(host.data_ptr()), n, C.byref(a))
            best = min(best, time.perf_counter() - t0)
        print(f"streams {streams} chunk 2^{lg}: {best*1e3:8.2f} ms  {n/best*1e-6:7.1f} Mrays/s", flush=True)
