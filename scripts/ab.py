"""A/B of library builds and tuning keys on the headline workload (test/bench tool; device-resident value only).

    python scripts/ab.py [--phi 1581] [--rays 33554432] name=lib.so[,key=val,...] ...

Every variant runs in its own process (a library is loaded once per process): scene commit, the bench's diffuse-bounce
stream, stat counters on a 1 Mi-ray sample, 3 warm-up + 5 timed passes (CUDA events), and a checksum of all hit records
so that variants can be compared for bit-identical results."""
import ctypes as C
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def worker(spec, phi, n):
    import torch
    name, rest = spec.split("=", 1)
    parts = rest.split(",")
    os.environ["EMBREE_B200_LIB"] = os.path.join(ROOT, parts[0])
    import embree_b200
    from embree_b200 import scenes
    import bench
    lib = embree_b200.load()
    for kv in parts[1:]:
        k, val = kv.split("=")
        assert lib.rtcb200SetTuning(k.encode(), int(val)) == 0, kv
    dev = lib.new_device(None)
    v, t = scenes.triangle_sphere(phi)
    devt = torch.device("cuda", 0)
    a = lib.args()
    stream = torch.cuda.current_stream().cuda_stream
    sc, keep, _ = bench.commit(lib, dev, v, t)
    st = lib.scene_stats(sc)

    def trace(tensor, count):
        lib.rtcb200Intersect1MDevice(sc, C.c_void_p(tensor.data_ptr()), count, C.byref(a), C.c_void_p(stream))

    prim = scenes.primary_rays(bench.PRIMARY_W, bench.PRIMARY_H, eye=bench.EYE, look=bench.LOOK, device=devt)
    trace(prim, prim.shape[0])
    torch.cuda.synchronize()
    stride = max(1, (1 << 26) // n)
    A = torch.empty((n, 24), dtype=torch.float32, device=devt)
    CH = 1 << 22
    for c0 in range(0, n, CH):
        ids = torch.arange(c0, min(c0 + CH, n), device=devt, dtype=torch.int64) * stride
        A[c0:c0 + len(ids)] = bench.bounce_rays(prim, ids)
    B = A.clone()
    lib.rtcb200SetSceneStatCounters(sc, 1)
    lib.rtcb200ResetSceneStatCounters(sc)
    S = A[:: max(1, n >> 20)].contiguous()
    trace(S, S.shape[0])
    torch.cuda.synchronize()
    s2 = lib.scene_stats(sc)
    lib.rtcb200SetSceneStatCounters(sc, 0)
    times = []
    for it in range(8):
        B.copy_(A)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        trace(B, n)
        e1.record()
        torch.cuda.synchronize()
        if it >= 3:
            times.append(e0.elapsed_time(e1))
    lib.check(dev)
    bi = B.view(torch.int32)
    w = torch.arange(1, 12, device=devt, dtype=torch.int64)
    chk = int((bi[:, 8:19].to(torch.int64) * w).sum().item())
    ms = sum(times) / len(times)
    print(json.dumps({"name": name, "Mrays_per_s": round(n / ms * 1e-3, 1), "ms": round(ms, 3), "min_ms": round(min(times), 3),
                      "nodes_per_ray": round(s2.trav_nodes / s2.trav_rays, 3), "tris_per_ray": round(s2.trav_tris / s2.trav_rays, 3),
                      "num_nodes": int(st.num_nodes), "build_ms": round(st.build_ms, 2), "sah": round(st.sah_cost, 3),
                      "hits": int((bi[:, 18] != -1).sum().item()), "checksum": chk}), flush=True)


if __name__ == "__main__":
    args = sys.argv[1:]
    phi, n = 1581, 1 << 25
    if args and args[0] == "--worker":
        worker(args[1], int(args[2]), int(args[3]))
        sys.exit(0)
    while args and args[0].startswith("--"):
        if args[0] == "--phi":
            phi = int(args[1])
        elif args[0] == "--rays":
            n = int(args[1])
        args = args[2:]
    for spec in args:
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--worker", spec, str(phi), str(n)], stdout=subprocess.PIPE,
                           stderr=subprocess.PIPE, text=True)
        out = [l for l in r.stdout.splitlines() if l.startswith("{")]
        print(out[-1] if out else f"{spec}: FAILED rc={r.returncode}\n{r.stderr[-1500:]}", flush=True)
