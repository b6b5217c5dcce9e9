"""Sweep of the leaf-test batching keys on the fur ball of bench.py's hair legs (flat and round Bezier curves): Mrays/s per setting."""
import json, sys, types, torch
sys.path.insert(0, '.')
import bench, embree_b200
lib = embree_b200.load()
dev = lib.new_device("verbose=0")
args = types.SimpleNamespace(no_cpu=True)
res = {}
for tb, tw in ((8, 4), (12, 6), (16, 8), (20, 12), (24, 16), (28, 24), (16, 4), (24, 8)):
    lib.rtcb200SetTuning(b"tri_batch_min", tb)
    lib.rtcb200SetTuning(b"tri_wait_max", tw)
    row = {}
    for rnd in (False, True):
        o = bench.hair_leg(lib, dev, torch.device("cuda:0"), torch.cuda.current_stream().cuda_stream, args, rnd=rnd)
        row["round" if rnd else "flat"] = [round(o["camera_1080p"]["Mrays_per_s"], 1), round(o["incoherent"]["Mrays_per_s"], 1),
                                           round(o["camera_1080p"]["occluded_Mrays_per_s"], 1), round(o["incoherent"]["occluded_Mrays_per_s"], 1)]
    res[f"tb{tb}_tw{tw}"] = row
    print(f"tb{tb}_tw{tw}", row, flush=True)
