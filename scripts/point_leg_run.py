"""Runs bench.py's point-primitive leg on its own (extras.points of the bench line) and prints its JSON."""
import json
import os
import sys
import types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import embree_b200
import bench

lib = embree_b200.load()
dev = lib.new_device(None)
devt = torch.device("cuda", 0)
args = types.SimpleNamespace(no_cpu="--no-cpu" in sys.argv)
out = bench.point_leg(lib, dev, devt, torch.cuda.current_stream().cuda_stream, args)
print(json.dumps(out))
