#!/bin/bash
# round-2 pass 16: flat cubic curves with one BVH primitive per tessellation segment -- curve tests + the hair_bezier leg alone
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -k "cubic or curve or filter or hair" > gpurun_out/r2_run16_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_run16_pytest.log
tail -6 gpurun_out/r2_run16_pytest.log
timeout 300 python - > gpurun_out/r2_run16_hair.json 2> gpurun_out/r2_run16_hair.err <<'PY'
import json, sys, types, torch
sys.path.insert(0, '.')
import bench, embree_b200
lib = embree_b200.load()
dev = lib.new_device("verbose=0")
args = types.SimpleNamespace(no_cpu=False)
devt = torch.device("cuda:0")
out = bench.hair_leg(lib, dev, devt, torch.cuda.current_stream().cuda_stream, args)
print(json.dumps(out, indent=1))
PY
echo "hair rc=$?"; tail -3 gpurun_out/r2_run16_hair.err; cat gpurun_out/r2_run16_hair.json
