#!/bin/bash
# round-2 pass 17: round cubic curves on the GPU + all curve / filter tests + the hair leg (regression check of the flat path)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "cubic or curve or filter or hair" > gpurun_out/r2_run17_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_run17_pytest.log
tail -15 gpurun_out/r2_run17_pytest.log
timeout 300 python - > gpurun_out/r2_run17_hair.json 2> gpurun_out/r2_run17_hair.err <<'PY'
import json, sys, types, torch
sys.path.insert(0, '.')
import bench, embree_b200
lib = embree_b200.load()
dev = lib.new_device("verbose=0")
args = types.SimpleNamespace(no_cpu=True)
out = bench.hair_leg(lib, dev, torch.device("cuda:0"), torch.cuda.current_stream().cuda_stream, args)
print(json.dumps({k: (v if not isinstance(v, dict) else {a: v[a] for a in ("Mrays_per_s", "occluded_Mrays_per_s")}) for k, v in out.items()}))
PY
echo "hair rc=$?"; tail -3 gpurun_out/r2_run17_hair.err; cat gpurun_out/r2_run17_hair.json
