#!/bin/bash
# round-2 pass 24: adaptive two-level regime: tests + bench leg + tutorial
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "two_level or dynamic_scene or refit or recommit" > gpurun_out/r2_run24_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_run24_pytest.log
tail -25 gpurun_out/r2_run24_pytest.log
timeout 600 python - > gpurun_out/r2_run24_two_level.json 2> gpurun_out/r2_run24_two_level.err <<'PY'
import json, sys, types
sys.path.insert(0, '.')
import bench, embree_b200
lib = embree_b200.load()
dev = lib.new_device("verbose=0")
print(json.dumps(bench.two_level_leg(lib, dev, types.SimpleNamespace(no_cpu=False)), indent=1))
PY
echo "leg rc=$?"; tail -3 gpurun_out/r2_run24_two_level.err; cat gpurun_out/r2_run24_two_level.json
