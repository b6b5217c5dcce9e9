#!/bin/bash
# round-2 pass 19: round cubic curves with one BVH primitive per first-level sub-segment: curve tests + both hair legs with parity
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "cubic or curve" > gpurun_out/r2_run19_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_run19_pytest.log
tail -25 gpurun_out/r2_run19_pytest.log
timeout 400 python - > gpurun_out/r2_run19_hair.json 2> gpurun_out/r2_run19_hair.err <<'PY'
import json, sys, types, torch
sys.path.insert(0, '.')
import bench, embree_b200
lib = embree_b200.load()
dev = lib.new_device("verbose=0")
args = types.SimpleNamespace(no_cpu=False)
out = {"hair_bezier": bench.hair_leg(lib, dev, torch.device("cuda:0"), torch.cuda.current_stream().cuda_stream, args),
       "hair_bezier_round": bench.hair_leg(lib, dev, torch.device("cuda:0"), torch.cuda.current_stream().cuda_stream, args, rnd=True)}
print(json.dumps(out, indent=1))
PY
echo "hair rc=$?"; tail -3 gpurun_out/r2_run19_hair.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_run19_hair.json'))
for k,h in d.items():
    print(k, h['commit_ms'], h['nodes'])
    for r in ('camera_1080p','incoherent'):
        print(' ', r, h[r]['Mrays_per_s'], h[r]['occluded_Mrays_per_s'], h[r].get('reference'), h[r].get('parity'))
PY
