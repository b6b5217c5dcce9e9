#!/bin/bash
# round-2 pass 18: full GPU suite + default bench line (with both hair legs) + path-tracer line
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2_run18_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_run18_pytest.log
tail -4 gpurun_out/r2_run18_pytest.log
timeout 700 python bench.py > gpurun_out/r2_run18_bench.json 2> gpurun_out/r2_run18_bench.err
echo "bench rc=$?"; tail -2 gpurun_out/r2_run18_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_run18_bench.json'))
print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['cpu_baseline']['value'])
for k in ('hair_bezier','hair_bezier_round'):
    h=d['extras'][k]
    for r in ('camera_1080p','incoherent'):
        print(k, r, h[r]['Mrays_per_s'], h[r]['occluded_Mrays_per_s'], h[r].get('reference'), h[r].get('parity'))
PY
