#!/bin/bash
# round-2 pass 25: final tree: full GPU suite, smoke(), the default bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2_run25_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_run25_pytest.log; tail -4 gpurun_out/r2_run25_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 700 python bench.py > gpurun_out/r2_run25_bench.json 2> gpurun_out/r2_run25_bench.err
echo "bench rc=$?"; tail -2 gpurun_out/r2_run25_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_run25_bench.json'))
print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['cpu_baseline']['value'], d['gpu_launches'])
print(json.dumps(d['extras']['dynamic_scene_two_level'])[:700])
for k in ('hair_bezier','hair_bezier_round'):
    h=d['extras'][k]; print(k, [round(h[r]['Mrays_per_s']) for r in ('camera_1080p','incoherent')])
PY
