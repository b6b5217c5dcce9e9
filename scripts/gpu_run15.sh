#!/bin/bash
# round-2 pass 15: full GPU suite (filters, cubic curves, hair shadows included) + the default bench line with the hair_bezier extra
mkdir -p gpurun_out
timeout 800 python -m pytest tests -m gpu -q > gpurun_out/r2_run15_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_run15_pytest.log
tail -6 gpurun_out/r2_run15_pytest.log
timeout 600 python bench.py > gpurun_out/r2_run15_bench.json 2> gpurun_out/r2_run15_bench.err
echo "bench rc=$?"; tail -5 gpurun_out/r2_run15_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_run15_bench.json'))
print(d['value'], d['e2e']['value'], d['roofline']['frac'])
print(json.dumps(d['extras']['hair_bezier'], indent=1))
PY
