#!/bin/bash
# round-2 pass 26: any-hit SPREAD: full GPU suite (occluded parity everywhere), A/B on the headline stream, path-tracer line
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2_run26_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_run26_pytest.log; tail -4 gpurun_out/r2_run26_pytest.log
timeout 400 python scripts/occ_ab.py 2>&1 | tail -6
timeout 400 python bench.py --workload pathtracer --no-cpu > gpurun_out/r2_run26_pathtracer.json 2> gpurun_out/r2_run26_pathtracer.err
echo "pathtracer rc=$?"; cut -c1-330 gpurun_out/r2_run26_pathtracer.json
