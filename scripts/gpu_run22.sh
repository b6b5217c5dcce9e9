#!/bin/bash
# round-2 pass 22 (final tree of this round): full GPU test suite, the default bench line, the path-tracer
# line and the ncu passes of scripts/profile.sh on the current tree
mkdir -p gpurun_out
date > gpurun_out/r2_run22_times.txt
timeout 700 python -m pytest tests -m gpu -q -x > gpurun_out/r2_run22_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_run22_pytest.log
tail -4 gpurun_out/r2_run22_pytest.log
date >> gpurun_out/r2_run22_times.txt
timeout 500 python bench.py > gpurun_out/r2_run22_bench.json 2> gpurun_out/r2_run22_bench.err
echo "bench rc=$?"; cut -c1-600 gpurun_out/r2_run22_bench.json
date >> gpurun_out/r2_run22_times.txt
timeout 600 bash scripts/profile.sh > gpurun_out/r2_run22_profile.log 2>&1
echo "profile rc=$?"
date >> gpurun_out/r2_run22_times.txt
timeout 400 python bench.py --workload pathtracer > gpurun_out/r2_run22_pathtracer.json 2> gpurun_out/r2_run22_pathtracer.err
echo "pathtracer rc=$?"; cut -c1-400 gpurun_out/r2_run22_pathtracer.json
date >> gpurun_out/r2_run22_times.txt
cat gpurun_out/r2_run22_times.txt
