#!/bin/bash
# round-2 fifth GPU pass: curves tests, TMA top-of-tree staging A/B, full bench lines (diffuse + pathtracer)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2_run5_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_run5_pytest.log
tail -25 gpurun_out/r2_run5_pytest.log
B=scripts/_build
L=embree_b200/csrc/libembree4_b200.so
timeout 600 python scripts/ab.py new=$L top9=$B/lib_top9.so top73=$B/lib_top73.so new2=$L > gpurun_out/r2_run5_ab.txt 2>&1
cat gpurun_out/r2_run5_ab.txt
timeout 1500 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_run5_bench.json 2> gpurun_out/r2_run5_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/r2_run5_bench.err; head -c 1200 gpurun_out/r2_run5_bench.json
timeout 900 python bench.py --workload pathtracer --steps 3 --warmup 2 > gpurun_out/r2_run5_pt.json 2> gpurun_out/r2_run5_pt.err
echo "pt rc=$?"; tail -2 gpurun_out/r2_run5_pt.err
