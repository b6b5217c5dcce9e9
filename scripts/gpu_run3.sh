#!/bin/bash
# round-2 third GPU pass: shared-memory stack / overlapped refill A/B, refit + pathstream tests, full bench lines
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_run3_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_run3_pytest.log
tail -15 gpurun_out/r2_run3_pytest.log
B=scripts/_build
L=embree_b200/csrc/libembree4_b200.so
timeout 1200 python scripts/ab.py \
  new=$L \
  nospread=$L,tri_spread=0 \
  ss0=$B/lib_ss0.so \
  ss4=$B/lib_ss4.so \
  ss6=$B/lib_ss6.so \
  ss12=$B/lib_ss12.so \
  mb9=$B/lib_mb9.so,blocks_per_sm=9 \
  new_tb8=$L,tri_batch_min=8,tri_wait_max=3 \
  new_tb8w4=$L,tri_batch_min=8,tri_wait_max=4 \
  new_rf6=$L,refill_min=6 \
  new_rf3=$L,refill_min=3 \
  > gpurun_out/r2_run3_ab.txt 2>&1
cat gpurun_out/r2_run3_ab.txt
timeout 1200 python bench.py --steps 3 --warmup 3 > gpurun_out/r2_run3_bench.json 2> gpurun_out/r2_run3_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/r2_run3_bench.err; head -c 3000 gpurun_out/r2_run3_bench.json
timeout 900 python bench.py --workload pathtracer --steps 2 --warmup 1 > gpurun_out/r2_run3_pt.json 2> gpurun_out/r2_run3_pt.err
echo "pt rc=$?"; tail -3 gpurun_out/r2_run3_pt.err; head -c 3000 gpurun_out/r2_run3_pt.json
