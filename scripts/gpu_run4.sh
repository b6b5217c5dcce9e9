#!/bin/bash
# round-2 fourth GPU pass: reverted refill overlap + composition-independent SPREAD, full test suite, full bench lines
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2_run4_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_run4_pytest.log
tail -15 gpurun_out/r2_run4_pytest.log
B=scripts/_build
L=embree_b200/csrc/libembree4_b200.so
timeout 1200 python scripts/ab.py \
  new=$L \
  ss0=$B/lib_ss0.so \
  ss4=$B/lib_ss4.so \
  ss12=$B/lib_ss12.so \
  tb6w3=$L,tri_batch_min=6,tri_wait_max=3 \
  tb10w4=$L,tri_batch_min=10,tri_wait_max=4 \
  tb8w6=$L,tri_batch_min=8,tri_wait_max=6 \
  rf6=$L,refill_min=6 \
  > gpurun_out/r2_run4_ab.txt 2>&1
cat gpurun_out/r2_run4_ab.txt
timeout 1500 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_run4_bench.json 2> gpurun_out/r2_run4_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/r2_run4_bench.err; head -c 1500 gpurun_out/r2_run4_bench.json
