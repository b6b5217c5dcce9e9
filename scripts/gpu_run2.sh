#!/bin/bash
# round-2 second GPU pass: gather modes + tightened parity tests, SPREAD tuning sweep, ncu profile of the SPREAD kernel
mkdir -p gpurun_out
RTCB200_TEST_SPREAD=1 timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_run2_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_run2_pytest.log
tail -15 gpurun_out/r2_run2_pytest.log
B=scripts/_build
L=embree_b200/csrc/libembree4_b200.so
timeout 1500 python scripts/ab.py \
  new=$L \
  spread=$L,tri_spread=1 \
  sp_tb1=$L,tri_spread=1,tri_batch_min=1,tri_wait_max=1 \
  sp_tb2=$L,tri_spread=1,tri_batch_min=2,tri_wait_max=1 \
  sp_tb3=$L,tri_spread=1,tri_batch_min=3,tri_wait_max=2 \
  sp_tb4=$L,tri_spread=1,tri_batch_min=4,tri_wait_max=2 \
  sp_tb8=$L,tri_spread=1,tri_batch_min=8,tri_wait_max=3 \
  sp_tb10=$L,tri_spread=1,tri_batch_min=10,tri_wait_max=4 \
  sp_rf2=$L,tri_spread=1,refill_min=2 \
  sp_rf6=$L,tri_spread=1,refill_min=6 \
  sp_rf8=$L,tri_spread=1,refill_min=8 \
  sp_rf12=$L,tri_spread=1,refill_min=12 \
  sp_mb9=$B/lib_mb9.so,tri_spread=1,blocks_per_sm=9 \
  sp_mb6=$B/lib_mb6.so,tri_spread=1,blocks_per_sm=6 \
  > gpurun_out/r2_run2_ab.txt 2>&1
cat gpurun_out/r2_run2_ab.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 3 -c 1 -o gpurun_out/r2_spread_full -f \
  python scripts/ab.py --worker spread=$L,tri_spread=1 1581 33554432 > gpurun_out/r2_spread_ncu.log 2>&1
ls -la gpurun_out | tail -5
