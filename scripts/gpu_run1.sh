#!/bin/bash
# round-2 first GPU pass: parity of the new kernel + A/B of the build variants (scripts/ab.py)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/r2_run1_gpu.txt 2>&1
RTCB200_TEST_SPREAD=1 timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_run1_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_run1_pytest.log
tail -5 gpurun_out/r2_run1_pytest.log
B=scripts/_build
timeout 1500 python scripts/ab.py \
  r1=$B/libembree4_b200_r1.so \
  new=embree_b200/csrc/libembree4_b200.so \
  tri2_0=$B/lib_tri2_0.so \
  spread=embree_b200/csrc/libembree4_b200.so,tri_spread=1 \
  mb10=$B/lib_mb10.so,blocks_per_sm=10 \
  mb10_tri2_0=$B/lib_mb10_tri2_0.so,blocks_per_sm=10 \
  mb6=$B/lib_mb6.so,blocks_per_sm=6 \
  mb6_tri2_0=$B/lib_mb6_tri2_0.so,blocks_per_sm=6 \
  new_refill2=embree_b200/csrc/libembree4_b200.so,refill_min=2 \
  new_refill8=embree_b200/csrc/libembree4_b200.so,refill_min=8 \
  new_tb4=embree_b200/csrc/libembree4_b200.so,tri_batch_min=4,tri_wait_max=2 \
  new_tb10=embree_b200/csrc/libembree4_b200.so,tri_batch_min=10,tri_wait_max=4 \
  new_notma=embree_b200/csrc/libembree4_b200.so,use_tma=0 \
  > gpurun_out/r2_run1_ab.txt 2>&1
cat gpurun_out/r2_run1_ab.txt
