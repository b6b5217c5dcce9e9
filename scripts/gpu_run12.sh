#!/bin/bash
# round-2 pass 12: ray direction / mask / tfar parked in shared memory; 56-register (9 CTAs/SM) variants on top of it
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -k "golden or robust or instances or quads or curve or tutorial" > gpurun_out/r2_run12_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_run12_pytest.log
tail -5 gpurun_out/r2_run12_pytest.log
B=scripts/_build
L=embree_b200/csrc/libembree4_b200.so
timeout 900 python scripts/ab.py new=$L ss4=$B/lib_ss4.so mb9=$B/lib_mb9.so,blocks_per_sm=9 mb9ss4=$B/lib_mb9ss4.so,blocks_per_sm=9 mb10ss4=$B/lib_mb10ss4.so,blocks_per_sm=10 new_b=$L > gpurun_out/r2_run12_ab.txt 2>&1
cat gpurun_out/r2_run12_ab.txt
