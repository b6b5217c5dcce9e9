#!/bin/bash
mkdir -p gpurun_out
L=embree_b200/csrc/libembree4_b200.so
timeout 600 ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 3 -c 1 -o gpurun_out/r2_v2_full -f \
  python scripts/ab.py --worker new=$L 1581 33554432 > gpurun_out/r2_v2_ncu.log 2>&1
tail -3 gpurun_out/r2_v2_ncu.log
