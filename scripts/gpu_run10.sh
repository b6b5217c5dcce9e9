#!/bin/bash
# round-2 pass 10: full GPU test suite (incl. the triangle_geometry tutorial frame), packed-FMA (FFMA2) A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2_run10_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_run10_pytest.log
tail -8 gpurun_out/r2_run10_pytest.log
B=scripts/_build
L=embree_b200/csrc/libembree4_b200.so
timeout 900 python scripts/ab.py f2_0=$B/lib_f2_0.so f2_3=$L f2_1=$B/lib_f2_1.so f2_2=$B/lib_f2_2.so f2_7=$B/lib_f2_7.so f2_0b=$B/lib_f2_0.so > gpurun_out/r2_run10_ab.txt 2>&1
cat gpurun_out/r2_run10_ab.txt
