#!/bin/bash
# round-2 pass 11: per-lane hit state in shared memory (zero spills) A/B, register-cap variants on top of it; tutorial frame test
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -k "tutorial or golden_all or gather" > gpurun_out/r2_run11_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_run11_pytest.log
tail -5 gpurun_out/r2_run11_pytest.log
B=scripts/_build
L=embree_b200/csrc/libembree4_b200.so
timeout 900 python scripts/ab.py ls0=$B/lib_ls0.so new=$L mb9=$B/lib_mb9.so,blocks_per_sm=9 mb10=$B/lib_mb10.so,blocks_per_sm=10 ss4=$B/lib_ss4.so new_b=$L > gpurun_out/r2_run11_ab.txt 2>&1
cat gpurun_out/r2_run11_ab.txt
