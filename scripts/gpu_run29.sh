#!/bin/bash
# round-2 pass 29: the one test whose new assertion compared instID of an instanced scene with the un-instanced one (test artefact, fixed)
mkdir -p gpurun_out
timeout 100 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "test_curves_large_vs_oracle_and_errors" > gpurun_out/r2_run29_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_run29_pytest.log; grep -E "^E |passed|failed|rc=" gpurun_out/r2_run29_pytest.log | cut -c1-400 | head -12
