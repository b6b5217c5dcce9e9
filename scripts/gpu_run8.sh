#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/gather_ab.py > gpurun_out/r2_run8_gather_ab.txt 2>&1
cat gpurun_out/r2_run8_gather_ab.txt
timeout 300 python -m pytest tests -m gpu -q -k "gather" 2>&1 | tail -3
