#!/bin/bash
# round-2 pass 27: point primitives (sphere / disc / oriented disc): their GPU tests, the reference's point_geometry tutorial, then the full suite
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_points.py tests/test_link_compat.py -m gpu -q -k "point" > gpurun_out/r2_run27_points.log 2>&1
echo "points rc=$?" >> gpurun_out/r2_run27_points.log; tail -25 gpurun_out/r2_run27_points.log | cut -c1-400
timeout 600 python -m pytest tests -m gpu -q -x --deselect tests/test_points.py > gpurun_out/r2_run27_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_run27_pytest.log; tail -4 gpurun_out/r2_run27_pytest.log
