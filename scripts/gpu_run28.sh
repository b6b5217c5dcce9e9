#!/bin/bash
# round-2 pass 28 (last): full GPU suite on the final tree (point primitives, geometry identity in the per-mesh BVH cache), point leg of the bench
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q --maxfail=6 > gpurun_out/r2_run28_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_run28_pytest.log; tail -6 gpurun_out/r2_run28_pytest.log | cut -c1-300
timeout 150 python scripts/point_leg_run.py > gpurun_out/r2_run28_points.json 2> gpurun_out/r2_run28_points.err
echo "point leg rc=$?"; cut -c1-1500 gpurun_out/r2_run28_points.json; tail -3 gpurun_out/r2_run28_points.err
timeout 60 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
