#!/bin/bash
# round-2 pass 23: the dynamic_scene tutorial test, then the whole GPU suite once more (repeatability)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_link_compat.py -m gpu -q > gpurun_out/r2_run23_link.log 2>&1
echo "link rc=$?" >> gpurun_out/r2_run23_link.log; tail -15 gpurun_out/r2_run23_link.log
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2_run23_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_run23_pytest.log; tail -4 gpurun_out/r2_run23_pytest.log
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_run23_refarm.json 2> gpurun_out/r2_run23_refarm.err
echo "reference arm rc=$?"; head -c 400 gpurun_out/r2_run23_refarm.json; echo
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
