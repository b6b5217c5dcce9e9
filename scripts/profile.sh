#!/bin/bash
# ncu passes of the bench command (run under gpurun). Outputs in gpurun_out/.
set -x
mkdir -p gpurun_out
RAYS=${RAYS:-67108864}
ARGS="--rays $RAYS --steps 2 --warmup 1 --no-e2e --no-cpu --no-extras"
# launch list (every kernel with its device time)
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches.csv python bench.py $ARGS > gpurun_out/launches.log 2>&1
# full set on the trace kernel: skip the primary-ray launch and the stats launch (ids 0,1), take the 3rd trace launch
ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 3 -c 1 -o gpurun_out/trace_full -f python bench.py $ARGS > gpurun_out/trace_full.log 2>&1
ls -la gpurun_out
