#!/bin/bash
# round-2 pass 20 (2 GPUs): the new single-GPU tests (hair tutorial, interpolate, gather, pathstream), then the driver's own N=2 launch lines
# of the default bench (exactly as the driver runs it: no flags beyond --gpus/--steps/--warmup) and of the reference arm, then the path tracer
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -k "hair_geometry or interpolate or gather or pathstream or tutorial" > gpurun_out/r2_run20_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_run20_pytest.log
tail -6 gpurun_out/r2_run20_pytest.log
N=2
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2_run20_n$N.json 2> gpurun_out/r2_run20_n$N.err
echo "bench N=$N rc=$?"; tail -2 gpurun_out/r2_run20_n$N.err; head -c 500 gpurun_out/r2_run20_n$N.json; echo
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 \
  bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/r2_run20_ref_n$N.json 2> gpurun_out/r2_run20_ref_n$N.err
echo "reference arm N=$N rc=$?"; tail -2 gpurun_out/r2_run20_ref_n$N.err; head -c 500 gpurun_out/r2_run20_ref_n$N.json; echo
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
  bench.py --workload pathtracer --gpus $N --steps 3 --warmup 2 --no-cpu > gpurun_out/r2_run20_pt_n$N.json 2> gpurun_out/r2_run20_pt_n$N.err
echo "pt N=$N rc=$?"; tail -2 gpurun_out/r2_run20_pt_n$N.err; head -c 400 gpurun_out/r2_run20_pt_n$N.json; echo
