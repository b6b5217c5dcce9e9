#!/bin/bash
# round-2 seventh GPU pass (2 GPUs): gather staging through a local buffer, unrolled SPREAD queue A/B
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -k "gather or curve" > gpurun_out/r2_run7_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_run7_pytest.log
tail -6 gpurun_out/r2_run7_pytest.log
B=scripts/_build
L=embree_b200/csrc/libembree4_b200.so
timeout 600 python scripts/ab.py prev=$B/lib_prev.so new=$L spl3=$B/lib_spl3.so spl6=$B/lib_spl6.so > gpurun_out/r2_run7_ab.txt 2>&1
cat gpurun_out/r2_run7_ab.txt
N=2
RTCB200_GATHER_AB=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 5 --warmup 3 --no-cpu --no-extras --no-e2e > gpurun_out/r2_run7_n$N.json 2> gpurun_out/r2_run7_n$N.err
echo "bench N=$N rc=$?"; tail -2 gpurun_out/r2_run7_n$N.err
python -c "
import json; d=json.loads(open('gpurun_out/r2_run7_n2.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['gather_verified'], d['gather_ab'], d['per_rank']['trace_ms'])"
