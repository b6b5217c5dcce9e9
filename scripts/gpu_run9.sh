#!/bin/bash
# round-2 multi-GPU pass: bench at N GPUs with both gather delivery modes measured in the same process, then the path-tracer stream
mkdir -p gpurun_out
N=${NGPU:-8}
nvidia-smi topo -m > gpurun_out/r2_topo_n$N.txt 2>&1
RTCB200_GATHER_AB=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 4 --warmup 3 --no-cpu --no-extras > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err
echo "bench N=$N rc=$?"; tail -2 gpurun_out/r2_bench_n$N.err
python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_n$N.json').read().strip().splitlines()[-1]); print('value', d['value'], 'ms', d['ms_per_step'], 'verified', d['gather_verified'], d['gather_ab']); print('trace_ms', d['per_rank']['trace_ms']); print('e2e', d['e2e']['value'], 'numa', d['numa'])"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
  bench.py --workload pathtracer --gpus $N --steps 2 --warmup 2 --no-cpu > gpurun_out/r2_bench_pathtracer_n$N.json 2> gpurun_out/r2_bench_pathtracer_n$N.err
echo "pt N=$N rc=$?"; tail -2 gpurun_out/r2_bench_pathtracer_n$N.err
python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_pathtracer_n$N.json').read().strip().splitlines()[-1]); print('pt value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'])"
