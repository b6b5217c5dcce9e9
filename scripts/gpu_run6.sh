#!/bin/bash
# round-2 sixth GPU pass (2 GPUs): curves tests after the rounding fix, gather modes A/B at N=2, pathtracer at N=2
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -k "curve or gather or pathstream" > gpurun_out/r2_run6_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_run6_pytest.log
tail -12 gpurun_out/r2_run6_pytest.log
N=${NGPU:-2}
RTCB200_GATHER_AB=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 5 --warmup 3 --no-cpu --no-extras > gpurun_out/r2_run6_n$N.json 2> gpurun_out/r2_run6_n$N.err
echo "bench N=$N rc=$?"; tail -2 gpurun_out/r2_run6_n$N.err; head -c 600 gpurun_out/r2_run6_n$N.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
  bench.py --workload pathtracer --gpus $N --steps 3 --warmup 2 --no-cpu > gpurun_out/r2_run6_pt_n$N.json 2> gpurun_out/r2_run6_pt_n$N.err
echo "pt N=$N rc=$?"; tail -2 gpurun_out/r2_run6_pt_n$N.err; head -c 400 gpurun_out/r2_run6_pt_n$N.json
nvidia-smi topo -m > gpurun_out/r2_topo_n$N.txt 2>&1
