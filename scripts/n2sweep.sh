for cfg in gather_chunks=4 gather_chunks=16 gather_chunks=32 gather_mode=0; do
RTCB200_TUNING=$cfg python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-e2e > gpurun_out/n2_$cfg.json 2> gpurun_out/n2_$cfg.err
python -c "
import json,sys; d=json.loads(open('gpurun_out/n2_$cfg.json').read().strip().splitlines()[-1]); print('$cfg', round(d['value'],1), round(d['ms_per_step'],2), d['gather_verified'], d['gpu_launches'])"
done
