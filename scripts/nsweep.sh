# usage: nsweep.sh N cfg1 cfg2 ...   (bench.py under torchrun at N GPUs with RTCB200_TUNING=cfg; "default" = none)
N=$1; shift
for cfg in "$@"; do
T=$cfg; [ "$cfg" = default ] && T=""
RTCB200_TUNING=$T python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 --no-e2e > gpurun_out/n${N}_$cfg.json 2> gpurun_out/n${N}_$cfg.err
python -c "
import json,sys; d=json.loads(open('gpurun_out/n${N}_$cfg.json').read().strip().splitlines()[-1]); print('N=$N $cfg', round(d['value'],1), 'Mrays/s', round(d['ms_per_step'],2), 'ms/step gather_ok', d['gather_verified'], 'kernel_ms', round(d['roofline']['kernel_ms'],2))"
done
