#!/bin/bash
# BASELINE configs[4]: the synthetic 1 M-point x 6-camera frame at N GPUs (one frame per GPU per step, no data-path collective).
#   gpurun --gpus N -- 'bash tools/c5_sweep.sh N'
N=${1:-1}
mkdir -p gpurun_out
if [ "$N" = "1" ]; then
  timeout 900 python bench.py --points 1000000 --sweeps 33 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_c5_n1.json 2> gpurun_out/r2_c5_n1.err
else
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --points 1000000 --sweeps 33 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_c5_n$N.json 2> gpurun_out/r2_c5_n$N.err
fi
echo "exit $?"; tail -c 400 gpurun_out/r2_c5_n$N.err
python -c "
import json; d=json.loads(open('gpurun_out/r2_c5_n$N.json').read().strip().splitlines()[-1]); print(d['n_gpus'], d['value'], d['ms_per_step'], d['hot_scope'], d['clocks'])"
