#!/bin/bash
# N GPUs of one box (gpurun --gpus N): the sharded public call over NCCL, then the bench contract line under torchrun.
N=${1:-2}
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 \
  tests/check_calc_all_nccl.py > gpurun_out/calc_all_nccl_${N}gpu.json 2> gpurun_out/calc_all_nccl_${N}gpu.err
tail -3 gpurun_out/calc_all_nccl_${N}gpu.err | cut -c1-300; cat gpurun_out/calc_all_nccl_${N}gpu.json | cut -c1-1200
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 \
  bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_config5_${N}gpu.json 2> gpurun_out/bench_config5_${N}gpu.err
tail -3 gpurun_out/bench_config5_${N}gpu.err | cut -c1-300
python -c "
import json
d=json.loads(open('gpurun_out/bench_config5_${N}gpu.json').read().strip().splitlines()[-1])
print('BENCH N=$N value %.3f ms/step %.3f e2e %.3f host_queue %.2f'%(d['value'], d['ms_per_step'], d['e2e']['value'], d['host_queue_ms']), d['config']['sharding'])
"
