#!/bin/bash
# last single-GPU check of the committed build: GPU suite, smoke, default bench line, small workloads, Laplacian
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_config5_1gpu.json 2> gpurun_out/bench_config5.err; tail -2 gpurun_out/bench_config5.err
for wl in config3 config2; do
  timeout 400 python bench.py --workload $wl --steps 10 --warmup 3 > gpurun_out/bench_${wl}_1gpu.json 2> gpurun_out/bench_${wl}.err
done
for f in gpurun_out/bench_config5_1gpu.json gpurun_out/bench_config3_1gpu.json gpurun_out/bench_config2_1gpu.json; do python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1])
print('$f', 'value %.3f e2e %.3f frac %.3f pipe %.3f launches %d'%(d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['fp64_pipe_utilisation'], d['gpu_launches']), {k:round(v,3) for k,v in d['phase_ms_per_step'].items()})
"; done
timeout 300 python tools/laplacian_bandwidth.py > gpurun_out/laplacian_bandwidth.json 2> gpurun_out/laplacian.err; cat gpurun_out/laplacian_bandwidth.json | tr -d '\n ' | cut -c1-900; echo
