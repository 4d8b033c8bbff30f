#!/bin/bash
# Round 2 final evidence run (one B200): GPU suite, smoke, the bench line of every BASELINE configuration with the library
# default, the reference arm, ncu launch list + full capture, error growth, Laplacian throughput.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_config5_1gpu.json 2> gpurun_out/bench_config5.err; tail -2 gpurun_out/bench_config5.err
for wl in config4 config3 config2; do
  timeout 400 python bench.py --workload $wl --steps 10 --warmup 3 > gpurun_out/bench_${wl}_1gpu.json 2> gpurun_out/bench_${wl}.err
done
timeout 400 python bench.py --workload config3 --generator displacement --distance 2 --steps 10 --warmup 3 > gpurun_out/bench_config3_displacement_1gpu.json 2> gpurun_out/bench_config3_disp.err
for f in gpurun_out/bench_*_1gpu.json; do python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1])
print('$f', 'value %.3f e2e %.3f frac %.3f pipe %.3f stencil_frac %.3f err %s'%(d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['fp64_pipe_utilisation'], d['roofline_stencil']['frac'], d['contraction'].get('parity_check',{}).get('worst_block_rel_err')), {k:round(v,3) for k,v in d['phase_ms_per_step'].items()})
"; done
timeout 900 python bench.py --impl reference > gpurun_out/bench_config5_reference_arm.json 2> gpurun_out/bench_ref.err; tail -c 600 gpurun_out/bench_config5_reference_arm.json
timeout 300 python tools/laplacian_bandwidth.py > gpurun_out/laplacian_bandwidth.json 2> gpurun_out/laplacian.err; cat gpurun_out/laplacian_bandwidth.json | tr -d '\n' | cut -c1-900; echo
timeout 600 python tools/error_growth.py > gpurun_out/error_growth.json 2> gpurun_out/error_growth.err
# launch list and full captures (numbers printed under ncu are not bench values)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_config5.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-file-leg --no-parity-check > gpurun_out/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:gram_sepx|sep_zfold_kernel|nabla3_kernel|combine_kernel' -s 45 -c 9 \
  -o gpurun_out/r02_config5_final -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-file-leg --no-parity-check > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out | tail -12
