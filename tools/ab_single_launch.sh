#!/bin/bash
# A/B: separable contraction launched once over the whole tile table (default) against one launch per tile shape
# (EDK_SEP_LAUNCHES=3), then parity, the GPU suite and the ncu captures of the default.
set -x
mkdir -p gpurun_out
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-file-leg"
for wl in config5 config4 config3 config2; do
  for n in 3 1; do
    EDK_SEP_LAUNCHES=$n timeout 300 $B --workload $wl > gpurun_out/ab_launches${n}_$wl.json 2> gpurun_out/ab_launches${n}_$wl.err
    python -c "
import json
d=json.loads(open('gpurun_out/ab_launches${n}_$wl.json').read().strip().splitlines()[-1])
print('AB $wl launches=$n', 'value %.3f ms/step %.3f pipe %.3f frac %.3f'%(d['value'], d['ms_per_step'], d['roofline']['fp64_pipe_utilisation'], d['roofline']['frac']), {k:round(v,3) for k,v in d['phase_ms_per_step'].items()}, 'err', d['contraction'].get('parity_check',{}).get('worst_block_rel_err'), 'launches', d['gpu_launches'])
"
  done
done
timeout 300 python tools/laplacian_bandwidth.py > gpurun_out/laplacian_bandwidth.json 2> gpurun_out/laplacian.err; cat gpurun_out/laplacian_bandwidth.json | tr -d '\n ' | cut -c1-700; echo
python tests/check_forms.py --form 4 > gpurun_out/check_form4_merged.log 2>&1; tail -1 gpurun_out/check_form4_merged.log
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_config5.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-file-leg --no-parity-check > gpurun_out/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:gram_sepx|sep_zfold_kernel|nabla3_kernel|combine_kernel' -s 35 -c 7 \
  -o gpurun_out/r02_config5_final -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-file-leg --no-parity-check > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
