#!/bin/bash
# A/B of the y-stage / warp-skew changes of gram_sepx_kernel at config 5, host-time diagnosis of the small workloads,
# and the full ncu capture of the default build.
set -x
mkdir -p gpurun_out
python tools/diag_value_path.py config2 > gpurun_out/diag_config2.json 2> gpurun_out/diag_config2.err; tail -5 gpurun_out/diag_config2.err
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-file-leg"
for lib in sm100a ypair; do
  for skew in 0 1; do
    EDK_LIBRARY=$PWD/easydistillation_b200/libedk_$lib.so EDK_SEP_SKEW=$skew timeout 300 $B > gpurun_out/ab_${lib}_skew$skew.json 2> gpurun_out/ab_${lib}_skew$skew.err
    python -c "
import json
d=json.loads(open('gpurun_out/ab_${lib}_skew$skew.json').read().strip().splitlines()[-1])
print('AB $lib skew$skew', 'value %.3f'%d['value'], {k:round(v,3) for k,v in d['phase_ms_per_step'].items()}, 'err', d['contraction'].get('parity_check',{}).get('worst_block_rel_err'))
"
  done
done
for wl in config4 config3; do
  for skew in 0 1; do
    EDK_SEP_SKEW=$skew timeout 300 $B --workload $wl > gpurun_out/ab_${wl}_skew$skew.json 2> gpurun_out/ab_${wl}_skew$skew.err
    python -c "
import json
d=json.loads(open('gpurun_out/ab_${wl}_skew$skew.json').read().strip().splitlines()[-1])
print('AB $wl skew$skew', 'value %.3f'%d['value'], {k:round(v,3) for k,v in d['phase_ms_per_step'].items()}, 'err', d['contraction'].get('parity_check',{}).get('worst_block_rel_err'))
"
  done
done
python tools/check_forms.py --form 4 > gpurun_out/check_form4_skew.log 2>&1; tail -2 gpurun_out/check_form4_skew.log
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_config5.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-file-leg --no-parity-check > gpurun_out/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:gram_sepx_kernel|sep_zfold_kernel|nabla3_kernel|combine_kernel' -s 27 -c 9 \
  -o gpurun_out/r02_config5_final -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-file-leg --no-parity-check > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out | tail -12
