#!/bin/bash
# Round 2, second GPU call: first hardware run of the separable contraction (form 4).
set -x
mkdir -p gpurun_out
timeout 900 python tools/check_forms.py --form 4 --bench --bench-shapes config2,config3,config4,config5 > gpurun_out/check_form4.log 2>&1; tail -8 gpurun_out/check_form4.log
timeout 400 python bench.py --workload config5 --contraction separable --no-cpu-baseline > gpurun_out/bench_config5_sep.json 2> gpurun_out/bench_config5_sep.err
tail -c 1500 gpurun_out/bench_config5_sep.json; tail -3 gpurun_out/bench_config5_sep.err
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:gram_sep_kernel|sep_zfold_kernel|combine_kernel' -s 3 -c 3 \
  -o gpurun_out/r02_config5_sep -f python bench.py --workload config5 --contraction separable --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_sep.log 2>&1
tail -5 gpurun_out/ncu_full_sep.log
ls -la gpurun_out | tail -12
