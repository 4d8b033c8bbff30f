#!/bin/bash
set -x
mkdir -p gpurun_out
EDK_SEP_VARIANT=6 timeout 600 python tools/check_forms.py --form 4 > gpurun_out/check_form4_v6.log 2>&1; tail -1 gpurun_out/check_form4_v6.log
for v in 6 5; do
  EDK_SEP_VARIANT=$v timeout 600 python tools/check_forms.py --form 4 --skip-cases --bench --bench-shapes config2,config3,config4,config5 > gpurun_out/bench_form4_v$v.log 2>&1
  grep -o '"form4_phase_ms": {[^}]*}\|"err_form4_vs_form[13]": [0-9.e-]*\|"workload": "[a-z0-9]*"' gpurun_out/bench_form4_v$v.log | tr '\n' ' '; echo
done
