#!/bin/bash
# lane remapping (2-cycle LDS.128): parity on the ragged cases for variants 0 and 2, timing of all variants
set -x
mkdir -p gpurun_out
for v in 0 2; do
  EDK_SEP_VARIANT=$v timeout 600 python tools/check_forms.py --form 4 > gpurun_out/check_form4_v$v.log 2>&1; tail -1 gpurun_out/check_form4_v$v.log
done
for v in 0 1 2 3 4; do
  EDK_SEP_VARIANT=$v timeout 600 python tools/check_forms.py --form 4 --skip-cases --bench --bench-shapes config4,config5 > gpurun_out/bench_form4_v$v.log 2>&1
  grep -o '"form4_phase_ms": {[^}]*}\|"err_form4_vs_form[13]": [0-9.e-]*\|"workload": "[a-z0-9]*"' gpurun_out/bench_form4_v$v.log | tr '\n' ' '; echo
done
