#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python tools/check_forms.py --form 4 --bench --bench-shapes config2,config3,config4,config5 > gpurun_out/check_form4_v6.log 2>&1; tail -1 gpurun_out/check_form4_v6.log
grep -o '"form4_phase_ms": {[^}]*}\|"err_form4_vs_form[13]": [0-9.e-]*\|"workload": "[a-z0-9]*"' gpurun_out/check_form4_v6.log | tr '\n' ' '; echo
