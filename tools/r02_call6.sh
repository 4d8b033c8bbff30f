#!/bin/bash
set -x
mkdir -p gpurun_out
EDK_SEP_VARIANT=6 timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:gram_sepx' -s 3 -c 3 -o gpurun_out/r02_config5_sepx -f \
    python tools/check_forms.py --form 4 --skip-cases --bench --bench-shapes config5 > gpurun_out/ncu_v6.log 2>&1
tail -2 gpurun_out/ncu_v6.log
