#!/bin/bash
# Round 2, third GPU call: tensor-memory scratch test, then the separable contraction with its y-stage accumulators in TMEM
# (variants 1-4) against the register variant (0): parity on the ragged cases, timing at the config-4 / config-5 shapes.
set -x
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/tmem_scratch tools/microbench/tmem_scratch.cu 2>/dev/null && timeout 60 /tmp/tmem_scratch > gpurun_out/tmem_scratch.log 2>&1; cat gpurun_out/tmem_scratch.log
for v in 1 2 3 4; do
  EDK_SEP_VARIANT=$v timeout 600 python tools/check_forms.py --form 4 > gpurun_out/check_form4_v$v.log 2>&1; tail -1 gpurun_out/check_form4_v$v.log
done
for v in 0 1 2 3 4; do
  EDK_SEP_VARIANT=$v timeout 600 python tools/check_forms.py --form 4 --skip-cases --bench --bench-shapes config3,config4,config5 > gpurun_out/bench_form4_v$v.log 2>&1
  grep -o '"form4_phase_ms": {[^}]*}\|"err_form4_vs_form[13]": [0-9.e-]*\|"workload": "[a-z0-9]*"' gpurun_out/bench_form4_v$v.log | tr '\n' ' '; echo
done
ls -la gpurun_out | tail -12
