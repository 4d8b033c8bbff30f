#!/bin/bash
# ncu launch list (gpu__time_duration.sum only) of the library's own kernels over whole timeslices of config 5
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none \
  -k 'regex:^(nabla3|gram_|sep_|combine|reorder_links|round_eigvecs|displace|phase_|pw_|stout|project)' -c 110 --csv \
  --log-file gpurun_out/launches_config5.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-file-leg --no-parity-check > gpurun_out/ncu_launches.log 2>&1
tail -2 gpurun_out/ncu_launches.log | cut -c1-200; wc -l gpurun_out/launches_config5.csv
