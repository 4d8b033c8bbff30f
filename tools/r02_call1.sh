#!/bin/bash
# Round 2, first GPU call: state of the round-1 code with the folded plane-wave form forced, plus the ncu evidence
# the round-1 verdict asked for (full capture of gram_pwf_kernel / pw_zfold_kernel / nabla3_kernel at config 5).
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 400 python bench.py --workload config5 --contraction planewave-folded --no-cpu-baseline > gpurun_out/bench_config5_pwf.json 2> gpurun_out/bench_config5_pwf.err
tail -c 600 gpurun_out/bench_config5_pwf.json
# launch list of one run (times under ncu are serialised / cold: only the kernel's share is meaningful)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_config5_pwf.csv \
  python bench.py --workload config5 --contraction planewave-folded --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
# full capture: one step's kernels (4 nabla3, gram_pwf, pw_zfold, combine), skipping the first step
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:gram_pwf_kernel|pw_zfold_kernel|nabla3_kernel|combine_kernel' -s 7 -c 7 \
  -o gpurun_out/r02_config5_pwf -f python bench.py --workload config5 --contraction planewave-folded --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -5 gpurun_out/ncu_full.log
ls -la gpurun_out | tail -20
