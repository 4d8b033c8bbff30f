#!/bin/bash
# First gpurun call of the next round: hardware validation + measurement of the plane-wave contraction
# (DESIGN.md 3.3b), everything into gpurun_out/.  One GPU, about 12 minutes.
#
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/round2_first_call.sh'
#
# Afterwards, here:  python tools/ncu_summary.py gpurun_out/pw_config4.ncu-rep > profiles/r02/gram_pw_config4_ncu.txt
set -x
mkdir -p gpurun_out
# 1. the whole GPU suite (the plane-wave check runs in its own process inside it)
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
# 2. plane-wave form against the oracle and the GEMM form, with timing at config 3 / config 4
timeout 600 python tools/check_plane_wave.py --bench > gpurun_out/plane_wave_check.log 2>&1; tail -4 gpurun_out/plane_wave_check.log
timeout 600 python tools/check_plane_wave.py --form3 --bench > gpurun_out/plane_wave_check_form3.log 2>&1; tail -4 gpurun_out/plane_wave_check_form3.log
# 3. both forms, both tile shapes, at the graded workloads
for wl in config4 config5; do
  timeout 400 python bench.py --workload $wl --contraction gemm --no-cpu-baseline > gpurun_out/bench_${wl}_gemm.json 2> gpurun_out/bench_${wl}_gemm.err
  EDK_PW_TILE=25 timeout 400 python bench.py --workload $wl --contraction planewave --no-cpu-baseline > gpurun_out/bench_${wl}_pw25.json 2> gpurun_out/bench_${wl}_pw25.err
  EDK_PW_TILE=24 timeout 400 python bench.py --workload $wl --contraction planewave --no-cpu-baseline > gpurun_out/bench_${wl}_pw24.json 2> gpurun_out/bench_${wl}_pw24.err
  timeout 400 python bench.py --workload $wl --contraction planewave-folded --no-cpu-baseline > gpurun_out/bench_${wl}_pwf24.json 2> gpurun_out/bench_${wl}_pwf24.err
  EDK_PW_TILE=25 timeout 400 python bench.py --workload $wl --contraction planewave-folded --no-cpu-baseline > gpurun_out/bench_${wl}_pwf25.json 2> gpurun_out/bench_${wl}_pwf25.err
done
# 4. the default line (auto selection, CPU baseline included)
timeout 600 python bench.py > gpurun_out/bench_config5_auto.json 2> gpurun_out/bench_config5_auto.err; tail -2 gpurun_out/bench_config5_auto.err
# 5. launch list and one full capture of the plane-wave kernel (numbers printed under ncu are not bench values)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_config3_pw.csv \
  python bench.py --workload config3 --contraction planewave --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gram_pw_kernel -s 1 -c 2 -o gpurun_out/pw_config4 -f \
  python bench.py --workload config4 --contraction planewave --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out | tail -20
