#!/bin/bash
# full GPU suite, smoke, default bench line (config 5) and the other BASELINE configurations with the product default
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -15 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_config5.json 2> gpurun_out/bench_config5.err; tail -c 2500 gpurun_out/bench_config5.json; tail -3 gpurun_out/bench_config5.err
