#!/bin/bash
# bench.py after the warm-up / clock-sampler changes: is the device-resident leg stable at small and large workloads?
set -x
mkdir -p gpurun_out
for i in 1 2; do
for wl in config2 config3 config5; do
  timeout 400 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline --no-file-leg > gpurun_out/q_${wl}_$i.json 2> gpurun_out/q_${wl}_$i.err
  python -c "
import json
d=json.loads(open('gpurun_out/q_${wl}_$i.json').read().strip().splitlines()[-1])
ph=d['phase_ms_per_step']
print('Q $wl $i', 'value %.3f ms/step %.3f phases %.3f host_queue %.2f e2e %.3f pipe %.3f frac %.3f'%(d['value'], d['ms_per_step'], sum(ph.values()), d['host_queue_ms'], d['e2e']['value'], d['roofline']['fp64_pipe_utilisation'], d['roofline']['frac']), {k:round(v,3) for k,v in ph.items()})
"
done
done
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
