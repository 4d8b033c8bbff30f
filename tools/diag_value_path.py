"""Where does the host time of the device-resident throughput leg go at a small workload?  (config 2: 0.5 ms of kernels per
timeslice.)  Times gen.calc_all() the way bench.py does - with and without the clock sampler thread, with and without
per-phase profiling events - and prints a cProfile of one run.  Measurement tool, not part of the product."""
import cProfile
import io
import json
import os
import pstats
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import easydistillation_b200 as edb  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "config2"
    K = 10
    Lx, Ly, Lz, Lt, Ne, nabla, nmom = bench.WORKLOADS[name]
    dev = torch.device("cuda", 0)
    moms = bench.momentum_set(nmom)
    inputs = [bench.synth_device_inputs(torch, dev, name, i) for i in range(2)]
    gen = edb.ElementalGenerator([Lx, Ly, Lz, K], edb.GaugeFieldDevice([U.reshape(Lz, Ly, Lx, 4, 3, 3) for U, _ in inputs], cyclic=True),
                                 edb.EigenvectorDevice([v.reshape(Ne, Lz, Ly, Lx, 3) for _, v in inputs], cyclic=True), nabla, moms, device=0)
    gen.load("diag")
    eng = gen._engine
    scratch = torch.empty(eng.out_shape, dtype=torch.complex128, device=dev)
    for i in range(3):
        gen.calc_device(i, out=scratch)
    torch.cuda.synchronize()
    out = {}

    def timed(label, sampler, profiling):
        eng.set_profiling(profiling)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ctx = bench.ClockSampler(0) if sampler else None
        if ctx:
            ctx.__enter__()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e0.record()
        g = gen.calc_all(dst=0)
        t1 = time.perf_counter()
        e1.record()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        if ctx:
            ctx.__exit__()
        if profiling:
            eng.get_profile()
        eng.set_profiling(False)
        del g
        out[label] = {"device_ms": e0.elapsed_time(e1), "host_queue_ms": 1e3 * (t1 - t0), "host_total_ms": 1e3 * (t2 - t0)}

    for rep in range(2):
        timed(f"plain_{rep}", False, False)
        timed(f"profiling_{rep}", False, True)
        timed(f"sampler_{rep}", True, False)
        timed(f"sampler+profiling_{rep}", True, True)
    pr = cProfile.Profile()
    pr.enable()
    g = gen.calc_all(dst=0)
    torch.cuda.synchronize()
    pr.disable()
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(18)
    out["cprofile"] = s.getvalue().splitlines()[:45]
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
