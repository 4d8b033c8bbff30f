"""Validation record for the composed CPU baseline.

bench.py's `cpu_baseline` / `--impl reference` legs do not run the reference algorithm in full at the graded sizes (hours
per timeslice): they time its two primitives at full size - one `_nD` hop (lattice/generator/elemental.py:279-288) and one
(pair, momentum) einsum (:324-329) - and compose  T = 78 t_hop + 43 Nmom t_pair  (num_nabla = 2).  This tool runs the
oracle's FAITHFUL restatement of the reference (78 hops, 43 x Nmom einsums, the reference's own loop structure) in full on
one timeslice of config 3 (24^3, Ne = 100, 33 momenta) on this machine's cores and prints both figures side by side.

    python tests/cpu_port_validation.py [--workload config3] > profiles/r02/cpu_port_config3_full.json
"""
import argparse
import importlib.util
import json
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np  # noqa: E402

spec = importlib.util.spec_from_file_location("edk_bench", os.path.join(REPO, "bench.py"))
bench = importlib.util.module_from_spec(spec)
spec.loader.exec_module(bench)
from oracle import elemental_oracle as orc  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="config3", choices=sorted(bench.WORKLOADS))
    args = ap.parse_args()
    cores = bench.use_all_host_cores()
    Lx, Ly, Lz, Lt, Ne, nabla, nmom = bench.WORKLOADS[args.workload]
    latt = [Lx, Ly, Lz, Lt]
    moms = bench.momentum_set(nmom)
    rng = np.random.default_rng(orc.SEED0)
    W0 = (rng.standard_normal((Ne, Lz, Ly, Lx, 3)) + 1j * rng.standard_normal((Ne, Lz, Ly, Lx, 3))).astype(np.complex64)
    U = orc.links_file_to_spatial(orc.synthetic_links(latt, 0))
    composed_rate, sample = bench.cpu_sample(args.workload, W0, U)
    t0 = time.perf_counter()
    E = orc.elemental_timeslice(W0, U, latt, nabla, moms)  # the reference's loop structure, in full
    full_s = time.perf_counter() - t0
    out = {"workload": bench.workload_desc(args.workload), "cores": cores, "composed_seconds_per_timeslice": 1.0 / composed_rate,
           "composed_sample": sample, "full_run_seconds_per_timeslice": full_s, "composed_over_full": (1.0 / composed_rate) / full_s,
           "result_checksum": float(np.abs(E[0, 0]).sum()), "numpy": np.__version__}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
