"""Two numbers of the rows next to the hot path, at config 5's lattice (measurement tool, not part of the product):
 (1) gauge preprocessing on the device: time of `stout_smear(20, 0.12)` + `project_SU3()` per timeslice,
     read from the library's own CUDA-event profile of the prepare phase, with and without the link operations;
 (2) the elemental file writer: `calc_to_file` of a few config-5 timeslices into a `.npy` on tmpfs,
     timeslices/s and GB/s written, next to `calc_range` on the same inputs.
Prints one JSON object."""
import json
import os
import shutil
import sys
import tempfile
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import easydistillation_b200 as edb  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "config5"
    K = 4
    Lx, Ly, Lz, Lt, Ne, nabla, nmom = bench.WORKLOADS[name]
    dev = torch.device("cuda", 0)
    moms = bench.momentum_set(nmom)
    U, v = bench.synth_device_inputs(torch, dev, name, 0)
    U_host = np.stack([U.cpu().numpy().reshape(Lz, Ly, Lx, 4, 3, 3)] * K)
    V_host = np.stack([v.cpu().numpy().reshape(Ne, Lz, Ly, Lx, 3)] * K)
    del U, v
    out = {"workload": name}
    gen = edb.ElementalGenerator([Lx, Ly, Lz, K], edb.GaugeFieldHostmem(U_host), edb.EigenvectorHostmem(V_host), nabla, moms, device=0)
    gen.load("aux")
    eng = gen._engine

    def prepare_ms():
        gen.calc_device(0)
        torch.cuda.synchronize()
        eng.set_profiling(True)
        for t in range(3):
            gen.calc_device(t % K)
        torch.cuda.synchronize()
        p = eng.get_profile()
        eng.set_profiling(False)
        return p["prepare"]["ms"] / 3, p["prepare"]["launches"] / 3

    base, nb = prepare_ms()
    gen.stout_smear(20, 0.12)
    gen.project_SU3()
    smeared, ns = prepare_ms()
    out["gauge_preprocessing"] = {"prepare_ms_plain": base, "prepare_ms_with_stout20_and_projection": smeared,
                                  "link_ops_ms_per_timeslice": smeared - base, "launches_plain": nb, "launches_with_ops": ns,
                                  "links_MB": Lx * Ly * Lz * 3 * 144 / 1e6}
    gen.load("aux")  # drops the link operations again

    gen.calc_range(0, 2)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    res = gen.calc_range(0, K)
    t_range = time.perf_counter() - t0
    tmp = tempfile.mkdtemp(prefix="edk_aux_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    try:
        handle = edb.ElementalNpy(os.path.join(tmp, "cfg_"), ".elemental.npy", list(res.shape[1:3]) + [K, Ne, Ne], Ne)
        gen.calc_to_file(handle, "warm")
        t0 = time.perf_counter()
        gen.calc_to_file(handle, "a")
        t_file = time.perf_counter() - t0
        back = np.load(os.path.join(tmp, "cfg_a.elemental.npy"), mmap_mode="r")
        same = bool(np.array_equal(np.asarray(back[:, :, 1]), res[1]))
        nbytes = os.path.getsize(os.path.join(tmp, "cfg_a.elemental.npy"))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    out["file_writer"] = {"timeslices": K, "calc_range_timeslices_per_s": K / t_range, "calc_to_file_timeslices_per_s": K / t_file,
                          "file_GB": nbytes / 1e9, "written_GBps": nbytes / 1e9 / t_file, "file_equals_calc_range": same}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
