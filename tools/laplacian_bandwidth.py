"""HBM throughput of the eigensolver's Laplacian operator (laplacian_kernel) at the BASELINE lattices:
algorithmic bytes = nvec x V x 48 B in + the same out + 3 x V x 144 B of links per application, timed with CUDA events
over inputs larger than L2.  Prints one JSON object.  Measurement tool, not part of the product."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import easydistillation_b200 as edb  # noqa: E402


def main():
    peak = 6453.4
    try:
        peak = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    out = {"peak_GBps": peak, "rows": []}
    dev = torch.device("cuda", 0)
    for L, nvec in ((24, 100), (32, 200), (48, 200)):
        latt = [L, L, L, 1]
        g0 = torch.Generator(device=dev)
        g0.manual_seed(1000 + L)
        U = bench.synth_links(torch, dev, L**3, g0).cpu().numpy().reshape(1, L, L, L, 4, 3, 3)
        lap = edb.Laplacian(latt, edb.GaugeFieldHostmem(U), device=0)
        lap.load("x")
        lap.set_timeslice(0)
        g = torch.Generator(device=dev)
        g.manual_seed(L)
        X = torch.randn((nvec, L, L, L, 3), dtype=torch.complex128, device=dev, generator=g)
        for _ in range(3):
            Y = lap.matmat(X)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        e0.record()
        for _ in range(reps):
            Y = lap.matmat(X)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        V = L**3
        nbytes = 2.0 * nvec * V * 48 + 3.0 * V * 144
        out["rows"].append({"lattice": [L, L, L], "nvec": nvec, "ms": ms, "algorithmic_GB": nbytes / 1e9, "GBps": nbytes / ms / 1e6,
                            "frac_of_peak": nbytes / ms / 1e6 / peak, "checksum": float(Y.abs().sum().item())})
        del lap, X, Y
        torch.cuda.empty_cache()
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
