"""Observed error of the planned contraction form (separable, form 4) and of the folded plane-wave form (3) against the
GEMM form (1) as the lattice grows, by |p|^2: the three forms sum the same products in different orders, so this is the
growth of the summation error with the volume.  Tolerance of the path: 1e-10 (block-wise Frobenius).

    python tools/error_growth.py > profiles/r02/error_growth.json
"""
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from easydistillation_b200 import _capi  # noqa: E402
from easydistillation_b200.engine import ElementalEngine  # noqa: E402
import bench  # noqa: E402


def block_errors(got, ref):
    norms = np.sqrt((np.abs(ref) ** 2).sum(axis=(-1, -2)))
    floor = 1e-4 * norms.max()
    return np.sqrt((np.abs(got - ref) ** 2).sum(axis=(-1, -2))) / np.maximum(norms, floor)  # [Nop, Nmom]


def main():
    moms = bench.momentum_set(33)
    p2 = np.array([sum(c * c for c in m) for m in moms])
    out = []
    for L, Ne in ((8, 32), (16, 32), (24, 32), (32, 32), (48, 32), (48, 200)):
        latt = [L, L, L]
        eng = ElementalEngine(latt, Ne, _capi.MODE_DERIVATIVE, 2, moms)
        g = torch.Generator(device="cuda").manual_seed(7)
        V = Ne * L ** 3 * 3
        v = torch.view_as_complex(torch.randn((V, 2), generator=g, device="cuda", dtype=torch.float32)).reshape(Ne, L, L, L, 3)
        v = (v / torch.linalg.vector_norm(v.reshape(Ne, -1), dim=1)[:, None, None, None, None]).contiguous()
        U = bench.synth_links(torch, torch.device("cuda"), L ** 3, g)
        eng.set_links(U, _capi.LINKS_FILE_T)
        eng.set_eigvecs(v)
        res = {}
        for form in (1, 3, 4):
            eng.debug_algo(form)
            res[form] = eng.calc().cpu().numpy()
        row = {"lattice": latt, "Ne": Ne, "sites": L ** 3}
        for form in (3, 4):
            e = block_errors(res[form], res[1])
            row[f"form{form}_worst"] = float(e.max())
            row[f"form{form}_worst_by_p2"] = {int(k): float(e[:, p2 == k].max()) for k in sorted(set(p2))}
            row[f"form{form}_worst_by_order"] = {"no derivative": float(e[0].max()), "first": float(e[1:4].max()), "second": float(e[4:].max())}
        out.append(row)
        eng.close()
        del eng
        torch.cuda.empty_cache()
        print(json.dumps(row), file=sys.stderr, flush=True)
    print(json.dumps({"tolerance": 1e-10, "reference": "GEMM form (3M arithmetic on DMMA)", "rows": out}, indent=1))


if __name__ == "__main__":
    main()
