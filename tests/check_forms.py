"""GPU check of one contraction form against the numpy oracle and the GEMM form, in its own process.

    python tests/check_forms.py --form 4 [--bench] [--bench-shapes config3,config4,config5]

Runs the form through the engine on ragged and multi-tile shapes, both pairing modes, derivative and displacement
jobs, every stage width of the separable kernel (8 / 6 / 4 site pairs) and every mode structure (5 / 9 / 13 modes),
and compares block by block (Frobenius, 1e-10) with the oracle and with form 1.  `--bench` adds a timing of the form
and of form 1 at the named workload shapes.  Writes gpurun_out/check_form<N>.json; exit status 1 on any mismatch.
"""
import argparse
import json
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from easydistillation_b200 import _capi  # noqa: E402
from easydistillation_b200.engine import ElementalEngine  # noqa: E402
from oracle import elemental_oracle as orc  # noqa: E402

TOL = 1e-10
SHAPES = {"config2": ([16, 16, 16], 100, 1, 9), "config3": ([24, 24, 24], 100, 2, 33), "config4": ([32, 32, 32], 200, 2, 33),
          "config5": ([48, 48, 48], 200, 2, 33)}


def worst_block_error(got, ref):
    norms = np.sqrt((np.abs(ref) ** 2).sum(axis=(-1, -2)))
    floor = 1e-4 * norms.max()
    w = 0.0
    for a in range(ref.shape[0]):
        for p in range(ref.shape[1]):
            w = max(w, float(np.linalg.norm(got[a, p] - ref[a, p]) / max(norms[a, p], floor)))
    return w


def run_case(form, latt, Ne, mode, order, moms, sym):
    U_file = orc.synthetic_links(latt + [1], 3)
    V = orc.synthetic_eigvecs(latt + [1], Ne, 3)
    U = orc.links_file_to_spatial(U_file)
    if mode == _capi.MODE_DERIVATIVE:
        ref = (orc.elemental_timeslice_closed_form if order <= 2 else orc.elemental_timeslice)(V, U, latt + [1], order, moms)
    else:
        ref = orc.displacement_timeslice(V, U, latt + [1], order, moms)
    eng = ElementalEngine(latt, Ne, mode, order, moms)
    planned = eng.query()["contraction_form"]
    if sym is not None:
        eng.debug_symmetry(sym)
    eng.set_links(torch.from_numpy(U_file).cuda(), _capi.LINKS_FILE_T)
    eng.set_eigvecs(torch.from_numpy(V).cuda())
    eng.debug_algo(1)
    gemm = eng.calc().cpu().numpy()
    eng.debug_algo(form)
    q = eng.query()
    assert q["contraction_form"] == form and q["plane_wave_modes"] >= 1, q
    got = eng.calc().cpu().numpy()
    again = eng.calc().cpu().numpy()
    res = {"form": form, "planned_form": planned, "latt": latt, "Ne": Ne, "mode": mode, "order": order, "nmom": len(moms), "sym": sym,
           "modes": q["plane_wave_modes"], "tile": q["plane_wave_tile"],
           "err_vs_oracle": worst_block_error(got, ref), "err_vs_gemm_form": worst_block_error(got, gemm),
           "gemm_vs_oracle": worst_block_error(gemm, ref), "deterministic": bool(np.array_equal(got, again))}
    res["ok"] = bool(res["err_vs_oracle"] < TOL and res["err_vs_gemm_form"] < TOL and res["deterministic"])
    eng.close()
    return res


def time_forms(form, latt, Ne, nabla, nmom, reps=3, with_gemm=True):
    moms = orc.momentum_set(nmom)
    eng = ElementalEngine(latt, Ne, _capi.MODE_DERIVATIVE, nabla, moms)
    g = torch.Generator(device="cuda").manual_seed(1)
    V = Ne * latt[0] * latt[1] * latt[2] * 3
    v = torch.view_as_complex(torch.randn((V, 2), generator=g, device="cuda", dtype=torch.float32)).reshape(
        Ne, latt[2], latt[1], latt[0], 3)
    v = v / torch.linalg.vector_norm(v.reshape(Ne, -1), dim=1)[:, None, None, None, None]
    U_file = torch.from_numpy(orc.synthetic_links(latt + [1], 1)).cuda()
    eng.set_links(U_file, _capi.LINKS_FILE_T)
    eng.set_eigvecs(v.contiguous())
    out = {"shape": {"latt": latt, "Ne": Ne, "num_nabla": nabla, "nmom": nmom}, "planned_form": eng.query()["contraction_form"]}
    results = {}
    for algo in ((1, form) if with_gemm else (3, form)):
        eng.debug_algo(algo)
        res = eng.calc()
        torch.cuda.synchronize()
        eng.set_profiling(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            eng.calc(res)
        e1.record()
        torch.cuda.synchronize()
        prof = eng.get_profile()
        eng.set_profiling(False)
        out[f"form{algo}_ms_per_timeslice"] = e0.elapsed_time(e1) / reps
        out[f"form{algo}_phase_ms"] = {k: v["ms"] / reps for k, v in prof.items()}
        results[algo] = res.cpu().numpy()
    base = 1 if with_gemm else 3
    out[f"err_form{form}_vs_form{base}"] = worst_block_error(results[form], results[base])
    eng.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--form", type=int, default=4, choices=[2, 3, 4])
    ap.add_argument("--bench", action="store_true")
    ap.add_argument("--bench-shapes", default="config3,config4")
    ap.add_argument("--skip-cases", action="store_true")
    args = ap.parse_args()
    assert torch.cuda.is_available(), "needs a CUDA device"
    D, X = _capi.MODE_DERIVATIVE, _capi.MODE_DISPLACEMENT
    form = args.form
    special = [(0, 0, 0), (0, 0, 1), (0, 1, 1), (1, 1, 1), (0, 0, 2), (0, 1, 2), (1, 1, 2)]  # the reference's test list
    if form == 4:
        cases = [
            ([8, 4, 4], 8, D, 0, orc.momentum_set(7), None),       # 4 pairs per stage, 5 modes
            ([8, 5, 2], 5, D, 1, orc.momentum_set(9), None),       # 9 modes
            ([16, 3, 2], 20, D, 1, orc.momentum_set(9), None),     # 8 pairs per stage
            ([12, 6, 4], 30, D, 2, orc.momentum_set(9), None),     # 6 pairs per stage, 2 x 1 tiles, pairing by cost
            ([12, 6, 4], 35, D, 2, orc.momentum_set(33), 1),       # 3 x 2 tiles, 13 modes, Hermitian pairing + half set
            ([8, 6, 4], 19, D, 2, orc.momentum_set(33), 0),        # direct pairs: multi-segment jobs with signs
            ([8, 4, 2], 21, D, 2, special, None),                  # non-closed list
            ([24, 2, 2], 9, D, 2, orc.momentum_set(33), None),     # two stages per row
            ([8, 4, 6], 13, D, 3, orc.momentum_set(7), None),
            ([8, 6, 8], 12, X, 3, orc.momentum_set(19), None),
            ([8, 3, 7], 110, D, 1, orc.momentum_set(9), None),     # 7 x 4 tiles, edge warps
            ([16, 4, 4], 70, D, 1, orc.momentum_set(33), None),    # mirror tiles of the self pair
        ]
    else:
        cases = [
            ([4, 4, 4], 8, D, 0, orc.momentum_set(7), None),
            ([3, 5, 2], 5, D, 1, orc.momentum_set(7), None),
            ([4, 6, 8], 30, D, 2, orc.momentum_set(9), None),
            ([4, 6, 8], 35, D, 2, orc.momentum_set(33), 1),
            ([4, 6, 8], 19, D, 2, orc.momentum_set(33), 0),
            ([6, 4, 2], 21, D, 2, [(0, 0, 1), (1, 2, 0), (3, -1, 2), (0, -2, 1)], None),
            ([2, 2, 2], 1, D, 2, orc.momentum_set(7), None),
            ([3, 3, 2], 9, D, 2, orc.momentum_set(33), None),
            ([4, 4, 6], 13, D, 3, orc.momentum_set(7), None),
            ([4, 6, 8], 12, X, 3, orc.momentum_set(9), None),
            ([5, 3, 7], 110, D, 1, orc.momentum_set(9), None),
        ]
    report = {"form": form, "cases": [], "ok": True}
    if not args.skip_cases:
        for c in cases:
            t0 = time.time()
            r = run_case(form, *c)
            r["seconds"] = time.time() - t0
            report["cases"].append(r)
            report["ok"] = report["ok"] and r["ok"]
            print(json.dumps(r), flush=True)
    if args.bench and report["ok"]:
        report["timing"] = []
        for name in args.bench_shapes.split(","):
            latt, Ne, nabla, nmom = SHAPES[name]
            t = time_forms(form, latt, Ne, nabla, nmom, reps=2 if name in ("config4", "config5") else 3, with_gemm=name != "config5")
            t["workload"] = name
            report["timing"].append(t)
            print(json.dumps(t), flush=True)
    os.makedirs(os.path.join(REPO, "gpurun_out"), exist_ok=True)
    with open(os.path.join(REPO, "gpurun_out", f"check_form{form}.json"), "w") as f:
        json.dump(report, f, indent=1)
    print(f"contraction form {form}:", "OK" if report["ok"] else "MISMATCH")
    return 0 if report["ok"] else 1


if __name__ == "__main__":
    sys.exit(main())
