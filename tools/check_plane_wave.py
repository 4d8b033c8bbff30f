"""GPU check of the plane-wave factorised contraction (edk_debug_algo 2, csrc/edk_gram_pw.cu) in its own process.

    python tools/check_plane_wave.py [--bench]

Compares form 2 with the numpy oracle and with the default GEMM form (3M) through the engine, on ragged and
multi-tile shapes, both pairing modes, derivative and displacement jobs; `--bench` adds a timing of both forms
at the config-3 and config-4 shapes.  Writes gpurun_out/plane_wave_check.json and exits non-zero on any
mismatch.  tests/test_gpu_parity.py runs it as a subprocess, so a fault in this not-yet-hardware-validated
kernel cannot poison the CUDA context of the other tests.
"""
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
FORM = 3 if "--form3" in sys.argv else 2  # the candidate: plane-wave form (2) or its folded variant (3)


def worst_block_error(got, ref):
    norms = np.sqrt((np.abs(ref) ** 2).sum(axis=(-1, -2)))
    floor = 1e-4 * norms.max()
    w = 0.0
    for a in range(ref.shape[0]):
        for p in range(ref.shape[1]):
            w = max(w, float(np.linalg.norm(got[a, p] - ref[a, p]) / max(norms[a, p], floor)))
    return w


def run_case(latt, Ne, mode, order, moms, sym):
    U_file = orc.synthetic_links(latt + [1], 3)
    V = orc.synthetic_eigvecs(latt + [1], Ne, 3)
    U = orc.links_file_to_spatial(U_file)
    if mode == _capi.MODE_DERIVATIVE:
        ref = (orc.elemental_timeslice_closed_form if order <= 2 else orc.elemental_timeslice)(V, U, latt + [1], order, moms)
    else:
        ref = orc.displacement_timeslice(V, U, latt + [1], order, moms)
    eng = ElementalEngine(latt, Ne, mode, order, moms)
    if sym is not None:
        eng.debug_symmetry(sym)
    eng.set_links(torch.from_numpy(U_file).cuda(), _capi.LINKS_FILE_T)
    eng.set_eigvecs(torch.from_numpy(V).cuda())
    eng.debug_algo(1)
    gemm = eng.calc().cpu().numpy()
    eng.debug_algo(FORM)
    q = eng.query()
    assert q["contraction_form"] == FORM and q["plane_wave_modes"] >= 1, q
    pw = eng.calc().cpu().numpy()
    pw_again = eng.calc().cpu().numpy()
    res = {"form": FORM, "latt": latt, "Ne": Ne, "mode": mode, "order": order, "nmom": len(moms), "sym": sym,
           "modes": q["plane_wave_modes"], "tile": q["plane_wave_tile"],
           "err_vs_oracle": worst_block_error(pw, ref), "err_vs_gemm_form": worst_block_error(pw, gemm),
           "gemm_vs_oracle": worst_block_error(gemm, ref), "deterministic": bool(np.array_equal(pw, pw_again))}
    res["ok"] = bool(res["err_vs_oracle"] < TOL and res["err_vs_gemm_form"] < TOL and res["deterministic"])
    eng.close()
    return res


def time_forms(latt, Ne, nabla, nmom, reps=3):
    moms = orc.momentum_set(nmom)
    eng = ElementalEngine(latt, Ne, _capi.MODE_DERIVATIVE, nabla, moms)
    g = torch.Generator(device="cuda").manual_seed(1)
    V = Ne * latt[0] * latt[1] * latt[2] * 3
    v = torch.view_as_complex(torch.randn((V, 2), generator=g, device="cuda", dtype=torch.float32)).reshape(
        Ne, latt[2], latt[1], latt[0], 3)
    U_file = torch.from_numpy(orc.synthetic_links(latt + [1], 1)).cuda()
    eng.set_links(U_file, _capi.LINKS_FILE_T)
    eng.set_eigvecs(v)
    out = {}
    results = {}
    for algo in (1, FORM):
        eng.debug_algo(algo)
        res = eng.calc()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            eng.calc(res)
        e1.record()
        torch.cuda.synchronize()
        out[f"form{algo}_ms_per_timeslice"] = e0.elapsed_time(e1) / reps
        results[algo] = res.cpu().numpy()
    out[f"err_form{FORM}_vs_form1"] = worst_block_error(results[FORM], results[1])
    out["shape"] = {"latt": latt, "Ne": Ne, "num_nabla": nabla, "nmom": nmom}
    eng.close()
    return out


def main():
    assert torch.cuda.is_available(), "needs a CUDA device"
    D, X = _capi.MODE_DERIVATIVE, _capi.MODE_DISPLACEMENT
    cases = [
        ([4, 4, 4], 8, D, 0, orc.momentum_set(7), None),
        ([3, 5, 2], 5, D, 1, orc.momentum_set(7), None),       # plane of 15 sites: ragged 4-site groups and stages
        ([4, 6, 8], 30, D, 2, orc.momentum_set(9), None),      # 2 x 1 tiles, pairing by cost
        ([4, 6, 8], 35, D, 2, orc.momentum_set(33), 1),        # 3 x 2 tiles, 13 modes, Hermitian pairing + half set
        ([4, 6, 8], 19, D, 2, orc.momentum_set(33), 0),        # direct pairs: multi-segment jobs with signs
        ([6, 4, 2], 21, D, 2, [(0, 0, 1), (1, 2, 0), (3, -1, 2), (0, -2, 1)], None),  # non-closed, larger momenta
        ([2, 2, 2], 1, D, 2, orc.momentum_set(7), None),      # planes of 4 sites (form 3: back run before the plane)
        ([3, 3, 2], 9, D, 2, orc.momentum_set(33), None),     # odd planes of 9 sites (form 3: self-paired middle site)
        ([4, 4, 6], 13, D, 3, orc.momentum_set(7), None),
        ([4, 6, 8], 12, X, 3, orc.momentum_set(9), None),
        ([5, 3, 7], 110, D, 1, orc.momentum_set(9), None),     # 7 x 4 tiles
    ]
    report = {"cases": [], "ok": True}
    for c in cases:
        t0 = time.time()
        r = run_case(*c)
        r["seconds"] = time.time() - t0
        report["cases"].append(r)
        report["ok"] = report["ok"] and r["ok"]
        print(json.dumps(r), flush=True)
    # every instantiated tile shape (the default pick above is by FP64-pipe time): multi-tile, mirror tiles, partial f-tiles
    for tile in ("24", "25", "17"):
        os.environ["EDK_PW_TILE"] = tile
        t0 = time.time()
        r = run_case([4, 4, 4], 70, D, 1, orc.momentum_set(7), None)
        r["seconds"] = time.time() - t0
        r["ok"] = bool(r["ok"] and r["tile"] == int(tile))
        report["cases"].append(r)
        report["ok"] = report["ok"] and r["ok"]
        print(json.dumps(r), flush=True)
    os.environ.pop("EDK_PW_TILE", None)
    if "--bench" in sys.argv and report["ok"]:
        report["timing"] = [time_forms([24, 24, 24], 100, 2, 33), time_forms([32, 32, 32], 200, 2, 33, reps=2)]
        print(json.dumps(report["timing"]), flush=True)
    os.makedirs(os.path.join(REPO, "gpurun_out"), exist_ok=True)
    with open(os.path.join(REPO, "gpurun_out", f"plane_wave_check_form{FORM}.json"), "w") as f:
        json.dump(report, f, indent=1)
    print(f"plane-wave form {FORM}:", "OK" if report["ok"] else "MISMATCH")
    return 0 if report["ok"] else 1


if __name__ == "__main__":
    sys.exit(main())
