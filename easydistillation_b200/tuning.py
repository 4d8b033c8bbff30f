"""Choice of the contraction form on THIS GPU, made by measurement in a separate process.

The library has three forms of the momentum-phased contraction:
  form 1  GEMM form, 3M arithmetic (`gram_tma_kernel`) - the default, validated on the B200 against the oracle;
  form 2  plane-wave factorised form (`gram_pw_kernel` + `pw_zfold_kernel`) - several times fewer FP64-pipe
          slots, newer;
  form 3  form 2 with centre-symmetric site pairs folded (`gram_pwf_kernel`) - half the DMMAs per site, newest.
`select_contraction` starts one child interpreter per candidate (forms 2 and 3; a fault in one cannot take the
other down) that, on the given device,
  (1) runs the candidate and form 1 on a few ragged shapes and on the caller's shape through the engine and compares
      them block by block (Frobenius, 1e-10 as everywhere in this package),
  (2) runs the candidate through the public class API (`ElementalGenerator.calc_range`, host arrays in, numpy out)
      and compares it with the engine's form-1 result,
  (3) times both at the caller's shape with CUDA events,
and reports a decision: the fastest candidate whose every comparison passed, if it beats form 1.  A crash, a trap
or a timeout of a child simply removes that candidate.  `apply` makes the decision the default of every handle created afterwards in
this process (the C library reads EDK_GRAM_ALGO in edk_create).  Nothing here uses a CPU implementation.

    python -m easydistillation_b200.tuning --latt 48 48 48 --Ne 200 --num-nabla 2 --momenta 33
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time

TOL = 1e-10
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worst_block_error(got, ref):
    import numpy as np

    norms = np.sqrt((np.abs(ref) ** 2).sum(axis=(-1, -2)))
    floor = 1e-4 * norms.max()
    w = 0.0
    for a in range(ref.shape[0]):
        for p in range(ref.shape[1]):
            w = max(w, float(np.linalg.norm(got[a, p] - ref[a, p]) / max(norms[a, p], floor)))
    return w


def momentum_set(count):
    """First `count` integer triples ordered by (|p|^2, p): 33 = all |p|^2 <= 4."""
    r = range(-3, 4)
    allp = sorted(((px * px + py * py + pz * pz, (px, py, pz)) for px in r for py in r for pz in r))
    return [p for _, p in allp[:count]]


def _inputs(torch, dev, latt3, Ne, seed):
    Lx, Ly, Lz = latt3
    V = Lx * Ly * Lz
    g = torch.Generator(device=dev)
    g.manual_seed(977 + seed)
    a = torch.randn((V, 4, 3, 3), dtype=torch.complex128, device=dev, generator=g)
    q, r = torch.linalg.qr(a)  # unitary links keep the derived fields at the eigenvectors' scale
    U = q.contiguous()
    v = torch.randn((Ne, V, 3), dtype=torch.complex64, device=dev, generator=g)
    v = v / torch.linalg.vector_norm(v.reshape(Ne, -1), dim=1)[:, None, None]
    return U, v.contiguous()


def _engine_forms(torch, dev, latt3, Ne, mode, order, moms, seed, reps, cand=2):
    """(form-1 result, candidate's result, ms of form 1, ms of the candidate, inputs) through the engine."""
    from .engine import ElementalEngine
    from . import _capi

    U, v = _inputs(torch, dev, latt3, Ne, seed)
    out, ms = {}, {}
    for form in (1, cand):  # one handle at a time: the big shapes need the memory
        eng = ElementalEngine(latt3, Ne, mode, order, moms, device=dev.index)
        eng.debug_algo(form)
        assert eng.query()["contraction_form"] == form
        eng.set_links(U, _capi.LINKS_FILE_T)
        eng.set_eigvecs(v)
        res = eng.calc()
        torch.cuda.synchronize(dev)
        if reps:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                eng.calc(res)
            e1.record()
            torch.cuda.synchronize(dev)
            ms[form] = e0.elapsed_time(e1) / reps
        out[form] = res.cpu().numpy()
        eng.close()
        del eng, res
        torch.cuda.empty_cache()
    return out[1], out[cand], ms.get(1), ms.get(cand), (U, v)


def _child(args):
    import numpy as np
    import torch

    from . import _capi

    dev = torch.device("cuda", args.device)
    torch.cuda.set_device(dev)
    D, X = _capi.MODE_DERIVATIVE, _capi.MODE_DISPLACEMENT
    report = {"ok": True, "cases": []}

    def record(tag, err):
        report["cases"].append({"case": tag, "err": err})
        if not (err < TOL):
            report["ok"] = False

    small = [([3, 5, 2], 5, D, 1, momentum_set(7)), ([4, 6, 8], 35, D, 2, momentum_set(33)), ([2, 2, 3], 3, D, 2, momentum_set(7)),
             ([6, 4, 2], 21, D, 2, [(0, 0, 1), (1, 2, 0), (3, -1, 2), (0, -2, 1)]), ([4, 6, 8], 12, X, 3, momentum_set(9))]
    for i, (latt3, Ne, mode, order, moms) in enumerate(small):
        r1, r2, _, _, _ = _engine_forms(torch, dev, latt3, Ne, mode, order, moms, i, 0, args.form)
        record(f"{latt3} Ne={Ne} mode={mode} order={order} nmom={len(moms)}", _worst_block_error(r2, r1))
    # the caller's shape: parity, timing, and the public class API on form 2
    latt3, Ne, mode, order = list(args.latt), args.Ne, args.mode, args.order
    moms = [tuple(m) for m in json.loads(args.momenta)]
    r1, r2, ms1, ms2, (U, v) = _engine_forms(torch, dev, latt3, Ne, mode, order, moms, 100, args.reps, args.form)
    record("caller's shape, engine", _worst_block_error(r2, r1))
    report["form1_ms"], report["form2_ms"] = ms1, ms2
    del r2
    if report["ok"]:
        import easydistillation_b200 as edb

        os.environ["EDK_GRAM_ALGO"] = str(args.form)
        Lx, Ly, Lz = latt3
        U_host = U.cpu().numpy().reshape(1, Lz, Ly, Lx, 4, 3, 3)
        U_host = np.concatenate([U_host, U_host])
        V_host = v.cpu().numpy().reshape(1, Ne, Lz, Ly, Lx, 3)
        V_host = np.concatenate([V_host, V_host])
        del U, v
        torch.cuda.empty_cache()
        cls = edb.ElementalGenerator if mode == D else edb.DisplacementElementalGenerator
        gen = cls([Lx, Ly, Lz, 2], edb.GaugeFieldHostmem(U_host), edb.EigenvectorHostmem(V_host), order, moms, device=args.device)
        gen.load("tune")
        assert gen._engine.query()["contraction_form"] == args.form
        res = gen.calc_range(0, 2)
        record("caller's shape, calc_range t=0", _worst_block_error(np.asarray(res[0]), r1))
        record("caller's shape, calc_range t=1", _worst_block_error(np.asarray(res[1]), r1))
    print("EDK_TUNING " + json.dumps(report), flush=True)
    return 0


def _run_candidate(form, tile, latt3, Ne, mode, order, momentum_list, device, reps, timeout):
    """One child process for one candidate (form, tile); returns its report dict or {"ok": False, "reason": ...}."""
    cmd = [sys.executable, "-m", "easydistillation_b200.tuning", "--child", "--form", str(int(form)), "--device", str(int(device)),
           "--latt", *[str(int(v)) for v in latt3], "--Ne", str(int(Ne)), "--mode", str(int(mode)), "--order", str(int(order)),
           "--reps", str(int(reps)), "--momenta", json.dumps([list(map(int, m)) for m in momentum_list])]
    env = dict(os.environ)
    for k in ("EDK_GRAM_ALGO", "EDK_PW_TILE", "RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT"):
        env.pop(k, None)  # the child is a plain single-GPU process that chooses the form itself
    if tile:
        env["EDK_PW_TILE"] = str(tile)
    env["PYTHONPATH"] = REPO + os.pathsep + env.get("PYTHONPATH", "")
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env, cwd=REPO)
    except subprocess.TimeoutExpired:
        return {"ok": False, "reason": f"tuning child timed out after {timeout:.0f} s"}
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("EDK_TUNING ")]
    if r.returncode != 0 or not line:
        return {"ok": False, "reason": f"tuning child failed (exit {r.returncode}): " + (r.stderr or r.stdout)[-400:].replace("\n", " | ")}
    rep = json.loads(line[-1][len("EDK_TUNING "):])
    if not rep["ok"]:
        bad = [c for c in rep["cases"] if not (c["err"] < TOL)]
        rep["reason"] = f"differs from form 1: {bad[:2]}"
    elif not (rep.get("form1_ms") and rep.get("form2_ms")):
        rep["ok"], rep["reason"] = False, "no timing"
    return rep


# candidates in the order they are tried: (form, EDK_PW_TILE or None = the library's own pick)
CANDIDATES = ((2, None), (3, None))  # pass e.g. ((3, "24"), (3, "25")) to A/B the tile shapes of one form


def select_contraction(latt3, Ne, mode, order, momentum_list, device: int = 0, reps: int = 2, timeout: float = 900.0,
                       candidates=CANDIDATES) -> dict:
    """Decide between the GEMM form (1) and the plane-wave forms (2, 3) for this shape on `device`; see the module text.
    `timeout` bounds the whole selection; the result carries one entry per candidate under "candidates"."""
    t0 = time.time()
    decision = {"form": 1, "tile": None, "validated": False, "form1_ms": None, "form2_ms": None, "worst_err": None, "reason": "",
                "candidates": []}
    best = None
    for form, tile in candidates:
        left = timeout - (time.time() - t0)
        if left < 30.0:
            decision["candidates"].append({"form": form, "tile": tile, "ok": False, "reason": "no time left"})
            continue
        rep = _run_candidate(form, tile, latt3, Ne, mode, order, momentum_list, device, reps, left)
        entry = {"form": form, "tile": tile, "ok": bool(rep.get("ok")), "ms": rep.get("form2_ms"), "form1_ms": rep.get("form1_ms"),
                 "worst_err": max((c["err"] for c in rep.get("cases", [])), default=None), "cases": len(rep.get("cases", [])),
                 "reason": rep.get("reason", "")}
        decision["candidates"].append(entry)
        if entry["ok"]:
            decision["form1_ms"] = entry["form1_ms"]
            if best is None or entry["ms"] < best["ms"]:
                best = entry
    decision["seconds"] = time.time() - t0
    if best is None:
        decision["reason"] = "no plane-wave candidate validated: " + "; ".join(f"form {c['form']}: {c['reason']}" for c in decision["candidates"])
        return decision
    decision.update(validated=True, form2_ms=best["ms"], worst_err=best["worst_err"], cases=best["cases"])
    if best["ms"] < best["form1_ms"]:
        decision.update(form=best["form"], tile=best["tile"])
        decision["reason"] = (f"form {best['form']}" + (f" (tile {best['tile']})" if best["tile"] else "") +
                              f" agrees with form 1 to {best['worst_err']:.1e} on {best['cases']} comparisons and is "
                              f"{best['form1_ms'] / best['ms']:.2f}x faster per timeslice at this shape")
    else:
        decision["reason"] = "the plane-wave forms are validated but not faster at this shape"
    return decision


def apply(decision: dict) -> int:
    """Make the decision the default contraction form of every handle created from now on in this process."""
    os.environ["EDK_GRAM_ALGO"] = str(int(decision["form"]))
    if decision.get("tile"):
        os.environ["EDK_PW_TILE"] = str(decision["tile"])
    else:
        os.environ.pop("EDK_PW_TILE", None)
    return int(decision["form"])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--child", action="store_true")
    ap.add_argument("--form", type=int, default=2, choices=[2, 3], help="--child: the candidate form")
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--latt", type=int, nargs=3, default=[24, 24, 24])
    ap.add_argument("--Ne", type=int, default=100)
    ap.add_argument("--mode", type=int, default=0)
    ap.add_argument("--order", "--num-nabla", type=int, default=2, dest="order")
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--momenta", default="33", help="a count (first N of the |p|^2-ordered set) or a JSON list of triples")
    args = ap.parse_args()
    if args.momenta.strip().isdigit():
        args.momenta = json.dumps([list(m) for m in momentum_set(int(args.momenta))])
    if args.child:
        return _child(args)
    d = select_contraction(args.latt, args.Ne, args.mode, args.order, json.loads(args.momenta), args.device, args.reps)
    print(json.dumps(d, indent=1))
    return 0


if __name__ == "__main__":
    sys.exit(main())
