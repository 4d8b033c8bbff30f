"""Choice of the contraction form on THIS GPU, made by measurement in a separate process.

The library has two forms of the momentum-phased contraction:
  form 1  GEMM form, 3M arithmetic (`gram_tma_kernel`) - the default, validated on the B200 against the oracle;
  form 2  plane-wave factorised form (`gram_pw_kernel` + `pw_zfold_kernel`) - several times fewer FP64-pipe
          slots, newer.
`select_contraction` starts a child interpreter that, on the given device,
  (1) runs both forms on a few ragged shapes and on the caller's shape through the engine and compares form 2
      with form 1 block by block (Frobenius, 1e-10 as everywhere in this package),
  (2) runs form 2 through the public class API (`ElementalGenerator.calc_range`, host arrays in, numpy out) and
      compares it with the engine's form-1 result,
  (3) times both forms at the caller's shape with CUDA events,
and reports a decision: form 2 only if every comparison passed AND it was faster.  A crash, a trap or a timeout
of the child simply means form 1.  `apply` makes the decision the default of every handle created afterwards in
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


def _engine_forms(torch, dev, latt3, Ne, mode, order, moms, seed, reps):
    """(form-1 result, form-2 result, ms of form 1, ms of form 2, inputs) through the engine."""
    from .engine import ElementalEngine
    from . import _capi

    U, v = _inputs(torch, dev, latt3, Ne, seed)
    out, ms = {}, {}
    for form in (1, 2):  # one handle at a time: the big shapes need the memory
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
    return out[1], out[2], ms.get(1), ms.get(2), (U, v)


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

    small = [([3, 5, 2], 5, D, 1, momentum_set(7)), ([4, 6, 8], 35, D, 2, momentum_set(33)),
             ([6, 4, 2], 21, D, 2, [(0, 0, 1), (1, 2, 0), (3, -1, 2), (0, -2, 1)]), ([4, 6, 8], 12, X, 3, momentum_set(9)),
             ([5, 3, 7], 90, D, 1, momentum_set(9))]
    for i, (latt3, Ne, mode, order, moms) in enumerate(small):
        r1, r2, _, _, _ = _engine_forms(torch, dev, latt3, Ne, mode, order, moms, i, 0)
        record(f"{latt3} Ne={Ne} mode={mode} order={order} nmom={len(moms)}", _worst_block_error(r2, r1))
    # the caller's shape: parity, timing, and the public class API on form 2
    latt3, Ne, mode, order = list(args.latt), args.Ne, args.mode, args.order
    moms = [tuple(m) for m in json.loads(args.momenta)]
    r1, r2, ms1, ms2, (U, v) = _engine_forms(torch, dev, latt3, Ne, mode, order, moms, 100, args.reps)
    record("caller's shape, engine", _worst_block_error(r2, r1))
    report["form1_ms"], report["form2_ms"] = ms1, ms2
    del r2
    if report["ok"]:
        import easydistillation_b200 as edb

        os.environ["EDK_GRAM_ALGO"] = "2"
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
        assert gen._engine.query()["contraction_form"] == 2
        res = gen.calc_range(0, 2)
        record("caller's shape, calc_range t=0", _worst_block_error(np.asarray(res[0]), r1))
        record("caller's shape, calc_range t=1", _worst_block_error(np.asarray(res[1]), r1))
    print("EDK_TUNING " + json.dumps(report), flush=True)
    return 0


def select_contraction(latt3, Ne, mode, order, momentum_list, device: int = 0, reps: int = 2, timeout: float = 900.0) -> dict:
    """Decide between the GEMM form (1) and the plane-wave form (2) for this shape on `device`; see the module text."""
    t0 = time.time()
    cmd = [sys.executable, "-m", "easydistillation_b200.tuning", "--child", "--device", str(int(device)), "--latt",
           *[str(int(v)) for v in latt3], "--Ne", str(int(Ne)), "--mode", str(int(mode)), "--order", str(int(order)),
           "--reps", str(int(reps)), "--momenta", json.dumps([list(map(int, m)) for m in momentum_list])]
    env = dict(os.environ)
    env.pop("EDK_GRAM_ALGO", None)
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT"):  # the child is a plain single-GPU process
        env.pop(k, None)
    env["PYTHONPATH"] = REPO + os.pathsep + env.get("PYTHONPATH", "")
    decision = {"form": 1, "validated": False, "form1_ms": None, "form2_ms": None, "worst_err": None, "reason": ""}
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env, cwd=REPO)
    except subprocess.TimeoutExpired:
        decision["reason"] = f"tuning child timed out after {timeout:.0f} s"
        return decision
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("EDK_TUNING ")]
    if r.returncode != 0 or not line:
        decision["reason"] = f"tuning child failed (exit {r.returncode}): " + (r.stderr or r.stdout)[-400:].replace("\n", " | ")
        return decision
    rep = json.loads(line[-1][len("EDK_TUNING "):])
    decision.update(form1_ms=rep.get("form1_ms"), form2_ms=rep.get("form2_ms"), validated=bool(rep["ok"]),
                    worst_err=max((c["err"] for c in rep["cases"]), default=None), cases=len(rep["cases"]),
                    seconds=time.time() - t0)
    if not rep["ok"]:
        bad = [c for c in rep["cases"] if not (c["err"] < TOL)]
        decision["reason"] = f"form 2 differs from form 1: {bad[:2]}"
    elif not (rep.get("form1_ms") and rep.get("form2_ms")):
        decision["reason"] = "no timing"
    elif rep["form2_ms"] < rep["form1_ms"]:
        decision["form"] = 2
        decision["reason"] = (f"form 2 agrees with form 1 to {decision['worst_err']:.1e} on {len(rep['cases'])} comparisons and is "
                              f"{rep['form1_ms'] / rep['form2_ms']:.2f}x faster per timeslice at this shape")
    else:
        decision["reason"] = "form 2 is validated but not faster at this shape"
    return decision


def apply(decision: dict) -> int:
    """Make the decision the default contraction form of every handle created from now on in this process."""
    os.environ["EDK_GRAM_ALGO"] = str(int(decision["form"]))
    return int(decision["form"])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--child", action="store_true")
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
