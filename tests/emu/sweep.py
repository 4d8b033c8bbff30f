"""Randomised sweep of the plane-wave contraction forms on the host emulator (TEST INFRASTRUCTURE ONLY, not part of the
pytest suite: run it by hand after touching csrc/edk_gram_pw.cu or the plane-wave part of csrc/edk_api.cu).

    python tests/emu/sweep.py [--seconds 600] [--seed 1] [--tiles | --separable]

Builds libedk_emu.so like tests/test_emu_library.py, then draws lattice extents, Ne, generator mode, order, momentum
lists (random triples incl. negative and larger-than-lattice components, or prefixes of the |p|^2-ordered set),
pairing mode and form (2 or 3), runs the C ABI and compares with the oracle (1e-10, block-wise).  `--tiles` instead
sweeps Ne = 17..120 on tiny lattices over the three tile shapes (EDK_PW_TILE = 24 / 25 / 17): multi-tile runs, mirror
tiles of self pairs, partial f-tiles, padding-only warps.  `--separable` sweeps the separable form (4, the planned
default on the production lattices) over Lx, Ne (all three warp grids), mode structures and both launch modes.
"""
import argparse
import os
import pathlib
import random
import sys
import tempfile
import time

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [REPO, os.path.join(REPO, "tests")]

import test_emu_library as T  # noqa: E402
from oracle import elemental_oracle as orc  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=600.0)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--tiles", action="store_true")
    ap.add_argument("--separable", action="store_true", help="sweep the separable form (4): Lx in {8, 12, 16, 24, 32}, all tile shapes")
    args = ap.parse_args()
    rnd = random.Random(args.seed)
    with tempfile.TemporaryDirectory() as tmp:
        lib = T.load_emulator_library(T.build_emulator_library(pathlib.Path(tmp)))
        t_end, n, worst = time.time() + args.seconds, 0, 0.0
        while time.time() < t_end:
            sym = None
            form = rnd.choice([2, 3, 3])
            if args.separable:
                # the planned default wherever plan_sep accepts the lattice and the list: 4 / 6 / 8 pairs per stage, one to
                # four stages per row, 5 / 9 / 13 modes, Ne across the three warp grids, both launch modes
                latt3 = (rnd.choice([8, 8, 12, 16, 24, 32]), rnd.randint(1, 5), rnd.randint(1, 3))
                Ne = rnd.choice([1, 3, 8, 9, 16, 31, 32, 33, 40, 47, 64, 70])
                mode = rnd.choice([T.D, T.D, T.X])
                order = rnd.randint(0, 2)
                kind = rnd.random()
                if kind < 0.5:
                    moms = orc.momentum_set(rnd.choice([1, 7, 9, 19, 27, 33]))
                else:  # any list inside the instantiated structures: |px|, |py| <= 2, px^2 + py^2 <= 4, pz free
                    moms = []
                    while len(moms) < rnd.randint(3, 8):
                        px, py = rnd.randint(-2, 2), rnd.randint(-2, 2)
                        if px * px + py * py <= 4:
                            moms.append((px, py, rnd.randint(-3, 3)))
                sym = rnd.choice([None, None, 0, 1]) if mode == T.D and order else None
                os.environ["EDK_SEP_LAUNCHES"] = rnd.choice(["1", "3"])
                form = 4
                nfield = {0: 1, 1: 4, 2: 13}[order] if mode == T.D else 2
                if latt3[0] * latt3[1] * latt3[2] * max(Ne, 32) ** 2 * nfield * nfield > 6e7:
                    continue  # keep a case within a few seconds of emulation
            elif args.tiles:
                latt3 = rnd.choice([(2, 2, 1), (4, 2, 1), (3, 3, 1), (4, 4, 1), (3, 5, 1)])
                Ne, mode, order = rnd.randint(17, 120), T.D, rnd.choice([0, 1])
                moms = rnd.choice([[(0, 0, 0)], [(0, 0, 0), (1, 0, 0), (-1, 0, 0)], [(1, 1, 0), (0, -1, 0)]])
                os.environ["EDK_PW_TILE"] = rnd.choice(["24", "25", "17"])
            else:
                latt3 = (rnd.randint(1, 6), rnd.randint(1, 6), rnd.randint(1, 4))
                Ne = rnd.choice([1, 2, 3, 5, 8, 9, 17, 20, 33])
                mode = rnd.choice([T.D, T.D, T.D, T.X])
                order = rnd.randint(0, 2) if mode == T.D else rnd.randint(0, 3)
                moms = [(rnd.randint(-3, 3), rnd.randint(-3, 3), rnd.randint(-2, 2)) for _ in range(rnd.randint(1, 6))]
                if rnd.random() < 0.3:
                    moms = orc.momentum_set(rnd.choice([1, 7, 9, 19]))
                sym = rnd.choice([None, None, 0, 1]) if mode == T.D else None
                if latt3[0] * latt3[1] * latt3[2] * Ne * Ne * len(moms) * (1 + order) ** 2 > 4e6:
                    continue  # keep a case within a few seconds of emulation
            U_file, V, ref = T.inputs_and_reference(latt3, Ne, mode, order, moms, seed=n)
            h = T.Handle(lib, latt3, Ne, mode, order, moms)
            if sym is not None:
                h.check(lib.edk_debug_symmetry(h.h, sym), "edk_debug_symmetry")
            h.check(lib.edk_debug_algo(h.h, form), "edk_debug_algo")
            if form == 4 and h.query(10) != 4:
                print(f"FAIL latt={latt3} moms={moms}: the separable form was not accepted")
                return 1
            h.set_inputs(U_file, V)
            err = T.worst_block_error(h.calc(), ref)
            tile = h.query(12)
            h.close()
            n, worst = n + 1, max(worst, err)
            if not err < 1e-10:
                print(f"FAIL latt={latt3} Ne={Ne} mode={mode} order={order} moms={moms} sym={sym} form={form} tile={tile}: {err:.3e}")
                return 1
            if n % 20 == 0:
                print(f"{n} cases ok, worst block error {worst:.2e}", flush=True)
    print(f"DONE: {n} cases, worst block error {worst:.2e}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
