"""Generate tests/golden/files_ildg_qdp.npz: a tiny ILDG gauge file and QDP timeslice eigenvector
file (written by easydistillation_b200.fileio), what the UNMODIFIED reference's own readers
(`GaugeFieldIldg`, `EigenvectorTimeSlice`: lattice/filedata/ildg.py, timeslice.py) return for them,
and the elementals the reference computes from those handles.

TEST INFRASTRUCTURE ONLY.  Run once from the repo root in the build container:

    python oracle/make_golden_files.py
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
sys.path.insert(0, REPO)

from oracle import elemental_oracle as orc  # noqa: E402
from oracle.make_golden import import_reference  # noqa: E402


def main():
    lattice = import_reference()
    from easydistillation_b200.fileio import write_ildg, write_qdp_timeslices

    latt = [4, 4, 6, 2]  # Lx, Ly, Lz, Lt
    Lx, Ly, Lz, Lt = latt
    Ne, num_nabla = 6, 1
    moms = [(0, 0, 0), (0, 1, 0), (-1, 0, 2)]
    U = np.stack([orc.synthetic_links(latt, 11 + t) for t in range(Lt)])     # [Lt,Lz,Ly,Lx,4,3,3] c16
    V = np.stack([orc.synthetic_eigvecs(latt, Ne, 20 + t) for t in range(Lt)]).astype("<c8")  # [Lt,Ne,Lz,Ly,Lx,3]
    with tempfile.TemporaryDirectory() as tmp:
        prefix = tmp + "/"
        write_ildg(prefix + "cfg.lime", U)
        write_qdp_timeslices(prefix + "cfg.mod", V.reshape(Lt, Ne, Lz * Ly * Lx, 3), latt)
        lime_bytes = np.fromfile(prefix + "cfg.lime", dtype=np.uint8)
        mod_bytes = np.fromfile(prefix + "cfg.mod", dtype=np.uint8)
        gauge = lattice.preset.GaugeFieldIldg(prefix, ".lime", [Lt, Lz, Ly, Lx, 4, 3, 3])
        evec = lattice.preset.EigenvectorTimeSlice(prefix, ".mod", [Lt, Ne, Lz, Ly, Lx, 3], Ne)
        U_ref = np.asarray(gauge.load("cfg")[:])
        ev = evec.load("cfg")
        V_ref = np.stack([np.stack([np.asarray(ev[t, e]) for e in range(Ne)]) for t in range(Lt)])
        assert U_ref.dtype == np.dtype("<c16") and np.array_equal(U_ref, U), "reference ILDG reader disagrees with the writer"
        assert V_ref.dtype == np.dtype("<c8") and np.array_equal(V_ref, V), "reference QDP reader disagrees with the writer"
        gen = lattice.ElementalGenerator(latt, gauge, evec, num_nabla, moms)
        gen.load("cfg")
        E = np.stack([gen.calc(t).copy() for t in range(Lt)])
    out = os.path.join(REPO, "tests", "golden", "files_ildg_qdp.npz")
    np.savez_compressed(out, lime_bytes=lime_bytes, mod_bytes=mod_bytes, U_ref=U_ref, V_ref=V_ref, E=E,
                        latt_size=np.array(latt), Ne=Ne, num_nabla=num_nabla, momentum_list=np.array(moms))
    print(f"files_ildg_qdp: lime {lime_bytes.size} B, mod {mod_bytes.size} B, E{E.shape} |E|={np.linalg.norm(E):.6e} -> {out}")


if __name__ == "__main__":
    main()
