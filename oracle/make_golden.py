"""Generate tests/golden/*.npz by running the UNMODIFIED reference in this container.

TEST INFRASTRUCTURE ONLY.  Run once from the repo root in the build container
(where /root/reference exists); the resulting small fixtures are committed so
that the GPU box, which has no /root/reference, can check against them.

    python oracle/make_golden.py                 # everything
    python oracle/make_golden.py NAME [NAME ...] # only the named elemental cases

How the reference is made importable without editing it:
  * oracle/shims/opt_einsum  -> numpy.einsum(optimize=True)
  * oracle/shims/mpi4py      -> single-process COMM_WORLD
  * numpy.lib.format._read_array_header alias (numpy >= 2 moved it)
Inputs go through the reference's own stock loaders (`GaugeFieldBinary`,
`EigenvectorNpy`), the generators are driven exactly like
tests/test_elemental.py:41-44 and tests/test_displacement_elemental.py:40-43.
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
REF = os.environ.get("EDK_REFERENCE_ROOT", "/root/reference")


def import_reference():
    sys.path.insert(0, os.path.join(HERE, "shims"))
    sys.path.insert(1, REF)
    import numpy.lib.format as fmt

    if not hasattr(fmt, "_read_array_header"):
        import numpy.lib._format_impl as impl

        fmt._read_array_header = impl._read_array_header
    import lattice

    lattice.set_backend("numpy")
    return lattice


def run_reference(lattice, latt_size, Ne, U_file, V_file, *, num_nabla=None, distance=None, momentum_list, dilution=None,
                  gauge_ops=(), return_links=False):
    """U_file [Lt,Lz,Ly,Lx,4,3,3] c16, V_file [Lt,Ne,Lz,Ly,Lx,3] c16 -> [Lt,Nop,Nmom,Ne,Ne]."""
    Lx, Ly, Lz, Lt = latt_size
    with tempfile.TemporaryDirectory() as tmp:
        prefix = tmp + "/"
        U_file.astype("<c16").tofile(prefix + "cfg.dat")
        np.save(prefix + "cfg.eigenvector.npy", V_file.astype("<c16"))
        gauge = lattice.preset.GaugeFieldBinary(prefix, ".dat", [Lt, Lz, Ly, Lx, 4, 3, 3], "<c16")
        evec = lattice.EigenvectorNpy(prefix, ".eigenvector.npy", [Lt, Ne, Lz, Ly, Lx, 3], Ne)
        if distance is None:
            if dilution is None:
                gen = lattice.ElementalGenerator(latt_size, gauge, evec, num_nabla, momentum_list)
            else:
                gen = lattice.ElementalGenerator(latt_size, gauge, evec, num_nabla, momentum_list, dilution, True)
        else:
            gen = lattice.DisplacementElementalGenerator(latt_size, gauge, evec, distance, momentum_list)
        gen.load("cfg")
        for op in gauge_ops:  # the gauge preprocessing methods of the reference classes
            if op[0] == "stout":
                gen.stout_smear(op[1], op[2])
            else:
                gen.project_SU3()
        links = np.array(gen._U, copy=True) if return_links else None
        out = []
        for t in range(Lt):
            out.append(np.array(gen.calc(t), copy=True))
    return (np.stack(out), links) if return_links else np.stack(out)


def main():
    sys.path.insert(0, REPO)
    from oracle import elemental_oracle as orc

    lattice = import_reference()
    outdir = os.path.join(REPO, "tests", "golden")
    os.makedirs(outdir, exist_ok=True)

    cases = {
        # name: (latt_size [Lx,Ly,Lz,Lt], Ne, link kind, kwargs)
        "deriv_weak_4x4x4x2": ([4, 4, 4, 2], 8, "weak", dict(num_nabla=2, momentum_list=[(0, 0, 0), (0, 0, 1), (0, 1, 1), (1, 1, 1), (0, 0, 2), (0, 1, 2), (1, 1, 2)])),
        "deriv_random_4x6x8x1": ([4, 6, 8, 1], 6, "random", dict(num_nabla=2, momentum_list=[(0, 0, 0), (1, 0, 0), (0, -1, 2), (-1, 1, -1), (3, 2, 1)])),
        "deriv_n1_random_6x4x2x1": ([6, 4, 2, 1], 5, "random", dict(num_nabla=1, momentum_list=[(0, 0, 0), (0, 1, 0), (-2, 0, 1)])),
        "deriv_blend_4x4x4x1": ([4, 4, 4, 1], 6, "random", dict(num_nabla=1, momentum_list=[(0, 0, 0), (1, 0, -1)], dilution=([10, 7], [4, 2]))),
        "disp_weak_4x4x4x2": ([4, 4, 4, 2], 8, "weak", dict(distance=8, momentum_list=[(0, 0, 0), (0, 0, 1), (0, 1, 1), (1, 1, 1), (0, 1, 2), (1, 1, 2)])),
        "disp_random_4x6x8x1": ([4, 6, 8, 1], 6, "random", dict(distance=3, momentum_list=[(0, 0, 0), (1, 0, 0), (0, -1, 2)])),
        # the two ends of num_nabla: no derivative on odd y, z extents (the reference needs an even Lx: phase.py:25), and all 40 third-order operators
        "deriv_n0_random_6x3x5x1": ([6, 3, 5, 1], 4, "random", dict(num_nabla=0, momentum_list=[(0, 0, 0), (1, -2, 0), (0, 0, 3)])),
        "deriv_n3_random_4x4x6x1": ([4, 4, 6, 1], 4, "random", dict(num_nabla=3, momentum_list=[(0, 0, 0), (1, 0, 0), (-1, 0, 0), (0, 1, -2)])),
        # config 1 at its own shape: the inputs of the reference's tests/test_elemental.py:15-24 and
        # tests/test_displacement_elemental.py:15-23 (4^3 x 8, Ne = 20, their momentum lists) on a synthetic weak field
        # (the stored weak_field.* files are git-LFS pointers); all 8 timeslices of input, results of three / two of them
        "config1_deriv_weak_4x4x4x8": ([4, 4, 4, 8], 20, "weak", dict(num_nabla=2, momentum_list=[(0, 0, 0), (0, 0, 1), (0, 1, 1), (1, 1, 1), (0, 0, 2), (0, 1, 2), (1, 1, 2)]), (0, 3, 7)),
        "config1_disp_weak_4x4x4x8": ([4, 4, 4, 8], 20, "weak", dict(distance=8, momentum_list=[(0, 0, 0), (0, 0, 1), (0, 1, 1), (1, 1, 1), (0, 1, 2), (1, 1, 2)]), (0, 5)),
        # lattices the separable contraction covers (Lx even, >= 8): 4 / 6 / 8 site pairs per stage, 13 / 9 / 9 xy-modes
        "deriv_sep_8x4x6x1": ([8, 4, 6, 1], 6, "random", dict(num_nabla=2, momentum_list=orc.momentum_set(33))),
        "deriv_sep_12x4x2x1": ([12, 4, 2, 1], 5, "random", dict(num_nabla=1, momentum_list=orc.momentum_set(9))),
        "disp_sep_16x2x4x1": ([16, 2, 4, 1], 5, "random", dict(distance=2, momentum_list=orc.momentum_set(19))),
    }
    only = set(sys.argv[1:])
    for name, case in cases.items():
        latt, Ne, kind, kw = case[:4]
        keep = case[4] if len(case) > 4 else None  # timeslices whose results are stored (all by default)
        if only and name not in only:
            continue
        Lx, Ly, Lz, Lt = latt
        U_file = np.stack([orc.synthetic_links(latt, t, kind) for t in range(Lt)])
        V_file = np.stack([orc.synthetic_eigvecs(latt, Ne, t) for t in range(Lt)])
        ref = run_reference(lattice, latt, Ne, U_file, V_file, **kw)
        meta = dict(latt_size=np.array(latt), Ne=Ne, momentum_list=np.array(kw["momentum_list"]))
        if keep is not None:
            ref = ref[list(keep)]
            meta["timeslices"] = np.array(keep)
        if "num_nabla" in kw:
            meta["num_nabla"] = kw["num_nabla"]
        if "distance" in kw:
            meta["distance"] = kw["distance"]
        if kw.get("dilution") is not None:
            meta["dilution_tot"] = np.array(kw["dilution"][0])
            meta["dilution_used"] = np.array(kw["dilution"][1])
        # eigenvectors are stored as complex64: the reference rounds them through
        # complex64 before use, so nothing the path sees is lost
        np.savez_compressed(
            os.path.join(outdir, name + ".npz"),
            U=U_file.astype("<c16"),
            V=V_file.astype("<c8"),
            E=ref.astype("<c16"),
            **meta,
        )
        print(f"{name}: U{U_file.shape} V{V_file.shape} -> E{ref.shape}  |E|={np.linalg.norm(ref):.6e}")

    if only:
        return

    # gauge preprocessing: stout smearing and the unitarity projection, on a weak
    # field and on a slightly non-unitary one; the processed links are stored next to the elementals
    rng = np.random.default_rng(orc.SEED0 + 99)
    latt, Ne = [4, 4, 6, 2], 5
    Lx, Ly, Lz, Lt = latt
    moms = [(0, 0, 0), (1, 0, -1)]
    U_file = np.stack([orc.synthetic_links(latt, t, "weak") for t in range(Lt)])
    V_file = np.stack([orc.synthetic_eigvecs(latt, Ne, t) for t in range(Lt)])
    noisy = U_file + 1e-3 * (rng.standard_normal(U_file.shape) + 1j * rng.standard_normal(U_file.shape))
    for name, Uin, ops in (("gauge_stout_4x4x6x2", U_file, [("stout", 3, 0.1)]),
                           ("gauge_project_4x4x6x2", noisy, [("project",)]),
                           ("gauge_project_stout_4x4x6x2", noisy, [("project",), ("stout", 2, 0.125)])):
        E, links = run_reference(lattice, latt, Ne, Uin, V_file, num_nabla=1, momentum_list=moms, gauge_ops=ops,
                                 return_links=True)
        np.savez_compressed(os.path.join(outdir, name + ".npz"), U=Uin.astype("<c16"), V=V_file.astype("<c8"), E=E.astype("<c16"),
                            links=links.astype("<c16"), latt_size=np.array(latt), Ne=Ne, momentum_list=np.array(moms), num_nabla=1,
                            ops=np.array([[1 if o[0] == "stout" else 2, o[1] if len(o) > 1 else 0] for o in ops]),
                            rhos=np.array([o[2] if len(o) > 2 else 0.0 for o in ops]))
        print(f"{name}: links{links.shape} E{E.shape} |E|={np.linalg.norm(E):.6e}")

    # the eigensolver's Laplacian, called exactly as eigenvector.py:243-250 does:
    # F is [Lz*Ly*Lx*Nc, nvec] (vector index fastest), U and U_dag are [3, Lz, Ly, Lx, 3, 3]
    from lattice.generator.eigenvector import _Laplacian

    latt = [4, 6, 8, 1]
    Lx, Ly, Lz, Lt = latt
    U_sp = orc.links_file_to_spatial(orc.synthetic_links(latt, 3))
    F = orc.synthetic_eigvecs(latt, 5, 3)                                    # [e, z, y, x, c]
    F_ref = np.ascontiguousarray(np.moveaxis(F, 0, -1)).reshape(Lz * Ly * Lx * 3, 5)
    LF = _Laplacian(F_ref, U_sp, U_sp.transpose(0, 1, 2, 3, 5, 4).conj(), latt)
    LF = np.moveaxis(LF.reshape(Lz, Ly, Lx, 3, 5), -1, 0)
    np.savez_compressed(os.path.join(outdir, "laplacian_4x6x8.npz"), U_file=orc.synthetic_links(latt, 3), F=F, LF=LF,
                        latt_size=np.array(latt))
    print(f"laplacian_4x6x8: F{F.shape} -> LF{LF.shape} |LF|={np.linalg.norm(LF):.6e}")

    # index-map and phase goldens straight from the reference's insertion module
    from lattice.insertion.derivative import derivative
    from lattice.insertion.phase import MomentumPhase

    tuples = [list(derivative(n)) + [-1] * (3 - len(derivative(n))) for n in range(40)]
    mp = MomentumPhase([4, 6, 8, 2])
    moms = [(0, 0, 0), (1, 0, 0), (0, -1, 2), (3, 2, 1), (-2, 5, -7)]
    np.savez_compressed(
        os.path.join(outdir, "insertion_maps.npz"),
        derivative_tuples=np.array(tuples),
        phase_latt=np.array([4, 6, 8, 2]),
        phase_moms=np.array(moms),
        phases=np.stack([mp.get(p) for p in moms]),
    )
    print("insertion_maps written")


if __name__ == "__main__":
    main()
