"""GPU parity: the sm_100a path (through the public classes / the C ABI) against
 (1) golden outputs of the unmodified reference (tests/golden, made by oracle/make_golden.py),
 (2) the numpy oracle on seeded synthetic inputs, including ragged / odd sizes,
 (3) size-independent properties at BASELINE.json sizes.
Tolerance: 1e-10 relative (Frobenius) per (operator, momentum) block, as north_star states."""
import numpy as np
import pytest

from conftest import golden_timeslices, load_golden, reference_weak_field_files, rel_err

pytestmark = pytest.mark.gpu

TOL = 1e-10


@pytest.fixture(scope="module")
def edb():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import easydistillation_b200 as m

    return m


def _orc():
    from oracle import elemental_oracle as orc

    return orc


def _blocks_close(got, ref, tol=TOL, what=""):
    """Per (operator, momentum) block Frobenius error <= tol * block norm.  Blocks that vanish
    analytically (e.g. p_z = 1 second derivatives on an L_z = 2 lattice) are measured against
    1e-4 of the largest block instead of their own rounding-noise norm."""
    assert got.shape == ref.shape, (got.shape, ref.shape)
    norms = np.sqrt((np.abs(ref) ** 2).sum(axis=(-1, -2)))
    floor = 1e-4 * norms.max()
    worst = 0.0
    for a in range(ref.shape[0]):
        for p in range(ref.shape[1]):
            worst = max(worst, float(np.linalg.norm(got[a, p] - ref[a, p]) / max(norms[a, p], floor)))
    assert worst < tol, f"{what}: worst block error {worst:.3e}"
    assert np.max(np.abs(got - ref)) <= tol * np.max(np.abs(ref)), what
    return worst


def _golden_case(name):
    g = load_golden(name)
    latt = [int(v) for v in g["latt_size"]]
    moms = [tuple(int(v) for v in p) for p in g["momentum_list"]]
    return g, latt, moms


def _forms_of(eng):
    """Every contraction form this handle can run, the planned one first: 1 = GEMM form (3M on DMMA), 2 / 3 = plane-wave
    forms, 4 = separable form where the lattice / momentum list allow it (easydistillation_b200._capi.plan_form)."""
    from easydistillation_b200 import _capi

    Lx, Ly, Lz = eng.latt3
    plan = _capi.plan_form((Lx, Ly, Lz), eng.mode, eng.order, [tuple(m) for m in eng.momenta])
    planned = eng.query()["contraction_form"]
    assert planned == plan["form"], (planned, plan)
    forms = [1, 2, 3] + ([4] if plan["separable_available"] else [])
    return [planned] + [f for f in forms if f != planned]


def _all_forms(gen, t, ref, what):
    """calc(t) with the planned form and then with every other form, each against `ref`; leaves the planned form on."""
    eng = gen._engine
    forms = _forms_of(eng)
    worst = {}
    for form in forms:
        eng.debug_algo(form)
        assert eng.query()["contraction_form"] == form
        worst[form] = _blocks_close(np.array(gen.calc(t)), ref, what=f"{what} form {form}")
    eng.debug_algo(-1)
    assert eng.query()["contraction_form"] == forms[0]
    return worst


# ---------------------------------------------------------------------------------------------
# (1) reference golden vectors
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["deriv_weak_4x4x4x2", "deriv_random_4x6x8x1", "deriv_n1_random_6x4x2x1",
                                  "deriv_n0_random_6x3x5x1", "deriv_n3_random_4x4x6x1",
                                  "config1_deriv_weak_4x4x4x8",  # config 1 at its own shape: tests/test_elemental.py:15-24
                                  "deriv_sep_8x4x6x1", "deriv_sep_12x4x2x1"])
def test_derivative_elementals_match_reference_golden(edb, name):
    g, latt, moms = _golden_case(name)
    gen = edb.ElementalGenerator(latt, edb.GaugeFieldHostmem(g["U"]), edb.EigenvectorHostmem(g["V"]),
                                 int(g["num_nabla"]), moms)
    assert gen.num_derivative == g["E"].shape[1] and gen.num_momentum == len(moms) and gen.Ne == int(g["Ne"])
    gen.load("cfg")
    data = np.zeros((latt[3],) + g["E"].shape[1:], "<c16")
    for t in range(latt[3]):
        data[t] = gen.calc(t)  # same idiom as the reference's test script
    stored = golden_timeslices(g)
    for i, t in stored:
        _blocks_close(data[t], g["E"][i], what=f"{name} t={t}")
    # every contraction form of the library against the reference's output, not only the planned one
    i, t = stored[-1]
    _all_forms(gen, t, g["E"][i], name)
    # complex128 input goes through the device-side complex64 rounding and must agree too
    gen2 = edb.ElementalGenerator(latt, edb.GaugeFieldHostmem(g["U"]),
                                  edb.EigenvectorHostmem(g["V"].astype(np.complex128)), int(g["num_nabla"]), moms)
    gen2.load("cfg")
    _blocks_close(gen2.calc(stored[0][1]), g["E"][stored[0][0]], what=name + " c16 input")


@pytest.mark.parametrize("name", ["disp_weak_4x4x4x2", "disp_random_4x6x8x1",
                                  "config1_disp_weak_4x4x4x8",  # tests/test_displacement_elemental.py:15-23: distance 8, 6 momenta
                                  "disp_sep_16x2x4x1"])
def test_displacement_elementals_match_reference_golden(edb, name):
    g, latt, moms = _golden_case(name)
    gen = edb.DisplacementElementalGenerator(latt, edb.GaugeFieldHostmem(g["U"]), edb.EigenvectorHostmem(g["V"]),
                                             int(g["distance"]), moms)
    assert gen.distance == int(g["distance"])
    gen.load("cfg")
    for i, t in golden_timeslices(g):
        _blocks_close(np.array(gen.calc(t)), g["E"][i], what=f"{name} t={t}")
    i, t = golden_timeslices(g)[-1]
    _all_forms(gen, t, g["E"][i], name)


def test_reference_stored_goldens_when_materialised(edb):
    """the reference's tests/weak_field.* are git-LFS pointers in its checkout.  Where they are the real
    files (sha256 of SURVEY section 4) the reference's own test scripts are replayed here: same loaders' inputs
    (ILDG gauge field, .npy eigenvectors), same constructor arguments, against its stored elemental files."""
    paths, why = reference_weak_field_files()
    if paths is None:
        pytest.skip("stored goldens of the reference unavailable: " + why)
    import os

    prefix = os.path.dirname(paths["weak_field.lime"]) + "/"
    latt, Ne = [4, 4, 4, 8], 20
    gauge = edb.GaugeFieldIldg(prefix, ".lime", [8, 4, 4, 4, 4, 3, 3])
    evec = edb.EigenvectorNpy(prefix, ".eigenvector.input.npy", [8, Ne, 4, 4, 4, 3], Ne)
    moms = [(0, 0, 0), (0, 0, 1), (0, 1, 1), (1, 1, 1), (0, 0, 2), (0, 1, 2), (1, 1, 2)]
    gen = edb.ElementalGenerator(latt, gauge, evec, 2, moms)
    gen.load("weak_field")
    ref = np.load(paths["weak_field.elemental.npy"], mmap_mode="r")
    for t in range(8):
        _blocks_close(np.array(gen.calc(t)), np.asarray(ref[:, :, t]), what=f"weak_field elemental t={t}")
    dmoms = [(0, 0, 0), (0, 0, 1), (0, 1, 1), (1, 1, 1), (0, 1, 2), (1, 1, 2)]
    dgen = edb.DisplacementElementalGenerator(latt, gauge, evec, 8, dmoms)
    dgen.load("weak_field")
    dref = np.load(paths["weak_field.displacement_elemental.npy"], mmap_mode="r")
    for t in range(8):
        _blocks_close(np.array(dgen.calc(t)), np.asarray(dref[:, :, t]), what=f"weak_field displacement t={t}")


def test_blending_matches_reference_golden(edb):
    g, latt, moms = _golden_case("deriv_blend_4x4x4x1")
    dil = ([int(v) for v in g["dilution_tot"]], [int(v) for v in g["dilution_used"]])
    gen = edb.ElementalGenerator(latt, edb.GaugeFieldHostmem(g["U"]), edb.EigenvectorHostmem(g["V"]),
                                 int(g["num_nabla"]), moms, dil, True)
    assert gen.stocastic_coeff.shape == (6, 6)
    gen.load("cfg")
    _blocks_close(np.array(gen.calc(0)), g["E"][0], what="blending")
    with pytest.raises(ValueError):
        edb.ElementalGenerator(latt, edb.GaugeFieldHostmem(g["U"]), edb.EigenvectorHostmem(g["V"]), 1, moms, None, True)


@pytest.mark.parametrize("name", ["gauge_stout_4x4x6x2", "gauge_project_4x4x6x2", "gauge_project_stout_4x4x6x2"])
def test_gauge_preprocessing_matches_reference_golden(edb, name):
    """stout_smear(nstep, rho) / project_SU3() of the generator classes (the reference's only CUDA
    kernel is the stout step): processed links and the elementals computed on them."""
    g, latt, moms = _golden_case(name)
    gen = edb.ElementalGenerator(latt, edb.GaugeFieldHostmem(g["U"]), edb.EigenvectorHostmem(g["V"]), int(g["num_nabla"]), moms)
    gen.load("cfg")
    for (kind, nstep), rho in zip(g["ops"], g["rhos"]):
        if kind == 1:
            gen.stout_smear(int(nstep), float(rho))
        else:
            gen.project_SU3()
    for t in range(latt[3]):
        E = np.array(gen.calc(t))
        links = gen._engine.debug_links().cpu().numpy()
        assert rel_err(links, g["links"][:, t]) < 1e-12, (name, t)
        _blocks_close(E, g["E"][t], what=f"{name} t={t}")
    # the displacement generator shares the preprocessing
    dgen = edb.DisplacementElementalGenerator(latt, edb.GaugeFieldHostmem(g["U"]), edb.EigenvectorHostmem(g["V"]), 1, moms)
    dgen.load("cfg")
    for (kind, nstep), rho in zip(g["ops"], g["rhos"]):
        dgen.stout_smear(int(nstep), float(rho)) if kind == 1 else dgen.project_SU3()
    dgen.calc(1)
    assert rel_err(dgen._engine.debug_links().cpu().numpy(), g["links"][:, 1]) < 1e-12
    # load() starts again from the unprocessed configuration
    gen.load("cfg")
    gen.calc(0)
    raw = np.moveaxis(g["U"][0], 3, 0)[:3]
    assert rel_err(gen._engine.debug_links().cpu().numpy(), raw) == 0.0


def test_stout_on_unit_links_is_identity(edb):
    """Q = 0: the reference divides 0/0 here, the kernel leaves the links unchanged."""
    orc = _orc()
    latt, Ne = [4, 4, 4, 1], 3
    U = np.zeros((1, 4, 4, 4, 4, 3, 3), np.complex128)
    U[...] = np.eye(3)
    V = orc.synthetic_eigvecs(latt, Ne, 0)[None]
    gen = edb.ElementalGenerator(latt, edb.GaugeFieldHostmem(U), edb.EigenvectorHostmem(V), 1, [(0, 0, 0)])
    gen.load("x")
    gen.stout_smear(2, 0.1)
    E = np.array(gen.calc(0))
    assert np.all(np.isfinite(E))
    assert np.array_equal(gen._engine.debug_links().cpu().numpy(), np.moveaxis(U[0], 3, 0)[:3])


def test_calc_returns_generator_owned_buffer_and_checks_state(edb):
    g, latt, moms = _golden_case("deriv_weak_4x4x4x2")
    gen = edb.ElementalGenerator(latt, edb.GaugeFieldHostmem(g["U"]), edb.EigenvectorHostmem(g["V"]), 1, moms[:2])
    with pytest.raises(RuntimeError):
        gen.calc(0)
    gen.load("cfg")
    a = gen.calc(0)
    first = a.copy()
    b = gen.calc(1)
    assert a is b and not np.array_equal(first, b)  # overwritten, like the reference's _VPV
    with pytest.raises(IndexError):
        gen.calc(latt[3])
    fresh = edb.ElementalGenerator(latt, edb.GaugeFieldHostmem(g["U"]), edb.EigenvectorHostmem(g["V"]), 0, moms[:1])
    with pytest.raises(RuntimeError):
        fresh.stout_smear(1, 0.1)  # nothing loaded yet


# ---------------------------------------------------------------------------------------------
# (2) oracle on synthetic inputs: kernels one by one, then edge shapes
# ---------------------------------------------------------------------------------------------
def test_phase_table_matches_oracle(edb):
    import torch

    orc = _orc()
    latt = [6, 4, 10, 1]
    mp = edb.MomentumPhase(latt)
    for p in [(0, 0, 0), (1, 0, 0), (0, -1, 2), (3, 2, 1), (-7, 11, 5)]:
        got = mp.get(p).cpu().numpy()
        assert got.shape == (10, 4, 6)
        # the oracle (like the reference) forms theta = 2 pi p.x/L in floating point, so ITS error
        # grows like |theta| * 2^-53; the kernel reduces p.x mod L in integers first
        theta_max = 2 * np.pi * sum(abs(c) for c in p)
        assert np.max(np.abs(got - orc.momentum_phase(latt, p))) < 4e-16 * (1 + theta_max)
        assert np.max(np.abs(np.abs(got) - 1.0)) < 3e-16
    assert mp.get((1, 0, 0)) is mp.get((1, 0, 0))
    g = load_golden("insertion_maps")
    mp = edb.MomentumPhase([int(v) for v in g["phase_latt"]])
    for p, ref in zip(g["phase_moms"], g["phases"]):
        assert np.max(np.abs(mp.get(tuple(int(v) for v in p)).cpu().numpy() - ref)) < 1e-14
    torch.cuda.synchronize()


@pytest.mark.parametrize("latt,Ne", [([4, 4, 4, 1], 5), ([6, 2, 4, 1], 3), ([3, 5, 7, 1], 4), ([8, 8, 8, 1], 9)])
def test_stencil_fields_match_oracle(edb, latt, Ne):
    """Every derived field W[seq] (bit-for-bit the same formula, so ~1e-15)."""
    import torch

    from easydistillation_b200 import _capi
    from easydistillation_b200.engine import ElementalEngine

    orc = _orc()
    U_file = orc.synthetic_links(latt, 0)
    V = orc.synthetic_eigvecs(latt, Ne, 0)
    U = orc.links_file_to_spatial(U_file)
    eng = ElementalEngine(latt[:3], Ne, _capi.MODE_DERIVATIVE, 2, [(0, 0, 0)])
    eng.set_links(torch.from_numpy(U_file).cuda(), _capi.LINKS_FILE_T)
    eng.set_eigvecs(torch.from_numpy(V).cuda())
    eng.calc()
    W0 = orc.round_through_c8(V).astype(np.complex128)
    assert np.array_equal(eng.debug_field(0).cpu().numpy(), W0)  # rounding is exact
    for a in range(3):
        W1 = orc.covariant_hop(W0, U, a)
        assert rel_err(eng.debug_field(1 + a).cpu().numpy(), W1) < 1e-14
        for b in range(3):
            W2 = orc.covariant_hop(W1, U, b)
            assert rel_err(eng.debug_field(4 + 3 * a + b).cpu().numpy(), W2) < 1e-14
    # direction-major link layout gives the same fields
    eng.set_links(torch.from_numpy(np.ascontiguousarray(U)).cuda(), _capi.LINKS_DIR_MAJOR)
    eng.calc()
    assert rel_err(eng.debug_field(2).cpu().numpy(), orc.covariant_hop(W0, U, 1)) < 1e-14


def test_displacement_fields_match_oracle(edb):
    import torch

    from easydistillation_b200 import _capi
    from easydistillation_b200.engine import ElementalEngine

    orc = _orc()
    latt, Ne, dist = [4, 6, 2, 1], 5, 4
    U_file = orc.synthetic_links(latt, 1)
    V = orc.synthetic_eigvecs(latt, Ne, 1)
    eng = ElementalEngine(latt[:3], Ne, _capi.MODE_DISPLACEMENT, dist, [(0, 0, 0)])
    eng.set_links(torch.from_numpy(U_file).cuda(), _capi.LINKS_FILE_T)
    eng.set_eigvecs(torch.from_numpy(V).cuda())
    eng.calc()
    for k, Dk in enumerate(orc.displacement_fields(V, orc.links_file_to_spatial(U_file), dist)):
        assert rel_err(eng.debug_field(k).cpu().numpy(), np.asarray(Dk, np.complex128)) < 1e-14, k


@pytest.mark.parametrize(
    "latt,Ne,nabla,nmom",
    [
        ([4, 4, 4, 1], 1, 2, 1),     # single eigenvector
        ([4, 4, 4, 1], 3, 0, 2),     # no derivative
        ([3, 5, 7, 1], 7, 2, 5),     # V = 105: not a multiple of the 8-site k stage, ragged Ne
        ([2, 2, 2, 1], 13, 1, 3),    # tiny volume, Ne not a multiple of 4
        ([8, 4, 6, 1], 21, 2, 4),    # several m-fragments with a ragged last one
        ([4, 4, 4, 1], 110, 1, 2),   # two row tiles (Ne > 104)
        ([4, 4, 2, 1], 6, 3, 2),     # third-order derivatives (40 operators)
    ],
)
def test_derivative_elementals_match_oracle_edge_shapes(edb, latt, Ne, nabla, nmom):
    orc = _orc()
    moms = orc.momentum_set(33)[::7][:nmom] if nmom > 1 else [(1, -1, 2)]
    U_file = orc.synthetic_links(latt, 2)[None]
    V = orc.synthetic_eigvecs(latt, Ne, 2)[None]
    gen = edb.ElementalGenerator(latt, edb.GaugeFieldHostmem(U_file), edb.EigenvectorHostmem(V), nabla, moms)
    gen.load("x")
    got = np.array(gen.calc(0))
    U = orc.links_file_to_spatial(U_file[0])
    if nabla <= 2:
        ref = orc.elemental_timeslice_closed_form(V[0], U, latt, nabla, moms)
    else:
        ref = orc.elemental_timeslice(V[0], U, latt, nabla, moms)
    _blocks_close(got, ref, what=f"{latt} Ne={Ne} nabla={nabla}")
    _all_forms(gen, 0, ref, f"{latt} Ne={Ne} nabla={nabla}")


def test_dmma_tile_variants_and_split_k_agree_with_scalar_kernel(edb):
    """Every instantiated tile height and several split-K factors against the scalar
    one-thread-per-output contraction and the oracle."""
    import torch

    from easydistillation_b200 import _capi
    from easydistillation_b200.engine import ElementalEngine

    orc = _orc()
    latt, Ne = [4, 6, 8, 1], 30
    moms = orc.momentum_set(9)
    U_file = orc.synthetic_links(latt, 5)
    V = orc.synthetic_eigvecs(latt, Ne, 5)
    ref = orc.elemental_timeslice_closed_form(V, orc.links_file_to_spatial(U_file), latt, 2, moms)
    eng = ElementalEngine(latt[:3], Ne, _capi.MODE_DERIVATIVE, 2, moms)
    eng.set_links(torch.from_numpy(U_file).cuda(), _capi.LINKS_FILE_T)
    eng.set_eigvecs(torch.from_numpy(V).cuda())
    eng.debug_use_naive_gram(True)
    naive = eng.calc().cpu().numpy()
    _blocks_close(naive, ref, what="scalar kernel")
    eng.debug_use_naive_gram(False)
    # product path = TMA producer warp + mbarrier ring with 3M arithmetic; the 4M arithmetic and
    # the cp.async loader are kept as cross-checks
    for loader, algo in ((0, 1), (0, 0), (1, 0)):
        eng.debug_loader(loader)
        eng.debug_algo(algo)
        for mfrag in range(2, 14):
            for ksplit in (1, 3):
                eng.debug_gram_config(mfrag, ksplit)
                q = eng.query()
                assert (q["tma_stages"] >= 2) == (loader == 0)
                assert q["real_mma_per_complex_block"] == (3 if algo else 4)
                got = eng.calc().cpu().numpy()
                _blocks_close(got, ref, what=f"loader={loader} algo={algo} mfrag={mfrag} ksplit={ksplit}")
                _blocks_close(got, naive, what=f"loader={loader} algo={algo} mfrag={mfrag} ksplit={ksplit} vs scalar")
        eng.debug_gram_config(0, 24)
        _blocks_close(eng.calc().cpu().numpy(), ref, what=f"loader={loader} algo={algo} ksplit=24")
    with pytest.raises(ValueError):
        eng.debug_gram_config(14, 1)


@pytest.mark.parametrize("nabla", [1, 2, 3])
def test_hermitian_pairing_and_direct_pairs_agree_with_oracle(edb, nabla):
    """The contraction either runs every (left, right) pair (34 for num_nabla=2) or only the
    canonical ones and reads G(R,L,-p)^dagger for the rest (19).  Both modes, on a momentum list
    that is closed under negation and on one that is not (missing -p are added internally)."""
    import torch

    from easydistillation_b200 import _capi
    from easydistillation_b200.engine import ElementalEngine

    orc = _orc()
    latt, Ne = ([4, 6, 8, 1], 10) if nabla < 3 else ([4, 4, 2, 1], 5)
    U_file = orc.synthetic_links(latt, 6)
    V = orc.synthetic_eigvecs(latt, Ne, 6)
    U = orc.links_file_to_spatial(U_file)
    expected_pairs = {1: (7, 4), 2: (34, 19)}
    for moms in (orc.momentum_set(7), [(0, 0, 0), (1, 0, 0), (0, -1, 2), (1, 1, 1), (-1, -1, -1)]):
        ref = orc.elemental_timeslice(V, U, latt, nabla, moms)
        eng = ElementalEngine(latt[:3], Ne, _capi.MODE_DERIVATIVE, nabla, moms)
        eng.set_links(torch.from_numpy(U_file).cuda(), _capi.LINKS_FILE_T)
        eng.set_eigvecs(torch.from_numpy(V).cuda())
        closed = set(moms) == {tuple(-c for c in p) for p in moms}
        if closed:
            assert eng.query()["hermitian_pairing"]  # auto mode picks the cheaper evaluation
        for mode in (0, 1):
            eng.debug_symmetry(mode)
            q = eng.query()
            assert q["hermitian_pairing"] == bool(mode)
            if nabla in expected_pairs:
                assert q["pair_gemms_per_momentum"] == expected_pairs[nabla][mode]
            assert q["internal_momenta"] == (len(moms) if (closed or not mode) else len(moms) + 2)
            if mode and closed and nabla == 2:
                # 4 self pairs run only p = 0 and one of each +-p couple (4 of 7 momenta)
                assert q["half_set_momenta"] == 4 and q["pair_momentum_gemms"] == 15 * 7 + 4 * 4
            _blocks_close(eng.calc().cpu().numpy(), ref, what=f"nabla={nabla} pairing={mode} closed={closed}")


def test_displacement_matches_oracle_ragged(edb):
    orc = _orc()
    latt, Ne, dist = [3, 4, 5, 1], 11, 3
    moms = [(0, 0, 0), (1, 2, -1), (0, -1, 0)]
    U_file = orc.synthetic_links(latt, 9)[None]
    V = orc.synthetic_eigvecs(latt, Ne, 9)[None]
    gen = edb.DisplacementElementalGenerator(latt, edb.GaugeFieldHostmem(U_file), edb.EigenvectorHostmem(V), dist, moms)
    gen.load("x")
    ref = orc.displacement_timeslice(V[0], orc.links_file_to_spatial(U_file[0]), latt, dist, moms)
    _blocks_close(np.array(gen.calc(0)), ref, what="displacement ragged")
    gen0 = edb.DisplacementElementalGenerator(latt, edb.GaugeFieldHostmem(U_file), edb.EigenvectorHostmem(V), 0, moms)
    gen0.load("x")
    _blocks_close(np.array(gen0.calc(0)), ref[:1], what="distance 0")


def test_file_handles_and_elemental_npy_roundtrip(edb, tmp_path):
    """Inputs through file handles, output saved in the reference's [Nop,Nmom,Lt,Ne,Ne] layout."""
    g, latt, moms = _golden_case("deriv_weak_4x4x4x2")
    prefix = str(tmp_path) + "/"
    g["U"].astype("<c16").tofile(prefix + "weak.dat")
    np.save(prefix + "weak.eigenvector.npy", g["V"].astype("<c16"))
    Lx, Ly, Lz, Lt = latt
    gauge = edb.GaugeFieldBinary(prefix, ".dat", [Lt, Lz, Ly, Lx, 4, 3, 3], "<c16")
    evec = edb.EigenvectorNpy(prefix, ".eigenvector.npy", [Lt, 8, Lz, Ly, Lx, 3], 8)
    gen = edb.ElementalGenerator(latt, gauge, evec, 2, moms)
    gen.load("weak")
    data = gen.calc_range(0, Lt)
    el = edb.ElementalNpy(prefix, ".elemental.npy", [13, len(moms), Lt, 8, 8], 8)
    mm = el.create("weak", [13, len(moms), Lt, 8, 8])
    mm[...] = data.transpose(1, 2, 0, 3, 4)
    mm.flush()
    back = el.load("weak")[:]
    _blocks_close(back[:, :, 1], g["E"][1], what="npy roundtrip")
    # streamed writer: whole file, then the same file written as two rank slabs, and a complex64 down-cast
    el2 = edb.ElementalNpy(prefix, ".elemental2.npy", None, 8)
    gen.calc_to_file(el2, "weak")
    assert np.array_equal(el2.load("weak")[:], back)
    el3 = edb.ElementalNpy(prefix, ".elemental3.npy", None, 8)
    gen.calc_to_file(el3, "weak", t_range=(0, 1))
    gen.calc_to_file(el3, "weak", t_range=(1, Lt))
    assert np.array_equal(el3.load("weak")[:], back)
    el4 = edb.ElementalNpy(prefix, ".elemental4.npy", None, 8)
    gen.calc_to_file(el4, "weak", dtype="<c8")
    got8 = el4.load("weak")[:]
    assert got8.dtype == np.complex64 and np.array_equal(got8, back.astype(np.complex64))
    # calc_all without torch.distributed = all timeslices on this GPU
    full = gen.calc_all()
    assert tuple(full.shape) == (Lt, 13, len(moms), 8, 8)
    _blocks_close(full[0].cpu().numpy(), g["E"][0], what="calc_all")


def test_big_endian_files_are_converted_on_the_device(edb, tmp_path):
    """ILDG gauge file + QDP timeslice eigenvector file (big-endian payloads, tests/golden made by
    oracle/make_golden_files.py: the reference read the same two files through its own readers) go to
    the GPU as raw bytes and are byte-swapped by the kernels; every calc flavour must match the
    reference's elementals."""
    import torch

    g, latt, moms = _golden_case("files_ildg_qdp")
    Lx, Ly, Lz, Lt = latt
    Ne = int(g["Ne"])
    g["lime_bytes"].tofile(tmp_path / "cfg.lime")
    g["mod_bytes"].tofile(tmp_path / "cfg.mod")
    prefix = str(tmp_path) + "/"
    gauge = edb.GaugeFieldIldg(prefix, ".lime")
    evec = edb.EigenvectorTimeSlice(prefix, ".mod", [Lt, Ne, Lz, Ly, Lx, 3], Ne)
    gen = edb.ElementalGenerator(latt, gauge, evec, int(g["num_nabla"]), moms)
    gen.load("cfg")
    assert gen._U.dtype == np.dtype(">c16") and gen._eigvecs_of(0).dtype == np.dtype(">c8")  # no host-side swap
    for t in range(Lt):
        _blocks_close(gen.calc(t), g["E"][t], what=f"calc({t}) from ILDG/QDP files")
        _blocks_close(gen.calc_device(t).cpu().numpy(), g["E"][t], what=f"calc_device({t}) from ILDG/QDP files")
    block = gen.calc_range(0, Lt)  # streamed pipeline
    for t in range(Lt):
        _blocks_close(block[t], g["E"][t], what=f"calc_range[{t}] from ILDG/QDP files")
    # in-memory big-endian arrays (page-locked in place by the pipeline), complex128 eigenvectors included
    U_be = np.ascontiguousarray(g["U_ref"], dtype=">c16")
    for vdt in (">c8", ">c16"):
        V_be = np.ascontiguousarray(g["V_ref"], dtype=vdt)
        gen2 = edb.ElementalGenerator(latt, edb.GaugeFieldHostmem(U_be), edb.EigenvectorHostmem(V_be), int(g["num_nabla"]), moms)
        gen2.load("cfg")
        _blocks_close(gen2.calc(1), g["E"][1], what=f"big-endian host arrays {vdt}")
        block = gen2.calc_range(0, Lt)
        for t in range(Lt):
            _blocks_close(block[t], g["E"][t], what=f"big-endian host arrays {vdt}, streamed t={t}")
    # the same through the displacement generator against the little-endian route
    args = (latt, 2, moms)
    d_be = edb.DisplacementElementalGenerator(latt, gauge, evec, 2, moms)
    d_le = edb.DisplacementElementalGenerator(latt, edb.GaugeFieldHostmem(g["U_ref"]), edb.EigenvectorHostmem(g["V_ref"]), 2, moms)
    d_be.load("cfg")
    d_le.load("cfg")
    assert np.array_equal(d_be.calc(0), d_le.calc(0)), args
    torch.cuda.synchronize()


def test_device_resident_inputs(edb):
    """the generators also take torch CUDA tensors (a handle whose load() returns a tensor)."""
    import torch

    g, latt, moms = _golden_case("deriv_weak_4x4x4x2")

    class TensorHandle:
        def __init__(self, tensor, Ne=None):
            self.tensor, self.Ne = tensor, Ne

        def load(self, key):
            return self.tensor

    U_dev = torch.from_numpy(g["U"]).cuda()
    V_dev = torch.from_numpy(g["V"]).cuda()
    for gauge, evec in ((TensorHandle(U_dev), TensorHandle(V_dev, 8)),                       # both on the device
                        (edb.GaugeFieldHostmem(g["U"]), TensorHandle(V_dev, 8)),             # mixed
                        (TensorHandle(U_dev), edb.EigenvectorHostmem(g["V"]))):
        gen = edb.ElementalGenerator(latt, gauge, evec, 2, moms)
        gen.load("cfg")
        _blocks_close(np.array(gen.calc(1)), g["E"][1], what="device-resident inputs, calc")
        out = gen.calc_device(0)
        assert out.is_cuda
        _blocks_close(out.cpu().numpy(), g["E"][0], what="device-resident inputs, calc_device")
        full = gen.calc_all()
        _blocks_close(full[1].cpu().numpy(), g["E"][1], what="device-resident inputs, calc_all")
        _blocks_close(gen.calc_range(0, 2)[0], g["E"][0], what="device-resident inputs, calc_range")


def test_laplacian_matches_reference_golden(edb):
    """The eigensolver's operator on the same links and stencil machinery."""
    import torch

    orc = _orc()
    g = load_golden("laplacian_4x6x8")
    latt = [int(v) for v in g["latt_size"]]
    lap = edb.Laplacian(latt, edb.GaugeFieldHostmem(g["U_file"][None]))
    lap.load("x")
    lap.set_timeslice(0)
    assert rel_err(lap.matmat(g["F"]), g["LF"]) < 1e-14
    Xd = torch.from_numpy(g["F"]).cuda()
    assert rel_err(lap.matmat(Xd).cpu().numpy(), g["LF"]) < 1e-14
    assert rel_err(lap.matvec(g["F"][2]), g["LF"][2]) < 1e-14
    # the reference's own convention (lattice/generator/eigenvector.py:11-26): flat vectors, vector index fastest
    N = lap.shape[0]
    F_flat = np.ascontiguousarray(np.moveaxis(g["F"], 0, -1)).reshape(N, -1)
    LF_flat = np.ascontiguousarray(np.moveaxis(g["LF"], 0, -1)).reshape(N, -1)
    assert rel_err(lap.matmat(F_flat), LF_flat) < 1e-14
    assert rel_err(lap.matvec(F_flat[:, 1]), LF_flat[:, 1]) < 1e-14
    assert rel_err(lap.matmat(torch.from_numpy(F_flat).cuda()).cpu().numpy(), LF_flat) < 1e-14
    op = lap.as_linear_operator()
    assert op.shape == (N, N) and rel_err(op @ F_flat[:, :2], LF_flat[:, :2]) < 1e-14
    # odd volume, many vectors (several vector chunks), smeared links
    latt2 = [3, 5, 7, 1]
    U_file = orc.synthetic_links(latt2, 1, "weak")
    F = orc.synthetic_eigvecs(latt2, 37, 1)
    lap2 = edb.Laplacian(latt2, edb.GaugeFieldHostmem(U_file[None]))
    lap2.load("x")
    lap2.stout_smear(2, 0.1)
    lap2.set_timeslice(0)
    U_s = orc.stout_smear_timeslice(orc.links_file_to_spatial(U_file), 2, 0.1)
    assert rel_err(lap2.matmat(F), orc.laplacian(F, U_s)) < 1e-13


def test_streamed_pipeline_matches_per_timeslice_calls(edb):
    """calc_range / calc_all overlap upload, kernels and download over double buffers: five
    timeslices so every buffer is reused, complex128 and complex64 eigenvector sources."""
    orc = _orc()
    latt, Ne = [4, 4, 6, 5], 7
    moms = orc.momentum_set(7)
    U = np.stack([orc.synthetic_links(latt, t, "weak") for t in range(latt[3])])
    V = np.stack([orc.synthetic_eigvecs(latt, Ne, t) for t in range(latt[3])])
    for Vsrc in (V, V.astype(np.complex64)):
        gen = edb.ElementalGenerator(latt, edb.GaugeFieldHostmem(U), edb.EigenvectorHostmem(Vsrc), 2, moms)
        gen.load("x")
        single = np.stack([np.array(gen.calc(t)) for t in range(latt[3])])
        ranged = gen.calc_range(0, latt[3])
        assert np.array_equal(ranged, single)  # same kernels, same order: bit-identical
        assert np.array_equal(gen.calc_range(1, 4), single[1:4])
        full = gen.calc_all().cpu().numpy()
        assert np.array_equal(full, single)
        ref = orc.elemental_timeslice_closed_form(V[3], orc.links_file_to_spatial(U[3]), latt, 2, moms)
        _blocks_close(ranged[3], ref, what="pipeline t=3")
    dgen = edb.DisplacementElementalGenerator(latt, edb.GaugeFieldHostmem(U), edb.EigenvectorHostmem(V), 2, moms[:2])
    dgen.load("x")
    assert np.array_equal(dgen.calc_range(0, 3), np.stack([np.array(dgen.calc(t)) for t in range(3)]))


# ---------------------------------------------------------------------------------------------
# (3) BASELINE.json sizes: direct oracle where it takes seconds, properties beyond
# ---------------------------------------------------------------------------------------------
def test_config2_shape_against_oracle(edb):
    """16^3, Ne=100, num_nabla=1, 9 momenta: one timeslice against the closed-form oracle, every contraction form."""
    orc = _orc()
    latt, Ne = [16, 16, 16, 1], 100
    moms = orc.momentum_set(9)
    U_file = orc.synthetic_links(latt, 0)[None]
    V = orc.synthetic_eigvecs(latt, Ne, 0)[None].astype(np.complex64)
    gen = edb.ElementalGenerator(latt, edb.GaugeFieldHostmem(U_file), edb.EigenvectorHostmem(V), 1, moms)
    gen.load("x")
    assert gen._engine.query()["contraction_form"] == 4  # what a user gets: the separable form
    got = np.array(gen.calc(0))
    ref = orc.elemental_timeslice_closed_form(V[0], orc.links_file_to_spatial(U_file[0]), latt, 1, moms)
    _blocks_close(got, ref, what="config 2")
    _all_forms(gen, 0, ref, "config 2")


def _hermiticity(E, moms, tol=TOL):
    """E[0,p]^dag = E[0,-p],  E[a,p]^dag = -E[a,-p],  E[(a,b),p]^dag = E[(b,a),-p] (num_nabla = 2)."""
    neg = [moms.index(tuple(-c for c in p)) for p in moms]
    for ip in range(len(moms)):
        assert rel_err(E[0, ip].conj().T, E[0, neg[ip]]) < tol
        for a in range(3):
            assert rel_err(-E[1 + a, ip].conj().T, E[1 + a, neg[ip]]) < tol
            for b in range(3):
                assert rel_err(E[4 + 3 * a + b, ip].conj().T, E[4 + 3 * b + a, neg[ip]]) < tol


def _large_shape_check(edb, latt, Ne, sel, what, cross_check_forms):
    """A BASELINE.json shape too slow for the full oracle (num_nabla = 2, 33 momenta), with the library default:
    (a) the form the library plans is the separable one (what ElementalGenerator.calc gives a user);
    (b) Hermiticity of every block and the unit Gram diagonal at p = 0 (complex64 accuracy of the inputs);
    (c) the sub-block of every (operator, momentum) on the eigenvectors `sel` against the closed-form oracle run on
        those vectors alone - for the planned form and for each form of `cross_check_forms`;
    (d) the whole result of each cross-check form against the planned form's, to 1e-10."""
    import torch

    orc = _orc()
    moms = orc.momentum_set(33)
    U_file = orc.synthetic_links(latt, 0)[None]
    V = orc.synthetic_eigvecs(latt, Ne, 0)[None].astype(np.complex64)
    gen = edb.ElementalGenerator(latt, edb.GaugeFieldHostmem(U_file), edb.EigenvectorHostmem(V), 2, moms)
    gen.load("x")
    q = gen._engine.query()
    assert q["contraction_form"] == 4 and q["form_requested"] == -1 and q["plane_wave_modes"] == 13, q
    assert q["hermitian_pairing"] and q["pair_gemms_per_momentum"] == 19
    E = np.array(gen.calc(0))
    _hermiticity(E, moms)
    assert np.max(np.abs(np.diag(E[0, 0]) - 1.0)) < 1e-6
    ref = orc.elemental_timeslice_closed_form(V[0][sel], orc.links_file_to_spatial(U_file[0]), latt, 2, moms)
    worst = {4: _blocks_close(E[:, :, sel][:, :, :, sel], ref, what=f"{what} sub-block, planned form")}
    for form in cross_check_forms:
        gen._engine.debug_algo(form)
        assert gen._engine.query()["contraction_form"] == form
        F = np.array(gen.calc(0))
        worst[form] = _blocks_close(F[:, :, sel][:, :, :, sel], ref, what=f"{what} sub-block, form {form}")
        _blocks_close(F, E, what=f"{what}: form {form} vs planned form, full result")
    print(f"{what}: worst sub-block error vs oracle by form: {worst}")
    gen._engine.debug_algo(-1)
    del gen
    torch.cuda.empty_cache()
    return E


def test_config3_shape_properties(edb):
    """24^3, Ne=100, num_nabla=2, 33 momenta (BASELINE config 3; 6 site pairs per stage): every form, plus the two
    evaluation orders (Hermitian pairing on: 19 pair contractions, off: 34 with multi-segment jobs) at full size."""
    orc = _orc()
    latt, Ne = [24, 24, 24, 1], 100
    E = _large_shape_check(edb, latt, Ne, [3, 17, 42, 64, 65, 99], "config 3", (3, 2, 1))
    moms = orc.momentum_set(33)
    U_file = orc.synthetic_links(latt, 0)[None]
    V = orc.synthetic_eigvecs(latt, Ne, 0)[None].astype(np.complex64)
    gen = edb.ElementalGenerator(latt, edb.GaugeFieldHostmem(U_file), edb.EigenvectorHostmem(V), 2, moms)
    gen.load("x")
    gen._engine.debug_symmetry(0)
    q = gen._engine.query()
    assert not q["hermitian_pairing"] and q["contraction_form"] == 4 and q["pair_gemms_per_momentum"] == 34
    _blocks_close(np.array(gen.calc(0)), E, tol=1e-11, what="config 3 direct pairs vs Hermitian pairing")


def test_config4_shape_properties(edb):
    """32^3, Ne=200, num_nabla=2, 33 momenta (BASELINE config 4: 13 x 7 tiles of 16 x 32 with idle edge warps and the
    self pairs' mirror tiles).  The full oracle would take ~50 min here: sub-block, properties, and the GEMM form with
    every pair contracted directly in 4M arithmetic (34 x 33 GEMMs) as an independent evaluation at full size."""
    import torch

    orc = _orc()
    latt, Ne = [32, 32, 32, 1], 200
    E = _large_shape_check(edb, latt, Ne, [0, 57, 103, 104, 199], "config 4", (3, 1))
    moms = orc.momentum_set(33)
    U_file = orc.synthetic_links(latt, 0)[None]
    V = orc.synthetic_eigvecs(latt, Ne, 0)[None].astype(np.complex64)
    gen = edb.ElementalGenerator(latt, edb.GaugeFieldHostmem(U_file), edb.EigenvectorHostmem(V), 2, moms)
    gen.load("x")
    gen._engine.debug_algo(0)
    gen._engine.debug_symmetry(0)
    q = gen._engine.query()
    assert q["pair_momentum_gemms"] == 34 * 33 and q["real_mma_per_complex_block"] == 4 and q["contraction_form"] == 0
    _blocks_close(np.array(gen.calc(0)), E, what="config 4: direct pairs / 4M GEMM form vs planned form")
    del gen
    torch.cuda.empty_cache()


def test_config5_shape_properties(edb):
    """48^3, Ne=200, num_nabla=2, 33 momenta: the graded configuration (BASELINE config 5, 8 site pairs per stage, three
    stages per row).  The reference would need hours per timeslice here (elemental.py:309-329 at K = 331 776): the
    library default against the closed-form oracle on 6 of the 200 eigenvectors, Hermiticity of all 429 blocks, and
    the folded plane-wave form as an independent evaluation of the whole result."""
    _large_shape_check(edb, [48, 48, 48, 1], 200, [0, 1, 77, 128, 198, 199], "config 5", (3,))


def test_linearity_and_scaling_property(edb):
    """E is sesquilinear in the eigenvectors: scaling vector f by c scales column f by c and row f by conj(c)
    (powers of two keep the complex64 staging exact)."""
    orc = _orc()
    latt, Ne = [8, 8, 8, 1], 12
    moms = orc.momentum_set(5)
    U_file = orc.synthetic_links(latt, 4)[None]
    V = orc.synthetic_eigvecs(latt, Ne, 4)[None].astype(np.complex64)
    V2 = V.copy()
    V2[0, 5] *= 4j
    outs = []
    for vv in (V, V2):
        gen = edb.ElementalGenerator(latt, edb.GaugeFieldHostmem(U_file), edb.EigenvectorHostmem(vv), 2, moms)
        gen.load("x")
        outs.append(np.array(gen.calc(0)))
    expect = outs[0].copy()
    expect[:, :, :, 5] *= 4j
    expect[:, :, 5, :] *= -4j
    _blocks_close(outs[1], expect, what="sesquilinearity")


# ---------------------------------------------------------------------------------------------
# (6) every contraction form, strictly: oracle and GEMM-form parity on ragged / multi-tile / multi-segment shapes
# ---------------------------------------------------------------------------------------------
_SPECIAL = [(0, 0, 0), (0, 0, 1), (0, 1, 1), (1, 1, 1), (0, 0, 2), (0, 1, 2), (1, 1, 2)]  # the reference's test list
_D, _X = 0, 1
FORM_CASES = {
    "plane-wave": [
        ([4, 4, 4], 8, _D, 0, 7, None),
        ([3, 5, 2], 5, _D, 1, 7, None),        # plane of 15 sites: ragged 4-site groups and stages
        ([4, 6, 8], 30, _D, 2, 9, None),       # 2 x 1 tiles, pairing by cost
        ([4, 6, 8], 35, _D, 2, 33, 1),         # 3 x 2 tiles, 13 modes, Hermitian pairing + half set
        ([4, 6, 8], 19, _D, 2, 33, 0),         # direct pairs: multi-segment jobs with signs
        ([6, 4, 2], 21, _D, 2, [(0, 0, 1), (1, 2, 0), (3, -1, 2), (0, -2, 1)], None),  # non-closed, larger momenta
        ([2, 2, 2], 1, _D, 2, 7, None),        # planes of 4 sites (form 3 runs them unfolded)
        ([3, 3, 2], 9, _D, 2, 33, None),       # odd planes of 9 sites (form 3: self-paired middle site)
        ([4, 4, 6], 13, _D, 3, 7, None),
        ([4, 6, 8], 12, _X, 3, 9, None),
        ([5, 3, 7], 110, _D, 1, 9, None),      # 7 x 4 tiles
    ],
    "separable": [
        ([8, 4, 4], 8, _D, 0, 7, None),        # 4 pairs per stage, 5 modes
        ([8, 5, 2], 5, _D, 1, 9, None),        # 9 modes
        ([16, 3, 2], 20, _D, 1, 9, None),      # 8 pairs per stage
        ([12, 6, 4], 30, _D, 2, 9, None),      # 6 pairs per stage, pairing by cost
        ([12, 6, 4], 35, _D, 2, 33, 1),        # 3 x 2 tiles, 13 modes, Hermitian pairing + half set
        ([8, 6, 4], 19, _D, 2, 33, 0),         # direct pairs: multi-segment jobs with signs
        ([8, 4, 2], 21, _D, 2, _SPECIAL, None),
        ([24, 2, 2], 9, _D, 2, 33, None),      # two stages per row
        ([8, 4, 6], 13, _D, 3, 7, None),
        ([8, 6, 8], 12, _X, 3, 19, None),
        ([8, 3, 7], 110, _D, 1, 9, None),      # 7 x 4 tiles, idle edge warps
        ([16, 4, 4], 70, _D, 1, 33, None),     # mirror tiles of the self pair
    ],
}


@pytest.mark.parametrize("form,case", [(f, c) for f in (2, 3) for c in range(len(FORM_CASES["plane-wave"]))] +
                         [(4, c) for c in range(len(FORM_CASES["separable"]))])
def test_contraction_forms_against_oracle_and_gemm_form(edb, form, case):
    """Forms 2 / 3 (csrc/edk_gram_pw.cu) and 4 (csrc/edk_gram_sep.cu) in process, no tolerance for failure: block-wise
    1e-10 against the numpy oracle and against the GEMM form, and bit-identical when repeated."""
    import torch

    from easydistillation_b200 import _capi
    from easydistillation_b200.engine import ElementalEngine

    orc = _orc()
    latt, Ne, mode, order, moms, sym = FORM_CASES["separable" if form == 4 else "plane-wave"][case]
    moms = orc.momentum_set(moms) if isinstance(moms, int) else moms
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
    eng.debug_algo(form)
    q = eng.query()
    assert q["contraction_form"] == form and q["plane_wave_modes"] >= 1 and q["ksplit"] == 1, q
    got = eng.calc().cpu().numpy()
    _blocks_close(got, ref, what=f"form {form} {latt} Ne={Ne} vs oracle")
    _blocks_close(got, gemm, what=f"form {form} {latt} Ne={Ne} vs GEMM form")
    assert np.array_equal(got, eng.calc().cpu().numpy())
    eng.close()


@pytest.mark.parametrize("tile", ["24", "25", "17"])
def test_plane_wave_tile_shapes(edb, monkeypatch, tile):
    """Every instantiated tile shape of forms 2 / 3 (EDK_PW_TILE, an A/B hook): multi-tile, mirror tiles, partial f-tiles."""
    import torch

    from easydistillation_b200 import _capi
    from easydistillation_b200.engine import ElementalEngine

    orc = _orc()
    latt, Ne, moms = [4, 4, 4], 70, orc.momentum_set(7)
    U_file = orc.synthetic_links(latt + [1], 3)
    V = orc.synthetic_eigvecs(latt + [1], Ne, 3)
    ref = orc.elemental_timeslice_closed_form(V, orc.links_file_to_spatial(U_file), latt + [1], 1, moms)
    monkeypatch.setenv("EDK_PW_TILE", tile)
    for form in (2, 3):
        monkeypatch.setenv("EDK_GRAM_ALGO", str(form))
        eng = ElementalEngine(latt, Ne, _capi.MODE_DERIVATIVE, 1, moms)
        assert eng.query()["contraction_form"] == form and eng.query()["plane_wave_tile"] == int(tile)
        eng.set_links(torch.from_numpy(U_file).cuda(), _capi.LINKS_FILE_T)
        eng.set_eigvecs(torch.from_numpy(V).cuda())
        _blocks_close(eng.calc().cpu().numpy(), ref, what=f"form {form} tile {tile}")
        eng.close()


def test_two_generators_on_one_device_and_the_callers_current_device(edb):
    """Every entry point runs on its handle's device and restores the caller's current device (two GPUs: one
    generator each in one process; one GPU: two handles interleaved)."""
    import torch

    orc = _orc()
    latt, Ne = [8, 4, 4, 2], 6
    moms = orc.momentum_set(7)
    U = np.stack([orc.synthetic_links(latt, t) for t in range(2)])
    V = np.stack([orc.synthetic_eigvecs(latt, Ne, t) for t in range(2)])
    ndev = torch.cuda.device_count()
    devs = [0, 1] if ndev > 1 else [0, 0]
    before = torch.cuda.current_device()
    gens = [edb.ElementalGenerator(latt, edb.GaugeFieldHostmem(U), edb.EigenvectorHostmem(V), 1, moms, device=d) for d in devs]
    assert torch.cuda.current_device() == before
    for g in gens:
        g.load("x")
    outs = [np.array(g.calc_device(1).cpu().numpy()) for g in gens]  # set_links / set_eigvecs first, on each device
    assert torch.cuda.current_device() == before
    ref = orc.elemental_timeslice_closed_form(V[1], orc.links_file_to_spatial(U[1]), latt, 1, moms)
    for o in outs:
        _blocks_close(o, ref, what="two generators")


def test_integration_md_ctypes_stub_on_the_real_library(edb):
    """INTEGRATION.md section B's ctypes stub (the file a reference maintainer would add as lattice/generator/_edk.py),
    executed as written against the real libedk_sm100a.so on the GPU: EdkHandle.calc(U_t, V_t, out) with the reference's
    host buffers (file-order links of one timeslice, complex64 staging buffer) against the oracle, at the shape of the
    reference's own tests (4^3, Ne = 20, num_nabla = 2, 7 momenta) and at a shape the separable form takes."""
    import os
    import re

    from easydistillation_b200 import _capi

    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    block = re.search(r"```python\n(# lattice/generator/_edk\.py.*?)```", open(os.path.join(repo, "INTEGRATION.md")).read(), re.S).group(1)
    assert 'C.CDLL("libedk_sm100a.so")' in block
    ns = {}
    exec(compile(block.replace('C.CDLL("libedk_sm100a.so")', f"C.CDLL({_capi.LIB_PATH!r})"), "INTEGRATION.md", "exec"), ns)
    orc = _orc()
    for latt, Ne, nabla, nmom in [([4, 4, 4, 1], 20, 2, 7), ([8, 6, 4, 1], 12, 1, 9)]:
        moms = orc.momentum_set(nmom)
        U_t = np.ascontiguousarray(orc.synthetic_links(latt, 0))
        V_t = orc.synthetic_eigvecs(latt, Ne, 0)
        V8 = np.ascontiguousarray(V_t.astype(np.complex64))  # the reference's staging buffer (elemental.py:55)
        ref = orc.elemental_timeslice_closed_form(V8.astype(np.complex128), orc.links_file_to_spatial(U_t), latt, nabla, moms)
        out = np.zeros(ref.shape, np.complex128)
        handle = ns["EdkHandle"](latt, Ne, ns["EDK_MODE_DERIVATIVE"], nabla, moms)
        handle.calc(U_t, V8, out)
        _blocks_close(out, ref, what=f"INTEGRATION.md stub {latt}")
        del handle
    with pytest.raises(ValueError):  # EDK_ERR_ARG maps to the reference's ValueError
        ns["EdkHandle"]([4, 4, 4, 1], 0, ns["EDK_MODE_DERIVATIVE"], 1, orc.momentum_set(1))
