"""The WHOLE library on a machine without a GPU: csrc/*.cu (host glue of edk_api.cu, every launcher, every kernel
source) compiled by g++ with -DEDK_HOST_EMU against tests/emu (kernel launches = one host thread per CUDA thread,
device memory = host memory, cuTensorMapEncodeTiled = the emulator's box descriptor) and driven through the C ABI
of include/edk.h exactly as easydistillation_b200/engine.py drives libedk_sm100a.so, against the numpy oracle.

What this covers that the kernel-level emulator tests (test_pw_model.py) do not: job lists, momentum bookkeeping,
buffer sizes, tensor-map encoding, launch configurations, the switching between the contraction forms
(edk_debug_algo / EDK_GRAM_ALGO / edk_debug_symmetry / loader / scalar kernel), link preprocessing, blending and
the host-buffer entry point - i.e. everything between the ctypes call and the kernel.  TEST INFRASTRUCTURE ONLY:
the emulator library is built into a temporary directory and is never loaded by the package."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import REPO
from easydistillation_b200 import _capi
from oracle import elemental_oracle as orc

SOURCES = ["edk_stencil.cu", "edk_gauge.cu", "edk_gram.cu", "edk_gram_pw.cu", "edk_gram_sep.cu", "edk_gram_sep_s11.cu", "edk_gram_sep_s12.cu", "edk_gram_sep_s24.cu", "edk_api.cu"]
D, X = _capi.MODE_DERIVATIVE, _capi.MODE_DISPLACEMENT


def build_emulator_library(out, extra_flags=()):
    flags = ["-std=c++17", "-O1", "-fPIC", "-w", "-DEDK_HOST_EMU", *extra_flags, "-I", os.path.join(REPO, "tests", "emu"),
             "-I", os.path.join(REPO, "easydistillation_b200", "csrc"), "-I", os.path.join(REPO, "include"),
             "-I", "/usr/local/cuda/include"]
    jobs = []
    for src in SOURCES:
        obj = str(out / (src + ".o"))
        jobs.append((obj, subprocess.Popen(["g++", *flags, "-x", "c++", "-c", os.path.join(REPO, "easydistillation_b200", "csrc", src),
                                            "-o", obj], stderr=subprocess.PIPE, text=True)))
    obj = str(out / "emu_runtime.o")
    jobs.append((obj, subprocess.Popen(["g++", *flags, "-c", os.path.join(REPO, "tests", "emu", "emu_runtime.cpp"), "-o", obj],
                                       stderr=subprocess.PIPE, text=True)))
    for obj, p in jobs:
        _, err = p.communicate()
        assert p.returncode == 0, f"{obj}: {err[-3000:]}"
    so = str(out / "libedk_emu.so")
    r = subprocess.run(["g++", "-shared", *extra_flags, "-o", so, *[o for o, _ in jobs], "-lpthread"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    return so


def load_emulator_library(so):
    lib = C.CDLL(so)
    for name, (res, args) in _capi.SIGNATURES.items():  # the emulator build exports the same C ABI
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    return lib


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    return load_emulator_library(build_emulator_library(tmp_path_factory.mktemp("edk_emu")))


class Handle:
    """The calls engine.py makes, on host arrays."""

    def __init__(self, lib, latt3, Ne, mode, order, moms):
        self.lib, self.latt3, self.Ne = lib, list(latt3), Ne
        mom = np.ascontiguousarray(np.asarray(moms, np.int32).reshape(-1, 3))
        self.h = C.c_void_p()
        self.check(lib.edk_create(*self.latt3, Ne, mode, order, len(mom), mom.ctypes.data_as(C.POINTER(C.c_int)), 0,
                                  C.byref(self.h)), "edk_create")
        self.nop, self.nmom = lib.edk_num_operators(self.h), len(mom)

    def check(self, rc, what):
        assert rc == _capi.EDK_OK, f"{what}: status {rc}: {self.lib.edk_last_error().decode()}"

    def set_inputs(self, U_file, V):
        self.U, self.V = np.ascontiguousarray(U_file), np.ascontiguousarray(V)  # kept alive
        self.check(self.lib.edk_set_links(self.h, self.U.ctypes.data, _capi.LINKS_FILE_T, None), "edk_set_links")
        flags = _capi.EIGVECS_C8 if self.V.dtype == np.complex64 else 0
        self.check(self.lib.edk_set_eigvecs(self.h, self.V.ctypes.data, flags, None), "edk_set_eigvecs")

    def calc(self):
        out = np.full((self.nop, self.nmom, self.Ne, self.Ne), np.nan + 0j, np.complex128)
        self.check(self.lib.edk_calc(self.h, out.ctypes.data, None), "edk_calc")
        return out

    def query(self, what):
        return self.lib.edk_query(self.h, what)

    def close(self):
        self.lib.edk_destroy(self.h)


def worst_block_error(got, ref):
    norms = np.sqrt((np.abs(ref) ** 2).sum(axis=(-1, -2)))
    floor = 1e-4 * norms.max()
    return max(float(np.linalg.norm(got[a, p] - ref[a, p]) / max(norms[a, p], floor))
               for a in range(ref.shape[0]) for p in range(ref.shape[1]))


def inputs_and_reference(latt3, Ne, mode, order, moms, seed=3):
    latt = list(latt3) + [1]
    U_file = orc.synthetic_links(latt, seed)
    V = orc.synthetic_eigvecs(latt, Ne, seed)
    U = orc.links_file_to_spatial(U_file)
    if mode == D:
        ref = (orc.elemental_timeslice_closed_form if order <= 2 else orc.elemental_timeslice)(V, U, latt, order, moms)
    else:
        ref = orc.displacement_timeslice(V, U, latt, order, moms)
    return U_file, V, ref


def test_every_contraction_path_through_the_c_abi(emu):
    """num_nabla = 1, 7 momenta: TMA 3M (default), TMA 4M, plane-wave form, cp.async loader, scalar kernel - each
    selected the way engine.py selects it, each against the oracle."""
    latt3, Ne, moms = (4, 4, 2), 5, orc.momentum_set(7)
    U_file, V, ref = inputs_and_reference(latt3, Ne, D, 1, moms)
    h = Handle(emu, latt3, Ne, D, 1, moms)
    h.set_inputs(U_file, V)
    # the library plans the form per handle (edk_plan_form): 7 momenta on planes of 16 sites -> folded plane-wave form
    # (Lx = 4 is below what the separable form covers)
    assert h.query(13) == -1 and h.query(10) == 3 == _capi.plan_form(latt3, D, 1, moms)["form"]
    seen = {"planned form": h.calc()}
    for algo, form, mmas in ((1, 1, 3), (0, 0, 4), (2, 2, 2)):
        h.check(emu.edk_debug_algo(h.h, algo), "edk_debug_algo")
        assert h.query(10) == form and h.query(7) == mmas
        seen[f"algo {algo}"] = h.calc()
    assert h.query(11) == 5 and h.query(3) == 1  # xy-couples (0,0), (0,1), (1,0): modes 1, cos/sin x 2; no split-K in form 2
    h.check(emu.edk_debug_loader(h.h, 1), "edk_debug_loader")
    assert h.query(10) == 0 and h.query(6) == 0
    seen["cp.async loader"] = h.calc()
    h.check(emu.edk_debug_loader(h.h, 0), "edk_debug_loader")
    h.check(emu.edk_debug_use_naive_gram(h.h, 1), "edk_debug_use_naive_gram")
    seen["scalar kernel"] = h.calc()
    h.check(emu.edk_debug_use_naive_gram(h.h, 0), "edk_debug_use_naive_gram")
    assert h.query(10) == 2  # back on the plane-wave form that was selected last
    seen["plane-wave again"] = h.calc()
    h.close()
    for tag, got in seen.items():
        assert worst_block_error(got, ref) < 1e-10, tag
    assert np.array_equal(seen["algo 2"], seen["plane-wave again"])


def test_gemm_form_tile_heights_and_split_k_through_the_launchers(emu):
    """The template dispatch of launch_gram_tma / launch_gram_dmma (smallest and largest tile height, split-K 2) for
    TMA 3M, TMA 4M and the cp.async loader; the full 12 x 2 x 3 sweep runs on the GPU (test_gpu_parity.py)."""
    latt3, Ne, moms = (4, 4, 1), 9, [(0, 0, 0), (1, 0, 0)]  # 16 sites = two 8-site stages, one per K segment
    U_file, V, ref = inputs_and_reference(latt3, Ne, D, 1, moms)
    h = Handle(emu, latt3, Ne, D, 1, moms)
    h.set_inputs(U_file, V)
    for mfrag in (2, 13):
        for algo, loader in ((1, 0), (0, 0), (0, 1)):
            h.check(emu.edk_debug_loader(h.h, loader), "edk_debug_loader")
            h.check(emu.edk_debug_algo(h.h, algo), "edk_debug_algo")
            h.check(emu.edk_debug_gram_config(h.h, mfrag, 2), "edk_debug_gram_config")
            assert h.query(4) == mfrag and h.query(3) == 2
            assert worst_block_error(h.calc(), ref) < 1e-10, (mfrag, algo, loader)
    assert emu.edk_debug_gram_config(h.h, 14, 0) == _capi.EDK_ERR_ARG  # not instantiated
    h.close()


def test_plane_wave_form_selected_by_environment_and_pairing_switches(emu, monkeypatch):
    """EDK_GRAM_ALGO=2 at edk_create (the A/B hook `bench.py --contraction` uses): configure() builds the plane-wave tables
    itself; then both pairing modes (edk_debug_symmetry re-configures: tables are rebuilt for the new job list and
    internal momentum list) and a switch back to the GEMM form.  num_nabla = 2, non-closed momentum list, ragged
    planes of 15 sites."""
    latt3, Ne = (3, 5, 2), 4
    moms = [(0, 0, 1), (1, 0, 0), (1, -1, 0), (0, 2, 1)]
    U_file, V, ref = inputs_and_reference(latt3, Ne, D, 2, moms)
    monkeypatch.setenv("EDK_GRAM_ALGO", "2")
    h = Handle(emu, latt3, Ne, D, 2, moms)
    monkeypatch.delenv("EDK_GRAM_ALGO")
    assert h.query(10) == 2 and h.query(11) >= 1
    h.set_inputs(U_file, V)
    results = {"created with form 2": h.calc()}
    for sym in (1, 0):
        h.check(emu.edk_debug_symmetry(h.h, sym), "edk_debug_symmetry")
        assert h.query(0) == sym and h.query(10) == 2 and h.query(11) >= 1
        assert h.query(2) == (19 if sym else 34)  # pair-GEMMs: Hermitian pairing / distinct direct pairs
        results[f"form 2, pairing {sym}"] = h.calc()
    h.check(emu.edk_debug_algo(h.h, 1), "edk_debug_algo")
    results["GEMM form, direct pairs"] = h.calc()
    h.close()
    for tag, got in results.items():
        assert worst_block_error(got, ref) < 1e-10, tag


def test_plane_wave_multi_tile_and_mirror_tiles(emu):
    """Ne = 64 -> 16 x 40 tiles, 4 x 2 of them; the self pair (W0, W0) of num_nabla = 1 skips the tile below the
    diagonal (e0 = 48 > 39) and the fold kernel reads its mirror; the second f-tile has 3 of 5 f-blocks inside Ne."""
    latt3, Ne, moms = (4, 2, 1), 64, orc.momentum_set(7)
    U_file, V, ref = inputs_and_reference(latt3, Ne, D, 1, moms)
    h = Handle(emu, latt3, Ne, D, 1, moms)
    h.check(emu.edk_debug_algo(h.h, 2), "edk_debug_algo")
    assert h.query(12) == 25 and h.query(0) == 1
    h.set_inputs(U_file, V)
    got = h.calc()
    h.close()
    assert worst_block_error(got, ref) < 1e-10


def test_displacement_mode_both_forms(emu):
    latt3, Ne, moms = (3, 5, 2), 7, orc.momentum_set(9)
    U_file, V, ref = inputs_and_reference(latt3, Ne, X, 3, moms)
    h = Handle(emu, latt3, Ne, X, 3, moms)
    h.set_inputs(U_file, V)
    for algo in (1, 2):
        h.check(emu.edk_debug_algo(h.h, algo), "edk_debug_algo")
        assert worst_block_error(h.calc(), ref) < 1e-10, algo
    h.close()


def test_link_preprocessing_blending_and_host_entry_point(emu):
    """stout smearing + SU(3) projection recorded with edk_set_link_ops, the blending matrix, and edk_calc_host
    (host buffers in, result out, staging inside) with big-endian complex64 eigenvectors and big-endian links, on the
    plane-wave form."""
    latt3, Ne, moms = (4, 4, 2), 4, orc.momentum_set(7)
    latt = list(latt3) + [1]
    U_file = orc.synthetic_links(latt, 5, kind="weak")
    V = orc.synthetic_eigvecs(latt, Ne, 5)
    U = orc.project_su3_timeslice(orc.stout_smear_timeslice(orc.links_file_to_spatial(U_file), 2, 0.1))
    coeff = orc.blending_matrix(Ne, ([6, 5], [2, 2]))
    ref = orc.elemental_timeslice_closed_form(V, U, latt, 1, moms, stocastic_coeff=coeff)
    h = Handle(emu, latt3, Ne, D, 1, moms)
    h.check(emu.edk_debug_algo(h.h, 2), "edk_debug_algo")
    kinds, nsteps, rhos = (C.c_int * 2)(1, 2), (C.c_int * 2)(2, 0), (C.c_double * 2)(0.1, 0.0)
    h.check(emu.edk_set_link_ops(h.h, 2, kinds, nsteps, rhos), "edk_set_link_ops")
    cf = np.ascontiguousarray(coeff, np.float64)
    h.check(emu.edk_set_blending(h.h, cf.ctypes.data, None), "edk_set_blending")
    U_be = np.ascontiguousarray(U_file.astype(">c16"))
    V_be = np.ascontiguousarray(orc.round_through_c8(V).astype(">c8"))
    out = np.full(ref.shape, np.nan + 0j, np.complex128)
    h.check(emu.edk_calc_host(h.h, U_be.ctypes.data, _capi.LINKS_FILE_T | _capi.LINKS_BIG_ENDIAN, V_be.ctypes.data,
                              _capi.EIGVECS_C8 | _capi.EIGVECS_BIG_ENDIAN, out.ctypes.data, None), "edk_calc_host")
    links = np.empty((3, latt3[2], latt3[1], latt3[0], 3, 3), np.complex128)
    h.check(emu.edk_debug_links(h.h, links.ctypes.data, None), "edk_debug_links")
    h.close()
    assert np.abs(links - U).max() < 1e-12
    assert worst_block_error(out, ref) < 1e-10


def test_package_refuses_the_emulator_build(emu):
    """EDK_LIBRARY may point the package at another build of the C ABI (A/B runs of kernel variants), but never at the
    host emulator: there is no CPU path through the product."""
    import sys

    code = "from easydistillation_b200 import _capi; _capi.lib()"
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=REPO,
                       env=dict(os.environ, EDK_LIBRARY=emu._name, PYTHONPATH=REPO))
    assert r.returncode != 0 and "host-emulator test build" in r.stderr, r.stderr[-800:]


def test_laplacian_and_state_errors(emu):
    latt3, Ne, moms = (4, 3, 2), 3, [(0, 0, 0)]
    latt = list(latt3) + [1]
    U_file, V, _ = inputs_and_reference(latt3, Ne, D, 0, moms)
    h = Handle(emu, latt3, Ne, D, 0, moms)
    out = np.empty((1, 1, Ne, Ne), np.complex128)
    assert emu.edk_calc(h.h, out.ctypes.data, None) == _capi.EDK_ERR_STATE  # nothing set yet
    assert b"must be set first" in emu.edk_last_error()
    assert h.query(10) == 1  # one momentum: the GEMM form is the cheapest
    assert emu.edk_debug_algo(h.h, 4) == _capi.EDK_ERR_ARG  # Lx = 4: not covered by the separable form
    assert emu.edk_debug_algo(h.h, 5) == _capi.EDK_ERR_ARG
    h.set_inputs(U_file, V)
    F = np.ascontiguousarray(orc.round_through_c8(V).astype(np.complex128))
    LF = np.empty_like(F)
    h.check(emu.edk_laplacian(h.h, F.ctypes.data, LF.ctypes.data, Ne, None), "edk_laplacian")
    h.close()
    ref = orc.laplacian(F, orc.links_file_to_spatial(U_file))
    assert np.linalg.norm(LF - ref) / np.linalg.norm(ref) < 1e-13


@pytest.mark.parametrize("latt3,Ne,order,moms", [
    ((2, 2, 2), 2, 3, [(0, 0, 0), (1, 0, 1)]),                      # num_nabla = 3: 40 operators, jobs of up to 8 segments
    ((4, 2, 2), 3, 1, [(0, 1, 0), (0, 1, 0), (0, -1, 0), (5, 0, -3)]),  # repeated momenta, |p| > L
    ((3, 2, 1), 1, 2, [(0, 0, 0)]),                                 # one eigenvector, one momentum, Lz = 1
    ((2, 3, 4), 3, 1, [(0, 1, k) for k in range(-5, 6)] + [(0, -1, 2)]),  # 12 momenta in one couple: the z fold works in chunks of 8
])
def test_plane_wave_form_edge_cases(emu, latt3, Ne, order, moms):
    U_file, V, ref = inputs_and_reference(latt3, Ne, D, order, moms)
    h = Handle(emu, latt3, Ne, D, order, moms)
    h.check(emu.edk_debug_algo(h.h, 2), "edk_debug_algo")
    h.set_inputs(U_file, V)
    got = h.calc()
    h.close()
    assert worst_block_error(got, ref) < 1e-10


@pytest.mark.parametrize("tile", ["24", "25", "17"])
def test_folded_form_both_tile_shapes(emu, monkeypatch, tile):
    if tile == "24":
        monkeypatch.setenv("EDK_PW_STAGES", "2")  # A/B hook: the shallowest ring (the wrap-around parity is exercised most)
    """EDK_PW_TILE forces the 16 x 32, the 16 x 40 (re-reads its L fragments per f-block) or the 8 x 56 instance; Ne = 45
    gives several tiles in each case, with the self pair's mirror tiles."""
    latt3, Ne, moms = (4, 2, 1), 45, [(0, 0, 0), (1, 0, 0), (0, -1, 0)]
    U_file, V, ref = inputs_and_reference(latt3, Ne, D, 1, moms)
    monkeypatch.setenv("EDK_PW_TILE", tile)
    monkeypatch.setenv("EDK_GRAM_ALGO", "3")
    h = Handle(emu, latt3, Ne, D, 1, moms)
    assert h.query(10) == 3 and h.query(12) == int(tile)
    h.set_inputs(U_file, V)
    got = h.calc()
    h.close()
    assert worst_block_error(got, ref) < 1e-10


MANY_COUPLES = [(0, 0, 0), (1, 0, 0), (0, 1, 1), (1, 1, 0), (1, -1, 0), (2, 0, 1), (0, 2, 0), (2, 1, 0), (1, 2, 0), (-2, 1, 0),
                (3, 0, 0), (-1, -1, 1)]  # 10 {+q, -q} couples of (px, py): two passes of the folded form, three m-blocks of form 2


@pytest.mark.parametrize("latt3,Ne,mode,order,moms,sym,switch", [
    ((4, 4, 2), 5, D, 1, orc.momentum_set(7), None, True),      # even planes of 16 sites
    ((3, 5, 2), 3, D, 2, [(0, 0, 1), (1, -1, 0), (0, 2, 1)], 1, True),   # odd planes: the middle site is its own partner
    ((3, 5, 1), 3, D, 2, [(1, -1, 0), (0, 2, 1)], 0, False),    # direct pairs: multi-segment jobs with signs
    ((2, 2, 2), 3, D, 1, orc.momentum_set(7), None, False),     # planes of 4 sites: too small to fold, run unfolded
    ((4, 2, 2), 3, D, 1, orc.momentum_set(7), None, False),     # planes of 8 sites: one stage, back run = the whole plane
    ((3, 3, 1), 2, D, 2, orc.momentum_set(33), None, False),    # 33 momenta: 7 couples, 13 modes, plane of 9 sites
    ((5, 4, 1), 2, D, 1, MANY_COUPLES, None, False),            # more than 8 couples: two passes
    ((3, 5, 2), 7, X, 3, orc.momentum_set(9), None, False),     # displacement lines
    ((4, 2, 2), 35, D, 1, orc.momentum_set(7), None, False),    # 3 x 2 tiles of 16 x 32, mirror tile of the self pair
])
def test_folded_plane_wave_form(emu, latt3, Ne, mode, order, moms, sym, switch):
    """Form 3 (centre-symmetric site pairs folded, gram_pwf_kernel) through the C ABI against the oracle, and (`switch`)
    switching between the two plane-wave forms on one handle: their weight tables and z-phase constants differ."""
    U_file, V, ref = inputs_and_reference(latt3, Ne, mode, order, moms)
    h = Handle(emu, latt3, Ne, mode, order, moms)
    if sym is not None:
        h.check(emu.edk_debug_symmetry(h.h, sym), "edk_debug_symmetry")
    h.set_inputs(U_file, V)
    h.check(emu.edk_debug_algo(h.h, 3), "edk_debug_algo")
    assert h.query(10) == 3 and h.query(12) in (24, 25, 17) and h.query(3) == 1
    folded = h.calc()
    assert worst_block_error(folded, ref) < 1e-10
    if switch:
        h.check(emu.edk_debug_algo(h.h, 2), "edk_debug_algo")
        assert worst_block_error(h.calc(), ref) < 1e-10
        h.check(emu.edk_debug_algo(h.h, 3), "edk_debug_algo")
        assert np.array_equal(folded, h.calc())
    h.close()


@pytest.mark.parametrize("latt3,Ne,mode,order,moms,sym", [
    ((8, 4, 2), 5, D, 1, orc.momentum_set(7), None),        # 4 site pairs per stage, 5 separable modes
    ((16, 3, 2), 20, D, 1, orc.momentum_set(9), None),      # 8 pairs per stage, 9 modes, 2 x 1 tiles with a mirror tile
    ((12, 2, 1), 35, D, 1, orc.momentum_set(33), 1),        # 6 pairs per stage, 13 modes, one 32 x 32 tile and both edge strips
    ((8, 4, 1), 4, D, 2, orc.momentum_set(33), 0),          # direct pairs: multi-segment jobs with signs
    ((8, 6, 2), 12, X, 2, orc.momentum_set(19), None),      # displacement lines
    ((8, 4, 1), 7, D, 2, [(0, 0, 0), (0, 0, 1), (0, 1, 1), (1, 1, 1), (0, 0, 2), (0, 1, 2), (1, 1, 2)], None),  # the reference's test list
    ((24, 2, 1), 3, D, 1, [(2, 0, 0), (-1, 1, 3), (0, -2, -1)], None),  # two stages per row, non-closed list, pz > Lz
    ((8, 2, 1), 43, D, 1, orc.momentum_set(7), None),     # Ne = 43: one 32 x 32 tile, two 8 x 128 strip tiles, one 64 x 16
])
def test_separable_form(emu, latt3, Ne, mode, order, moms, sym):
    """Form 4 (gram_sep_kernel + sep_zfold_kernel: swizzled TMA tiles, x transform per site pair, y transform per row)
    through the C ABI against the oracle; it is what edk_create plans for these shapes.  Switching to the folded
    plane-wave form and back on the same handle rebuilds the other form's tables and reproduces the result bit for bit."""
    U_file, V, ref = inputs_and_reference(latt3, Ne, mode, order, moms)
    h = Handle(emu, latt3, Ne, mode, order, moms)
    plan = _capi.plan_form(latt3, mode, order, moms)
    assert plan["separable_available"] and h.query(10) == plan["form"] == 4 and h.query(11) == plan["separable_modes"]
    assert h.query(14) == plan["pairs_per_stage"] and h.query(12) == 3232 and h.query(15) == 6 and h.query(3) == 1 and h.query(7) == 0
    if sym is not None:
        h.check(emu.edk_debug_symmetry(h.h, sym), "edk_debug_symmetry")
        assert h.query(10) == 4 and h.query(0) == sym
    h.set_inputs(U_file, V)
    sep = h.calc()
    assert worst_block_error(sep, ref) < 1e-10
    if Ne <= 12:  # switching forms on a live handle (the larger cases only run the planned form: CPU time)
        h.check(emu.edk_debug_algo(h.h, 3), "edk_debug_algo")
        assert h.query(10) == 3 and h.query(13) == 3
        assert worst_block_error(h.calc(), ref) < 1e-10
        h.check(emu.edk_debug_algo(h.h, -1), "edk_debug_algo")  # back to the planned form
        assert h.query(10) == 4 and h.query(13) == -1
        assert np.array_equal(sep, h.calc())
    h.close()


def test_separable_form_shallow_ring_and_environment(emu, monkeypatch):
    """EDK_SEP_STAGES=2 (A/B hook): the shallowest operand ring, where the wrap-around parity of the full / empty
    barriers is exercised most; EDK_GRAM_ALGO=4 asks for the form at edk_create, and is refused for a momentum list
    outside the instantiated mode structures."""
    latt3, Ne, moms = (16, 2, 2), 9, orc.momentum_set(9)
    U_file, V, ref = inputs_and_reference(latt3, Ne, D, 1, moms)
    monkeypatch.setenv("EDK_SEP_STAGES", "2")
    monkeypatch.setenv("EDK_GRAM_ALGO", "4")
    h = Handle(emu, latt3, Ne, D, 1, moms)
    assert h.query(10) == 4 and h.query(13) == 4
    h.set_inputs(U_file, V)
    assert worst_block_error(h.calc(), ref) < 1e-10
    h.close()
    # EDK_SEP_VARIANT=0 (A/B hook): the first kernel of the form, 1 x 2 elements per lane, accumulators in registers
    monkeypatch.setenv("EDK_SEP_VARIANT", "0")
    h = Handle(emu, latt3, Ne, D, 1, moms)
    assert h.query(10) == 4 and h.query(15) == 0 and h.query(12) == 1632
    h.set_inputs(U_file, V)
    assert worst_block_error(h.calc(), ref) < 1e-10
    h.close()
    monkeypatch.delenv("EDK_SEP_VARIANT")
    # EDK_SEP_LAUNCHES=3 (A/B hook): one launch per tile shape instead of one over the whole tile table
    monkeypatch.setenv("EDK_SEP_LAUNCHES", "3")
    lt, Nt = (8, 2, 1), 43  # Ne = 43: all three tile shapes
    Ut, Vt, reft = inputs_and_reference(lt, Nt, D, 1, orc.momentum_set(7))
    h = Handle(emu, lt, Nt, D, 1, orc.momentum_set(7))
    h.set_inputs(Ut, Vt)
    n0 = emu.edk_launch_count(h.h)
    per_shape = h.calc()
    n_per_shape = emu.edk_launch_count(h.h) - n0
    assert worst_block_error(per_shape, reft) < 1e-10
    h.close()
    monkeypatch.delenv("EDK_SEP_LAUNCHES")
    h = Handle(emu, lt, Nt, D, 1, orc.momentum_set(7))
    h.set_inputs(Ut, Vt)
    n0 = emu.edk_launch_count(h.h)
    assert np.array_equal(h.calc(), per_shape)  # same tiles, same arithmetic: bit-identical
    assert n_per_shape - (emu.edk_launch_count(h.h) - n0) == 2
    h.close()
    mom = np.ascontiguousarray(np.asarray([(3, 0, 0), (0, 0, 1)], np.int32))
    hh = C.c_void_p()
    rc = emu.edk_create(16, 2, 2, 3, D, 1, 2, mom.ctypes.data_as(C.POINTER(C.c_int)), 0, C.byref(hh))
    assert rc == _capi.EDK_ERR_ARG and b"separable" in emu.edk_last_error()
    monkeypatch.delenv("EDK_GRAM_ALGO")
    many = [(3, 0, 0), (0, 0, 1), (1, 0, 0), (0, 1, 0), (0, 0, 2), (1, 1, 0)]
    hp = Handle(emu, (16, 2, 2), 3, D, 1, many)  # planned: |px| = 3 is outside the structures -> folded plane-wave form
    assert hp.query(10) == 3
    hp.close()
    hp = Handle(emu, (16, 2, 2), 3, D, 1, many[:2])  # two momenta: the GEMM form is the cheapest
    assert hp.query(10) == 1
    hp.close()


@pytest.mark.parametrize("name,forms,timeslices", [
    ("deriv_sep_8x4x6x1", (4,), (0,)),          # lattices the separable form covers: 13 / 9 / 9 modes, 4 / 6 / 8 pairs per stage
    ("deriv_sep_12x4x2x1", (4,), (0,)),
    ("disp_sep_16x2x4x1", (4, 2), (0,)),
    ("config1_deriv_weak_4x4x4x8", (3,), (7,)),  # config 1 at its own shape (4^3 x 8, Ne = 20, 7 momenta; distance 8, 6 momenta)
    ("config1_disp_weak_4x4x4x8", (3,), (5,)),
    ("deriv_n1_random_6x4x2x1", (1, 2, 3), (0,)),
    ("deriv_n0_random_6x3x5x1", (2, 3), (0,)),
    ("deriv_blend_4x4x4x1", (2, 3), (0,)),
    ("deriv_weak_4x4x4x2", (2, 3), (1,)),       # the shape of the reference's own tests/test_elemental.py: num_nabla = 2, 7 momenta
    ("disp_random_4x6x8x1", (2, 3), (0,)),
    ("deriv_n3_random_4x4x6x1", (3,), (0,)),    # all 40 third-order operators
    ("deriv_random_4x6x8x1", (3,), (0,)),
    ("disp_weak_4x4x4x2", (2,), (0,)),          # distance 8, the shape of tests/test_displacement_elemental.py
])
def test_reference_goldens_through_every_contraction_form(emu, name, forms, timeslices):
    """Outputs of the UNMODIFIED reference (tests/golden/*.npz, made by oracle/make_golden.py through the reference's
    own loaders and generators) against the library on the host emulator, for the plane-wave forms that have not run
    on hardware yet: parity with the reference itself, not only with the oracle."""
    from conftest import load_golden

    g = load_golden(name)
    latt = [int(v) for v in g["latt_size"]]
    Ne, moms = int(g["Ne"]), [tuple(int(c) for c in m) for m in g["momentum_list"]]
    disp = "distance" in g.files
    h = Handle(emu, latt[:3], Ne, X if disp else D, int(g["distance"] if disp else g["num_nabla"]), moms)
    if "dilution_tot" in g.files:
        coeff = np.ascontiguousarray(orc.blending_matrix(Ne, (list(g["dilution_tot"]), list(g["dilution_used"]))), np.float64)
        h.check(emu.edk_set_blending(h.h, coeff.ctypes.data, None), "edk_set_blending")
    from conftest import golden_timeslices

    stored = {t: i for i, t in golden_timeslices(g)}
    for t in timeslices:
        h.set_inputs(g["U"][t], g["V"][t])
        for form in forms:
            h.check(emu.edk_debug_algo(h.h, form), "edk_debug_algo")
            assert worst_block_error(h.calc(), g["E"][stored[t]]) < 1e-10, (name, t, form)
    h.close()


def test_integration_md_ctypes_stub_runs_as_written(emu):
    """INTEGRATION.md section B shows the ctypes stub a reference maintainer would add (lattice/generator/_edk.py).
    The code block is executed here verbatim, bound to the emulator build of the same C ABI, and its
    EdkHandle.calc(U_t, V_t, out) must reproduce the reference algorithm for one timeslice."""
    import re

    text = open(os.path.join(REPO, "INTEGRATION.md")).read()
    block = re.search(r"```python\n(# lattice/generator/_edk\.py.*?)```", text, re.S).group(1)
    assert 'C.CDLL("libedk_sm100a.so")' in block
    ns = {}
    exec(compile(block.replace('C.CDLL("libedk_sm100a.so")', f"C.CDLL({emu._name!r})"), "INTEGRATION.md", "exec"), ns)
    latt, Ne, moms = [4, 2, 2, 1], 3, [(0, 0, 0), (0, 1, 1)]
    U_file, V, ref = inputs_and_reference(latt[:3], Ne, D, 1, moms)
    V8 = np.ascontiguousarray(V.astype(np.complex64))  # the reference's staging buffer is complex64 (elemental.py:55)
    out = np.zeros(ref.shape, np.complex128)
    handle = ns["EdkHandle"](latt, Ne, ns["EDK_MODE_DERIVATIVE"], 1, moms)
    handle.calc(np.ascontiguousarray(U_file), V8, out)
    assert worst_block_error(out, ref) < 1e-10
    with pytest.raises(ValueError):  # EDK_ERR_ARG maps to the reference's ValueError
        ns["EdkHandle"](latt, 0, ns["EDK_MODE_DERIVATIVE"], 1, moms)


def test_reference_class_bound_to_the_c_abi_matches_its_own_numpy_path(emu):
    """Where the reference checkout exists (the build container): the UNMODIFIED lattice.ElementalGenerator with the
    three touch points of INTEGRATION.md B and the stub of that section, fed by the reference's own loaders, against
    the reference's own numpy calc(t) (tests/emu/reference_binding_check.py, in a child process because importing
    the reference needs the oracle's shims on sys.path)."""
    import sys

    if not os.path.isdir(os.environ.get("EDK_REFERENCE_ROOT", "/root/reference")):
        pytest.skip("no reference checkout on this machine")
    r = subprocess.run([sys.executable, os.path.join(REPO, "tests", "emu", "reference_binding_check.py"), emu._name],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "EDK_BINDING_OK" in r.stdout, (r.stdout[-1000:], r.stderr[-3000:])


def memcheck_cases(lib):
    """Small runs that touch every kernel family; executed by the AddressSanitizer test below in a child process."""
    cases = [((4, 4, 2), 5, D, 1, orc.momentum_set(7), (1, 0, 2)),            # stencil + GEMM forms + plane-wave form
             ((4, 2, 2), 3, D, 1, orc.momentum_set(7), (3,)),                   # folded form, planes of exactly one stage
             ((3, 5, 1), 3, D, 2, [(1, -1, 0), (0, 2, 1)], (3, 1)),            # ragged / odd planes, second-order fields
             ((3, 5, 2), 7, X, 2, orc.momentum_set(9), (2,)),                   # displacement lines
             ((4, 2, 1), 35, D, 1, orc.momentum_set(7), (3,)),                  # multi-tile plane-wave runs with mirror tiles
             ((8, 2, 1), 35, D, 1, orc.momentum_set(9), (4,)),                  # separable form: multi-tile, mirror tile, idle edge warps
             ((12, 2, 1), 5, X, 2, orc.momentum_set(33), (4,))]                 # separable form: 6 pairs per stage, 13 modes
    worst = 0.0
    for latt3, Ne, mode, order, moms, algos in cases:
        U_file, V, ref = inputs_and_reference(latt3, Ne, mode, order, moms)
        h = Handle(lib, latt3, Ne, mode, order, moms)
        h.set_inputs(U_file, V)
        for algo in algos:
            h.check(lib.edk_debug_algo(h.h, algo), "edk_debug_algo")
            worst = max(worst, worst_block_error(h.calc(), ref))
        h.close()
    return worst


def test_emulated_library_is_clean_under_address_sanitizer(tmp_path):
    """compute-sanitizer memcheck without a GPU: the emulator build with -fsanitize=address (every cudaMalloc is a
    separate redzoned heap block, kernels touch it through plain pointers) runs the stencil, both GEMM arithmetics,
    the plane-wave form and the displacement lines; any out-of-bounds global access of a kernel or of the host glue
    aborts the child."""
    import sys

    libasan = subprocess.run(["g++", "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip()
    if not os.path.isabs(libasan) or not os.path.exists(libasan):
        pytest.skip("no AddressSanitizer runtime for this g++")
    so = build_emulator_library(tmp_path, ("-g1", "-fsanitize=address", "-fno-omit-frame-pointer"))
    code = ("import sys; sys.path[:0] = [%r, %r]; import test_emu_library as T; "
            "w = T.memcheck_cases(T.load_emulator_library(%r)); print('EDK_ASAN_OK', w); assert w < 1e-10"
            % (REPO, os.path.join(REPO, "tests"), so))
    env = dict(os.environ, LD_PRELOAD=libasan, ASAN_OPTIONS="detect_leaks=0:halt_on_error=1:abort_on_error=0")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=1500)
    assert r.returncode == 0 and "EDK_ASAN_OK" in r.stdout, (r.stdout[-1500:], r.stderr[-4000:])

