"""CPU-only checks: C-ABI surface, host-side logic, sharding over gloo (no kernels are run)."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from conftest import REPO, load_golden

import easydistillation_b200 as edb
from easydistillation_b200 import _capi
from easydistillation_b200.generator.elemental import blending_matrix
from easydistillation_b200.insertion.derivative import derivative, num_derivative
from easydistillation_b200.sharding import timeslice_range


def _header_symbols():
    text = open(os.path.join(REPO, "include", "edk.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(edk_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from easydistillation_b200.build import build

    build()
    lib = ctypes.CDLL(_capi.LIB_PATH)
    names = _header_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/edk.h but not exported"
    assert sorted(_capi.SIGNATURES) == names, "ctypes table and header disagree"
    lib.edk_version.restype = ctypes.c_int
    assert lib.edk_version() == 1


def test_library_is_sm100a_dmma():
    """The shipped cubin is sm_100a and the contraction really is DMMA (FP64 tensor pipe)."""
    out = subprocess.run(["cuobjdump", "-lelf", _capi.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    sass = subprocess.run(["cuobjdump", "-sass", _capi.LIB_PATH], capture_output=True, text=True).stdout
    assert sass.count("DMMA.8x8x4") >= 312


def test_missing_gpu_fails_loudly():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    U = np.zeros((1, 2, 2, 2, 4, 3, 3), np.complex128)
    V = np.zeros((1, 2, 2, 2, 2, 3), np.complex128)
    with pytest.raises(_capi.EdkError):
        edb.ElementalGenerator([2, 2, 2, 1], edb.GaugeFieldHostmem(U), edb.EigenvectorHostmem(V), 1)
    with pytest.raises(_capi.EdkError):
        edb.MomentumPhase([2, 2, 2, 1]).get((0, 0, 1))


def test_derivative_map_matches_reference_golden():
    g = load_golden("insertion_maps")
    for n, row in enumerate(g["derivative_tuples"]):
        assert derivative(n) == tuple(int(v) for v in row if v >= 0)
    assert [num_derivative(n) for n in range(4)] == [1, 4, 13, 40]
    with pytest.raises(ValueError):
        derivative(-1)


def test_blending_matrix_matches_reference_output():
    # printed by the reference for dilution=([10, 7], [4, 2]) when the golden was generated
    c = blending_matrix(6, ([10, 7], [4, 2]))
    assert c[0, 0] == 2.5 and c[0, 1] == 7.5 and c[0, 4] == 8.75 and c[4, 4] == 3.5 and c[4, 5] == 21.0
    assert np.array_equal(c, c.T)
    with pytest.raises(AssertionError):
        blending_matrix(5, ([10, 7], [4, 2]))
    from oracle.elemental_oracle import blending_matrix as ob

    assert np.array_equal(ob(6, ([10, 7], [4, 2])), c)
    assert np.array_equal(ob(6, ([9, 9, 9], 2)), blending_matrix(6, ([9, 9, 9], 2)))


def test_timeslice_ranges_partition():
    for Lt in (1, 7, 8, 96, 128):
        for size in (1, 2, 3, 4, 8):
            got = [timeslice_range(Lt, r, size) for r in range(size)]
            assert got[0][0] == 0 and got[-1][1] == Lt
            assert all(a[1] == b[0] for a, b in zip(got, got[1:]))
            lens = [b - a for a, b in got]
            assert max(lens) - min(lens) <= 1


def test_presets_roundtrip(tmp_path):
    rng = np.random.default_rng(0)
    U = (rng.standard_normal((2, 2, 2, 2, 4, 3, 3)) + 0j).astype("<c16")
    V = (rng.standard_normal((2, 3, 2, 2, 2, 3)) + 0j).astype("<c16")
    prefix = str(tmp_path) + "/"
    U.tofile(prefix + "c.dat")
    np.save(prefix + "c.npy", U)
    np.save(prefix + "c.evec.npy", V)
    g = edb.GaugeFieldBinary(prefix, ".dat", list(U.shape), "<c16").load("c")
    assert np.array_equal(g[:], U) and g.file.endswith("c.dat")
    assert np.array_equal(edb.GaugeFieldNpy(prefix, ".npy").load("c")[1], U[1])
    ev = edb.EigenvectorNpy(prefix, ".evec.npy", list(V.shape), 3)
    assert ev.Ne == 3 and np.array_equal(ev.load("c")[1, 2], V[1, 2])
    el = edb.ElementalNpy(prefix, ".elemental.npy", [4, 1, 2, 3, 3], 3)
    mm = el.create("c", [4, 1, 2, 3, 3])
    mm[:, :, 1] = 2.0 + 1j
    mm.flush()
    back = el.load("c")[:]
    assert back.shape == (4, 1, 2, 3, 3) and back[0, 0, 1, 0, 0] == 2.0 + 1j and back[0, 0, 0, 0, 0] == 0
    with pytest.raises(ValueError):
        edb.GaugeFieldHostmem(np.zeros((2, 2, 2, 2, 3, 3, 3)))
    with pytest.raises(ValueError):
        edb.EigenvectorHostmem(np.zeros((2, 3, 2, 2, 2, 3)), totNe=5)


GLOO_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.environ["EDB_REPO"])
from easydistillation_b200.sharding import gather_timeslices, timeslice_range, world
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:" + os.environ["EDB_PORT"],
                        rank=int(os.environ["EDB_RANK"]), world_size=2)
rank, size = world()
Lt = 5
t0, t1 = timeslice_range(Lt, rank, size)
full = (torch.arange(Lt * 6, dtype=torch.float64).reshape(Lt, 3, 2) * (1 + 0.5j)).to(torch.complex128)
local = full[t0:t1].clone()
got = gather_timeslices(local, Lt, dst=0)
if rank == 0:
    assert got.shape == full.shape and torch.equal(got, full), "gather to rank 0"
else:
    assert got is None
got = gather_timeslices(local, Lt, dst=None)
assert torch.equal(got, full), "all-gather"
dist.destroy_process_group()
print("OK", rank)
"""


def test_gather_timeslices_gloo_world2(tmp_path):
    import socket

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    script = tmp_path / "worker.py"
    script.write_text(GLOO_WORKER)
    procs = []
    for r in range(2):
        env = dict(os.environ, EDB_REPO=REPO, EDB_PORT=str(port), EDB_RANK=str(r))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=180)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and f"OK {r}" in o, o


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (CPU oracle port on a bounded sample) on the smallest workload."""
    import json

    for extra in ([], ["--generator", "displacement"]):
        out = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--workload", "tiny",
                              "--steps", "1", "--warmup", "0"] + extra, capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, out.stderr[-2000:]
        line = json.loads(out.stdout.strip().splitlines()[-1])
        assert line["impl"] == "reference" and line["metric"] == "elemental_timeslices_per_sec"
        assert line["unit"] == "timeslices/s" and line["higher_is_better"] is True and line["value"] > 0
        assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
        assert line["cpu_baseline"]["value"] == line["value"] == line["e2e"]["value"]
        assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
        assert "workload" in line["config"]
    # non-zero ranks of a torchrun launch print nothing and exit 0
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--workload", "tiny",
                          "--steps", "1", "--warmup", "0", "--gpus", "2"], capture_output=True, text=True, env=env, timeout=120)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_contraction_plan_host_logic():
    """edk_plan: pairing decision and pair / momentum counts (pure host code of the library)."""
    from oracle.elemental_oracle import momentum_set

    m33, m9 = momentum_set(33), momentum_set(9)
    p = _capi.plan(_capi.MODE_DERIVATIVE, 2, m33)
    assert p == {"hermitian_pairing": True, "internal_momenta": 33, "half_set_momenta": 17, "pairs_direct": 34,
                 "pairs_paired": 19, "pair_momentum_gemms": 15 * 33 + 4 * 17, "operators": 13, "self_pairs": 4}
    # the reference contracts 43 pairs x 33 momenta = 1419 for the same timeslice
    assert _capi.plan(_capi.MODE_DERIVATIVE, 2, m33, 0)["pair_momentum_gemms"] == 34 * 33
    p = _capi.plan(_capi.MODE_DERIVATIVE, 1, m9)  # 9 momenta are not closed under negation: 2 are added
    assert p["hermitian_pairing"] and p["internal_momenta"] == 11 and p["pairs_direct"] == 7 and p["pairs_paired"] == 4
    assert p["half_set_momenta"] == 6 and p["pair_momentum_gemms"] == 3 * 11 + 1 * 6 and p["operators"] == 4
    # a list with no negatives at all: closing it doubles the momenta, 19 * 8 > 34 * 4, so pairing does not pay
    one_sided = [(1, 0, 0), (0, 1, 0), (0, 0, 1), (1, 1, 0)]
    p = _capi.plan(_capi.MODE_DERIVATIVE, 2, one_sided)
    assert not p["hermitian_pairing"] and p["internal_momenta"] == 4 and p["pair_momentum_gemms"] == 34 * 4
    assert not _capi.plan(_capi.MODE_DERIVATIVE, 1, one_sided)["hermitian_pairing"]  # 4 * 8 > 7 * 4
    p = _capi.plan(_capi.MODE_DERIVATIVE, 2, one_sided, 1)  # forced
    assert p["hermitian_pairing"] and p["internal_momenta"] == 8 and p["half_set_momenta"] == 4
    # no derivative / displacement: nothing to pair
    assert not _capi.plan(_capi.MODE_DERIVATIVE, 0, m33)["hermitian_pairing"]
    p = _capi.plan(_capi.MODE_DISPLACEMENT, 8, m9)
    assert not p["hermitian_pairing"] and p["operators"] == 9 and p["pair_momentum_gemms"] == 9 * 9
    assert _capi.plan(_capi.MODE_DERIVATIVE, 3, m33)["operators"] == 40
    with pytest.raises(ValueError):
        _capi.plan(_capi.MODE_DERIVATIVE, 4, m9)


# ---------------------------------------------------------------------------------------------
# big-endian file formats (ILDG gauge field, QDP timeslice eigenvectors): reader parity against what
# the reference's own readers returned for the same files (oracle/make_golden_files.py)
# ---------------------------------------------------------------------------------------------
def _ildg_qdp_files(tmp_path):
    g = load_golden("files_ildg_qdp")
    g["lime_bytes"].tofile(tmp_path / "cfg.lime")
    g["mod_bytes"].tofile(tmp_path / "cfg.mod")
    return g, str(tmp_path) + "/"


def test_ildg_reader_matches_reference_reader(tmp_path):
    from easydistillation_b200 import preset
    from easydistillation_b200.fileio import ildg_layout, lime_records, write_ildg

    g, prefix = _ildg_qdp_files(tmp_path)
    Lx, Ly, Lz, Lt = (int(v) for v in g["latt_size"])
    assert [r[0] for r in lime_records(prefix + "cfg.lime")] == ["ildg-format", "ildg-binary-data", "ildg-data-lfn"]
    assert ildg_layout(prefix + "cfg.lime")[2:] == ([Lx, Ly, Lz, Lt], 64)
    data = preset.GaugeFieldIldg(prefix, ".lime").load("cfg")
    U = data[:]
    assert U.dtype == np.dtype(">c16") and U.shape == (Lt, Lz, Ly, Lx, 4, 3, 3) and data.latt_size == [Lx, Ly, Lz, Lt]
    assert np.array_equal(U, g["U_ref"])  # value comparison across byte orders
    assert np.array_equal(data[1], g["U_ref"][1])
    # the reference's flattened default shape [Lt, V, 4, 3, 3] and a wrong element count
    flat = preset.GaugeFieldIldg(prefix, ".lime", [Lt, Lz * Ly * Lx, 4, 3, 3]).load("cfg")[:]
    assert np.array_equal(flat.reshape(U.shape), g["U_ref"])
    with pytest.raises(ValueError):
        preset.GaugeFieldIldg(prefix, ".lime", [Lt + 1, Lz, Ly, Lx, 4, 3, 3]).load("cfg")
    # single precision payload, byte-identical rewrite of the double precision file
    write_ildg(prefix + "again.lime", g["U_ref"])
    assert np.array_equal(np.fromfile(prefix + "again.lime", np.uint8), g["lime_bytes"])
    write_ildg(prefix + "single.lime", g["U_ref"], precision=32)
    single = preset.GaugeFieldIldg(prefix, ".lime").load("single")[:]
    assert single.dtype == np.dtype(">c8") and np.array_equal(single, g["U_ref"].astype("<c8"))
    (tmp_path / "bad.lime").write_bytes(b"\0" * 200)
    with pytest.raises(ValueError):
        preset.GaugeFieldIldg(prefix, ".lime").load("bad")


def test_qdp_timeslice_reader_matches_reference_reader(tmp_path):
    import struct

    from easydistillation_b200 import preset
    from easydistillation_b200.fileio import QDP_MAGIC, TimesliceRecords

    g, prefix = _ildg_qdp_files(tmp_path)
    Lx, Ly, Lz, Lt = (int(v) for v in g["latt_size"])
    Ne = int(g["Ne"])
    handle = preset.EigenvectorTimeSlice(prefix, ".mod", [Lt, Ne, Lz, Ly, Lx, 3], Ne)
    data = handle.load("cfg")
    assert handle.Ne == Ne and data.latt_size == [Lx, Ly, Lz, Lt] and handle.load("cfg") is data
    for t in range(Lt):
        block = data[t]
        assert block.dtype == np.dtype(">c8") and block.shape == (Ne, Lz, Ly, Lx, 3)
        assert np.array_equal(block, g["V_ref"][t])
        assert np.array_equal(data[t, Ne - 1], g["V_ref"][t, Ne - 1])
        assert np.array_equal(data[t, 2, 1], g["V_ref"][t, 2, 1])
    with pytest.raises(IndexError):
        data[Lt, 0]
    with pytest.raises(IndexError):
        data[Lt]
    # records stored out of order (no back-to-back run): the gather path must give the same timeslice
    V = g["V_ref"]
    rec = V[0, 0].size * 8
    magic, xml = QDP_MAGIC.encode(), b"<m><lattSize>%d %d %d %d</lattSize><decay_dir>3</decay_dir></m>" % (Lx, Ly, Lz, Lt)
    head = struct.pack(">i", len(magic)) + magic + struct.pack(">ii", 1, len(xml)) + xml
    order = [(t, e) for e in reversed(range(Ne)) for t in range(Lt)]
    data_pos = len(head) + 16
    with open(prefix + "shuffled.mod", "wb") as f:
        f.write(head + struct.pack(">qq", 0, data_pos + len(order) * rec))
        for t, e in order:
            f.write(V[t, e].astype(">c8").tobytes())
        f.write(struct.pack(">I", len(order)))
        for i, (t, e) in enumerate(order):
            f.write(struct.pack(">iii", 8, t, e) + struct.pack(">qq", 0, data_pos + i * rec))
    shuffled = TimesliceRecords(prefix + "shuffled.mod", [Lt, Ne, Lz, Ly, Lx, 3], ">c8")
    assert np.array_equal(shuffled[1], V[1]) and np.array_equal(shuffled[0, 3], V[0, 3])


def test_raw_view_keeps_bytes_and_flags_big_endian():
    from easydistillation_b200 import _capi

    a = np.arange(6, dtype=">c8")
    v, be = _capi.raw_view(a)
    assert be and v.dtype == np.dtype("<c8") and v.tobytes() == a.tobytes() and np.shares_memory(v, a)
    b = np.arange(6, dtype="<c16")
    v, be = _capi.raw_view(b)
    assert not be and v is b


def test_calc_to_file_layout_slabs_and_downcast(tmp_path):
    """Host logic of the elemental file writer (reference layout [Nop, Nmom, Lt, Ne, Ne],
    tests/test_elemental.py:47) with the device part replaced by a stand-in `calc_range`."""
    import types

    from easydistillation_b200.generator._base import _TimesliceGenerator

    Nop, Nmom, Lt, Ne = 3, 2, 10, 4
    rng = np.random.default_rng(5)
    full = rng.standard_normal((Lt, Nop, Nmom, Ne, Ne)) + 1j * rng.standard_normal((Lt, Nop, Nmom, Ne, Ne))
    calls = []

    class Fake(_TimesliceGenerator):
        def calc_range(self, t0, t1):
            calls.append((t0, t1))
            return full[t0:t1].copy()

    gen = object.__new__(Fake)
    gen.latt_size, gen.Ne = [4, 4, 4, Lt], Ne
    gen._engine = types.SimpleNamespace(out_shape=(Nop, Nmom, Ne, Ne))
    handle = edb.ElementalNpy(str(tmp_path) + "/", ".elemental.npy")
    gen.calc_to_file(handle, "cfg")
    assert calls == [(0, 4), (4, 8), (8, 10)]
    stored = np.load(tmp_path / "cfg.elemental.npy")
    assert stored.dtype == np.dtype("<c16") and np.array_equal(stored, full.transpose(1, 2, 0, 3, 4))
    assert np.array_equal(handle.load("cfg")[1, 0, 7], full[7, 1, 0])
    # two ranks writing their slabs into one file, down-cast to the complex64 the reference declares
    calls.clear()
    gen.calc_to_file(handle, "two", t_range=(0, 5), dtype="<c8")
    gen.calc_to_file(handle, "two", t_range=(5, 10), dtype="<c8")
    assert calls == [(0, 4), (4, 5), (5, 9), (9, 10)]
    stored = np.load(tmp_path / "two.elemental.npy")
    assert stored.dtype == np.dtype("<c8") and np.array_equal(stored, full.transpose(1, 2, 0, 3, 4).astype("<c8"))
    with pytest.raises(ValueError):
        gen.calc_to_file(handle, "two", t_range=(5, 10), dtype="<c16")  # existing file has another dtype

    # the reference's other elemental preset: headerless raw binary, shape and dtype declared by the handle
    # (lattice/preset.py:129-137); written here, read back by the REFERENCE's byte layout (numpy.fromfile)
    raw = edb.ElementalBinary(str(tmp_path) + "/", ".meson", [Nop, Nmom, Lt, Ne, Ne], Ne)
    gen.calc_to_file(raw, "bin", t_range=(0, 5))
    gen.calc_to_file(raw, "bin", t_range=(5, 10))
    flat = np.fromfile(tmp_path / "bin.meson", dtype="<c16").reshape(Nop, Nmom, Lt, Ne, Ne)
    assert np.array_equal(flat, full.transpose(1, 2, 0, 3, 4))
    assert np.array_equal(raw.load("bin")[2, 1, 9], full[9, 2, 1])
    with pytest.raises(ValueError):
        gen.calc_to_file(raw, "bin", dtype="<c8")  # a headerless file cannot change dtype silently

    class Broken(Fake):
        def calc_range(self, t0, t1):
            if t0 >= 4:
                raise RuntimeError("device failure")
            return full[t0:t1].copy()

    bad = object.__new__(Broken)
    bad.__dict__.update(gen.__dict__)
    with pytest.raises(RuntimeError, match="device failure"):
        bad.calc_to_file(handle, "broken")


def test_contraction_tuning_falls_back_to_the_gemm_form_without_a_gpu():
    """tuning.select_contraction runs its trial in a child process; anything but a clean, faster, validated
    plane-wave run - here: no CUDA device - must leave the validated GEMM form (1) in place."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from easydistillation_b200 import tuning

    d = tuning.select_contraction((4, 4, 4), 8, _capi.MODE_DERIVATIVE, 2, tuning.momentum_set(7), device=0, timeout=300)
    assert d["form"] == 1 and not d["validated"] and "child failed" in d["reason"]
    saved = os.environ.pop("EDK_GRAM_ALGO", None)
    try:
        assert tuning.apply(d) == 1 and os.environ["EDK_GRAM_ALGO"] == "1"
    finally:
        os.environ.pop("EDK_GRAM_ALGO", None)
        if saved is not None:
            os.environ["EDK_GRAM_ALGO"] = saved


def test_contraction_selection_picks_the_fastest_validated_candidate(monkeypatch):
    """Decision logic of tuning.select_contraction with the child processes replaced by canned reports: a candidate
    that failed validation or crashed is never chosen, the fastest validated one wins only if it beats the GEMM form,
    and the chosen tile shape travels with the decision (tuning.apply exports EDK_GRAM_ALGO / EDK_PW_TILE)."""
    from easydistillation_b200 import tuning

    ok = lambda ms: {"ok": True, "cases": [{"case": "x", "err": 3e-14}], "form1_ms": 100.0, "form2_ms": ms}  # noqa: E731
    canned = {(2, None): ok(40.0), (3, "25"): {"ok": False, "reason": "tuning child failed (exit -6): trap"}, (3, "24"): ok(25.0)}
    calls = []

    def fake(form, tile, *a):
        calls.append((form, tile))
        return dict(canned[(form, tile)])

    monkeypatch.setattr(tuning, "_run_candidate", fake)
    three = ((2, None), (3, "25"), (3, "24"))
    d = tuning.select_contraction((8, 8, 8), 16, 0, 2, [(0, 0, 0)], timeout=900, candidates=three)
    assert calls == [(2, None), (3, "25"), (3, "24")]
    assert d["form"] == 3 and d["tile"] == "24" and d["validated"] and d["form2_ms"] == 25.0 and "4.00x faster" in d["reason"]
    assert [c["ok"] for c in d["candidates"]] == [True, False, True]
    for k in ("EDK_GRAM_ALGO", "EDK_PW_TILE"):
        monkeypatch.delenv(k, raising=False)
    assert tuning.apply(d) == 3 and os.environ["EDK_GRAM_ALGO"] == "3" and os.environ["EDK_PW_TILE"] == "24"
    # validated but slower than the GEMM form: stay on form 1, no tile
    canned[(2, None)], canned[(3, "24")] = ok(140.0), ok(120.0)
    d = tuning.select_contraction((8, 8, 8), 16, 0, 2, [(0, 0, 0)], timeout=900, candidates=three)
    assert d["form"] == 1 and d["tile"] is None and d["validated"] and "not faster" in d["reason"]
    assert tuning.apply(d) == 1 and os.environ["EDK_GRAM_ALGO"] == "1" and "EDK_PW_TILE" not in os.environ
    # a differing result is never chosen, however fast
    canned[(2, None)] = {"ok": False, "cases": [{"case": "x", "err": 2e-3}], "form1_ms": 100.0, "form2_ms": 1.0, "reason": "differs"}
    canned[(3, "24")] = {"ok": False, "reason": "tuning child timed out after 10 s"}
    d = tuning.select_contraction((8, 8, 8), 16, 0, 2, [(0, 0, 0)], timeout=900, candidates=three)
    assert d["form"] == 1 and not d["validated"] and "no plane-wave candidate validated" in d["reason"]
    # no time left: candidates are skipped, not started
    calls.clear()
    d = tuning.select_contraction((8, 8, 8), 16, 0, 2, [(0, 0, 0)], timeout=5)
    assert calls == [] and d["form"] == 1
    os.environ.pop("EDK_GRAM_ALGO", None)


def test_parallel_copy_matches_plain_copy():
    """Host side of the streamed pipeline: the threaded copy of a result into the caller's array equals numpy's own
    copy for contiguous and strided destinations, small and large, and refuses mismatched shapes."""
    from easydistillation_b200.pipeline import parallel_copy

    rng = np.random.default_rng(3)
    for shape in [(13, 5, 40, 40), (1, 3, 8, 8), (3, 2, 200, 200), (7,)]:
        src = (rng.standard_normal(shape) + 1j * rng.standard_normal(shape)).astype(np.complex128)
        dst = np.full(shape, np.nan + 0j)
        parallel_copy(dst, src, min_bytes=1)
        assert np.array_equal(dst, src)
        big = np.full((2,) + shape, np.nan + 0j)
        parallel_copy(big[1], src, min_bytes=1)  # a slab of a larger array, as calc_range uses it
        assert np.array_equal(big[1], src) and np.isnan(big[0]).all()
    wide = np.zeros((4, 6), np.complex128)
    parallel_copy(wide[:, ::2], np.ones((4, 3), np.complex128), min_bytes=1)  # non-contiguous destination
    assert wide[:, ::2].sum() == 12 and wide[:, 1::2].sum() == 0
    with pytest.raises(ValueError):
        parallel_copy(np.zeros((4, 3)), np.zeros((3, 4)))


def test_bench_contraction_accounting():
    """The flop accounting behind bench.py's roofline object, for the three contraction forms (config-5 numbers)."""
    import importlib.util

    spec = importlib.util.spec_from_file_location("edk_bench", os.path.join(REPO, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    Ne, V = 200, 48 ** 3
    gemm = bench.contraction_accounting({"contraction_form": 1, "tma_stages": 3, "real_mma_per_complex_block": 3,
                                         "pair_momentum_gemms": 563}, Ne, V)
    assert gemm["kernel"] == "gram_tma_kernel" and not gemm["plane_wave"] and gemm["padded"] is None
    assert abs(gemm["executed"] / 4.48e13 - 1) < 2e-3  # 4.48e13 executed flops per timeslice at config 5
    q = {"contraction_form": 2, "pair_gemms_per_momentum": 19, "plane_wave_modes": 13}
    pw = bench.contraction_accounting(q, Ne, V)
    assert pw["kernel"] == "gram_pw_kernel" and pw["plane_wave"] and not pw["folded"]
    assert pw["executed"] == 19.0 * Ne * Ne * V * (24 + 4 * 13) and pw["padded"] == 19.0 * Ne * Ne * V * (24 + 4 * 16)
    q["contraction_form"] = 3
    pwf = bench.contraction_accounting(q, Ne, V)
    assert pwf["kernel"] == "gram_pwf_kernel" and pwf["folded"]
    assert pwf["executed"] == 19.0 * Ne * Ne * V * (24 + 2 + 2 * 13) and pwf["padded"] == 19.0 * Ne * Ne * V * (24 + 2 + 2 * 16)
    q["plane_wave_modes"] = 19  # 10 couples: two passes of a cos and a sin block
    assert bench.contraction_accounting(q, Ne, V)["padded"] == 19.0 * Ne * Ne * V * (24 + 2 + 2 * 32)
    assert bench.stencil_bytes_moved("config5", planes=False) == bench.algorithmic("config5")[1]
    assert bench.stencil_bytes_moved("config5") == bench.algorithmic("config5")[1] + 3 * Ne * V * 24.0
    old = bench.contraction_accounting({"contraction_form": 0, "tma_stages": 0, "real_mma_per_complex_block": 4,
                                        "pair_momentum_gemms": 10}, 8, 64)
    assert old["kernel"] == "gram_dmma_kernel" and old["executed"] == 2.0 * 4 * 64 * 3 * 64 * 10


@pytest.mark.parametrize("form,dist_", [(1, None), (2, None), (3, None), (1, 2), (3, 2)])
def test_bench_result_line_assembles_for_every_contraction_form(form, dist_):
    """bench.py's native arm needs a GPU, but the block that turns its measurements into the contract's JSON line is
    plain Python: it is cut out of run_native() here and executed on stand-in measurements for each contraction form,
    so that a renamed variable or a missing key cannot cost the round-end benchmark its result."""
    import ast
    import importlib.util
    import json
    import types

    path = os.path.join(REPO, "bench.py")
    spec = importlib.util.spec_from_file_location("edk_bench_line", path)
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    tree = ast.parse(open(path).read())
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "run_native")
    block = next(n for n in ast.walk(fn) if isinstance(n, ast.If) and ast.unparse(n.test) == "rank == 0"
                 and any("json.dumps" in ast.unparse(b) for b in n.body))
    code = compile(ast.Module(body=block.body, type_ignores=[]), path, "exec")
    name = "config5"
    Lx, Ly, Lz, Lt, Ne, nabla, nmom = bench.WORKLOADS[name]
    q = {"hermitian_pairing": True, "internal_momenta": 33, "pair_gemms_per_momentum": 19, "ksplit": 4 if form == 1 else 1, "mfrag": 13,
         "jobs": 19, "tma_stages": 3, "real_mma_per_complex_block": 3 if form == 1 else 4 - form, "pair_momentum_gemms": 563,
         "half_set_momenta": 17, "contraction_form": form, "plane_wave_modes": 13 if form > 1 else 0, "plane_wave_tile": 25 if form > 1 else 0}
    K = 3
    prof = {"prepare": {"ms": 1.5, "launches": 2 * K}, "stencil": {"ms": 12.0, "launches": 4 * K},
            "contraction": {"ms": {1: 4173.0, 2: 600.0, 3: 450.0}[form], "launches": K}, "combine": {"ms": 20.0, "launches": 2 * K}}
    ns = dict(vars(bench))
    ns.update(name=name, dist_=dist_, Lx=Lx, Ly=Ly, Lz=Lz, Ne=Ne, nabla=nabla, nmom=nmom, V=Lx * Ly * Lz, K=K, W=3, world=1, rank=0,
              prof=prof, q=q, ms=640.0, ms_max=640.0, value=K / 0.64, e2e_value=4.4, h2d=595000000, d2h=274560000, checksum=1.0,
              launches=27, workspace_mb=31000.0, W0_host=None, U_sp_host=None, dmma_tf=37.0, dfma_tf=36.3,
              contraction={"requested": "auto", "form": form, "reason": "stand-in"}, torch=None, dev=None, args=types.SimpleNamespace(),
              clocks=types.SimpleNamespace(summary=lambda: {"sm_mhz": 1965.0, "sm_max_mhz": 1965, "reasons": []}),
              fp64_gemm_peak=lambda torch, dev: {"dgemm": 35.5, "zgemm": 36.8}, print=lambda *a, **k: None)
    exec(code, ns)
    line = json.loads(json.dumps(ns["line"]))
    assert line["metric"] == "elemental_timeslices_per_sec" and line["unit"] == "timeslices/s" and line["n_gpus"] == 1
    for key in ("value", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config", "e2e",
                "gpu_launches", "clocks", "roofline", "cpu_baseline"):
        assert key in line, key
    roof = line["roofline"]
    assert roof["bound"] == "tensor" and 0 < roof["frac"] < 1.5 and roof["unit"] == "TFLOP/s" and roof["peak"] == 36.8
    assert roof["kernel"].startswith({1: "gram_tma_kernel", 2: "gram_pw_kernel", 3: "gram_pwf_kernel"}[form])
    assert ("executed_tflops_incl_padded_mode_rows" in roof) == (form > 1)
    st = line["roofline_stencil"]
    planes = 3 * Ne * Lx * Ly * Lz * 24.0
    assert st["algorithmic_bytes_per_launch"] == st["survey_bytes_per_launch"] + (planes if form == 1 and dist_ is None else 0.0)
    assert st["kernel"].startswith("nabla3_kernel" if dist_ is None else "displace_step6_kernel")
    assert line["e2e"]["h2d_bytes_per_step"] == 595000000 and line["cpu_baseline"]["kind"] == "port"

