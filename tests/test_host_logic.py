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


def test_preset_signatures_follow_the_reference():
    """Constructor signatures of lattice/preset.py: EigenvectorHostmem(host_ndarray, shape, totNe) (:77-89, ValueError on
    a shape mismatch), GaugeFieldBinary(prefix, suffix, shape=[128, 16^3, 4, 3, 3], dtype="<f8") (:162-170)."""
    import inspect

    V = np.zeros((2, 5, 2, 2, 2, 3), np.complex64)
    ev = edb.EigenvectorHostmem(V, [2, 5, 2, 2, 2, 3], 4)       # the reference's positional form (a list is accepted)
    assert ev.Ne == 4 and ev.load("any")[1, 2].shape == (2, 2, 2, 3)
    assert edb.EigenvectorHostmem(V, (2, 5, 2, 2, 2, 3)).Ne == 5  # totNe defaults to what is stored
    assert edb.EigenvectorHostmem(V, totNe=3).Ne == 3 and edb.EigenvectorHostmem(V, 3).Ne == 3  # this package's earlier form
    with pytest.raises(ValueError, match="does not match expected shape"):
        edb.EigenvectorHostmem(V, [2, 5, 8, 3], 5)
    flat = np.zeros((2, 5, 8, 3), np.complex64)                  # the flattened [Lt, Ne, Lz*Ly*Lx, Nc] layout
    assert edb.EigenvectorHostmem(flat, [2, 5, 8, 3], 5).load()[0].shape == (5, 8, 3)
    sig = inspect.signature(edb.GaugeFieldBinary.__init__).parameters
    assert sig["shape"].default == [128, 16**3, 4, 3, 3] and sig["dtype"].default == "<f8"
    assert list(inspect.signature(edb.EigenvectorHostmem.__init__).parameters)[1:] == ["host_ndarray", "shape", "totNe"]


def test_device_resident_handles_index_like_the_file_handles():
    """GaugeFieldDevice / EigenvectorDevice wrap per-timeslice tensors (any object with the tensor interface the
    generators use); `[:]` returns the indexable, `[t]` a timeslice, `[t, e]` one eigenvector; cyclic repeats."""
    import torch

    U = [torch.zeros((2, 2, 2, 4, 3, 3), dtype=torch.complex128) + t for t in range(2)]
    V = [torch.zeros((3, 2, 2, 2, 3), dtype=torch.complex64) + t for t in range(2)]
    g = edb.GaugeFieldDevice(U, cyclic=True).load("k")
    assert g[:] is g and g.device_resident and g[5] is U[1] and g.file == "<device>"
    ev = edb.EigenvectorDevice(V, cyclic=True)
    assert ev.Ne == 3 and ev.load("k")[4] is V[0] and ev.load("k")[3, 1].shape == (2, 2, 2, 3)
    with pytest.raises(IndexError):
        edb.EigenvectorDevice(V).load("k")[2]
    with pytest.raises(ValueError):
        edb.EigenvectorDevice(V, totNe=4)


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
# the streamed form calc_all() uses: chunks of finished timeslices travel while the next ones are "computed"; the
# destination (here rank 1, which owns fewer timeslices than rank 0) writes its own in place
from easydistillation_b200.sharding import TimesliceGatherer
Lt = 7
full = (torch.arange(Lt * 4, dtype=torch.float64).reshape(Lt, 2, 2) * (2 - 1j)).to(torch.complex128)
g = TimesliceGatherer(Lt, (2, 2), torch.complex128, "cpu", dst=1, chunk=2)
assert (g.t0, g.t1) == timeslice_range(Lt, rank, size) and g.n_local == (4 if rank == 0 else 3)
sent = 0
for i in range(g.n_local):
    g.local[i] = full[g.t0 + i]
    if (i + 1) % g.chunk == 0 or i + 1 == g.n_local:
        g.push(sent, i + 1)
        sent = i + 1
out = g.finish()
if rank == 1:
    assert torch.equal(out, full), "streamed gather"
    assert out.data_ptr() == g.out.data_ptr() and g.local.data_ptr() == out[g.t0:].data_ptr()  # one buffer, own slab in place
else:
    assert out is None
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
    # slabs in any order: creation is separate from slab writing, and an existing file of the right shape is never
    # truncated (a later slab - or the slab that holds t = 0 - must not wipe what another rank has written)
    gen.calc_to_file(handle, "ooo", t_range=(5, 10))
    gen.calc_to_file(handle, "ooo", t_range=(0, 5))
    assert np.array_equal(np.load(tmp_path / "ooo.elemental.npy"), full.transpose(1, 2, 0, 3, 4))
    gen.calc_to_file(handle, "ooo", t_range=(0, 2))  # rewriting a slab leaves the others alone
    assert np.array_equal(np.load(tmp_path / "ooo.elemental.npy"), full.transpose(1, 2, 0, 3, 4))
    assert handle.ensure("ooo", [Nop, Nmom, Lt, Ne, Ne]) is False and handle.ensure("fresh", [Nop, Nmom, Lt, Ne, Ne]) is True
    assert not [f for f in os.listdir(tmp_path) if ".tmp." in f]
    with pytest.raises(IndexError):
        gen.calc_to_file(handle, "ooo", t_range=(8, 12))
    with pytest.raises(ValueError):
        gen.calc_to_file(handle, "ooo", t_range=(0, 5), shard=True)
    # a whole-file run (the one coordinating process) replaces a stale file of another shape
    np.save(tmp_path / "stale.elemental.npy", np.zeros((2, 2), np.complex128))
    with pytest.raises(ValueError):
        gen.calc_to_file(handle, "stale", t_range=(0, 5))
    gen.calc_to_file(handle, "stale")
    assert np.array_equal(np.load(tmp_path / "stale.elemental.npy"), full.transpose(1, 2, 0, 3, 4))

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


def test_contraction_form_plan_host_logic():
    """edk_plan_form: the form a handle runs is planned from the FP64-pipe work per (e, f, site) of each form - no
    environment variable, no run-time trial.  BASELINE.json's configurations all take the separable form; one or two
    momenta take the GEMM form; lattices / momentum lists the separable kernel does not cover take the folded
    plane-wave form, toy planes of fewer than 8 sites the GEMM form."""
    D, X = _capi.MODE_DERIVATIVE, _capi.MODE_DISPLACEMENT
    from oracle.elemental_oracle import momentum_set

    for latt3, order, nmom, pairs in (((16, 16, 16), 1, 9, 8), ((24, 24, 24), 2, 33, 6), ((32, 32, 32), 2, 33, 8), ((48, 48, 48), 2, 33, 8)):
        p = _capi.plan_form(latt3, D, order, momentum_set(nmom))
        assert p["form"] == 4 and p["separable_available"] and p["pairs_per_stage"] == pairs
        assert p["separable_modes"] == (13 if nmom == 33 else 9) and p["separable_qmax"] == (2 if nmom == 33 else 1)
    assert _capi.plan_form((24, 24, 24), X, 2, momentum_set(33))["form"] == 4
    assert _capi.plan_form((48, 48, 48), D, 2, [(0, 0, 0)])["form"] == 1                     # one momentum: 9 slots per pair
    assert _capi.plan_form((48, 48, 48), D, 0, momentum_set(7))["separable_modes"] == 5      # |p|^2 <= 1: the 5-mode structure
    config1 = [(0, 0, 0), (0, 0, 1), (0, 1, 1), (1, 1, 1), (0, 0, 2), (0, 1, 2), (1, 1, 2)]  # the reference's test list
    assert _capi.plan_form((4, 4, 4), D, 2, config1)["form"] == 3 and not _capi.plan_form((4, 4, 4), D, 2, config1)["separable_available"]
    assert _capi.plan_form((8, 8, 8), D, 2, config1) == {"form": 4, "separable_available": True, "separable_qmax": 1, "separable_r2": 2,
                                                         "pairs_per_stage": 4, "separable_modes": 9}
    assert _capi.plan_form((20, 20, 20), D, 2, momentum_set(33))["form"] == 3                 # Lx/2 = 10: no stage width divides it
    assert _capi.plan_form((16, 16, 16), D, 2, [(1, 2, 0), (0, 0, 1), (2, 2, 1), (3, 0, 0), (0, 1, 0)])["form"] == 3  # outside |p_xy|^2 <= 4
    assert _capi.plan_form((2, 2, 8), D, 2, momentum_set(33))["form"] == 1                    # planes of 4 sites
    with pytest.raises(ValueError):
        _capi.plan_form((0, 4, 4), D, 1, [(0, 0, 0)])


def test_separable_tile_table_covers_every_element_once():
    """edk_plan_tiles: 32 x 32 tiles where whole tiles fit, 8 x 128 / 64 x 16 tiles along the ragged edges; every
    element in exactly one tile, and (the point of the table) few tiles with idle warps: the CTA time it stands for
    stays within a few per cent of the useful work at the production Ne."""
    for Ne in (1, 7, 8, 20, 33, 64, 70, 96, 97, 100, 129, 200, 256, 300):
        t = _capi.plan_tiles(Ne)
        cover = np.zeros((Ne, Ne), int)
        for e0, f0, rows, cols in t:
            assert rows >= 1 and cols >= 1 and e0 + rows <= Ne and f0 + cols <= Ne
            assert (rows, cols) <= (32, 32) or rows <= 8 or cols <= 16
            cover[e0:e0 + rows, f0:f0 + cols] += 1
        assert (cover == 1).all(), Ne
    assert len(_capi.plan_tiles(256)) == 64 and len(_capi.plan_tiles(200)) == 36 + 2 + 3

    def cta_time(Ne):  # a tile costs the time of its busiest warp: a lane owns 2 x 2 elements, 4 rows and 8 columns apart
        total = 0.0
        for e0, f0, rows, cols in _capi.plan_tiles(Ne):
            total += (0.5 if rows <= 4 else 1.0) * (0.5 if cols <= 8 else 1.0)
        return total

    assert cta_time(200) / (200 * 200 / 1024) < 1.02 and cta_time(100) / (100 * 100 / 1024) < 1.10


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
    # what calc_to_file's writer does: a batch [t, Nop, Nmom, Ne, Ne] dropped, transposed, into the time slab of the
    # [Nop, Nmom, Lt, Ne, Ne] file mapping; also down-cast to the complex64 the reference stores
    batch = (rng.standard_normal((3, 5, 4, 12, 12)) + 1j * rng.standard_normal((3, 5, 4, 12, 12))).astype(np.complex128)
    for dt in (np.complex128, np.complex64):
        filed = np.zeros((5, 4, 7, 12, 12), dt)
        parallel_copy(filed[:, :, 2:5], batch.transpose(1, 2, 0, 3, 4), min_bytes=1)
        assert np.array_equal(filed[:, :, 2:5], batch.transpose(1, 2, 0, 3, 4).astype(dt))
        assert not filed[:, :, :2].any() and not filed[:, :, 5:].any()


def test_bench_contraction_accounting():
    """The work accounting behind bench.py's roofline object, for every contraction form (config-5 numbers)."""
    import importlib.util

    spec = importlib.util.spec_from_file_location("edk_bench", os.path.join(REPO, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    Ne, latt3 = 200, (48, 48, 48)
    V = 48 ** 3
    gemm = bench.contraction_accounting({"contraction_form": 1, "tma_stages": 3, "real_mma_per_complex_block": 3,
                                         "pair_momentum_gemms": 563, "pair_gemms_per_momentum": 19}, Ne, latt3)
    assert gemm["kernel"] == "gram_tma_kernel" and gemm["padded"] is None and gemm["slots"] == gemm["executed"] / 2
    assert abs(gemm["executed"] / 4.48e13 - 1) < 2e-3  # 4.48e13 executed flops per timeslice at config 5
    q = {"contraction_form": 2, "pair_gemms_per_momentum": 19, "plane_wave_modes": 13, "plane_wave_tile": 24}
    pw = bench.contraction_accounting(q, Ne, latt3)
    assert pw["kernel"] == "gram_pw_kernel"
    assert pw["executed"] == 19.0 * Ne * Ne * V * (24 + 4 * 13) and pw["padded"] == 19.0 * Ne * Ne * V * (24 + 4 * 16)
    q["contraction_form"] = 3
    pwf = bench.contraction_accounting(q, Ne, latt3)
    assert pwf["kernel"] == "gram_pwf_kernel"
    assert pwf["executed"] == 19.0 * Ne * Ne * V * (24 + 2 + 2 * 13) and pwf["padded"] == 19.0 * Ne * Ne * V * (24 + 2 + 2 * 16)
    assert pwf["slots"] == 19.0 * Ne * Ne * V * (12 + 2 + 16)  # 12 DFMA + 2 DADD + 2 x 8 DMMA lane-slots per site
    # the 4 self pairs skip their tiles below the diagonal (16 x 32 tiles: 16 896 of 40 000 elements each)
    skipped = bench.contraction_accounting(q, Ne, latt3, self_pairs=4)
    elems = bench._executed_elements(Ne, 16, 32, True)
    assert elems == 40000 - 16896 and skipped["executed"] == (15 * 40000 + 4 * elems) * V * (24 + 2 + 2 * 13.0)
    q["plane_wave_modes"] = 19  # 10 couples: two passes of a cos and a sin block
    assert bench.contraction_accounting(q, Ne, latt3)["padded"] == 19.0 * Ne * Ne * V * (24 + 2 + 2 * 32)
    # separable form: 19 operations per site (12 site product, 2 sum / difference, 1 + 4 x-stage), 26 per row and element
    sep = bench.contraction_accounting({"contraction_form": 4, "pair_gemms_per_momentum": 19, "plane_wave_modes": 13,
                                        "plane_wave_tile": 3232}, Ne, latt3)
    assert sep["kernel"] == "gram_sepx_kernel" and sep["padded"] is None
    assert sep["slots"] == 19.0 * Ne * Ne * V * (19 + 26 / 48) and sep["executed"] == 19.0 * Ne * Ne * V * (33 + 42 / 48)
    sep_self = bench.contraction_accounting({"contraction_form": 4, "pair_gemms_per_momentum": 19, "plane_wave_modes": 13,
                                             "plane_wave_tile": 3232}, Ne, latt3, self_pairs=4)
    assert sep_self["slots"] == (15 * 40000 + 4 * bench._executed_elements(Ne, 8, 16, True)) * V * (19 + 26 / 48)
    assert bench._executed_elements(Ne, 8, 8, True) == 25 * 26 // 2 * 64  # 8 x 8 blocks on and above the diagonal ...
    assert bench._executed_elements(Ne, 8, 16, True) == 20800 + 12 * 64   # ... plus the lower block of the 12 warp tiles across it
    assert bench.stencil_bytes_moved("config5", planes=False) == bench.algorithmic("config5")[1]
    assert bench.stencil_bytes_moved("config5") == bench.algorithmic("config5")[1] + 3 * Ne * V * 24.0
    old = bench.contraction_accounting({"contraction_form": 0, "tma_stages": 0, "real_mma_per_complex_block": 4,
                                        "pair_momentum_gemms": 10, "pair_gemms_per_momentum": 1}, 8, (4, 4, 4))
    assert old["kernel"] == "gram_dmma_kernel" and old["executed"] == 2.0 * 4 * 64 * 3 * 64 * 10
    # both arms print the same workload description
    assert bench.config_dict("config5", None, 20, 8) == bench.config_dict("config5", None, 20, 8)
    assert set(bench.config_dict("config3", 2, 3, 1)) == {"workload", "lattice", "Ne", "distance", "momenta", "sharding", "l2"}


@pytest.mark.parametrize("form,dist_", [(1, None), (2, None), (3, None), (4, None), (1, 2), (4, 2)])
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
    q = {"hermitian_pairing": dist_ is None, "internal_momenta": 33, "pair_gemms_per_momentum": 19 if dist_ is None else 3,
         "ksplit": 4 if form == 1 else 1, "mfrag": 13, "jobs": 19, "tma_stages": 3,
         "real_mma_per_complex_block": 3 if form == 1 else max(0, 4 - form), "pair_momentum_gemms": 563,
         "half_set_momenta": 17, "contraction_form": form, "plane_wave_modes": 13 if form > 1 else 0,
         "plane_wave_tile": {1: 0, 2: 25, 3: 24, 4: 3232}[form], "form_requested": -1, "pairs_per_stage": 8 if form == 4 else 0}
    K = 3
    prof = {"prepare": {"ms": 1.5, "launches": 2 * K}, "stencil": {"ms": 12.0, "launches": 4 * K},
            "contraction": {"ms": {1: 4173.0, 2: 600.0, 3: 450.0, 4: 360.0}[form], "launches": K}, "combine": {"ms": 20.0, "launches": 2 * K}}
    mode_, order_ = (_capi.MODE_DERIVATIVE, nabla) if dist_ is None else (_capi.MODE_DISPLACEMENT, dist_)
    ns = dict(vars(bench))
    ns.update(name=name, dist_=dist_, Lx=Lx, Ly=Ly, Lz=Lz, Ne=Ne, nabla=nabla, nmom=nmom, V=Lx * Ly * Lz, K=K, W=3, world=1, rank=0,
              prof=prof, q=q, ms=640.0, ms_max=640.0, value=K / 0.64, e2e_value=4.4, h2d=595000000, d2h=274560000, checksum=1.0,
              launches=27, workspace_mb=31000.0, host_queue_ms=1.2, W0_host=None, U_sp_host=None, dmma_tf=37.0, dfma_tf=36.3, _capi=_capi,
              mode_=mode_, order_=order_, moms=bench.momentum_set(nmom), files={"timeslices": 2}, forced=None,
              parity={"against": "GEMM form", "tolerance": 1e-10, "worst_block_rel_err": 3e-14},
              contraction={"requested": "auto", "reason": "stand-in"}, torch=None, dev=None, args=types.SimpleNamespace(),
              clocks=types.SimpleNamespace(summary=lambda: {"sm_mhz": 1965.0, "sm_max_mhz": 1965, "reasons": []}),
              fp64_gemm_peak=lambda torch, dev: {"dgemm": 35.5, "zgemm": 36.8}, print=lambda *a, **k: None)
    exec(code, ns)
    line = json.loads(json.dumps(ns["line"]))
    assert line["metric"] == "elemental_timeslices_per_sec" and line["unit"] == "timeslices/s" and line["n_gpus"] == 1
    for key in ("value", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config", "e2e",
                "gpu_launches", "clocks", "roofline", "cpu_baseline", "workspace_MB", "host_queue_ms", "api"):
        assert key in line, key
    assert line["config"] == bench.config_dict(name, dist_, K, 1)  # what `--impl reference` prints too
    roof = line["roofline"]
    assert roof["bound"] == "tensor" and 0 < roof["frac"] < 1.5 and roof["unit"] == "TFLOP/s" and roof["peak"] == 36.8
    assert 0 < roof["fp64_pipe_utilisation"] < 1.2
    assert roof["kernel"].startswith({1: "gram_tma_kernel", 2: "gram_pw_kernel", 3: "gram_pwf_kernel", 4: "gram_sepx_kernel"}[form])
    assert ("executed_tflops_incl_padded_mode_rows" in roof) == (form in (2, 3))
    assert line["contraction"]["form"] == form and line["contraction"]["parity_check"]["worst_block_rel_err"] == 3e-14
    st = line["roofline_stencil"]
    planes = 3 * Ne * Lx * Ly * Lz * 24.0
    assert st["algorithmic_bytes_per_launch"] == st["survey_bytes_per_launch"] + (planes if form == 1 and dist_ is None else 0.0)
    assert st["kernel"].startswith("nabla3_kernel" if dist_ is None else "displace_step6_kernel")
    assert line["e2e"]["h2d_bytes_per_step"] == 595000000 and line["cpu_baseline"]["kind"] == "port"
    assert line["e2e"]["from_files"] == {"timeslices": 2}
