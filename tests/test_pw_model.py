"""CPU-only checks of the plane-wave factorised contraction (csrc/edk_gram_pw.cu, edk_debug_algo 2).

The CUDA kernels cannot run here, so two things are pinned instead:
  * the host-side mode plan the library really uses (edk_plan_modes through the C ABI) reproduces the
    reference's phase exp(+2 pi i p.x/L) (lattice/insertion/phase.py:41-46, via the oracle);
  * a lane-level numpy model of gram_pw_kernel / pw_zfold_kernel that transcribes the kernel's index
    arithmetic one to one (shared-memory tile layout filled by the TMA boxes, per-lane byte offsets, the
    m8n8k4 fragment ownership, plane-boundary zero weights, epilogue indices) equals the direct contraction
    sum_x conj(L) phase R of the oracle on ragged shapes.  The model is the specification the kernel was
    written against; the kernel itself is checked on the GPU (tests/test_gpu_parity.py).
"""
import numpy as np
import pytest

from easydistillation_b200 import _capi
from oracle import elemental_oracle as orc

PW_WARPS, KG = 8, 6


def _mode_value(mode, x, y, Lx, Ly):
    qx, qy, kind = (int(v) for v in mode)
    turns = ((qx * x) % Lx) / Lx + ((qy * y) % Ly) / Ly
    return np.sin(2 * np.pi * turns) if kind else np.cos(2 * np.pi * turns)


@pytest.mark.parametrize("count", [1, 7, 9, 33])
def test_mode_plan_reproduces_the_phase(count):
    latt = [6, 4, 5, 1]
    Lx, Ly, Lz = latt[:3]
    moms = orc.momentum_set(count)
    modes, momode = _capi.plan_modes(moms)
    if count == 33:
        assert len(modes) == 13  # 6 couples x (cos, sin) + the constant mode
    z, y, x = np.meshgrid(np.arange(Lz), np.arange(Ly), np.arange(Lx), indexing="ij")
    for p, mom in enumerate(moms):
        mc, ms, sg = (int(v) for v in momode[p])
        w = _mode_value(modes[mc], x, y, Lx, Ly).astype(complex)
        if ms >= 0:
            w = w + 1j * sg * _mode_value(modes[ms], x, y, Lx, Ly)
        else:
            assert mom[0] == 0 and mom[1] == 0
        ph = w * np.exp(2j * np.pi * ((mom[2] * z) % Lz) / Lz)
        assert np.abs(ph - orc.momentum_phase(latt, mom)).max() < 1e-14


def test_mode_plan_general_momenta():
    moms = [(3, -1, 2), (-3, 1, 0), (0, -2, 1), (0, 2, 1), (0, 0, -4), (5, 0, 0)]
    modes, momode = _capi.plan_modes(moms)
    assert modes.tolist() == [[3, -1, 0], [3, -1, 1], [0, 2, 0], [0, 2, 1], [0, 0, 0], [5, 0, 0], [5, 0, 1]]
    assert momode.tolist() == [[0, 1, 1], [0, 1, -1], [2, 3, -1], [2, 3, 1], [4, -1, 0], [5, 6, 1]]


# ---------------------------------------------------------------------------------------------------------
# lane-level model
# ---------------------------------------------------------------------------------------------------------
def _tma_box(fields, field, row0, kd0, rows):
    """cp.async.bulk.tensor.3d box {8 doubles, rows, 1} of the array [nfield][Ne][2*Kc doubles]: out-of-range
    rows / doubles are zero-filled.  Returns [rows][4 complex]."""
    nf, Ne, Kc = fields.shape
    out = np.zeros((rows, 4), complex)
    for r in range(rows):
        for k in range(4):
            kc = kd0 // 2 + k
            assert kd0 % 2 == 0
            if row0 + r < Ne and kc < Kc:
                out[r, k] = fields[field, row0 + r, kc]
    return out


def _weight_tiles(modes, Lx, Ly, kplane, mbtot):
    """pw_weights_kernel: wtiles[kstep][grp][mblock][lane]."""
    wt = np.zeros((kplane, 2, mbtot, 32))
    for k in range(kplane):
        for grp in range(2):
            for mbt in range(mbtot):
                for lane in range(32):
                    mode = mbt * 8 + (lane >> 2)
                    sxy = 8 * k + 4 * grp + (lane & 3)
                    if mode < len(modes) and sxy < Lx * Ly:
                        wt[k, grp, mbt, lane] = _mode_value(modes[mode], sxy % Lx, sxy // Lx, Lx, Ly)
    return wt


def _gram_pw_model(fields, jobs, latt3, moms_int, max_mb=2, PW_EL=2, PW_FL=4):
    """jobs: list of (segments [(Lf, Rf, sign)], nmom).  Returns partial[job][p][e][f]."""
    ROWS_L, ROWS_R = PW_WARPS * PW_EL, 8 * PW_FL
    L_BYTES, R_BYTES = KG * ROWS_L * 64, KG * ROWS_R * 64
    Lx, Ly, Lz = latt3
    nf, Ne, Kc = fields.shape
    A = Lx * Ly
    kplane = (A + 7) // 8
    modes, momode = _capi.plan_modes(moms_int)
    nmodes = len(modes)
    mbtot = (nmodes + 7) // 8
    wt = _weight_tiles(modes, Lx, Ly, kplane, mbtot)
    n_et, n_ft = (Ne + ROWS_L - 1) // ROWS_L, (Ne + ROWS_R - 1) // ROWS_R
    Y = np.full((len(jobs), Lz, nmodes, Ne, Ne), np.nan, complex)
    lane = np.arange(32)
    sidx, n = lane & 3, lane >> 2
    for mb0 in range(0, mbtot, max_mb):
        MB = min(max_mb, mbtot - mb0)
        for job_id, (segs, _) in enumerate(jobs):
            for z in range(Lz):
                for et in range(n_et):
                    for ft in range(n_ft):
                        e0, f0 = et * ROWS_L, ft * ROWS_R
                        if len(segs) == 1 and segs[0][0] == segs[0][1] and e0 > f0 + ROWS_R - 1:
                            continue  # self pair: tile below the diagonal, the fold reads the mirror
                        yre = np.zeros((PW_WARPS, PW_EL, PW_FL, MB, 2, 32))
                        yim = np.zeros_like(yre)
                        cur_sign = 1
                        for (Lf, Rf, sgn) in segs:
                            for kstep in range(kplane):
                                if sgn != cur_sign:
                                    yre, yim, cur_sign = -yre, -yim, sgn
                                # ---- producer: one stage of shared memory, addressed in 16-byte units
                                smem = np.zeros((L_BYTES + R_BYTES) // 16, complex)
                                kd = (z * A + 8 * kstep) * 6
                                for kg in range(KG):
                                    box = _tma_box(fields, Lf, e0, kd + 8 * kg, ROWS_L)
                                    smem[kg * ROWS_L * 4:(kg + 1) * ROWS_L * 4] = box.reshape(-1)
                                    box = _tma_box(fields, Rf, f0, kd + 8 * kg, ROWS_R)
                                    o = L_BYTES // 16 + kg * ROWS_R * 4
                                    smem[o:o + ROWS_R * 4] = box.reshape(-1)
                                wst = wt[kstep][:, mb0:mb0 + MB, :]  # two bulk copies: [grp][MB][32]
                                # ---- consumers
                                for warp in range(PW_WARPS):
                                    for grp in range(2):
                                        offL, offR = [], []
                                        for c in range(3):
                                            kc = (4 * grp + sidx) * 3 + c
                                            kg, kin = kc >> 2, kc & 3
                                            offL.append(kg * (ROWS_L * 64) + warp * PW_EL * 64 + kin * 16)
                                            offR.append(L_BYTES + kg * (ROWS_R * 64) + n * 64 + kin * 16)
                                        for j in range(PW_FL):
                                            rv = [smem[(offR[c] + j * 512) // 16] for c in range(3)]
                                            for i in range(PW_EL):
                                                lv = [smem[(offL[c] + i * 64) // 16] for c in range(3)]
                                                cval = sum(np.conj(lv[c]) * rv[c] for c in range(3))  # per lane
                                                for mb in range(MB):
                                                    a = wst[grp, mb]  # lane holds A[row = lane>>2][k = lane&3]
                                                    Amat = np.zeros((8, 4))
                                                    Amat[n, sidx] = a
                                                    for part, acc in ((cval.real, yre), (cval.imag, yim)):
                                                        Bmat = np.zeros((4, 8))
                                                        Bmat[sidx, n] = part  # lane holds B[k = lane&3][col = lane>>2]
                                                        D = Amat @ Bmat
                                                        for q in range(2):  # lane holds D[lane>>2][2 (lane&3) + q]
                                                            acc[warp, i, j, mb, q] += D[n, 2 * sidx + q]
                        # ---- epilogue
                        for warp in range(PW_WARPS):
                            for mb in range(MB):
                                for ln in range(32):
                                    mode = (mb0 + mb) * 8 + n[ln]
                                    if mode >= nmodes:
                                        continue
                                    for i in range(PW_EL):
                                        e = e0 + warp * PW_EL + i
                                        if e >= Ne:
                                            continue
                                        for j in range(PW_FL):
                                            for q in range(2):
                                                f = f0 + 8 * j + 2 * sidx[ln] + q
                                                if f < Ne:
                                                    Y[job_id, z, mode, e, f] = cur_sign * (
                                                        yre[warp, i, j, mb, q, ln] + 1j * yim[warp, i, j, mb, q, ln])
    for job_id, (segs, _) in enumerate(jobs):  # what pw_zfold_kernel reads must have been written
        self_pair = len(segs) == 1 and segs[0][0] == segs[0][1]
        for e in range(Ne):
            for f in range(Ne):
                mirror = self_pair and (e // ROWS_L) * ROWS_L > (f // ROWS_R) * ROWS_R + ROWS_R - 1
                src = Y[job_id, :, :, f, e] if mirror else Y[job_id, :, :, e, f]
                assert not np.isnan(src).any(), "the fold would read an element of Y that was never written"
    # ---- pw_zfold_kernel
    nmom = len(moms_int)
    partial = np.zeros((len(jobs), nmom, Ne, Ne), complex)
    for job_id, (_, nmom_job) in enumerate(jobs):
        for p in range(nmom_job):
            mc, ms, sg = (int(v) for v in momode[p])
            segs = jobs[job_id][0]
            self_pair = len(segs) == 1 and segs[0][0] == segs[0][1]
            e_idx, f_idx = np.meshgrid(np.arange(Ne), np.arange(Ne), indexing="ij")
            mirror = self_pair & ((e_idx // ROWS_L) * ROWS_L > (f_idx // ROWS_R) * ROWS_R + ROWS_R - 1)

            def read(m, z):
                return np.where(mirror, np.conj(Y[job_id, z, m].T), Y[job_id, z, m])

            for z in range(Lz):
                u = read(mc, z)
                if ms >= 0:
                    u = u + 1j * sg * read(ms, z)
                r = (moms_int[p][2] * z) % Lz
                partial[job_id, p] += np.exp(2j * np.pi * r / Lz) * u
    return partial


@pytest.mark.parametrize("latt3,Ne,nmom,max_mb,fl", [((3, 5, 2), 5, 7, 2, 4), ((4, 2, 3), 19, 33, 2, 4), ((5, 3, 1), 35, 9, 2, 4),
                                                     ((2, 3, 2), 3, 33, 1, 4),   # max_mb = 1: two passes over 13 modes
                                                     ((3, 2, 2), 43, 9, 2, 5)])  # 16 x 40 tiles
def test_lane_model_equals_direct_contraction(latt3, Ne, nmom, max_mb, fl):
    Lx, Ly, Lz = latt3
    V = Lx * Ly * Lz
    rng = np.random.default_rng(1234 + Ne)
    nfield = 3
    fields = rng.standard_normal((nfield, Ne, 3 * V)) + 1j * rng.standard_normal((nfield, Ne, 3 * V))
    moms = orc.momentum_set(nmom)
    jobs = [([(0, 1, 1)], nmom), ([(2, 0, -1), (1, 1, 1), (0, 2, -1)], nmom), ([(2, 2, 1)], max(1, nmom // 2))]
    got = _gram_pw_model(fields, jobs, latt3, moms, max_mb, 2, fl)
    f4 = fields.reshape(nfield, Ne, Lz, Ly, Lx, 3)
    for job_id, (segs, nmom_job) in enumerate(jobs):
        for p in range(nmom_job):
            ph = orc.momentum_phase([Lx, Ly, Lz, 1], moms[p])
            ref = sum(s * orc.gram(f4[a], f4[b], ph) for a, b, s in segs)
            err = np.linalg.norm(got[job_id, p] - ref) / np.linalg.norm(ref)
            assert err < 1e-12, (job_id, p, err)


# ---------------------------------------------------------------------------------------------------------
# the kernel SOURCES on a host emulator (tests/emu): TMA producer warp, mbarrier ring, DMMA fragments, fold
# ---------------------------------------------------------------------------------------------------------
def _build_emulator(tmp_path):
    return _build(tmp_path, "pw_emu.cpp", "pw_emu")


def _run_emulator(exe, tmp_path, fields, jobs, latt3, moms, max_mb, nstages, el=2, fl=4):
    import subprocess

    Lx, Ly, Lz = latt3
    nfield, Ne, Kc = fields.shape
    modes, momode = _capi.plan_modes(moms)
    jraw = np.zeros((len(jobs), 26), np.int32)
    for j, (segs, nmom_job) in enumerate(jobs):
        jraw[j, 0], jraw[j, 1] = len(segs), nmom_job
        for s, (a, b, sg) in enumerate(segs):
            jraw[j, 2 + s], jraw[j, 10 + s], jraw[j, 18 + s] = a, b, sg
    zphase = np.array([[np.exp(2j * np.pi * ((m[2] * z) % Lz) / Lz) for z in range(Lz)] for m in moms])
    inp, out = tmp_path / "in.bin", tmp_path / "out.bin"
    with open(inp, "wb") as f:
        np.array([Lx, Ly, Lz, Ne, nfield, len(jobs), len(moms), len(modes), max_mb, nstages, el, fl], np.int32).tofile(f)
        jraw.tofile(f)
        modes.astype(np.int32).tofile(f)
        momode.astype(np.int32).tofile(f)
        np.ascontiguousarray(zphase).view(np.float64).tofile(f)
        np.ascontiguousarray(fields).view(np.float64).tofile(f)
    r = subprocess.run([exe, str(inp), str(out)], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, f"emulated kernels failed ({r.returncode}): {r.stderr[-2000:]}"
    return np.fromfile(out, np.complex128).reshape(len(jobs), len(moms), Ne, Ne)


@pytest.mark.parametrize("latt3,Ne,nmom,max_mb,nstages,fl", [
    ((5, 8, 2), 35, 33, 2, 0, 4),   # 3 x 2 tiles, 13 modes in one pass, 15 stages through the 8-deep ring, mirror tiles
    ((3, 5, 2), 5, 33, 1, 3, 4),    # ragged plane (15 sites), two passes of one m-block, 3-deep ring
    ((5, 8, 1), 43, 33, 2, 0, 5),   # 16 x 40 tiles: 3 x 2 tiles, mirror tile (e0 = 32 > f0 + 39 is never true: none skipped)
    ((3, 3, 1), 90, 7, 2, 2, 5),    # 16 x 40 tiles with skipped mirror tiles (e0 >= 48), 2-deep ring
    ((2, 2, 2), 1, 1, 2, 0, 4),     # one eigenvector, p = 0 only (a single constant mode), planes of 4 sites
    ((4, 3, 2), 7, [(0, 0, 1), (1, 2, 0), (3, -1, 2), (0, -2, 1), (-5, 0, 7)], 2, 0, 5),  # non-closed list, |p| > L
])
def test_kernel_sources_on_host_emulator(tmp_path, latt3, Ne, nmom, max_mb, nstages, fl):
    """edk_gram_pw.cu itself (not a transcription), compiled by g++ against tests/emu/edk_emu.h and run with one
    host thread per CUDA thread, equals the direct contraction."""
    Lx, Ly, Lz = latt3
    V = Lx * Ly * Lz
    rng = np.random.default_rng(99 + Ne)
    nfield = 3
    fields = rng.standard_normal((nfield, Ne, 3 * V)) + 1j * rng.standard_normal((nfield, Ne, 3 * V))
    moms = orc.momentum_set(nmom) if isinstance(nmom, int) else nmom
    nmom = len(moms)
    jobs = [([(0, 1, 1)], nmom), ([(2, 0, -1), (1, 1, 1), (0, 2, -1)], nmom), ([(2, 2, 1)], max(1, nmom // 2))]
    exe = _build_emulator(tmp_path)
    got = _run_emulator(exe, tmp_path, fields, jobs, latt3, moms, max_mb, nstages, 2, fl)
    f4 = fields.reshape(nfield, Ne, Lz, Ly, Lx, 3)
    for job_id, (segs, nmom_job) in enumerate(jobs):
        for p in range(nmom_job):
            ph = orc.momentum_phase([Lx, Ly, Lz, 1], moms[p])
            ref = sum(s * orc.gram(f4[a], f4[b], ph) for a, b, s in segs)
            err = np.linalg.norm(got[job_id, p] - ref) / np.linalg.norm(ref)
            assert err < 1e-12, (job_id, p, err)


def _build(tmp_path, src, name):
    import os
    import subprocess

    from conftest import REPO

    exe = str(tmp_path / name)
    cmd = ["g++", "-std=c++17", "-O1", "-DEDK_HOST_EMU", "-I", os.path.join(REPO, "tests", "emu"),
           "-I", os.path.join(REPO, "easydistillation_b200", "csrc"), "-I", os.path.join(REPO, "include"),
           "-I", "/usr/local/cuda/include", "-x", "c++", os.path.join(REPO, "tests", "emu", src), "-o", exe, "-lpthread"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    return exe


@pytest.mark.parametrize("kernel", [1, 0, 2])  # TMA 3M (the product path), TMA 4M, cp.async 4M
def test_emulator_calibration_on_gpu_validated_kernels(tmp_path, kernel):
    """The emulator must also reproduce the contraction through the kernels that ARE validated on the B200
    (gram_tma_kernel / gram_dmma_kernel of edk_gram.cu): this pins its TMA-box, mbarrier-parity and m8n8k4 fragment
    semantics to the hardware's, which is what gives the emulated run of edk_gram_pw.cu above its weight."""
    import subprocess

    Lx, Ly, Lz = 3, 5, 2  # V = 30: the last 8-site stage is ragged
    V, Ne, nfield, ksplit = Lx * Ly * Lz, 21, 3, 2
    Vpad = (V + 7) // 8 * 8
    rng = np.random.default_rng(7 + kernel)
    fields = rng.standard_normal((nfield, Ne, 3 * V)) + 1j * rng.standard_normal((nfield, Ne, 3 * V))
    moms = orc.momentum_set(7)
    jobs = [([(0, 1, 1)], 7), ([(2, 0, -1), (1, 1, 1), (0, 2, -1)], 7), ([(2, 2, 1)], 4)]
    jraw = np.zeros((len(jobs), 26), np.int32)
    for j, (segs, nmom_job) in enumerate(jobs):
        jraw[j, 0], jraw[j, 1] = len(segs), nmom_job
        for s, (a, b, sg) in enumerate(segs):
            jraw[j, 2 + s], jraw[j, 10 + s], jraw[j, 18 + s] = a, b, sg
    phase = np.zeros((2, len(moms), Vpad), complex)
    for p, m in enumerate(moms):
        phase[0, p, :V] = orc.momentum_phase([Lx, Ly, Lz, 1], m).reshape(-1)
    phase[1] = -1j * phase[0]
    exe = _build(tmp_path, "gram_emu.cpp", "gram_emu")
    inp, out = tmp_path / "gin.bin", tmp_path / "gout.bin"
    with open(inp, "wb") as f:
        np.array([Lx, Ly, Lz, Ne, nfield, len(jobs), len(moms), kernel, ksplit], np.int32).tofile(f)
        jraw.tofile(f)
        phase.view(np.float64).tofile(f)
        np.ascontiguousarray(fields).view(np.float64).tofile(f)
    r = subprocess.run([exe, str(inp), str(out)], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, f"emulated kernels failed ({r.returncode}): {r.stderr[-2000:]}"
    got = np.fromfile(out, np.complex128).reshape(ksplit, len(jobs), len(moms), Ne, Ne).sum(axis=0)
    f4 = fields.reshape(nfield, Ne, Lz, Ly, Lx, 3)
    for job_id, (segs, nmom_job) in enumerate(jobs):
        for p in range(nmom_job):
            ph = orc.momentum_phase([Lx, Ly, Lz, 1], moms[p])
            ref = sum(s * orc.gram(f4[a], f4[b], ph) for a, b, s in segs)
            err = np.linalg.norm(got[job_id, p] - ref) / np.linalg.norm(ref)
            assert err < 1e-12, (kernel, job_id, p, err)


def test_shared_memory_pipeline_is_race_free_under_thread_sanitizer(tmp_path):
    """racecheck without a GPU: the emulated CTA (one host thread per CUDA thread, TMA writes as plain stores of the
    producer thread, mbarriers as mutex-protected counters) runs under ThreadSanitizer.  The real full/empty protocol
    of gram_pw_kernel must be silent; the same build with waits that do not wait must be reported (self-test)."""
    import os
    import subprocess

    from conftest import REPO

    def build(name, extra):
        exe = str(tmp_path / name)
        cmd = ["g++", "-std=c++17", "-O1", "-g", "-fsanitize=thread", "-DEDK_HOST_EMU", *extra, "-I", os.path.join(REPO, "tests", "emu"),
               "-I", os.path.join(REPO, "easydistillation_b200", "csrc"), "-I", os.path.join(REPO, "include"),
               "-I", "/usr/local/cuda/include", "-x", "c++", os.path.join(REPO, "tests", "emu", "pw_emu.cpp"), "-o", exe, "-lpthread"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            pytest.skip("no ThreadSanitizer runtime for this g++: " + r.stderr[-300:])
        return exe

    Lx, Ly, Lz, Ne, nfield = 3, 5, 1, 20, 3
    rng = np.random.default_rng(5)
    fields = rng.standard_normal((nfield, Ne, 3 * Lx * Ly * Lz)) + 1j * rng.standard_normal((nfield, Ne, 3 * Lx * Ly * Lz))
    moms = orc.momentum_set(7)
    jobs = [([(0, 1, 1)], 7), ([(2, 0, -1), (1, 1, 1)], 7), ([(2, 2, 1)], 4)]
    env = dict(os.environ, TSAN_OPTIONS="halt_on_error=0 exitcode=0")
    counts = {}
    for name, extra in (("good", []), ("broken", ["-DEDK_EMU_BREAK_PROTOCOL"])):
        exe = build("pw_emu_tsan_" + name, extra)
        sub = tmp_path / name
        sub.mkdir()
        # same input writer as the plain emulator run; 2-deep ring so that slots are reused many times
        modes, momode = _capi.plan_modes(moms)
        jraw = np.zeros((len(jobs), 26), np.int32)
        for j, (segs, nmom_job) in enumerate(jobs):
            jraw[j, 0], jraw[j, 1] = len(segs), nmom_job
            for s, (a, b, sg) in enumerate(segs):
                jraw[j, 2 + s], jraw[j, 10 + s], jraw[j, 18 + s] = a, b, sg
        zphase = np.array([[np.exp(2j * np.pi * ((m[2] * z) % Lz) / Lz) for z in range(Lz)] for m in moms])
        with open(sub / "in.bin", "wb") as f:
            np.array([Lx, Ly, Lz, Ne, nfield, len(jobs), len(moms), len(modes), 2, 2, 2, 4], np.int32).tofile(f)
            jraw.tofile(f)
            modes.astype(np.int32).tofile(f)
            momode.astype(np.int32).tofile(f)
            np.ascontiguousarray(zphase).view(np.float64).tofile(f)
            np.ascontiguousarray(fields).view(np.float64).tofile(f)
        r = subprocess.run([exe, str(sub / "in.bin"), str(sub / "out.bin")], capture_output=True, text=True, timeout=900, env=env)
        counts[name] = (r.stdout + r.stderr).count("WARNING: ThreadSanitizer")
    assert counts["good"] == 0, "data race in the kernel's shared-memory pipeline"
    assert counts["broken"] > 0, "the race detection did not see a deliberately broken protocol"
