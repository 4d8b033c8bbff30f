#!/usr/bin/env python
"""Elemental-generation benchmark: timeslices/s on N B200s, roofline and CPU baseline beside it.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload config5] [--impl reference]

A "step" is one timeslice of the workload: all Nop x Nmom matrices E[d,p](t) of size Ne x Ne.
Native arm: `value` is measured through the public sharded call gen.calc_all() with the step's inputs already
resident in HBM (device handles), `e2e` through gen.calc_range() on host arrays (streamed pipeline over the
C ABI) with H2D/D2H inside the timed region, and `e2e.from_files` through the file presets (ILDG, .npy, QDP).
For N > 1 each rank owns its own timeslices (weak scaling); finished chunks travel to rank 0 over NCCL inside
the timed region while the next are computed - the only exchange step this path has.
Contraction form: the one the library plans for the handle (`edk_plan_form`; `--contraction X` forces another
through the A/B hook for measurements).  Every run also checks the planned form against the GEMM form on one
timeslice and reports the observed error under "contraction".
Reference arm (`--impl reference`): the numpy restatement of the reference algorithm
(oracle/elemental_oracle.py, kind "port": the reference itself is a Python package that does
not exist on the GPU box) timed on the host cores on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

WORKLOADS = {
    # name: (Lx, Ly, Lz, Lt, Ne, num_nabla, nmom)  -- BASELINE.json configs[1..4]; tiny is for smoke runs
    "tiny": (8, 8, 8, 8, 16, 2, 7),
    "config2": (16, 16, 16, 128, 100, 1, 9),
    "config3": (24, 24, 24, 72, 100, 2, 33),
    "config4": (32, 32, 32, 64, 200, 2, 33),
    "config5": (48, 48, 48, 96, 200, 2, 33),
}
PAIRS_DISTINCT = {0: 1, 1: 7, 2: 34}     # distinct (left, right) pair-GEMMs per momentum
PAIRS_REFERENCE = {0: 1, 1: 7, 2: 43}    # pair-einsums the reference executes: sum_k 3^k 2^k
HOPS_REFERENCE = {0: 0, 1: 6, 2: 78}     # single-direction _nD applications per timeslice
SRC_OUT = {0: (0, 0), 1: (1, 3), 2: (4, 12)}  # stencil source / output fields per timeslice


def workload_desc(name, distance=None):
    Lx, Ly, Lz, Lt, Ne, nabla, nmom = WORKLOADS[name]
    if distance is not None:
        return (f"{name}: {Lx}x{Ly}x{Lz}x{Lt} synthetic gauge field, Ne={Ne}, distance={distance}, {nmom} momenta, "
                "DisplacementElementalGenerator, one timeslice per step")
    return (f"{name}: {Lx}x{Ly}x{Lz}x{Lt} synthetic gauge field, Ne={Ne}, num_nabla={nabla}, {nmom} momenta, "
            "ElementalGenerator, one timeslice per step")


def config_dict(name, distance, K, world):
    """The workload description both arms print (identical for `--impl native` and `--impl reference`)."""
    Lx, Ly, Lz, Lt, Ne, nabla, nmom = WORKLOADS[name]
    V = Lx * Ly * Lz
    fields_mb = (Ne * V * 3 * 8 + Ne * V * 48 * (SRC_OUT[nabla][1] if distance is None else 12)) / 1e6
    return {
        "workload": workload_desc(name, distance), "lattice": [Lx, Ly, Lz], "Ne": Ne,
        **({"num_nabla": nabla} if distance is None else {"distance": distance}), "momenta": nmom,
        "sharding": f"timeslices, {K} per rank, {world} rank(s)" + (", results gathered on rank 0 inside the timed region" if world > 1 else ""),
        "l2": f"step inputs+fields ({fields_mb:.0f} MB) exceed the 126 MB L2; two input sets alternated",
    }


def algorithmic(name, distance=None):
    """contraction flops per timeslice and bytes of ONE stencil launch."""
    Lx, Ly, Lz, Lt, Ne, nabla, nmom = WORKLOADS[name]
    V = Lx * Ly * Lz
    if distance is not None:
        return 8.0 * Ne * Ne * 3 * V * nmom * (distance + 1), 13 * Ne * V * 48.0 + 3 * V * 144.0
    flops = 8.0 * Ne * Ne * 3 * V * nmom * PAIRS_DISTINCT[nabla]
    # nabla3 launch: 1 field in, 3 fields out, links once ...
    return flops, 4 * Ne * V * 48.0 + 3 * V * 144.0


def stencil_bytes_moved(name, distance=None, planes=True):
    """... plus what this build's nabla3 additionally has to write when the GEMM form of the contraction is in
    use: the Re+Im plane of each output field (8 B per element), which the 3M arithmetic reads as its third A
    operand.  With the plane-wave form (`planes=False`) the pass moves SURVEY's bytes and nothing else."""
    Lx, Ly, Lz, Lt, Ne, nabla, nmom = WORKLOADS[name]
    V = Lx * Ly * Lz
    base = algorithmic(name, distance)[1]
    return base if (distance is not None or not planes) else base + 3 * Ne * V * 24.0


def measured_traffic(kernel, name):
    """DRAM bytes per launch from the committed ncu --set full captures (profiles/traffic.json), or None."""
    try:
        return json.load(open(os.path.join(REPO, "profiles", "traffic.json")))[kernel][name]["dram_bytes_per_launch"]
    except Exception:
        return None


def _executed_elements(Ne, te, tf, self_pair):
    """(e, f) elements a plane kernel with te x tf tiles really computes for one job: every valid element, except that a
    self pair (L == R) skips its tiles entirely below the diagonal (the fold kernel reads the mirror element)."""
    if not self_pair:
        return Ne * Ne
    n = 0
    for e0 in range(0, Ne, te):
        for f0 in range(0, Ne, tf):
            if e0 > f0 + tf - 1:
                continue
            n += (min(e0 + te, Ne) - e0) * (min(f0 + tf, Ne) - f0)
    return n


def contraction_accounting(q, Ne, latt3, self_pairs=0):
    """Work the contraction kernel of this handle EXECUTES per timeslice, from `ElementalEngine.query()`.

    Returns a dict: form, kernel, executed (real flops: FMA = 2, add / multiply = 1), slots (FP64-pipe lane-operations:
    one per DFMA / DADD / DMUL lane, 8 per lane of a DMMA.8x8x4 - the unit the shared FP64 pipe of B200 retires 16 of per
    cycle and SM sub-partition, so slots / (592 x 16 x clock x time) is the pipe utilisation ncu reports as
    sm__pipe_fp64_cycles_active + sm__pipe_tensor_subpipe_dmma_cycles_active), padded (forms 2 / 3: flops including the
    DMMA rows that pad the mode blocks).
      GEMM forms: 2 x (real MMAs per complex block) x Ne^2 x 3V per (pair, momentum) - the Hermitian pairing contracts 19
        instead of SURVEY 8d's 34 pairs (self pairs only one momentum of each +-p couple), 3M needs 3 real MMAs, not 4.
      Plane-wave form (2): per (pair, e, f, site) 12 DFMA for the colour-summed site product and, for each real xy-mode,
        one multiply-add on its real and imaginary part.  Folded (3): the two sites of a centre-symmetric pair share that
        multiply-add after one add / subtract per part.
      Separable form (4): per pair of sites (x, Lx-1-x): 2 x (10 DFMA + 2 DMUL) site products, 4 DADD (sum, difference),
        2 DADD + 8 DFMA (5 x-modes; 4 DFMA with 3); per row and (e, f) one multiply-add (or add) per separable xy-mode on
        the real and imaginary part.
    Self pairs (L == R) skip the tiles below the diagonal in forms 2 - 4: only executed tiles are counted."""
    Lx, Ly, Lz = latt3
    V = Lx * Ly * Lz
    form = int(q.get("contraction_form", 1))
    segs = float(q["pair_gemms_per_momentum"])  # (left, right) segments contracted
    if form < 2:
        kernel = "gram_tma_kernel" if q.get("tma_stages") else "gram_dmma_kernel"
        executed = 2.0 * q["real_mma_per_complex_block"] * Ne * Ne * 3 * V * q["pair_momentum_gemms"]
        return {"form": form, "kernel": kernel, "executed": executed, "slots": executed / 2.0, "padded": None}
    tile = int(q["plane_wave_tile"])
    # forms 2 / 3 skip whole CTA tiles of a self pair below the diagonal, form 4 (gram_sepx_kernel) warp tiles of 8 x 16 elements
    te, tf = (8, 16) if form == 4 else (8 * (tile // 10), 8 * (tile % 10))
    elems = (segs - self_pairs) * Ne * Ne + self_pairs * _executed_elements(Ne, te, tf, True)
    base = elems * V
    modes = int(q["plane_wave_modes"])
    if form == 4:
        nx = {5: 3, 9: 3, 13: 5}[modes]                    # x modes: constant + cos / sin of 1 (and 2)
        ymul = {5: 2, 9: 6, 13: 8}[modes]                  # separable modes with a non-constant y weight
        per_site_flops = 22.0 + 2.0 + 1.0 + 2.0 * (nx - 1)  # site product, sum / difference, X[1], X[c_q] / X[s_q]
        per_site_slots = 12.0 + 2.0 + 1.0 + 1.0 * (nx - 1)
        per_row_flops = 4.0 * ymul + 2.0 * (modes - ymul)
        per_row_slots = 2.0 * modes
        return {"form": 4, "kernel": "gram_sepx_kernel", "executed": base * (per_site_flops + per_row_flops / Lx),
                "slots": base * (per_site_slots + per_row_slots / Lx), "padded": None}
    folded = form == 3
    per_mode, fold_adds = (2.0, 2.0) if folded else (4.0, 0.0)
    # rows the DMMAs run over: blocks of 8 modes (form 2); per pass of 8 couples a cos block and a sin block (form 3)
    rows = 16 * (((modes + 1) // 2 + 7) // 8) if folded else 8 * ((modes + 7) // 8)
    return {"form": form, "kernel": "gram_pwf_kernel" if folded else "gram_pw_kernel",
            "executed": base * (24.0 + fold_adds + per_mode * modes), "padded": base * (24.0 + fold_adds + per_mode * rows),
            "slots": base * (12.0 + fold_adds + per_mode * rows / 2.0)}


def momentum_set(count):
    r = range(-3, 4)
    allp = sorted(((px * px + py * py + pz * pz, (px, py, pz)) for px in r for py in r for pz in r))
    return [p for _, p in allp[:count]]


# --------------------------------------------------------------------------------------------
# clocks during the timed region
# --------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {
        0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
        0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
        0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting",
    }

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._first = threading.Event()
        self._thr = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.dev = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.dev, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.dev, nv.NVML_CLOCK_SM))
                get = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
                mask = get(self.dev)
                for bit, name in self.REASONS.items():
                    if mask & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            self._first.set()  # NVML's first queries take tens of ms and hold the driver: the timed region starts after them
            self._stop.wait(0.1)

    def __enter__(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()
            self._first.wait(5.0)
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thr is not None:
            self._thr.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


# --------------------------------------------------------------------------------------------
# CPU arm: the oracle port on a bounded sample
# --------------------------------------------------------------------------------------------
def cpu_sample(name, W0=None, U=None, hop_frac=8):
    """Time the reference's two primitives at the workload's size and compose one timeslice
    (SURVEY 8d): T = hops * t_hop + pairs * Nmom * t_pair.  One hop is run on Ne/hop_frac of the
    eigenvectors (its cost is linear in Ne) and scaled; the pair contraction is run in full.
    Validated in the build container on config 3: composed 335 s vs 335.7 s for the complete
    reference run."""
    from oracle import elemental_oracle as orc

    use_all_host_cores()
    Lx, Ly, Lz, Lt, Ne, nabla, nmom = WORKLOADS[name]
    latt = [Lx, Ly, Lz, Lt]
    rng = np.random.default_rng(orc.SEED0)
    if W0 is None:
        W0 = (rng.standard_normal((Ne, Lz, Ly, Lx, 3)) + 1j * rng.standard_normal((Ne, Lz, Ly, Lx, 3))).astype(np.complex64)
    if U is None:
        U = orc.links_file_to_spatial(orc.synthetic_links(latt, 0))
    ne_hop = max(1, Ne // hop_frac)
    # each primitive is run once untimed first: the reference calls them 78 / 43*Nmom times per
    # timeslice, so the warm figure (BLAS threads up, einsum path cached) is the representative one
    orc.covariant_hop(W0[:ne_hop], U, 0)
    t0 = time.perf_counter()
    W1 = orc.covariant_hop(W0[:ne_hop], U, 0)
    t_hop = (time.perf_counter() - t0) * (Ne / ne_hop)
    right = np.ascontiguousarray(np.broadcast_to(W1[:1], W0.shape)) if ne_hop < Ne else W1
    phase = orc.momentum_phase(latt, (0, 0, 1))
    orc.gram(W0, right, -1 * phase)
    t0 = time.perf_counter()
    orc.gram(W0, right, -1 * phase)
    t_pair = time.perf_counter() - t0
    T = HOPS_REFERENCE[nabla] * t_hop + PAIRS_REFERENCE[nabla] * nmom * t_pair
    sample = (f"1 _nD hop on {ne_hop}/{Ne} eigenvectors (scaled x{Ne / ne_hop:.0f}) = {t_hop:.2f}s, "
              f"1 (pair, momentum) einsum at full size = {t_pair:.2f}s; composed "
              f"{HOPS_REFERENCE[nabla]} hops + {PAIRS_REFERENCE[nabla]}x{nmom} einsums = {T:.1f}s per timeslice")
    return 1.0 / T, sample


def cpu_sample_displacement(name, distance, W0=None, U=None, hop_frac=8):
    """Reference composition for the displacement generator (displacement_elemental.py:78-96):
    per timeslice `distance` line-extension steps (_D) and (distance+1) x Nmom einsums."""
    from oracle import elemental_oracle as orc

    use_all_host_cores()
    Lx, Ly, Lz, Lt, Ne, nabla, nmom = WORKLOADS[name]
    latt = [Lx, Ly, Lz, Lt]
    rng = np.random.default_rng(orc.SEED0)
    if W0 is None:
        W0 = (rng.standard_normal((Ne, Lz, Ly, Lx, 3)) + 1j * rng.standard_normal((Ne, Lz, Ly, Lx, 3))).astype(np.complex64)
    if U is None:
        U = orc.links_file_to_spatial(orc.synthetic_links(latt, 0))
    ne_hop = max(1, Ne // hop_frac)
    list(orc.displacement_fields(W0[:ne_hop], U, 1))
    t0 = time.perf_counter()
    D1 = list(orc.displacement_fields(W0[:ne_hop], U, 1))[1]
    t_step = (time.perf_counter() - t0) * (Ne / ne_hop)
    right = np.ascontiguousarray(np.broadcast_to(D1[:1], W0.shape)) if ne_hop < Ne else D1
    phase = orc.momentum_phase(latt, (0, 0, 1))
    orc.gram(W0, right, phase)
    t0 = time.perf_counter()
    orc.gram(W0, right, phase)
    t_pair = time.perf_counter() - t0
    T = distance * t_step + (distance + 1) * nmom * t_pair
    sample = (f"1 _D step on {ne_hop}/{Ne} eigenvectors (scaled x{Ne / ne_hop:.0f}) = {t_step:.2f}s, 1 (distance, momentum) "
              f"einsum at full size = {t_pair:.2f}s; composed {distance} steps + {distance + 1}x{nmom} einsums = {T:.1f}s per timeslice")
    return 1.0 / T, sample


def use_all_host_cores():
    """Give the BLAS behind numpy every core this process may run on (torchrun exports
    OMP_NUM_THREADS=1, which would otherwise cripple the CPU arm) and return the count in use."""
    try:
        want = len(os.sched_getaffinity(0))
    except Exception:
        want = os.cpu_count() or 1
    try:
        from threadpoolctl import threadpool_info, threadpool_limits

        threadpool_limits(limits=want)
        got = max([p.get("num_threads", 1) for p in threadpool_info()] or [1])
        return int(got)
    except Exception:
        return 1


def host_threads():
    return use_all_host_cores()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle.elemental_oracle  # noqa: F401  (loads numpy/BLAS before threads are counted)

    name = args.workload
    vals = []
    sample = ""
    dist_ = args.distance if args.generator == "displacement" else None
    # synthetic inputs are made once; every step then times the same bounded sample of the workload
    from oracle import elemental_oracle as orc

    Lx, Ly, Lz, Lt, Ne, nabla, nmom = WORKLOADS[name]
    rng = np.random.default_rng(orc.SEED0)
    W0 = (rng.standard_normal((Ne, Lz, Ly, Lx, 3)) + 1j * rng.standard_normal((Ne, Lz, Ly, Lx, 3))).astype(np.complex64)
    U = orc.links_file_to_spatial(orc.synthetic_links([Lx, Ly, Lz, Lt], 0))
    for i in range(args.warmup + args.steps):
        v, sample = cpu_sample(name, W0, U) if dist_ is None else cpu_sample_displacement(name, dist_, W0, U)
        if i >= args.warmup:
            vals.append(v)
    value = float(len(vals) / sum(1.0 / v for v in vals))
    cores = host_threads()
    line = {
        "impl": "reference", "metric": "elemental_timeslices_per_sec", "value": value, "unit": "timeslices/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 / value,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "c128 (f64)", "data": "synthetic",
        "config": config_dict(name, dist_, args.steps, args.gpus),
        "cpu_baseline": {"value": value, "unit": "timeslices/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "timeslices/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------
# native arm
# --------------------------------------------------------------------------------------------
def synth_links(torch, dev, V, g):
    """Random SU(3) links [V,4,3,3] c128 (file order of one timeslice) from the generator g, on the device."""
    a = torch.randn((V, 4, 3, 3), dtype=torch.complex128, device=dev, generator=g)
    q, r = torch.linalg.qr(a)
    d = torch.diagonal(r, dim1=-2, dim2=-1)
    q = q * (d / d.abs()).unsqueeze(-2)
    det = torch.linalg.det(q)
    return (q / det.pow(1.0 / 3.0)[..., None, None]).contiguous()


def synth_device_inputs(torch, dev, name, seed):
    """Random SU(3) links [V,4,3,3] c128 (file order of one timeslice) and unit-norm complex64
    eigenvectors [Ne,V,3], generated on the device (bench only; tests use the oracle's numpy ones)."""
    Lx, Ly, Lz, Lt, Ne, nabla, nmom = WORKLOADS[name]
    V = Lx * Ly * Lz
    g = torch.Generator(device=dev)
    g.manual_seed(20261017 + seed)
    U = synth_links(torch, dev, V, g)
    v = torch.randn((Ne, V, 3), dtype=torch.complex64, device=dev, generator=g)
    v = v / torch.linalg.vector_norm(v.reshape(Ne, -1), dim=1)[:, None, None]
    return U, v.contiguous()


def fp64_gemm_peak(torch, dev):
    """cuBLAS FP64 GEMM on this device, best of 10 (same method as MEASURED_PEAKS.json):
    the denominator of the contraction roofline."""
    n = 8192
    out = {}
    for label, dt, flop in (("dgemm", torch.float64, 2.0 * n**3), ("zgemm", torch.complex128, 8.0 * n**3)):
        m = n if dt == torch.float64 else n // 2
        fl = flop if dt == torch.float64 else 8.0 * m**3
        a = torch.randn((m, m), dtype=dt, device=dev)
        b = torch.randn((m, m), dtype=dt, device=dev)
        torch.matmul(a, b)
        best = 0.0
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            torch.matmul(a, b)
            e1.record()
            e1.synchronize()
            best = max(best, fl / (e0.elapsed_time(e1) * 1e-3) / 1e12)
        out[label] = best
        del a, b
    return out


def worst_block_error(got, ref):
    """Largest Frobenius error of an (operator, momentum) block relative to the block's norm (blocks that vanish
    analytically are measured against 1e-4 of the largest block): the measure of the parity tests."""
    norms = np.sqrt((np.abs(ref) ** 2).sum(axis=(-1, -2)))
    floor = 1e-4 * norms.max()
    err = np.sqrt((np.abs(got - ref) ** 2).sum(axis=(-1, -2)))
    return float((err / np.maximum(norms, floor)).max())


def file_backed_leg(edb, torch, dev, name, dist_, moms, U_two, V_two, K, barrier):
    """End to end from FILES: the timeslices are read through the typed file handles - an ILDG gauge
    configuration, eigenvectors once as a `.npy` file and once as a QDP timeslice `.mod` file (big-endian complex64,
    byte-swapped on the device) - by the streamed pipeline of calc_range, results to host arrays.  Replaces the reference's
    per-eigenvector open + mmap + copy loop (filedata/ndarray.py:17-47, timeslice.py:65-99, elemental.py:297-298)."""
    import shutil
    import tempfile

    from easydistillation_b200.fileio import write_ildg, write_qdp_timeslices

    Lx, Ly, Lz, Lt, Ne, nabla, nmom = WORKLOADS[name]
    Kf = max(2, min(K, 4))
    need = Kf * (V_two[0].nbytes * 3 + U_two[0].nbytes) * 1.2
    root = "/dev/shm" if os.path.isdir("/dev/shm") and shutil.disk_usage("/dev/shm").free > need else None
    tmp = tempfile.mkdtemp(prefix="edk_bench_", dir=root)
    out = {"timeslices": Kf, "directory": "tmpfs" if root else "default temporary directory"}
    try:
        prefix = tmp + "/"
        t0 = time.perf_counter()
        write_ildg(prefix + "cfg.lime", np.stack([U_two[i % 2] for i in range(Kf)]))
        Vc16 = np.lib.format.open_memmap(prefix + "cfg.eigenvector.npy", mode="w+", dtype="<c16", shape=(Kf, Ne, Lz, Ly, Lx, 3))
        for i in range(Kf):
            Vc16[i] = V_two[i % 2]
        Vc16.flush()
        del Vc16
        write_qdp_timeslices(prefix + "cfg.mod", np.stack([V_two[i % 2] for i in range(Kf)]).reshape(Kf, Ne, Lz * Ly * Lx, 3), [Lx, Ly, Lz, Kf])
        out["write_seconds"] = time.perf_counter() - t0
        gauge = edb.GaugeFieldIldg(prefix, ".lime", [Kf, Lz, Ly, Lx, 4, 3, 3])
        legs = {"npy_c16": edb.EigenvectorNpy(prefix, ".eigenvector.npy", [Kf, Ne, Lz, Ly, Lx, 3], Ne),
                "qdp_mod_c8_big_endian": edb.EigenvectorTimeSlice(prefix, ".mod", [Kf, Ne, Lz * Ly * Lx, 3], Ne)}
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for tag, evec in legs.items():
            if dist_ is None:
                gen = edb.ElementalGenerator([Lx, Ly, Lz, Kf], gauge, evec, nabla, moms, device=dev.index)
            else:
                gen = edb.DisplacementElementalGenerator([Lx, Ly, Lz, Kf], gauge, evec, dist_, moms, device=dev.index)
            gen.load("cfg")
            gen.calc_range(0, 2)  # warm-up: staging buffers
            pipe = gen._pipeline
            h2d0 = pipe.h2d_bytes
            barrier()
            t0 = time.perf_counter()
            e0.record()
            res = gen.calc_range(0, Kf)
            e1.record()
            barrier()
            wall = time.perf_counter() - t0
            read_bytes = (pipe.h2d_bytes - h2d0)  # every byte uploaded was read from the files first
            out[tag] = {"value": Kf / (e0.elapsed_time(e1) * 1e-3), "unit": "timeslices/s", "file_read_GBps": read_bytes / wall / 1e9,
                        "file_bytes_per_step": read_bytes // Kf, "checksum": float(np.abs(res[0, 0, 0]).sum())}
            del gen, res
            torch.cuda.empty_cache()
    except Exception as exc:  # the file leg must never cost the main line
        out["error"] = repr(exc)[:300]
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return out


def run_native(args):
    import torch
    import torch.distributed as dist

    import easydistillation_b200 as edb
    from easydistillation_b200 import _capi
    from easydistillation_b200.engine import microbench_fp64
    from easydistillation_b200.sharding import gather_timeslices

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the native arm has no CPU fallback")
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("bench.py: --gpus N > 1 must be launched with torch.distributed.run (one rank per GPU)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    name = args.workload
    Lx, Ly, Lz, Lt, Ne, nabla, nmom = WORKLOADS[name]
    V = Lx * Ly * Lz
    K, W = args.steps, args.warmup
    moms = momentum_set(nmom)
    dist_ = args.distance if args.generator == "displacement" else None
    mode_, order_ = (_capi.MODE_DERIVATIVE, nabla) if dist_ is None else (_capi.MODE_DISPLACEMENT, dist_)

    # ---- contraction form: the library plans it per handle (edk_plan_form); `--contraction X` asks for one form through
    # the A/B hook EDK_GRAM_ALGO, for measurements only ----
    contraction = {"requested": args.contraction}
    forced = {"gemm": 1, "planewave": 2, "planewave-folded": 3, "separable": 4}.get(args.contraction)
    if forced is not None:
        os.environ["EDK_GRAM_ALGO"] = str(forced)
        contraction["reason"] = "forced by --contraction"
    elif os.environ.get("EDK_GRAM_ALGO"):
        contraction["reason"] = "EDK_GRAM_ALGO set by the caller"
    else:
        contraction["reason"] = "planned by the library (edk_plan_form)"
        contraction["plan"] = _capi.plan_form((Lx, Ly, Lz), mode_, order_, moms)

    def make_generator(gauge, evec, Lt_run):
        if dist_ is None:
            return edb.ElementalGenerator([Lx, Ly, Lz, Lt_run], gauge, evec, nabla, moms, device=local)
        return edb.DisplacementElementalGenerator([Lx, Ly, Lz, Lt_run], gauge, evec, dist_, moms, device=local)

    # ---- device-resident throughput through the public class: K timeslices per rank, sharded by calc_all() ----
    # two resident input sets, alternated, each far larger than L2 at the graded workloads
    inputs = [synth_device_inputs(torch, dev, name, 1000 * rank + i) for i in range(2)]
    gen = make_generator(edb.GaugeFieldDevice([U.reshape(Lz, Ly, Lx, 4, 3, 3) for U, _ in inputs], cyclic=True),
                         edb.EigenvectorDevice([v.reshape(Ne, Lz, Ly, Lx, 3) for _, v in inputs], cyclic=True), K * world)
    gen.load("bench")
    eng = gen._engine

    if os.environ.get("EDK_BENCH_GRAM"):  # tuning hook: "mfrag,ksplit" (0 = auto)
        mf, ks = (int(v) for v in os.environ["EDK_BENCH_GRAM"].split(","))
        eng.debug_gram_config(mf, ks)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    scratch = torch.empty(eng.out_shape, dtype=torch.complex128, device=dev)
    for i in range(W):
        gen.calc_device(i % (K * world), out=scratch)
    # one untimed pass through the timed call itself: the [Lt, ...] result buffer then comes from torch's caching
    # allocator, as in any second use of the generator, and the NCCL point-to-point channels exist
    del_me = gen.calc_all(dst=0)
    del del_me
    if world > 1:  # first use of the communicator sets up the NCCL channels: not part of a step
        gather_timeslices(scratch[None, :1, :1].contiguous().expand(1, 1, 1, Ne, Ne).contiguous(), world, dst=0)
    barrier()

    launches0 = eng.launch_count
    eng.set_profiling(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        barrier()
        h0 = time.perf_counter()
        e0.record()
        gathered = gen.calc_all(dst=0)  # this rank's K timeslices; finished chunks travel to rank 0 while the next are computed
        e1.record()
        host_queue_ms = 1e3 * (time.perf_counter() - h0)  # host time to queue the K timeslices (diagnostic)
        barrier()
    ms = e0.elapsed_time(e1)
    prof = eng.get_profile()
    eng.set_profiling(False)
    launches = eng.launch_count - launches0
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * K / (ms_max * 1e-3)
    first = gathered[0].clone() if rank == 0 else None  # timeslice 0 = rank 0's first: input set 0
    del gathered

    # ---- observed error of the planned form at this shape: the same timeslice through the GEMM form (untimed) ----
    q = eng.query()
    parity = None
    if rank == 0 and q["contraction_form"] != 1 and not args.no_parity_check:
        eng.debug_algo(1)
        ref1 = gen.calc_device(0, out=scratch)
        parity = {"against": "GEMM form (3M arithmetic on DMMA) on the same timeslice", "tolerance": 1e-10,
                  "worst_block_rel_err": worst_block_error(first.cpu().numpy(), ref1.cpu().numpy())}
        eng.debug_algo(-1 if forced is None and not os.environ.get("EDK_GRAM_ALGO") else int(os.environ["EDK_GRAM_ALGO"]))
    del first

    # everything that needs the first handle, then free its workspace for the end-to-end legs
    workspace_mb = eng.workspace_bytes / 1e6
    W0_host = eng.debug_field(0).cpu().numpy() if (rank == 0 and world == 1 and not args.no_cpu_baseline) else None
    U0, v0 = inputs[0]
    U_sp_host = None
    if W0_host is not None:
        U_sp_host = np.ascontiguousarray(np.moveaxis(U0.cpu().numpy().reshape(Lz, Ly, Lx, 4, 3, 3), 3, 0)[:3])
    dmma_tf, dfma_tf = microbench_fp64(local) if rank == 0 else (0.0, 0.0)
    U_two = [inputs[i][0].cpu().numpy().reshape(Lz, Ly, Lx, 4, 3, 3) for i in range(2)]
    V_two = [inputs[i][1].cpu().numpy().reshape(Ne, Lz, Ly, Lx, 3) for i in range(2)]
    del gen, eng, scratch, inputs, U0, v0
    torch.cuda.empty_cache()

    # ---- end to end through the public class API: host arrays in, numpy out ---------------------
    # ElementalGenerator.calc_range streams the timeslices: pinned staging + H2D of t+1 and D2H of
    # t-1 overlap the kernels of t (easydistillation_b200/pipeline.py); every step uploads its links
    # and eigenvectors and downloads its result.  With N ranks each streams its own K timeslices.
    class CyclicEigenvectors:  # duck-typed eigenvector handle: two distinct host timeslices, repeated
        def __init__(self, two):
            self.two, self.Ne = two, Ne

        def load(self, key):
            return self

        def __getitem__(self, key):
            t = key[0] if isinstance(key, tuple) else key
            blk = self.two[t % 2]
            return blk[key[1]] if isinstance(key, tuple) and len(key) > 1 else blk

    U_host = np.stack([U_two[i % 2] for i in range(K)])
    if K * V_two[0].nbytes <= 8 << 30:  # small enough to hold K timeslices: the ordinary in-memory handle
        evec = edb.EigenvectorHostmem(np.stack([V_two[i % 2] for i in range(K)]))
    else:
        evec = CyclicEigenvectors(V_two)
    gen = make_generator(edb.GaugeFieldHostmem(U_host), evec, K)
    gen.load("bench")
    gen.calc_range(0, min(K, 2))  # warm-up: allocates the staging buffers
    pipe = gen._pipeline
    h2d0, d2h0 = pipe.h2d_bytes, pipe.d2h_bytes
    barrier()
    e0.record()
    res = gen.calc_range(0, K)
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * K / (float(t.item()) * 1e-3)
    h2d = (pipe.h2d_bytes - h2d0) // K
    d2h = (pipe.d2h_bytes - d2h0) // K
    checksum = float(np.abs(res[0, 0, 0]).sum())
    del res, gen, pipe, evec, U_host
    torch.cuda.empty_cache()
    files = None
    if world == 1 and not args.no_file_leg:
        files = file_backed_leg(edb, torch, dev, name, dist_, moms, U_two, V_two, K, barrier)

    line = None
    if rank == 0:
        flops, st_bytes_launch = algorithmic(name, dist_)
        # the contraction of one timeslice: one launch, or one per tile shape (separable form: 32 x 32 tiles, then the two
        # edge strips) - timed together by the CUDA events around the phase, accounted per timeslice
        gram_launches = max(1, prof["contraction"]["launches"]) / K
        gram_ms = prof["contraction"]["ms"] / K
        st_launch = max(1, prof["stencil"]["launches"])
        st_ms = prof["stencil"]["ms"] / st_launch
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        hbm_src = "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
        gemm = fp64_gemm_peak(torch, dev)
        fp64_peak = max(gemm.values())
        self_pairs = _capi.plan(mode_, order_, moms)["self_pairs"] if (dist_ is None and q["hermitian_pairing"]) else 0
        acct = contraction_accounting(q, Ne, (Lx, Ly, Lz), self_pairs)
        exec_flops, pw_padded_flops, form_used = acct["executed"], acct["padded"], acct["form"]
        pw_form = form_used >= 2
        contraction["form"] = form_used
        if parity is not None:
            contraction["parity_check"] = parity
        # FP64-pipe utilisation: lane-operations retired / (592 sub-partitions x 16 lanes per cycle x SM clock x time)
        sm_clock_hz = 1e6 * float(clocks.summary().get("sm_mhz") or 1965.0)
        pipe_util = acct["slots"] / (gram_ms * 1e-3) / (592.0 * 16.0 * sm_clock_hz)
        achieved_tf = exec_flops / (gram_ms * 1e-3) / 1e12
        survey_tf = flops / (gram_ms * 1e-3) / 1e12
        # stencil: bytes of ONE launch (nabla3: 1 source, 3 outputs, links once; displacement step: 6 in, 6 + mean out)
        st_moved = stencil_bytes_moved(name, dist_, planes=not pw_form)
        st_gbs = st_moved / (st_ms * 1e-3) / 1e9 if prof["stencil"]["launches"] else None
        st_survey_gbs = st_bytes_launch / (st_ms * 1e-3) / 1e9 if prof["stencil"]["launches"] else None
        cpu_val, cpu_smp = (None, "skipped (--no-cpu-baseline)")
        if W0_host is not None and dist_ is None:
            cpu_val, cpu_smp = cpu_sample(name, W0_host.astype(np.complex64), U_sp_host)
        elif W0_host is not None:
            cpu_val, cpu_smp = cpu_sample_displacement(name, dist_, W0_host.astype(np.complex64), U_sp_host)
        gram_name = acct["kernel"]
        st_name = "nabla3_kernel" if dist_ is None else "displace_step6_kernel"
        line = {
            "metric": "elemental_timeslices_per_sec", "value": value, "unit": "timeslices/s", "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": ms_max / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "c128 (f64)", "data": "synthetic",
            "config": config_dict(name, dist_, K, world),
            "api": "ElementalGenerator.calc_all() over device-resident inputs (GaugeFieldDevice / EigenvectorDevice): timeslices sharded "
                   "over the ranks, finished chunks gathered on rank 0 while the next are computed",
            "workspace_MB": workspace_mb, "host_queue_ms": host_queue_ms,
            "e2e": {"value": e2e_value, "unit": "timeslices/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "api": "ElementalGenerator.calc_range over host arrays (streamed: pinned staging, H2D/D2H overlapped with the kernels)"
                           + ("; every rank streams its own timeslices" if world > 1 else ""),
                    "checksum": checksum, **({"from_files": files} if files is not None else {})},
            "gpu_launches": int(launches),
            "contraction": contraction,
            "clocks": clocks.summary(),
            "roofline": {
                "kernel": gram_name + {
                    4: " (separable contraction: site products, x transform per site pair and y transform per row by DFMA, y-stage "
                       "accumulators in tensor memory; swizzled TMA tiles, producer warp + mbarrier ring; z fold timed with the combine step)",
                    3: " (plane-wave factorised contraction: site products by DFMA, real xy-mode transform by DMMA.8x8x4, "
                       "centre-symmetric site pairs folded, TMA producer warp + mbarrier ring; z fold timed with the combine step)",
                    2: " (plane-wave factorised contraction: site products by DFMA, real xy-mode transform by DMMA.8x8x4, "
                       "TMA producer warp + mbarrier ring; z fold timed with the combine step)",
                }.get(form_used, " (momentum-phased contraction, DMMA.8x8x4, TMA producer warp + mbarrier ring)"),
                "bound": "tensor",
                "achieved": achieved_tf, "peak": fp64_peak, "unit": "TFLOP/s", "frac": achieved_tf / fp64_peak,
                "traffic": measured_traffic(gram_name, name if dist_ is None else name + "_displacement"),
                "peak_source": f"cuBLAS FP64 GEMM measured in this run (dgemm {gemm['dgemm']:.1f}, zgemm {gemm['zgemm']:.1f} TFLOP/s); "
                               f"DMMA issue-rate microbench {dmma_tf:.1f}, DFMA {dfma_tf:.1f} TFLOP/s; nominal 37-40",
                "algorithmic_flops_per_launch": exec_flops, "ms_per_launch": gram_ms, "launches_per_step": gram_launches,
                "per_launch_means": "the contraction of one timeslice" + (f" ({gram_launches:.0f} launches of the kernel, one per tile shape, timed together)" if gram_launches > 1 else ""),
                "fp64_pipe_utilisation": pipe_util,
                "pairing": q, "survey_flops_per_launch": flops, "survey_equivalent_tflops": survey_tf,
                **({"executed_tflops_incl_padded_mode_rows": pw_padded_flops / (gram_ms * 1e-3) / 1e12} if pw_padded_flops else {}),
                "note": ("bound = the FP64 pipe that DFMA and DMMA share on B200 (same rate: microbench above); achieved = real "
                         "flops executed / time (FMA = 2, add / multiply = 1; tiles a self pair skips are not counted); "
                         "fp64_pipe_utilisation = FP64 lane-operations / (592 x 16 x SM clock x time), the figure ncu reports as "
                         "sm__pipe_fp64_cycles_active (+ the DMMA sub-pipe); the phase factorises over the axes, so these forms "
                         "need far fewer flops than SURVEY 8d's GEMM count: survey_equivalent_tflops = that count / time") if pw_form else
                        ("achieved = real flops the DMMA kernel executes / time (useful rows only); the Hermitian pairing "
                         "G(L,R,p)^dag = G(R,L,-p) contracts 19 instead of SURVEY 8d's 34 pairs and the 3M product uses 3 "
                         "instead of 4 real MMAs per complex block; survey_equivalent_tflops = SURVEY 8d flops / time"),
                "share_of_step": prof["contraction"]["ms"] / ms,
            },
            "roofline_stencil": {
                "kernel": st_name + (" (covariant central differences, 3 directions per pass)" if dist_ is None
                                     else " (six straight Wilson lines extended by one link + their mean)"), "bound": "hbm",
                "achieved": st_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": (st_gbs / hbm_peak) if st_gbs else None,
                "traffic": measured_traffic(st_name, name if dist_ is None else name + "_displacement"), "peak_source": hbm_src, "algorithmic_bytes_per_launch": st_moved,
                "survey_bytes_per_launch": st_bytes_launch, "survey_equivalent_gbs": st_survey_gbs,
                "note": ("algorithmic bytes = 1 field in + 3 fields out + links (SURVEY 8d); the Re+Im planes are not written "
                         "when a plane-wave / separable form of the contraction is in use") if (pw_form or dist_ is not None) else
                        ("algorithmic bytes = 1 field in + 3 fields out + links (SURVEY 8d) + the 3 Re+Im planes (8 B per element) "
                         "this build's stencil writes for the 3M contraction; survey_equivalent_gbs counts SURVEY's bytes only"),
                "ms_per_launch": st_ms, "launches_per_step": st_launch / K,
                "share_of_step": prof["stencil"]["ms"] / ms,
            },
            "phase_ms_per_step": {k: v["ms"] / K for k, v in prof.items()},
            "cpu_baseline": {"value": cpu_val, "unit": "timeslices/s", "cores": host_threads(), "kind": "port", "sample": cpu_smp},
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default=os.environ.get("EDK_BENCH_WORKLOAD", "config5"), choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-file-leg", action="store_true", help="skip the file-backed end-to-end leg (ILDG / .npy / QDP .mod files)")
    ap.add_argument("--no-parity-check", action="store_true", help="skip the untimed comparison of the planned form with the GEMM form")
    ap.add_argument("--contraction", default=os.environ.get("EDK_BENCH_CONTRACTION", "auto"), choices=["auto", "gemm", "planewave", "planewave-folded", "separable"],
                    help="contraction form: auto = the form the library plans for the shape (edk_plan_form); the others ask "
                         "for one form through the A/B hook EDK_GRAM_ALGO")
    ap.add_argument("--generator", default="derivative", choices=["derivative", "displacement"],
                    help="ElementalGenerator (default, the graded workload) or DisplacementElementalGenerator")
    ap.add_argument("--distance", type=int, default=2, help="displacement generator: number of link steps")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "native" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
