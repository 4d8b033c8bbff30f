#!/usr/bin/env python
"""Elemental-generation benchmark: timeslices/s on N B200s, roofline and CPU baseline beside it.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload config5] [--impl reference]

A "step" is one timeslice of the workload: all Nop x Nmom matrices E[d,p](t) of size Ne x Ne.
Native arm: `value` is measured with the step's inputs already resident in HBM, `e2e` through
the host-buffer C-ABI call (edk_calc_host) with H2D/D2H inside the timed region.  For N > 1 each
rank owns its own timeslices (weak scaling), and the timed region ends with the NCCL gather of
all results onto rank 0 - the only exchange step this path has.
Contraction form: by default (`--contraction auto`) an untimed set-up step validates the newer plane-wave
factorised form against the GEMM form at the workload's shape in a child process and uses it only if it agrees to
1e-10 and is faster (easydistillation_b200/tuning.py); the line reports the decision under "contraction".
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


def contraction_accounting(q, Ne, V):
    """Flops the contraction kernel of this handle EXECUTES per timeslice, from `ElementalEngine.query()`.

    GEMM forms: 2 x (real MMAs per complex block) x Ne^2 x 3V per (pair, momentum) - the Hermitian pairing contracts
    19 instead of SURVEY 8d's 34 pairs (self pairs only for one momentum of each +-p couple) and the 3M product needs
    3 real MMAs instead of 4.  Plane-wave form (2): per (pair, e, f, site) 12 DFMA for the colour-summed site product
    and, for each real xy-mode, one multiply-add on its real and imaginary part.  Folded form (3): the two sites of a
    centre-symmetric pair share that multiply-add after one add / subtract per real and imaginary part - 2 flops per
    site for the folding, 2 per mode and site.  `padded` additionally counts the DMMA rows that pad the mode blocks
    (None for the GEMM forms).  Returns a dict: form, plane_wave, folded, kernel, executed, padded."""
    form = int(q.get("contraction_form", 1))
    pw_form, folded = form in (2, 3), form == 3
    if not pw_form:
        kernel = "gram_tma_kernel" if q.get("tma_stages") else "gram_dmma_kernel"
        executed = 2.0 * q["real_mma_per_complex_block"] * Ne * Ne * 3 * V * q["pair_momentum_gemms"]
        return {"form": form, "plane_wave": False, "folded": False, "kernel": kernel, "executed": executed, "padded": None}
    modes = int(q["plane_wave_modes"])
    per_mode, fold_adds = (2.0, 2.0) if folded else (4.0, 0.0)
    # rows the DMMAs run over: blocks of 8 modes (form 2); per pass of 8 couples a cos block and a sin block (form 3)
    rows = 16 * (((modes + 1) // 2 + 7) // 8) if folded else 8 * ((modes + 7) // 8)
    base = float(q["pair_gemms_per_momentum"]) * Ne * Ne * V
    return {"form": form, "plane_wave": True, "folded": folded, "kernel": "gram_pwf_kernel" if folded else "gram_pw_kernel",
            "executed": base * (24.0 + fold_adds + per_mode * modes), "padded": base * (24.0 + fold_adds + per_mode * rows)}


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
            self._stop.wait(0.1)

    def __enter__(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()
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
        "config": {"workload": workload_desc(name, dist_)},
        "cpu_baseline": {"value": value, "unit": "timeslices/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "timeslices/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------
# native arm
# --------------------------------------------------------------------------------------------
def synth_device_inputs(torch, dev, name, seed):
    """Random SU(3) links [V,4,3,3] c128 (file order of one timeslice) and unit-norm complex64
    eigenvectors [Ne,V,3], generated on the device (bench only; tests use the oracle's numpy ones)."""
    Lx, Ly, Lz, Lt, Ne, nabla, nmom = WORKLOADS[name]
    V = Lx * Ly * Lz
    g = torch.Generator(device=dev)
    g.manual_seed(20261017 + seed)
    a = torch.randn((V, 4, 3, 3), dtype=torch.complex128, device=dev, generator=g)
    q, r = torch.linalg.qr(a)
    d = torch.diagonal(r, dim1=-2, dim2=-1)
    q = q * (d / d.abs()).unsqueeze(-2)
    det = torch.linalg.det(q)
    U = (q / det.pow(1.0 / 3.0)[..., None, None]).contiguous()
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


def run_native(args):
    import torch
    import torch.distributed as dist

    from easydistillation_b200 import _capi
    from easydistillation_b200.engine import ElementalEngine, microbench_fp64
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

    # ---- contraction form (untimed set-up, like a user's one-off `python -m easydistillation_b200.tuning`) ----
    # "auto": a child process validates the plane-wave form against the GEMM form at this workload's shape and
    # through the public class API, times both, and the faster validated one becomes the default of this process.
    contraction = {"requested": args.contraction}
    if os.environ.get("EDK_GRAM_ALGO"):
        contraction.update(form=int(os.environ["EDK_GRAM_ALGO"]), reason="EDK_GRAM_ALGO set by the caller")
    elif args.contraction == "auto":
        from easydistillation_b200 import tuning

        decision = tuning.select_contraction((Lx, Ly, Lz), Ne, mode_, order_, moms, device=local, reps=2, timeout=420.0)
        form = int(decision["form"])
        if world > 1:  # every rank uses the same form: the slowest decision wins
            tf = torch.tensor([form], dtype=torch.int32, device=dev)
            dist.all_reduce(tf, op=dist.ReduceOp.MIN)
            form = int(tf.item())
        if form != decision["form"]:
            decision.update(form=form, tile=None)
        tuning.apply(decision)
        contraction.update(decision)
        if rank == 0:
            print(f"bench.py: contraction form {form} ({decision['reason']})", file=sys.stderr, flush=True)
    else:
        form = {"planewave": 2, "planewave-folded": 3}.get(args.contraction, 1)
        os.environ["EDK_GRAM_ALGO"] = str(form)
        contraction.update(form=form, reason="forced by --contraction")
    eng = ElementalEngine((Lx, Ly, Lz), Ne, mode_, order_, moms, device=local)

    if os.environ.get("EDK_BENCH_GRAM"):  # tuning hook: "mfrag,ksplit" (0 = auto)
        mf, ks = (int(v) for v in os.environ["EDK_BENCH_GRAM"].split(","))
        eng.debug_gram_config(mf, ks)

    # two resident input sets, alternated, each far larger than L2 at the graded workloads
    inputs = [synth_device_inputs(torch, dev, name, 1000 * rank + i) for i in range(2)]
    outs = torch.empty((K,) + eng.out_shape, dtype=torch.complex128, device=dev)

    def step(i, out):
        U, v = inputs[i % 2]
        eng.set_links(U, _capi.LINKS_FILE_T)
        eng.set_eigvecs(v)
        eng.calc(out)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    scratch = torch.empty(eng.out_shape, dtype=torch.complex128, device=dev)
    for i in range(W):
        step(i, scratch)
    if world > 1:  # first use of the collective sets up the NCCL channels: not part of a step
        gather_timeslices(outs[:1], world, dst=0)
    barrier()

    # ---- device-resident throughput --------------------------------------------------------
    launches0 = eng.launch_count
    eng.set_profiling(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        barrier()
        e0.record()
        for i in range(K):
            step(i, outs[i])
        gathered = gather_timeslices(outs, K * world, dst=0) if world > 1 else outs
        e1.record()
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
    del gathered

    # everything that needs the first handle, then free its 14 GB workspace for the end-to-end leg
    q = eng.query()
    workspace_mb = eng.workspace_bytes / 1e6
    W0_host = eng.debug_field(0).cpu().numpy() if (rank == 0 and world == 1 and not args.no_cpu_baseline) else None
    U0, v0 = inputs[0]
    U_sp_host = None
    if W0_host is not None:
        U_sp_host = np.ascontiguousarray(np.moveaxis(U0.cpu().numpy().reshape(Lz, Ly, Lx, 4, 3, 3), 3, 0)[:3])
    dmma_tf, dfma_tf = microbench_fp64(local) if rank == 0 else (0.0, 0.0)
    eng.close()
    del eng, outs, scratch

    # ---- end to end through the public class API: host arrays in, numpy out ---------------------
    # ElementalGenerator.calc_range streams the timeslices: pinned staging + H2D of t+1 and D2H of
    # t-1 overlap the kernels of t (easydistillation_b200/pipeline.py); every step uploads its links
    # and eigenvectors and downloads its result.
    import easydistillation_b200 as edb

    class CyclicEigenvectors:  # duck-typed eigenvector handle: two distinct host timeslices, repeated
        def __init__(self, two):
            self.two, self.Ne = two, Ne

        def load(self, key):
            return self

        def __getitem__(self, key):
            t = key[0] if isinstance(key, tuple) else key
            blk = self.two[t % 2]
            return blk[key[1]] if isinstance(key, tuple) and len(key) > 1 else blk

    U_two = [inputs[i][0].cpu().numpy().reshape(Lz, Ly, Lx, 4, 3, 3) for i in range(2)]
    U_host = np.stack([U_two[i % 2] for i in range(K)])
    V_two = [inputs[i][1].cpu().numpy().reshape(Ne, Lz, Ly, Lx, 3) for i in range(2)]
    if K * V_two[0].nbytes <= 8 << 30:  # small enough to hold K timeslices: the ordinary in-memory handle
        evec = edb.EigenvectorHostmem(np.stack([V_two[i % 2] for i in range(K)]))
    else:
        evec = CyclicEigenvectors(V_two)
    del inputs, U0, v0
    torch.cuda.empty_cache()
    if dist_ is None:
        gen = edb.ElementalGenerator([Lx, Ly, Lz, K], edb.GaugeFieldHostmem(U_host), evec, nabla, moms, device=local)
    else:
        gen = edb.DisplacementElementalGenerator([Lx, Ly, Lz, K], edb.GaugeFieldHostmem(U_host), evec, dist_, moms, device=local)
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
    del res

    line = None
    if rank == 0:
        flops, st_bytes_launch = algorithmic(name, dist_)
        gram_ms = prof["contraction"]["ms"] / max(1, prof["contraction"]["launches"])
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
        # flops the kernel EXECUTES: 2 x (real MMAs per complex block) x Ne^2 x 3V per (pair, momentum).
        # Algorithmic savings make this smaller than SURVEY 8d's count (8 Ne^2 3V x 34 pairs x Nmom): the
        # Hermitian pairing contracts 19 pairs (self pairs only for one momentum of each +-p couple), and
        # the 3M complex product needs 3 real MMAs instead of 4.
        # `achieved`/`frac` are the executed rate (what the FP64 pipe really does); the SURVEY-counted
        # rate is reported beside it as survey_equivalent_tflops.
        acct = contraction_accounting(q, Ne, V)
        exec_flops, pw_padded_flops, pw_form, folded = acct["executed"], acct["padded"], acct["plane_wave"], acct["folded"]
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
            "config": {
                "workload": workload_desc(name, dist_), "lattice": [Lx, Ly, Lz], "Ne": Ne,
                **({"num_nabla": nabla} if dist_ is None else {"distance": dist_}), "momenta": nmom,
                "sharding": f"timeslices, {K} per rank, {world} rank(s)" + (", NCCL gather to rank 0 inside the timed region" if world > 1 else ""),
                "l2": f"step inputs+fields ({(Ne * V * 3 * 8 + Ne * V * 48 * (SRC_OUT[nabla][1] if dist_ is None else 12)) / 1e6:.0f} MB) exceed the 126 MB L2; two input sets alternated",
                "workspace_MB": workspace_mb,
            },
            "e2e": {"value": e2e_value, "unit": "timeslices/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "api": "ElementalGenerator.calc_range over host arrays (streamed: pinned staging, H2D/D2H overlapped with the kernels)",
                    "checksum": checksum},
            "gpu_launches": int(launches),
            "contraction": contraction,
            "clocks": clocks.summary(),
            "roofline": {
                "kernel": gram_name + (" (plane-wave factorised contraction: site products by DFMA, real xy-mode transform by "
                                       "DMMA.8x8x4" + (", centre-symmetric site pairs folded" if folded else "") +
                                       ", TMA producer warp + mbarrier ring; z fold timed with the combine step)" if pw_form
                                       else " (momentum-phased contraction, DMMA.8x8x4, TMA producer warp + mbarrier ring)"),
                "bound": "tensor",
                "achieved": achieved_tf, "peak": fp64_peak, "unit": "TFLOP/s", "frac": achieved_tf / fp64_peak,
                "traffic": measured_traffic(gram_name, name if dist_ is None else name + "_displacement"),
                "peak_source": f"cuBLAS FP64 GEMM measured in this run (dgemm {gemm['dgemm']:.1f}, zgemm {gemm['zgemm']:.1f} TFLOP/s); "
                               f"DMMA issue-rate microbench {dmma_tf:.1f}, DFMA {dfma_tf:.1f} TFLOP/s; nominal 37-40",
                "algorithmic_flops_per_launch": exec_flops, "ms_per_launch": gram_ms,
                "pairing": q, "survey_flops_per_launch": flops, "survey_equivalent_tflops": survey_tf,
                **({"executed_tflops_incl_padded_mode_rows": pw_padded_flops / (gram_ms * 1e-3) / 1e12} if pw_form else {}),
                "note": ("achieved = flops the plane-wave form needs / time: the phase factorises, so the site product is formed once "
                         "for all momenta and only the real xy-modes are transformed per plane; the FP64 pipe is the bound, "
                         "survey_equivalent_tflops = SURVEY 8d flops (GEMM form, 34 pairs x Nmom) / time") if pw_form else
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
                         "when the plane-wave form of the contraction is in use") if (pw_form or dist_ is not None) else
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
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default=os.environ.get("EDK_BENCH_WORKLOAD", "config5"), choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--contraction", default=os.environ.get("EDK_BENCH_CONTRACTION", "auto"), choices=["auto", "gemm", "planewave", "planewave-folded"],
                    help="contraction form: auto = validate and time the plane-wave form against the GEMM form in a child "
                         "process first and use the faster validated one; gemm / planewave force one")
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
