#!/usr/bin/env python
# -*- coding: utf-8 -*-
"""bench.py -- ADI cell-steps/s (fp64) of the B200 engine on BASELINE.json's workloads.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload ...]

Prints ONE JSON line (rank 0).

N=1 (default): BASELINE.json configs[1], `single_track_on_plate` at 512^3 -- Cartesian plate + track mask,
Robin on six faces with per-face variable h (dense coefficient field per axis: 75 algorithmic bytes per
cell-step), theta=0.5, dt=0.02 s.  A "step" is one full ADI time step (explicit stage + x, y, z
implicit sweeps) of the whole grid.  The same line carries sub-records for the other single-GPU configs, each
timed in the same run: `c3_cyl` (configs[2], cylindrical 256x1024x512), `c4_waam` (configs[3], 1024^3 deposition
with births) and `c5_n1` (configs[4], 2048x2048x1024 on ONE GPU: the T1 of the strong-scaling target).

N>1 (one process per GPU under torchrun): `value` is the same plate at 512 x 512 x (512*N), z-slab partitioned,
one 512^3 slab per GPU ("scaling": "weak", so N=1 of a scaling run equals the N=1 line).  The line also carries
`c5_strong` -- configs[4] at 2048x2048x1024 z-slab sharded over the N GPUs, with the one-GPU time measured on
rank 0 in the same run and the parallel efficiency T1/(N*TN) -- and `parity`: a 256x256x128 miniature of both
workloads stepped over NCCL and checked against the oracle on the undivided grid.

`value` is device-resident throughput (inputs in HBM, CUDA events).  `e2e.value` is the same metric through
the host-array C-ABI call, one blocking call per step as the reference's time loop T = step(T) does it
(H2D + step + D2H inside the timed region); the two-slot pipelined figure for independent fields is reported
next to it as `e2e.pipelined_value`.  `roofline` is for the slowest kernel of the step, timed live with CUDA
events recorded inside the engine on the launching stream.  `cpu_baseline`: the reference's own serial Numba
path (baseline/_ref, unmodified; `numba_value`, 1 core) when it is present, and its C restatement
(oracle/adi_oracle.c, bit-identical on the golden vectors) with OpenMP on the box's cores; the same run checks
GPU-vs-oracle parity on the sample.

--impl reference times the reference itself on the host cores (Numba, serial: the reference has no threads to
use) on a bounded y-slab sample of the workload; when baseline/_ref or numba is missing, its C restatement.
"""
from __future__ import annotations

import argparse
import ctypes as C
import importlib.util
import json
import math
import os
import statistics
import subprocess
import sys
import tempfile
import time
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "ADI cell-steps/sec (fp64)"
UNIT = "cell-steps/s"
RHO, CP, K = 7800.0, 500.0, 25.0
DX, DT, THETA, TINF = 1.0e-3, 0.02, 0.5, 20.0
FACES = ("x-", "x+", "y-", "y+", "z-", "z+")
REF_DIR = os.path.join(ROOT, "baseline", "_ref")


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def traffic_of(kernel):
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        return json.load(open(tp)).get(kernel, {}).get("dram_bytes_per_launch")
    except Exception:
        return None


# ------------------------------------------------------------------------------------------
# workload: single_track_on_plate (single_track_on_plate.py:113-114,159)
# ------------------------------------------------------------------------------------------
def plate_track_mask_np(n, ny=None):
    ny = ny or n
    m = np.zeros((n, ny, n), dtype=bool)
    nzp = n - max(1, n // 64)
    m[:, :, :nzp] = True
    m[: max(1, n // 32), : ny // 2, nzp:] = True
    return m


def host_inputs(n, ny, seed=0):
    """Seeded host inputs (sample sub-box for the CPU legs and the parity check)."""
    rng = np.random.default_rng(seed)
    mask = plate_track_mask_np(n, ny)
    T0 = np.full(mask.shape, TINF)
    T0[mask] = 20.0 + 1380.0 * rng.random(int(mask.sum()))
    rng2 = np.random.default_rng(1234)
    h = {f: 10.0 * (0.3 + rng2.random(mask.shape)) for f in FACES}
    return mask, T0, h


def rel_l2(a, b, where=None):
    if where is not None:
        a, b = a[where], b[where]
    den = float(np.sqrt(np.sum(b * b)))
    return float(np.sqrt(np.sum((a - b) ** 2))) / (den if den > 0 else 1.0)


# ------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "50", "-i", str(self.idx)],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            out, _ = self.p.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
            out, _ = self.p.communicate()
        sm, smax, reasons, power = [], [], set(), []
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------
# CPU legs -- the only places bench.py touches oracle/ and baseline/_ref
# ------------------------------------------------------------------------------------------
def host_threads():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_leg(n, ny_sample, threads, steps=1):
    """The C restatement of the reference's Numba path (oracle/adi_oracle.c) on a y-slab sample."""
    from oracle import cart
    mask, T0, h = host_inputs(n, ny_sample)
    cart.set_threads(threads)
    grid = cart.Grid3D(n, ny_sample, n, DX, mask)
    mat = cart.Material(RHO, CP, K)
    prm = cart.Params(DT, THETA)
    packs = cart.precompute_coeff_packs_unified(grid, mat, robin_h=h)
    work = np.empty(3 * T0.size)
    T = T0
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        T = cart.adi_step_numba_coeff(T, grid, mat, prm, packs, Tinf=TINF, work=work)
        times.append(time.perf_counter() - t0)
    return T, times, (mask, T0, h)


def load_reference(module):
    """Import an UNMODIFIED reference module from baseline/_ref (git-ignored copy of /root/reference made by
    __graft_entry__.build(); it travels to the GPU box with the snapshot).  Returns (module, None) or (None, why)."""
    path = os.path.join(REF_DIR, module + ".py")
    if not os.path.exists(path):
        return None, f"baseline/_ref/{module}.py is absent (build() copies it where /root/reference exists)"
    os.environ.setdefault("NUMBA_CACHE_DIR", os.path.join(tempfile.gettempdir(), "adi_b200_numba_cache"))
    os.environ.setdefault("PYTHONDONTWRITEBYTECODE", "1")
    try:
        spec = importlib.util.spec_from_file_location("adi_reference_" + module, path)
        mod = importlib.util.module_from_spec(spec)
        sys.modules[spec.name] = mod
        spec.loader.exec_module(mod)
        return mod, None
    except Exception as e:  # noqa: BLE001  (numba missing, ...)
        return None, f"import of baseline/_ref/{module}.py failed: {type(e).__name__}: {e}"


def numba_reference_leg(n, ny_sample, steps, warm=1):
    """adi3d_numba_coeff.adi_step_numba_coeff (adi3d_numba_coeff.py:290) itself, serial, on the y-slab sample.
    The JIT is warmed on an 8^3 grid first.  Returns (field, [seconds per step], None) or (None, None, why)."""
    ref, why = load_reference("adi3d_numba_coeff")
    if ref is None:
        return None, None, why
    try:
        m8 = np.ones((8, 8, 8), dtype=bool)
        g8, mt = ref.Grid3D(8, 8, 8, DX, m8), ref.Material(RHO, CP, K)
        p8 = ref.precompute_coeff_packs_unified(g8, mt, robin_h={f: 10.0 for f in FACES})
        ref.adi_step_numba_coeff(np.full((8, 8, 8), 300.0), g8, mt, ref.Params(DT, THETA), p8, Tinf=TINF)
        mask, T0, h = host_inputs(n, ny_sample)
        grid = ref.Grid3D(n, ny_sample, n, DX, mask)
        packs = ref.precompute_coeff_packs_unified(grid, mt, robin_h=h)
        prm = ref.Params(DT, THETA)
        T, times = T0, []
        for i in range(warm + steps):
            t0 = time.perf_counter()
            T = ref.adi_step_numba_coeff(T, grid, mt, prm, packs, Tinf=TINF)
            if i >= warm:
                times.append(time.perf_counter() - t0)
        return T, times, None
    except Exception as e:  # noqa: BLE001
        return None, None, f"reference run failed: {type(e).__name__}: {e}"


def cyl_setup(nr, nphi, nz):
    R = 0.02
    dr = R / nr
    return dict(nr=nr, nphi=nphi, nz=nz, R=R, dr=dr, dz=dr, dphi=2.0 * math.pi / nphi, rho=7800.0, cp=490.0, k=54.0)


def cyl_host_field(c, seed=2):
    rng = np.random.default_rng(seed)
    T = 20.0 + 5.0 * rng.random((c["nr"], c["nphi"], c["nz"]))
    T[:, :, c["nz"] - max(1, c["nz"] // 32):] = 1000.0
    return T


def cyl_cpu_legs(c, want_reference=True):
    """One backward-Euler step of the configs[2] problem at size c on the host: the NumPy restatement
    (oracle/cyl.py) and, when present, the reference's own adi3d_cyl_phi_v3.adi_step.  Both are serial."""
    from oracle import cyl
    alpha = c["k"] / (c["rho"] * c["cp"])
    dt = min(c["dr"] ** 2, c["dz"] ** 2, (c["R"] * c["dphi"]) ** 2) / alpha
    T0 = cyl_host_field(c)
    args = (c["nr"], c["nphi"], c["nz"], c["dr"], c["dphi"], c["dz"], c["R"])
    zk = dict(kind_bot="neumann0", kind_top="robin", h_top=500.0, T_inf_top=20.0)
    t0 = time.perf_counter()
    out = cyl.adi_step(T0, cyl.GridCyl(*args), cyl.Material(c["rho"], c["cp"], c["k"]), cyl.Params(dt, 1.0, "be"),
                       cyl.RobinR(500.0, 20.0), cyl.ZBC(**zk))
    t_port = time.perf_counter() - t0
    cells = c["nr"] * c["nphi"] * c["nz"]
    rec = {"value": cells / t_port, "unit": UNIT, "cores": 1, "kind": "port",
           "sample": f"one step at {c['nr']}x{c['nphi']}x{c['nz']} (oracle/cyl.py, NumPy, serial like the reference)"}
    if want_reference:
        ref, why = load_reference("adi3d_cyl_phi_v3")
        if ref is None:
            rec["reference_unavailable"] = why
        else:
            try:
                t0 = time.perf_counter()
                o2 = ref.adi_step(T0, ref.GridCyl(*args), ref.Material(c["rho"], c["cp"], c["k"]),
                                  ref.Params(dt, 1.0, "be"), ref.RobinR(500.0, 20.0), ref.ZBC(**zk))
                t_ref = time.perf_counter() - t0
                rec["reference_value"] = cells / t_ref
                rec["reference_note"] = "adi3d_cyl_phi_v3.adi_step of the unmodified reference (baseline/_ref), 1 core"
                rec["port_vs_reference_rel_l2"] = rel_l2(out, o2)
            except Exception as e:  # noqa: BLE001
                rec["reference_unavailable"] = f"{type(e).__name__}: {e}"
    return rec, (T0, out, dt)


def run_reference(args):
    """--impl reference: the reference's own CPU path on the host cores (rank 0 only)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    n = args.size
    total = args.warmup + args.steps
    # bounded sample: a y-slab of the workload (x and z lines keep their full length), sized so that the whole
    # run stays near a minute at the reference's ~4.5 M cell-steps/s
    ny_s = args.ref_ny if args.ref_ny > 0 else int(max(4, min(64, 2.6e8 / (total * n * n))))
    ny_s = min(n, ny_s)
    cells = n * ny_s * n
    kind, cores, note = "reference", 1, None
    _, times, why = numba_reference_leg(n, ny_s, args.steps, warm=args.warmup)
    port_value = None
    try:
        _, tp, _ = cpu_leg(n, ny_s, host_threads(), steps=2)
        port_value = cells / tp[-1]
    except Exception:  # noqa: BLE001
        pass
    if times is None:
        # the reference is not on this box: its C restatement with all host threads
        kind, cores, note = "port", host_threads(), why
        _, t_all, _ = cpu_leg(n, ny_s, cores, steps=total)
        times = t_all[args.warmup:]
    dt_step = sum(times) / len(times)
    value = cells / dt_step
    sample = f"{n}x{ny_s}x{n} y-slab of the workload per step"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt_step * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        # the arm's config is our arm's config at this N, verbatim; what was actually stepped is cpu_baseline.sample
        "data": "synthetic", "config": workload_config(args, max(1, args.gpus)),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample,
                         "host_cpus": os.cpu_count(),
                         "what": ("adi3d_numba_coeff.adi_step_numba_coeff of the unmodified reference (baseline/_ref): serial "
                                  "Numba, 1 core -- the reference has no threaded path") if kind == "reference" else
                                 "C restatement of the reference (oracle/adi_oracle.c), OpenMP over lines",
                         "port_value_all_threads": port_value, "port_threads": host_threads(), "note": note},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def workload_config(args, world):
    n = args.size
    return {"workload": f"single_track_on_plate {n}^3 (BASELINE configs[1]): plate+track mask, Robin x6, "
                        f"per-face variable h (dense coeff per axis), theta={THETA}, dt={DT}",
            "grid": [n, n, n * world], "cells": n ** 3 * world, "bytes_per_cell_step": 75,
            "l2": "fields (1.07 GB each at 512^3) exceed the 126 MB L2; no flush needed",
            "parallelism": "single GPU" if world == 1 else
            f"z-slab x{world}: grid {n}x{n}x{n * world}, one {n}^3 slab per GPU; per step 2 T-plane halo "
            f"messages per interior boundary + all-gather of 2 doubles per z line and rank (NCCL)"}


def guarded(fn, *a, **kw):
    """Sub-records must not take the headline down with them."""
    try:
        return fn(*a, **kw)
    except Exception as e:  # noqa: BLE001
        return {"error": f"{type(e).__name__}: {e}", "trace": traceback.format_exc(limit=3)}


def free_cuda():
    import gc
    import torch
    gc.collect()
    torch.cuda.empty_cache()


# ------------------------------------------------------------------------------------------
# GPU arm, N = 1: plate 512^3
# ------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the engine has no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    if args.workload == "c4":
        line = bench_c4(args, local)
    elif args.workload == "cyl":
        line = bench_cyl(args, local)
    elif world > 1 or args.workload == "c5":
        line = bench_slab(args, rank, world, local, c5=args.workload == "c5")
        if world > 1 and args.workload == "plate" and not args.no_extras:
            free_cuda()
            c5 = guarded(bench_slab, args, rank, world, local, c5=True, sub=True)
            free_cuda()
            t1 = guarded(bench_c5_single, args, local) if rank == 0 else None
            dist.barrier()
            free_cuda()
            par = guarded(slab_parity, rank, world, local)
            if rank == 0:
                if isinstance(c5, dict) and isinstance(t1, dict) and "ms_per_step" in c5 and "ms_per_step" in t1:
                    c5["t1_ms_per_step"] = t1["ms_per_step"]
                    c5["t1_step_frac"] = t1.get("step_frac")
                    c5["efficiency_vs_c5_n1"] = t1["ms_per_step"] / (world * c5["ms_per_step"])
                elif isinstance(c5, dict):
                    c5["t1"] = t1
                line["c5_strong"] = c5
                line["parity"] = par
    else:
        line = bench_plate(args, local)
        if not args.no_extras:
            free_cuda()
            line["c3_cyl"] = guarded(bench_cyl, args, local, sub=True)
            free_cuda()
            line["c4_waam"] = guarded(bench_c4, args, local, sub=True)
            free_cuda()
            line["c5_n1"] = guarded(bench_c5_single, args, local)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def bench_plate(args, local):
    import torch
    from adi_thermal_fields_b200 import _capi, adi3d_gpu_coeff as g, devarray as cp

    n = args.size
    dev = torch.device("cuda", local)
    # ---- synthetic inputs, created on the device (outside any timed region) ----
    mask = torch.zeros((n, n, n), dtype=torch.bool, device=dev)
    nzp = n - max(1, n // 64)
    mask[:, :, :nzp] = True
    mask[: max(1, n // 32), : n // 2, nzp:] = True
    gen = torch.Generator(device=dev).manual_seed(0)
    T0 = torch.full((n, n, n), TINF, dtype=torch.float64, device=dev)
    T0 = torch.where(mask, 20.0 + 1380.0 * torch.rand((n, n, n), dtype=torch.float64, device=dev, generator=gen), T0)
    gen2 = torch.Generator(device=dev).manual_seed(1234)
    grid = g.Grid3D.__new__(g.Grid3D)
    grid.nx = grid.ny = grid.nz = n
    grid.dx = DX
    grid.mask = cp.ndarray(mask)
    mat = g.Material(RHO, CP, K)
    prm = g.Params(DT, THETA)
    h = {}
    for f in FACES:
        h[f] = cp.ndarray(10.0 * (0.3 + torch.rand((n, n, n), dtype=torch.float64, device=dev, generator=gen2)))
    packs = g.precompute_coeff_packs_unified(grid, mat, robin_h=h)
    del h
    torch.cuda.synchronize()

    L = _capi.load()
    eng = g._engine
    eng.bind(grid); eng.set_mask(grid); eng.set_packs(packs)
    ctx = eng.context()
    kappa = K / (RHO * CP)
    for o in args.opt:
        name, _, val = o.partition("=")
        _capi.check(L.adi_set_option(ctx, name.encode(), int(val)), "adi_set_option")
    A = T0.clone()
    B = torch.empty_like(A)
    stream = torch.cuda.current_stream()

    def step(src, dst):
        _capi.check(L.adi_cart_step(ctx, src.data_ptr(), dst.data_ptr(), DT, THETA, kappa, TINF,
                                    stream.cuda_stream), "adi_cart_step")

    for _ in range(args.warmup):
        step(A, B); A, B = B, A
    torch.cuda.synchronize()

    # ---- timed region: exactly K steps, device-resident ----
    L.adi_set_option(ctx, b"profile", 1)
    L.adi_profile_reset(ctx)
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = L.adi_launch_count(ctx)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(stream)
    for _ in range(args.steps):
        step(A, B); A, B = B, A
    e1.record(stream)
    torch.cuda.synchronize()
    ms_total = e0.elapsed_time(e1)
    launches = L.adi_launch_count(ctx) - launches0
    ms3 = (C.c_double * 4)()
    nst = C.c_long()
    L.adi_profile_read(ctx, ms3, C.byref(nst))
    L.adi_set_option(ctx, b"profile", 0)
    # a longer run of the same loop (not the reported value): steadier clocks sample, stability check
    long_steps = max(args.steps, args.long_steps)
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record(stream)
    for _ in range(long_steps):
        step(A, B); A, B = B, A
    e3.record(stream)
    torch.cuda.synchronize()
    ms_long = e2.elapsed_time(e3) / long_steps
    clocks = sampler.stop()
    ms_per_step = ms_total / args.steps
    cells = n ** 3
    value = cells * args.steps / (ms_total * 1e-3)

    # ---- e2e: host arrays through the C ABI (H2D + step + D2H per step) ----
    # e2e.value: one blocking adi_cart_step_host call per step -- the reference's calling convention for a time
    # loop (T = step(T): step n+1 needs the result of step n on the host).  pipelined_value: independent host
    # fields (an ensemble) through two staging slots on two streams (adi_cart_step_host_async), upload of one
    # overlapping compute of another and the download of a third.
    e2e_steps = max(2, min(args.steps, args.e2e_steps))
    hin = [torch.empty((n, n, n), dtype=torch.float64, pin_memory=True) for _ in range(2)]
    hout = [torch.empty((n, n, n), dtype=torch.float64, pin_memory=True) for _ in range(2)]
    for hbuf in hin:
        hbuf.copy_(T0)
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    torch.cuda.synchronize()

    def host_step(i=0):
        _capi.check(L.adi_cart_step_host(ctx, hin[i].data_ptr(), hout[i].data_ptr(), 1, DT, THETA, kappa, TINF,
                                         stream.cuda_stream), "adi_cart_step_host")

    def host_step_async(i):
        _capi.check(L.adi_cart_step_host_async(ctx, i & 1, hin[i & 1].data_ptr(), hout[i & 1].data_ptr(), DT, THETA,
                                               kappa, TINF, streams[i & 1].cuda_stream), "adi_cart_step_host_async")

    host_step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for s in range(e2e_steps):
        host_step(0)
        hin[0], hout[0] = hout[0], hin[0]          # dependent steps: the result is the next input
    torch.cuda.synchronize()
    t_serial = time.perf_counter() - t0
    for hbuf in hin:
        hbuf.copy_(T0)
    for i in range(2):
        host_step_async(i)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        host_step_async(i)
    torch.cuda.synchronize()
    t_pipe = time.perf_counter() - t0
    e2e_ok = bool(torch.equal(hout[0], hout[1]))   # both slots stepped the same input
    del hin, hout

    # ---- roofline of the slowest kernel (live CUDA events inside the engine) ----
    peak, peak_src = peaks()
    per = [ms3[i] / max(1, nst.value) for i in range(4)]
    sparse = int(L.adi_get_option(ctx, b"sparse_active"))
    # the z sweep of lines longer than 128 cells runs k_sweep_zt unless its coefficient field must be read densely
    zt = bool(L.adi_get_option(ctx, b"zt")) and n > 128 and ((sparse >> 2) & 1)
    names = ["k_explicit", "k_sweep_xy<x>", "k_sweep_xy<y>", "k_sweep_zt" if zt else "k_sweep_z"]
    # algorithmic bytes per cell: explicit stage T in 8 + code 1 + out 8; a sweep
    # in 8 + out 8 + code 1 + dense coeff 8.  The 75 B/cell-step of the metric counts the fused
    # form (3 sweeps); the separate explicit pass is extra real traffic, not extra credit.
    bpc = [17.0, 25.0, 25.0, 25.0]
    # what the kernels really fetch: a sweep whose coefficient field was verified surface-only (option
    # sparse_coeff, x / y sweeps) skips the 8 B/cell of interior coefficient reads
    moved = [17.0] + [17.0 if (sparse >> a) & 1 else 25.0 for a in range(3)]
    dom = int(np.argmax(per))
    bytes_per_launch = bpc[dom] * cells
    achieved = bytes_per_launch / (per[dom] * 1e-3) / 1e9
    kern_key = ["k_explicit", "k_sweep_xy<0", "k_sweep_xy<1", "k_sweep_zt" if zt else "k_sweep_z"][dom]
    traffic = traffic_of(kern_key) or traffic_of(kern_key.split("<")[0])
    roofline = {"bound": "hbm", "kernel": names[dom], "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": bytes_per_launch,
                # the same kernel against the bytes it really moves (its share of moved_bytes_per_cell_step; `traffic`
                # is the ncu DRAM figure of the same launch shape): the DRAM-side utilisation
                "moved_bytes_per_launch": moved[dom] * cells,
                "achieved_moved": moved[dom] * cells / (per[dom] * 1e-3) / 1e9,
                "frac_moved": moved[dom] * cells / (per[dom] * 1e-3) / 1e9 / peak,
                "kernel_ms": {"explicit": per[0], "x": per[1], "y": per[2], "z": per[3]},
                "kernel_GBs": {k: b * cells / (t * 1e-3) / 1e9 if t > 0 else None
                               for k, b, t in zip(("explicit", "x", "y", "z"), bpc, per)},
                "kernel_moved_GBs": {k: b * cells / (t * 1e-3) / 1e9 if t > 0 else None
                                     for k, b, t in zip(("explicit", "x", "y", "z"), moved, per)},
                "kernel_moved_frac": {k: b * cells / (t * 1e-3) / 1e9 / peak if t > 0 else None
                                      for k, b, t in zip(("explicit", "x", "y", "z"), moved, per)},
                "step_achieved_GBs": 75.0 * cells / (ms_per_step * 1e-3) / 1e9,
                "step_frac": 75.0 * cells / (ms_per_step * 1e-3) / 1e9 / peak,
                "moved_bytes_per_cell_step": sum(moved),
                "step_frac_moved": sum(moved) * cells / (ms_per_step * 1e-3) / 1e9 / peak,
                "note": "achieved/frac use SURVEY 8(d)'s algorithmic bytes (75 B/cell-step = 3 sweeps x 25 B, explicit "
                        "stage counted as fused).  The kernels move moved_bytes_per_cell_step: +17 B for the separate "
                        "explicit pass, -8 B for each sweep that reads its verified surface-only coefficient field at "
                        "exposed cells only (sparse_active bitmask %d); kernel_moved_* are per-kernel DRAM-side figures" % sparse}

    # ---- CPU baseline + parity on a bounded sample ----
    cpu = None
    parity = None
    if not args.no_cpu:
        nth = host_threads()
        ny_s = min(n, args.cpu_ny)
        Tref, times, (m_s, T0_s, h_s) = cpu_leg(n, ny_s, nth, steps=1)
        v_all = n * ny_s * n / times[0]
        ny_1 = max(8, ny_s // 8)
        _, t1, _ = cpu_leg(n, ny_1, 1, steps=1)
        v_one = n * ny_1 * n / t1[0]
        cpu = {"value": v_all, "unit": UNIT, "cores": nth, "kind": "port", "host_cpus": os.cpu_count(),
               "sample": f"one step of a {n}x{ny_s}x{n} y-slab of the workload, OpenMP over lines",
               "value_1core": v_one,
               "note": "value / value_1core: C restatement of the reference (bit-identical on the golden vectors); "
                       "numba_value: the reference's own serial Numba path on this box (1 core -- it has no threads)"}
        ny_n = max(4, min(ny_s, args.numba_ny))
        Tnb, tn, why = numba_reference_leg(n, ny_n, 2, warm=1)
        if tn is not None:
            cpu["numba_value"] = n * ny_n * n / min(tn)
            cpu["numba_sample"] = f"best of 2 steps of a {n}x{ny_n}x{n} y-slab, JIT warmed on 8^3, 1 core"
            if ny_n == ny_s:
                cpu["numba_vs_port_rel_l2"] = rel_l2(Tnb, Tref, m_s)
        else:
            cpu["numba_unavailable"] = why
        # the same sample through the GPU engine
        gs = g.Grid3D(n, ny_s, n, DX, m_s)
        ps = g.precompute_coeff_packs_unified(gs, mat, robin_h=h_s)
        out = g.adi_step_gpu_coeff(cp.asarray(T0_s), gs, mat, prm, ps, Tinf=TINF)
        Tg = cp.asnumpy(out)
        parity = {"rel_l2_vs_oracle": rel_l2(Tg, Tref, m_s), "void_bit_equal": bool(np.array_equal(Tg[~m_s], T0_s[~m_s])),
                  "sample": f"{n}x{ny_s}x{n}", "tol": 1e-12}

    return {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(args, 1),
        "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks,
        "long_run": {"steps": long_steps, "ms_per_step": ms_long},
        "e2e": {"value": cells * e2e_steps / t_serial, "unit": UNIT, "h2d_bytes_per_step": 8 * cells,
                "d2h_bytes_per_step": 8 * cells, "steps": e2e_steps,
                "api": "adi_cart_step_host: one blocking call per step, each step's input is the previous result on the host "
                       "(pinned arrays; the reference's T = step(T) loop)",
                "pcie_floor_ms": None, "pipelined_value": cells * e2e_steps / t_pipe,
                "pipelined_api": "adi_cart_step_host_async: independent fields, two staging slots on two streams",
                "slots_agree": e2e_ok},
        "gpu_launches": int(launches), "parity": parity,
    }


# ------------------------------------------------------------------------------------------
# z-slab runs (N > 1 plate weak scaling; configs[4] strong scaling)
# ------------------------------------------------------------------------------------------
class _Mat:
    rho, cp, k = RHO, CP, K


class _Prm:
    dt, theta = DT, THETA


def bench_slab(args, rank, world, local, c5=False, sub=False):
    """z-slab partitioned step over `world` ranks.  c5: BASELINE configs[4] (strong scaling); else one n^3 plate
    slab per rank (weak scaling).  Returns the JSON record on rank 0 (None elsewhere)."""
    import torch
    import torch.distributed as dist
    from adi_thermal_fields_b200 import slab

    dev = torch.device("cuda", local)
    comm = slab.TorchDistComm() if world > 1 else slab.LocalComm(1).view(0)
    steps = args.steps if not sub else max(3, min(args.steps, args.c5_steps))
    warmup = args.warmup if not sub else 3
    if c5:
        # BASELINE configs[4] / SURVEY 8d C5: full mask, scalar Robin h = 10 on the six faces
        # (coefficients derived from the mask on the fly: 51 B/cell-step), strong scaling
        nx = ny = 4 * args.size
        nzg = 2 * args.size
        nzl = nzg // world
        mask = torch.ones((nx, ny, nzl), dtype=torch.bool, device=dev)
        gen = torch.Generator(device=dev).manual_seed(5 + rank)
        T0 = 20.0 + 1380.0 * torch.rand((nx, ny, nzl), dtype=torch.float64, device=dev, generator=gen)
        grid = slab.SlabGrid3D(nx, ny, nzl, DX, mask, comm)
        packs = slab.precompute_coeff_packs_unified(grid, _Mat, robin_h={f: 10.0 for f in FACES})
        bpc = 51.0
        wl = {"workload": f"synthetic Cartesian {nx}x{ny}x{nzg} (BASELINE configs[4]): full mask, scalar Robin h=10 x6 "
                          f"(on-the-fly coefficients), theta={THETA}, dt={DT}; strong scaling",
              "grid": [nx, ny, nzg], "cells": nx * ny * nzg, "bytes_per_cell_step": 51,
              "l2": "fields (34 GB) exceed the 126 MB L2; no flush needed",
              "parallelism": f"z-slab x{world}: {nzl} planes per GPU; per step T-plane halos + all-gather of 2 doubles "
                             f"per z line and rank (NCCL)"}
        scaling = "strong"
    else:
        n = args.size
        nx = ny = nzl = n
        nzg = n * world
        z0 = rank * n
        # plate + track on the global grid (single_track_on_plate.py:113-114,159): plate below
        # nzg - n/64, track on top of it in the last slab
        nzp = nzg - max(1, n // 64)
        kz = torch.arange(z0, z0 + n, device=dev)
        mask = (kz < nzp)[None, None, :].expand(n, n, n).clone()
        if rank == world - 1:
            mask[: max(1, n // 32), : n // 2, (nzp - z0):] = True
        gen = torch.Generator(device=dev).manual_seed(rank)
        T0 = torch.full((n, n, n), TINF, dtype=torch.float64, device=dev)
        T0 = torch.where(mask, 20.0 + 1380.0 * torch.rand((n, n, n), dtype=torch.float64, device=dev, generator=gen), T0)
        grid = slab.SlabGrid3D(n, n, n, DX, mask, comm)
        gen2 = torch.Generator(device=dev).manual_seed(1234 + rank)
        h = {f: 10.0 * (0.3 + torch.rand((n, n, n), dtype=torch.float64, device=dev, generator=gen2)) for f in FACES}
        packs = slab.precompute_coeff_packs_unified(grid, _Mat, robin_h=h)
        del h
        bpc = 75.0
        wl = workload_config(args, world)
        scaling = "weak"
    cells = nx * ny * nzl          # per rank
    for o in args.opt:
        name, _, val = o.partition("=")
        grid.be.set_option(name, int(val))
    torch.cuda.synchronize()
    A, B = T0, torch.empty_like(T0)
    stream = torch.cuda.current_stream()

    def step(src, dst):
        slab.adi_step_gpu_coeff(src, grid, _Mat, _Prm, packs, Tinf=TINF, out=dst)

    for _ in range(warmup):
        step(A, B); A, B = B, A

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    grid.be.profile(True)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = grid.be.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(steps):
        step(A, B); A, B = B, A
    e1.record(stream)
    barrier()
    ms_total = e0.elapsed_time(e1)
    launches = grid.be.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    ms3, nst = grid.be.profile_read()
    grid.be.profile(False)
    ms_total = allmax(ms_total)
    ms_per_step = ms_total / steps
    value = world * cells * steps / (ms_total * 1e-3)

    # e2e: pinned host slabs in and out every step (H2D + step + D2H inside the timed region), one dependent
    # step after the other; pipelined: two device slots, copies on their own streams (independent fields)
    e2e = None
    if not sub:
        e2e_steps = max(2, min(steps, args.e2e_steps))
        if 8 * cells > 12 * 2 ** 30:   # two pinned slabs of > 12 GB each: not attempted
            e2e_steps = 0
        if e2e_steps:
            hin = [torch.empty((nx, ny, nzl), dtype=torch.float64, pin_memory=True) for _ in range(2)]
            hout = [torch.empty((nx, ny, nzl), dtype=torch.float64, pin_memory=True) for _ in range(2)]
            for hb in hin:
                hb.copy_(A)
            torch.cuda.synchronize()

            def host_step():
                A.copy_(hin[0], non_blocking=True)
                step(A, B)
                hout[0].copy_(B, non_blocking=True)
                torch.cuda.synchronize()
                hin[0], hout[0] = hout[0], hin[0]

            host_step()
            barrier()
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                host_step()
            t_serial = allmax(time.perf_counter() - t0)
            # pipelined: in-slot s is filled on the copy-in stream while the step of the other slot runs
            cin, cout = torch.cuda.Stream(), torch.cuda.Stream()
            Ain = [A, torch.empty_like(A)]
            Bout = [B, torch.empty_like(B)]
            ev_in = [torch.cuda.Event() for _ in range(2)]
            ev_cmp = [torch.cuda.Event() for _ in range(2)]
            ev_out = [torch.cuda.Event() for _ in range(2)]

            def pipe(nsteps):
                for i in range(nsteps):
                    s = i & 1
                    with torch.cuda.stream(cin):
                        cin.wait_event(ev_cmp[s])          # the slot's previous step has consumed its input
                        Ain[s].copy_(hin[s], non_blocking=True)
                        ev_in[s].record(cin)
                    stream.wait_event(ev_in[s])
                    stream.wait_event(ev_out[s])           # the slot's previous result has left
                    step(Ain[s], Bout[s])
                    ev_cmp[s].record(stream)
                    with torch.cuda.stream(cout):
                        cout.wait_event(ev_cmp[s])
                        hout[s].copy_(Bout[s], non_blocking=True)
                        ev_out[s].record(cout)
                torch.cuda.synchronize()

            pipe(2)
            barrier()
            t0 = time.perf_counter()
            pipe(e2e_steps)
            t_pipe = allmax(time.perf_counter() - t0)
            e2e = {"value": world * cells * e2e_steps / t_serial, "unit": UNIT, "h2d_bytes_per_step": 8 * cells * world,
                   "d2h_bytes_per_step": 8 * cells * world, "steps": e2e_steps,
                   "api": "slab.adi_step_gpu_coeff on pinned host slabs, one dependent step after the other "
                          "(H2D + step + D2H per rank, blocking)",
                   "pipelined_value": world * cells * e2e_steps / t_pipe,
                   "pipelined_api": "independent fields: two device slots per rank, copy-in / step / copy-out on three streams"}
            del hin, hout, Ain, Bout
        else:
            e2e = {"value": None, "unit": UNIT, "h2d_bytes_per_step": 8 * cells * world, "d2h_bytes_per_step": 8 * cells * world,
                   "steps": 0, "api": "not attempted: pinned host slabs of more than 12 GB per rank"}

    if rank != 0:
        return None
    peak, peak_src = peaks()
    per = [ms3[i] / max(1, nst) for i in range(4)]
    solve_first = grid.z_form() == "solve-first"
    names = ["k_explicit", "k_sweep_xy<x>", "k_sweep_xy<y>",
             "k_sweep_z solve-first + all-gather + k_spike_apply" if solve_first
             else ("k_sweep_z" if world == 1 else "k_sweep_z pass1 + all-gather + pass2")]
    # N>1, two-pass form: the z sweep reads the slab twice (pass 1 without the write); the solve-first
    # form (steady stepping) makes one pass and touches ~20 cells per line next to each slab face
    s1 = bpc / 3.0   # one sweep: 25 B/cell dense, 17 B/cell scalar Robin
    alg = [17.0 * cells, s1 * cells, s1 * cells,
           (s1 + (s1 - 8.0 if (world > 1 and not solve_first) else 0.0)) * cells]
    dom = int(np.argmax(per))
    achieved = alg[dom] / (per[dom] * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": names[dom], "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg[dom], "per": "GPU (rank 0)",
                "kernel_ms": {"explicit": per[0], "x": per[1], "y": per[2], "z": per[3]},
                "step_achieved_GBs": bpc * cells / (ms_per_step * 1e-3) / 1e9,
                "step_frac": bpc * cells / (ms_per_step * 1e-3) / 1e9 / peak}
    rec = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps,
        "warmup": warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": scaling,
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": wl,
        "roofline": roofline, "cpu_baseline": None, "clocks": clocks, "e2e": e2e,
        "gpu_launches": int(launches), "parity": None,
        "exchange": {"halo_bytes_per_step_per_boundary": 2 * 8 * nx * ny,
                     "allgather_bytes_per_rank_per_step": 2 * 8 * nx * ny,
                     "allgather_bytes_per_rank_on_matrix_change": 4 * 8 * nx * ny,
                     "z_form": "solve-first" if solve_first else ("single GPU" if world == 1 else "two-pass")},
    }
    if sub:
        for k in ("metric", "unit", "higher_is_better", "vs_baseline", "dtype", "data", "cpu_baseline", "e2e", "parity", "clocks"):
            rec.pop(k, None)
        rec["step_frac"] = roofline["step_frac"]
        rec["kernel_ms"] = roofline["kernel_ms"]
    return rec


def bench_c5_single(args, local):
    """BASELINE configs[4] on ONE GPU (the T1 of the strong-scaling target): 2048 x 2048 x 1024, full mask, scalar
    Robin (51 B/cell-step), through the plain single-GPU step."""
    import torch
    from adi_thermal_fields_b200 import _capi, adi3d_gpu_coeff as g, devarray as cp

    dev = torch.device("cuda", local)
    nx = ny = 4 * args.size
    nz = 2 * args.size
    cells = nx * ny * nz
    mask = torch.ones((nx, ny, nz), dtype=torch.bool, device=dev)
    grid = g.Grid3D.__new__(g.Grid3D)
    grid.nx, grid.ny, grid.nz, grid.dx, grid.mask = nx, ny, nz, DX, cp.ndarray(mask)
    mat = g.Material(RHO, CP, K)
    packs = g.precompute_coeff_packs_unified(grid, mat, robin_h={f: 10.0 for f in FACES})
    gen = torch.Generator(device=dev).manual_seed(5)
    A = torch.empty((nx, ny, nz), dtype=torch.float64, device=dev)
    for i0 in range(0, nx, 256):     # filled in pieces: no 34 GB temporaries
        A[i0:i0 + 256] = 20.0 + 1380.0 * torch.rand((min(256, nx - i0), ny, nz), dtype=torch.float64, device=dev, generator=gen)
    B = torch.empty_like(A)
    eng = g._engine
    eng.bind(grid); eng.set_mask(grid); eng.set_packs(packs)
    L, ctx = _capi.load(), eng.context()
    kappa = K / (RHO * CP)
    st = torch.cuda.current_stream()

    def step(s, d):
        _capi.check(L.adi_cart_step(ctx, s.data_ptr(), d.data_ptr(), DT, THETA, kappa, TINF, st.cuda_stream), "adi_cart_step")

    for _ in range(2):
        step(A, B); A, B = B, A
    torch.cuda.synchronize()
    steps = max(3, min(args.steps, args.c5_steps))
    L.adi_set_option(ctx, b"profile", 1)
    L.adi_profile_reset(ctx)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(steps):
        step(A, B); A, B = B, A
    e1.record(st)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    ms4 = (C.c_double * 4)()
    nst = C.c_long()
    L.adi_profile_read(ctx, ms4, C.byref(nst))
    L.adi_set_option(ctx, b"profile", 0)
    peak, _ = peaks()
    per = [ms4[i] / max(1, nst.value) for i in range(4)]
    del A, B, mask, packs
    eng.bound = None   # the engine context is shared with the other workloads of this process
    return {"workload": f"synthetic Cartesian {nx}x{ny}x{nz} (BASELINE configs[4]) on one GPU, scalar Robin, 51 B/cell-step",
            "steps": steps, "ms_per_step": ms, "value": cells / (ms * 1e-3), "unit": UNIT,
            "kernel_ms": {"explicit": per[0], "x": per[1], "y": per[2], "z": per[3]},
            "step_achieved_GBs": 51.0 * cells / (ms * 1e-3) / 1e9, "step_frac": 51.0 * cells / (ms * 1e-3) / 1e9 / peak,
            "kernel_GBs": {k: 17.0 * cells / (t * 1e-3) / 1e9 if t > 0 else None for k, t in zip(("explicit", "x", "y", "z"), per)}}


def slab_parity(rank, world, local):
    """256 x 256 x 128 miniatures of the two slab workloads (configs[4]: full mask, scalar Robin; plate: plate+track,
    dense per-face h) stepped over NCCL on the `world` ranks and checked on rank 0 against the oracle on the
    undivided grid.  The oracle is the checker here, after every timed region."""
    import torch
    import torch.distributed as dist
    from adi_thermal_fields_b200 import slab
    from oracle import cart

    dev = torch.device("cuda", local)
    comm = slab.TorchDistComm()
    nx = ny = 256
    nz = 128
    out = {"grid": [nx, ny, nz], "tol_per_step": 1e-12, "steps": 4,
           "sequencing": "adi_cart_slab_step (inside the library, own NCCL communicator)" if slab.USE_LIBRARY_SEQUENCING
           else "adi_thermal_fields_b200/slab.py (torch.distributed collectives)"}
    ext = slab.split_z(nz, world, multiple=16)
    z0, z1 = ext[rank]
    for name in ("c5_scalar_full", "plate_dense"):
        rng = np.random.default_rng(77)
        if name == "c5_scalar_full":
            mask = np.ones((nx, ny, nz), dtype=bool)
            bcs = dict(robin_h={f: 10.0 for f in FACES})
        else:
            mask = np.zeros((nx, ny, nz), dtype=bool)
            mask[:, :, :nz - 4] = True
            mask[:8, :ny // 2, nz - 4:] = True
            bcs = dict(robin_h={f: 10.0 * (0.3 + rng.random((nx, ny, nz))) for f in FACES})
        T0 = np.full((nx, ny, nz), TINF)
        T0[mask] = 20.0 + 1380.0 * rng.random(int(mask.sum()))

        def cut(v):
            return np.ascontiguousarray(v[:, :, z0:z1]) if isinstance(v, np.ndarray) else v
        grid = slab.SlabGrid3D(nx, ny, z1 - z0, DX, mask[:, :, z0:z1], comm)
        packs = slab.precompute_coeff_packs_unified(grid, _Mat, robin_h={f: cut(v) for f, v in bcs["robin_h"].items()})
        T = torch.from_numpy(np.ascontiguousarray(T0[:, :, z0:z1])).to(dev)
        for _ in range(out["steps"]):
            T = slab.adi_step_gpu_coeff(T, grid, _Mat, _Prm, packs, Tinf=TINF)
        parts = [None] * world
        dist.all_gather_object(parts, (z0, z1, T.cpu().numpy(), grid.z_form() == "solve-first"))
        if rank == 0:
            full = np.empty((nx, ny, nz))
            for a, b, t, _ in parts:
                full[:, :, a:b] = t
            cart.set_threads(host_threads())
            hg, hm = cart.Grid3D(nx, ny, nz, DX, mask), cart.Material(RHO, CP, K)
            hp = cart.precompute_coeff_packs_unified(hg, hm, **bcs)
            ref = T0
            for _ in range(out["steps"]):
                ref = cart.adi_step_numba_coeff(ref, hg, hm, cart.Params(DT, THETA), hp, Tinf=TINF)
            out[name] = {"rel_l2_vs_oracle": rel_l2(full, ref, mask),
                         "void_bit_equal": bool(np.array_equal(full[~mask], T0[~mask])),
                         "z_form_last_step": "solve-first" if parts[0][3] else "two-pass"}
        del grid, packs, T
    if rank != 0:
        return None
    worst = max(out[k]["rel_l2_vs_oracle"] for k in ("c5_scalar_full", "plate_dense"))
    out["rel_l2_vs_oracle"] = worst
    out["ok"] = bool(worst <= out["steps"] * 1e-12 and all(out[k]["void_bit_equal"] for k in ("c5_scalar_full", "plate_dense")))
    return out


# ------------------------------------------------------------------------------------------
# configs[3]: waam 1024^3
# ------------------------------------------------------------------------------------------
def waam_head(n, dev):
    import torch
    ax = (torch.arange(n, device=dev, dtype=torch.float64) + 0.5) / n - 0.5
    X, Y, Z = ax[:, None, None], ax[None, :, None], ax[None, None, :]
    return ((X / 0.35) ** 2 + (Y / 0.42) ** 2 + ((Z - 0.05) / 0.45) ** 2 <= 1.0) | \
           ((X * X + Y * Y <= 0.12 ** 2) & (Z < -0.3))


def bench_c4(args, local, sub=False):
    """BASELINE configs[3] (waam_from_stl_v7_mm at 1024^3; the STL is not in the tree): a
    synthetic head (ellipsoid + neck cylinder) is deposited bottom-up, `n_per_layer` z planes per birth
    (activate_layer, waam_from_stl_v7_mm.py:487-495), the packs are rebuilt on the device after every birth
    (precompute_coeff_packs_unified with per-face dense h fields standing in for voxel_bc_correction's
    output) and `steps_per_layer` ADI steps follow.  Timed: births + pack rebuilds + steps."""
    import torch
    from adi_thermal_fields_b200 import adi3d_gpu_coeff as g, devarray as cp

    n = 2 * args.size
    dev = torch.device("cuda", local)
    rho, cp_, k = 7800.0, 490.0, 54.0
    kappa = k / (rho * cp_)
    dt = 2000.0 * DX * DX / kappa           # cfl = 2000 (waam_from_stl_v7_mm.py:355-363)
    full = waam_head(n, dev)
    gen = torch.Generator(device=dev).manual_seed(3)
    h = {f: cp.ndarray(40.0 * (0.3 + torch.rand((n, n, n), dtype=torch.float64, device=dev, generator=gen)))
         for f in FACES}
    n_per_layer, steps_per_layer = max(1, n // 64), args.c4_steps_per_layer
    act = torch.zeros((n, n, n), dtype=torch.bool, device=dev)
    T = cp.full((n, n, n), TINF, dtype=cp.float64)
    grid = g.Grid3D.__new__(g.Grid3D)
    grid.nx = grid.ny = grid.nz = n
    grid.dx = DX
    grid.mask = cp.ndarray(act)
    mat, prm = g.Material(rho, cp_, k), g.Params(dt, THETA)
    zs = torch.nonzero(full.any(0).any(0)).flatten()
    k0, k1 = int(zs[0]), int(zs[-1]) + 1
    layers = [(a, min(a + n_per_layer, k1)) for a in range(k0, k1, n_per_layer)]

    def birth(ks, ke):
        born = full[:, :, ks:ke] & ~act[:, :, ks:ke]
        T._t[:, :, ks:ke][born] = 1000.0        # Ts
        act[:, :, ks:ke] |= born
        grid.mask = cp.ndarray(act)             # rebinding, as the driver does
        return g.precompute_coeff_packs_unified(grid, mat, robin_h=h)

    kacc = [0.0, 0.0, 0.0, 0.0, 0]     # per-kernel times of the steady steps (engine events), their count
    per_layer = []                     # (mask update + pack rebuild, first step) ms of every birth

    def run(layer_list, prof=None):
        """-> steps, ms spent on births (mask update + pack rebuild), ms of the FIRST step of every layer (it rebuilds the
        neighbour code, its transposed copies and the tile lists for the new mask), ms of the other steps."""
        nonlocal T
        nsteps = 0
        tb = t1 = ts = 0.0
        for ks, ke in layer_list:
            e0, e1, e1b, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(4))
            e0.record()
            packs = birth(ks, ke)
            e1.record()
            T = g.adi_step_gpu_coeff(T, grid, mat, prm, packs, Tinf=TINF)
            e1b.record()
            if prof:
                torch.cuda.synchronize()
                prof[0].adi_profile_reset(prof[1])
            for _ in range(steps_per_layer - 1):
                T = g.adi_step_gpu_coeff(T, grid, mat, prm, packs, Tinf=TINF)
            nsteps += steps_per_layer
            e2.record()
            torch.cuda.synchronize()
            if prof:
                ms4 = (C.c_double * 4)()
                nst = C.c_long()
                prof[0].adi_profile_read(prof[1], ms4, C.byref(nst))
                for i in range(4):
                    kacc[i] += ms4[i]
                kacc[4] += nst.value
            tb += e0.elapsed_time(e1)
            t1 += e1.elapsed_time(e1b)
            ts += e1b.elapsed_time(e2)
            per_layer.append((round(e0.elapsed_time(e1), 3), round(e1.elapsed_time(e1b), 3)))
        return nsteps, tb, t1, ts

    # Warm-up = the births right below the timed ones, at least two: the second birth of a run still allocates (the
    # first generation of pack arrays is alive while the second is built; from the third on the caching allocator
    # hands the blocks back and forth), and the kernels a half-built part needs (all-uniform x / y tiles, trimmed z
    # lines) are first loaded there, not inside the timed region.
    nwarm = max(2, args.warmup // steps_per_layer)
    c4_steps = args.steps if not sub else min(args.steps, 16)
    nlay = max(1, c4_steps // steps_per_layer)
    m0 = max(nwarm, len(layers) // 2)
    mid = layers[m0: m0 + nlay]                               # mid-build: half of the head is active
    for ks, ke in layers[:m0 - nwarm]:                        # fast-forward the activation (untimed)
        born = full[:, :, ks:ke]
        T._t[:, :, ks:ke][born] = 1000.0
        act[:, :, ks:ke] |= born
    from adi_thermal_fields_b200 import _capi
    Lc, cctx = _capi.load(), g._engine.context()
    for o in args.opt:
        name, _, val = o.partition("=")
        _capi.check(Lc.adi_set_option(cctx, name.encode(), int(val)), "adi_set_option")
    run(layers[m0 - nwarm: m0])
    per_layer.clear()
    sampler = ClockSampler(local)
    sampler.start()
    l0 = g.launch_count()
    Lc.adi_set_option(cctx, b"profile", 1)
    Lc.adi_profile_reset(cctx)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    nsteps, tb, t1, ts = run(mid, prof=(Lc, cctx))
    wall = time.perf_counter() - t0
    launches = g.launch_count() - l0
    clocks = sampler.stop()
    Lc.adi_set_option(cctx, b"profile", 0)
    kernel_ms = {k: kacc[i] / max(1, kacc[4]) for i, k in enumerate(("explicit", "x", "y", "z"))}
    nlay_t = len(mid)
    steady = ts / max(1, nsteps - nlay_t) if steps_per_layer > 1 else t1 / nlay_t
    birth_ms = (tb + max(0.0, t1 - steady * nlay_t)) / nlay_t    # per birth: mask + packs + the first step's rebuilds
    tiles = {"active": int(Lc.adi_get_option(cctx, b"tiles_active")), "total": int(Lc.adi_get_option(cctx, b"tiles_total"))}
    cells = n ** 3
    active_frac = float(act.sum().item()) / cells
    peak, peak_src = peaks()
    total_ms = tb + t1 + ts
    del full, h, act, T
    g._engine.bound = None
    free_cuda()
    # CPU baseline + parity: the same deposition loop on a miniature (oracle = the reference's Numba path restated)
    cpu, parity = (None, None) if args.no_cpu else c4_miniature(local, args.c4_mini)
    line = {
        "metric": METRIC, "value": cells * nsteps / (total_ms * 1e-3), "unit": UNIT, "n_gpus": 1, "steps": nsteps,
        "warmup": nwarm * steps_per_layer, "ms_per_step": total_ms / nsteps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"waam_from_stl_v7_mm {n}^3 (BASELINE configs[3]), synthetic head (ellipsoid + neck; the STL is "
                               f"not in the tree), {n_per_layer} z planes per birth, {steps_per_layer} steps per layer, "
                               f"per-face dense h fields, theta={THETA}, cfl=2000; births + device pack rebuilds inside the timed region",
                   "grid": [n, n, n], "cells": cells, "bytes_per_cell_step": 75,
                   "active_fraction_mid_build": active_frac, "parallelism": "single GPU"},
        "roofline": {"bound": "hbm", "kernel": "whole step (explicit + x + y + z), steady state between births",
                     "achieved": 75.0 * cells / (steady * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                     "frac": 75.0 * cells / (steady * 1e-3) / 1e9 / peak, "traffic": None, "peak_source": peak_src,
                     "steady_ms_per_step": steady, "birth_ms": birth_ms, "kernel_ms": kernel_ms,
                     "birth_ms_each": [{"mask_and_packs": a, "first_step": b} for a, b in per_layer],
                     "sweep_tiles": tiles,
                     "active_cell_steps_per_s": active_frac * cells / (steady * 1e-3),
                     "birth_note": "per birth: mask update + k_build_packs (6 dense h fields -> 3 coeff fields) + what the first "
                                   "step after it spends on the neighbour code, its transposed copies and the tile lists "
                                   "(first step minus a steady step); steady_ms_per_step and kernel_ms are the other steps",
                     "note": "achieved/frac restate the metric (all cells of the box, void included, at SURVEY 8(d)'s "
                             "75 B/cell-step) in GB/s; they are not DRAM utilisation: sweep tiles without an active "
                             "cell are skipped and coefficient fields are read at exposed cells only, so the bytes "
                             "moved per step are well below 75 B x cells while the part is being built"},
        "cpu_baseline": cpu, "clocks": clocks,
        "e2e": {"value": cells * nsteps / wall, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                "api": "adi3d_gpu_coeff.precompute_coeff_packs_unified + adi_step_gpu_coeff (device arrays, host wall clock)"},
        "gpu_launches": int(launches), "parity": parity,
    }
    if sub:
        for k_ in ("metric", "unit", "higher_is_better", "vs_baseline", "dtype", "data", "scaling", "n_gpus"):
            line.pop(k_, None)
    return line


def c4_miniature(local, n):
    """The deposition loop of configs[3] on an n^3 miniature: GPU engine vs the oracle (births, pack rebuilds, steps),
    and the oracle's time for it as the CPU baseline."""
    import torch
    from adi_thermal_fields_b200 import adi3d_gpu_coeff as g, devarray as cp
    from oracle import cart
    dev = torch.device("cuda", local)
    rho, cp_, k = 7800.0, 490.0, 54.0
    kappa = k / (rho * cp_)
    dt = 2000.0 * DX * DX / kappa
    full = waam_head(n, dev).cpu().numpy()
    rng = np.random.default_rng(3)
    h = {f: 40.0 * (0.3 + rng.random((n, n, n))) for f in FACES}
    zs = np.nonzero(full.any(0).any(0))[0]
    k0, k1 = int(zs[0]), int(zs[-1]) + 1
    npl = max(1, n // 16)
    layers = [(a, min(a + npl, k1)) for a in range(k0, k1, npl)][:4]
    # GPU
    act = np.zeros((n, n, n), dtype=bool)
    Tg = cp.full((n, n, n), TINF, dtype=cp.float64)
    grid = g.Grid3D(n, n, n, DX, act)
    mat, prm = g.Material(rho, cp_, k), g.Params(dt, THETA)
    hd = {f: cp.asarray(v) for f, v in h.items()}
    dact = cp.asarray(act)
    # CPU
    cart.set_threads(host_threads())
    Th = np.full((n, n, n), TINF)
    hact = act.copy()
    hm = cart.Material(rho, cp_, k)
    t_cpu = 0.0
    nsteps = 0
    for ks, ke in layers:
        born = np.zeros_like(act)
        born[:, :, ks:ke] = full[:, :, ks:ke] & ~hact[:, :, ks:ke]
        Tg[cp.asarray(born)] = 1000.0
        dact[cp.asarray(born)] = True
        grid.mask = dact
        packs = g.precompute_coeff_packs_unified(grid, mat, robin_h=hd)
        t0 = time.perf_counter()
        Th[born] = 1000.0
        hact |= born
        hg = cart.Grid3D(n, n, n, DX, hact)
        hp = cart.precompute_coeff_packs_unified(hg, hm, robin_h=h)
        t_cpu += time.perf_counter() - t0
        for _ in range(2):
            Tg = g.adi_step_gpu_coeff(Tg, grid, mat, prm, packs, Tinf=TINF)
            t0 = time.perf_counter()
            Th = cart.adi_step_numba_coeff(Th, hg, hm, cart.Params(dt, THETA), hp, Tinf=TINF)
            t_cpu += time.perf_counter() - t0
            nsteps += 1
    out = cp.asnumpy(Tg)
    g._engine.bound = None
    cpu = {"value": n ** 3 * nsteps / t_cpu, "unit": UNIT, "cores": host_threads(), "kind": "port",
           "sample": f"{n}^3 miniature of the deposition loop: {len(layers)} births (pack rebuild on the host) x 2 steps, "
                     f"C restatement of the reference with OpenMP (the reference itself is serial: about 1/8 of this per core)"}
    parity = {"rel_l2_vs_oracle": rel_l2(out, Th, hact), "void_bit_equal": bool(np.array_equal(out[~hact], Th[~hact])),
              "sample": f"{n}^3 miniature, {len(layers)} births x 2 steps, cfl 2000", "tol": 1e-12 * nsteps}
    return cpu, parity


# ------------------------------------------------------------------------------------------
# configs[2]: cylindrical 256 x 1024 x 512
# ------------------------------------------------------------------------------------------
def bench_cyl(args, local, sub=False):
    """BASELINE configs[2]: cylindrical r-phi-z grid 256 x 1024 x 512 (at --size 512), periodic phi,
    RobinR(500, 20), ZBC('neumann0', 'robin'), scheme 'be', dt = min(dr^2, dz^2, (R dphi)^2) / alpha
    (quick_compare_layer_birth_robin_cyl_v3.py:115,127); layer births grow nz 448 -> 512 by 16 planes on a
    buffer pre-pitched at 512 before the timed region; timed at nz = 512.  48 B/cell-step."""
    import torch
    from adi_thermal_fields_b200 import _capi, adi3d_cyl_phi_v3 as gc

    nr, nphi, nzf = args.size // 2, 2 * args.size, args.size
    dev = torch.device("cuda", local)
    c = cyl_setup(nr, nphi, nzf)
    R, dr, dz, dphi = c["R"], c["dr"], c["dz"], c["dphi"]
    mat = gc.Material(7800.0, 490.0, 54.0)
    dt = min(dr * dr, dz * dz, (R * dphi) ** 2) / mat.alpha
    rob, zbc = gc.RobinR(500.0, 20.0), gc.ZBC("neumann0", "robin", h_top=500.0, T_inf_top=20.0)
    prm = gc.Params(dt, 1.0, "be")
    gen = torch.Generator(device=dev).manual_seed(2)
    A = torch.full((nr, nphi, nzf), 20.0, dtype=torch.float64, device=dev)
    B = torch.empty_like(A)
    nz = nzf - 4 * (nzf // 32)
    A[:, :, :nz] += 5.0 * torch.rand((nr, nphi, nz), dtype=torch.float64, device=dev, generator=gen)
    while True:   # births: 16 planes at Ts on top, two steps each, in place on the pitched buffer
        A[:, :, nz - nzf // 32:nz] = 1000.0
        gg = gc.GridCyl(nr, nphi, nz, dr, dphi, dz, R)
        for _ in range(2):
            B.copy_(A)
            gc.adi_step_device(A, gg, mat, prm, rob, zbc, out=B, nz_pitch=nzf)
            A, B = B, A
        if nz >= nzf:
            break
        nz += nzf // 32
    grid = gc.GridCyl(nr, nphi, nzf, dr, dphi, dz, R)
    L, ctx = _capi.load(), gc._engine.context()
    for o in args.opt:
        name, _, val = o.partition("=")
        _capi.check(L.adi_set_option(ctx, name.encode(), int(val)), "adi_set_option")
    stream = torch.cuda.current_stream()
    for _ in range(args.warmup):
        gc.adi_step_device(A, grid, mat, prm, rob, zbc, out=B); A, B = B, A
    torch.cuda.synchronize()
    L.adi_set_option(ctx, b"profile", 1)
    L.adi_profile_reset(ctx)
    sampler = ClockSampler(local)
    sampler.start()
    l0 = gc.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        gc.adi_step_device(A, grid, mat, prm, rob, zbc, out=B); A, B = B, A
    e1.record(stream)
    torch.cuda.synchronize()
    ms_total = e0.elapsed_time(e1)
    launches = gc.launch_count() - l0
    clocks = sampler.stop()
    ms4 = (C.c_double * 4)()
    nst = C.c_long()
    L.adi_profile_read(ctx, ms4, C.byref(nst))
    L.adi_set_option(ctx, b"profile", 0)
    cells = nr * nphi * nzf
    ms_per_step = ms_total / args.steps
    # e2e: the reference's own calling convention -- host NumPy in, host NumPy out (adi_cyl_step_host),
    # each step's input the previous step's output
    host = A.cpu().numpy()
    host = gc.adi_step(host, grid, mat, prm, rob, zbc)
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        host = gc.adi_step(host, grid, mat, prm, rob, zbc)
    t_e2e = time.perf_counter() - t0
    del A, B
    peak, peak_src = peaks()
    per = [ms4[i] / max(1, nst.value) for i in range(1, 4)]
    names = ["k_cyl_strided<r>", "k_cyl_strided<phi>", "k_sweep_zt (cylindrical rows)"]
    dom = int(np.argmax(per))
    achieved = 16.0 * cells / (per[dom] * 1e-3) / 1e9
    cpu = parity = None
    if not args.no_cpu:
        # CPU legs at 128 x 512 x 256 + parity of the GPU step on those very inputs
        cs = cyl_setup(max(4, nr // 2), max(8, nphi // 2), max(8, nzf // 2))
        cpu, (T0s, outs, dts) = cyl_cpu_legs(cs)
        gs = gc.GridCyl(cs["nr"], cs["nphi"], cs["nz"], cs["dr"], cs["dphi"], cs["dz"], cs["R"])
        og = gc.adi_step(T0s, gs, mat, gc.Params(dts, 1.0, "be"), rob, zbc)
        parity = {"rel_l2_vs_oracle": rel_l2(og, outs), "sample": f"{cs['nr']}x{cs['nphi']}x{cs['nz']}, one step", "tol": 1e-12}
    line = {
        "metric": METRIC, "value": cells * args.steps / (ms_total * 1e-3), "unit": UNIT, "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"cylindrical r-phi-z {nr}x{nphi}x{nzf} (BASELINE configs[2]): periodic phi, RobinR(500,20), "
                               f"ZBC(neumann0, robin), scheme be, cfl 1; nz grown {nzf - 4 * (nzf // 32)}->{nzf} by births on a pitched "
                               f"buffer before timing", "grid": [nr, nphi, nzf], "cells": cells, "bytes_per_cell_step": 48,
                   "l2": "fields (1.07 GB) exceed the 126 MB L2; no flush needed", "parallelism": "single GPU"},
        "roofline": {"bound": "hbm", "kernel": names[dom], "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic_of(names[dom].split("<")[0]), "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": 16.0 * cells,
                     "kernel_ms": {"r": per[0], "phi": per[1], "z": per[2]},
                     "step_achieved_GBs": 48.0 * cells / (ms_per_step * 1e-3) / 1e9,
                     "step_frac": 48.0 * cells / (ms_per_step * 1e-3) / 1e9 / peak},
        "cpu_baseline": cpu, "clocks": clocks,
        "e2e": {"value": cells * e2e_steps / t_e2e, "unit": UNIT, "h2d_bytes_per_step": 8 * cells, "d2h_bytes_per_step": 8 * cells,
                "steps": e2e_steps, "api": "adi3d_cyl_phi_v3.adi_step (host NumPy in / out, the reference's calling convention; "
                                           "dependent steps)"},
        "gpu_launches": int(launches), "parity": parity,
    }
    if sub:
        for k_ in ("metric", "unit", "higher_is_better", "vs_baseline", "dtype", "data", "scaling", "n_gpus"):
            line.pop(k_, None)
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--workload", default="plate", choices=["plate", "c5", "c4", "cyl"],
                    help="plate: BASELINE configs[1] + sub-records of the other configs (N>1: one size^3 slab per GPU, weak "
                         "scaling, + c5_strong + parity); c5: BASELINE configs[4] alone, z-slab over N GPUs; "
                         "cyl: BASELINE configs[2] alone; c4: BASELINE configs[3] alone")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--long-steps", type=int, default=300, help="steps of the extra stability run of the plate arm")
    ap.add_argument("--c5-steps", type=int, default=5, help="timed steps of the configs[4] sub-records")
    ap.add_argument("--c4-mini", type=int, default=96, help="edge of the configs[3] parity / CPU miniature")
    ap.add_argument("--cpu-ny", type=int, default=128, help="y extent of the CPU-baseline sample slab")
    ap.add_argument("--numba-ny", type=int, default=16, help="y extent of the Numba reference sample inside cpu_baseline")
    ap.add_argument("--ref-ny", type=int, default=0, help="y extent of the --impl reference sample slab (0: from the step count)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="headline workload only (no c3 / c4 / c5 sub-records)")
    ap.add_argument("--c4-steps-per-layer", type=int, default=4)
    ap.add_argument("--opt", action="append", default=[], metavar="NAME=VALUE",
                    help="engine tuning option (adi_set_option), e.g. --opt m=16 --opt kt=8")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
