#!/usr/bin/env python
# -*- coding: utf-8 -*-
"""bench.py -- ADI cell-steps/s (fp64) of the B200 engine on BASELINE.json's workload.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload ...]

N=1 workload: BASELINE.json configs[1], `single_track_on_plate` at 512^3 -- Cartesian plate
+ track mask, Robin on six faces with per-face variable h (dense coefficient field per axis:
75 algorithmic bytes per cell-step), theta=0.5, dt=0.02 s.
A "step" is one full ADI time step (explicit stage + x, y, z implicit sweeps) of the whole grid.

N>1 workload (one process per GPU, launched by torchrun): the same plate at 512 x 512 x (512*N),
z-slab partitioned, one 512^3 slab per GPU ("scaling": "weak").  x and y sweeps are rank-local;
per step the ranks exchange one T plane per side (explicit stage) and all-gather the interface
relations of the partitioned z sweep over NCCL (adi_thermal_fields_b200/slab.py).

Prints ONE JSON line (rank 0).  `value` is device-resident throughput (inputs in HBM,
CUDA events); `e2e` is the same metric through the host-array C-ABI call
(adi_cart_step_host: H2D + step + D2H inside the timed region).  `roofline` is for the
slowest of the three sweep kernels, timed live with CUDA events recorded inside the engine on
the launching stream.  `cpu_baseline` is the oracle (C restatement of the reference's Numba
path, oracle/adi_oracle.c) timed on this box's host cores on a bounded sample; the same run
also checks GPU-vs-oracle parity on that sample.

--impl reference times the CPU restatement alone (the reference is pure Python + Numba and
does not travel to the GPU box; the oracle is bit-identical to it on the golden vectors).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "ADI cell-steps/sec (fp64)"
UNIT = "cell-steps/s"
RHO, CP, K = 7800.0, 500.0, 25.0
DX, DT, THETA, TINF = 1.0e-3, 0.02, 0.5, 20.0
FACES = ("x-", "x+", "y-", "y+", "z-", "z+")


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


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
                                       "-lms", "100", "-i", str(self.idx)],
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
# CPU legs (oracle) -- the only place bench.py touches oracle/
# ------------------------------------------------------------------------------------------
def host_threads():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_leg(n, ny_sample, threads, steps=1, check_against=None):
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


def run_reference(args):
    """--impl reference: the CPU restatement of the reference path, all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import cart
    n = args.size
    # all host cores this process may use (torchrun exports OMP_NUM_THREADS=1: not a property of the box)
    threads = host_threads()
    # bounded sample: a y-slab of the workload (x and z lines keep their full length)
    ny_s = min(n, args.ref_ny)
    cells = n * ny_s * n
    _, times, _ = cpu_leg(n, ny_s, threads, steps=args.warmup + args.steps)
    tt = times[args.warmup:]
    dt_step = sum(tt) / len(tt)
    value = cells / dt_step
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt_step * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": workload_config(args, 1) | {"sample": f"{n}x{ny_s}x{n} y-slab of the workload per step"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{n}x{ny_s}x{n} y-slab, {args.steps} steps, OpenMP over lines "
                                   f"(the reference itself is serial Numba)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def workload_config(args, world):
    n = args.size
    return {"workload": f"single_track_on_plate {n}^3 (BASELINE configs[1]): plate+track mask, Robin x6, "
                        f"per-face variable h (dense coeff per axis), theta={THETA}, dt={DT}",
            "grid": [n, n, n], "cells": n ** 3, "bytes_per_cell_step": 75,
            "l2": "fields (1.07 GB each at 512^3) exceed the 126 MB L2; no flush needed",
            "parallelism": "single GPU" if world == 1 else
            f"z-slab x{world}: grid {n}x{n}x{n * world}, one {n}^3 slab per GPU; per step 2 T-plane halo "
            f"messages per interior boundary + all-gather of 6 doubles per z line and rank (NCCL)"}


# ------------------------------------------------------------------------------------------
# GPU arm
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

    from adi_thermal_fields_b200 import _capi, adi3d_gpu_coeff as g, devarray as cp

    if args.workload == "c4":
        return run_ours_c4(args, local)
    if args.workload == "cyl":
        return run_ours_cyl(args, local)
    if world > 1 or args.workload == "c5":
        return run_ours_slab(args, rank, world, local)
    n = args.size
    dev = torch.device("cuda", local)
    # ---- synthetic inputs, created on the device (outside any timed region) ----
    mask = torch.zeros((n, n, n), dtype=torch.bool, device=dev)
    nzp = n - max(1, n // 64)
    mask[:, :, :nzp] = True
    mask[: max(1, n // 32), : n // 2, nzp:] = True
    gen = torch.Generator(device=dev).manual_seed(rank)
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

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- timed region: K steps, device-resident ----
    L.adi_set_option(ctx, b"profile", 1)
    L.adi_profile_reset(ctx)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = L.adi_launch_count(ctx)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        step(A, B); A, B = B, A
    e1.record(stream)
    barrier()
    ms_total = e0.elapsed_time(e1)
    launches = L.adi_launch_count(ctx) - launches0
    clocks = sampler.stop() if rank == 0 else None
    ms3 = (C.c_double * 4)()
    nst = C.c_long()
    L.adi_profile_read(ctx, ms3, C.byref(nst))
    L.adi_set_option(ctx, b"profile", 0)
    if world > 1:
        t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_per_step = ms_total / args.steps
    cells = n ** 3
    value = world * cells * args.steps / (ms_total * 1e-3)

    # ---- e2e: host arrays through the C ABI (H2D + step + D2H per step) ----
    # Every step uploads its input from pinned host memory and downloads its result.  The
    # steps are independent host fields, so two of them are kept in flight on two streams
    # (adi_cart_step_host_async, two staging slots): the upload of one overlaps the compute of
    # the other and the download of the previous result.  The serial form (one blocking
    # adi_cart_step_host call per step) is timed as well and reported next to it.
    e2e_steps = max(2, min(args.steps, args.e2e_steps))
    hin = [torch.empty((n, n, n), dtype=torch.float64, pin_memory=True) for _ in range(2)]
    hout = [torch.empty((n, n, n), dtype=torch.float64, pin_memory=True) for _ in range(2)]
    for hbuf in hin:
        hbuf.copy_(T0)
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    torch.cuda.synchronize()

    def host_step():
        _capi.check(L.adi_cart_step_host(ctx, hin[0].data_ptr(), hout[0].data_ptr(), 1, DT, THETA, kappa, TINF,
                                         stream.cuda_stream), "adi_cart_step_host")

    def host_step_async(i):
        _capi.check(L.adi_cart_step_host_async(ctx, i & 1, hin[i & 1].data_ptr(), hout[i & 1].data_ptr(), DT, THETA,
                                               kappa, TINF, streams[i & 1].cuda_stream), "adi_cart_step_host_async")

    host_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        host_step()
    torch.cuda.synchronize()
    t_serial = time.perf_counter() - t0
    for i in range(2):
        host_step_async(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        host_step_async(i)
    torch.cuda.synchronize()
    t_e2e = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([t_e2e], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_e2e = float(t.item())
    e2e_value = world * cells * e2e_steps / t_e2e
    e2e_serial = world * cells * e2e_steps / t_serial
    e2e_ok = bool(torch.equal(hout[0], hout[1]))   # both slots stepped the same input

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the slowest sweep kernel (live CUDA events inside the engine) ----
    peak, peak_src = peaks()
    per = [ms3[i] / max(1, nst.value) for i in range(4)]
    names = ["k_explicit", "k_sweep_strided<x>", "k_sweep_strided<y>", "k_sweep_z"]
    # algorithmic bytes per cell: explicit stage T in 8 + code 1 + out 8; a sweep
    # in 8 + out 8 + code 1 + dense coeff 8.  The 75 B/cell-step of the metric counts the fused
    # form (3 sweeps); the separate explicit pass is extra real traffic, not extra credit.
    bpc = [17.0, 25.0, 25.0, 25.0]
    # what the kernels really fetch: a sweep whose coefficient field was verified surface-only (option
    # sparse_coeff, x / y sweeps) skips the 8 B/cell of interior coefficient reads
    sparse = int(L.adi_get_option(ctx, b"sparse_active"))
    moved = [17.0] + [17.0 if (sparse >> a) & 1 else 25.0 for a in range(3)]
    dom = int(np.argmax(per))
    bytes_per_launch = bpc[dom] * cells
    achieved = bytes_per_launch / (per[dom] * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": names[dom], "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": bytes_per_launch,
                "kernel_ms": {"explicit": per[0], "x": per[1], "y": per[2], "z": per[3]},
                "kernel_GBs": {k: b * cells / (t * 1e-3) / 1e9 if t > 0 else None
                               for k, b, t in zip(("explicit", "x", "y", "z"), bpc, per)},
                "step_achieved_GBs": 75.0 * cells / (ms_per_step * 1e-3) / 1e9,
                "step_frac": 75.0 * cells / (ms_per_step * 1e-3) / 1e9 / peak,
                "moved_bytes_per_cell_step": sum(moved),
                "step_frac_moved": sum(moved) * cells / (ms_per_step * 1e-3) / 1e9 / peak,
                "note": "achieved/frac use SURVEY 8(d)'s algorithmic bytes (75 B/cell-step = 3 sweeps x 25 B, explicit "
                        "stage counted as fused).  The kernels move moved_bytes_per_cell_step: +17 B for the separate "
                        "explicit pass, -8 B for each sweep that reads its verified surface-only coefficient field at "
                        "exposed cells only (sparse_active bitmask %d)" % sparse}
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            roofline["traffic"] = json.load(open(tp)).get(names[dom].split("<")[0], {}).get("dram_bytes_per_launch")
        except Exception:
            pass

    # ---- CPU baseline + parity on a bounded sample (rank 0, N=1 only) ----
    cpu = None
    parity = None
    if world == 1 and not args.no_cpu:
        from oracle import cart
        nth = host_threads()
        ny_s = min(n, args.cpu_ny)
        Tref, times, (m_s, T0_s, h_s) = cpu_leg(n, ny_s, nth, steps=1)
        v_all = n * ny_s * n / times[0]
        ny_1 = max(8, ny_s // 8)
        _, t1, _ = cpu_leg(n, ny_1, 1, steps=1)
        v_one = n * ny_1 * n / t1[0]
        cpu = {"value": v_all, "unit": UNIT, "cores": nth, "kind": "port",
               "sample": f"one step of a {n}x{ny_s}x{n} y-slab of the workload, OpenMP over lines",
               "value_1core": v_one,
               "note": "the reference's Numba path is serial (1 core); value_1core is the like-for-like figure"}
        # the same sample through the GPU engine
        gs = g.Grid3D(n, ny_s, n, DX, m_s)
        ps = g.precompute_coeff_packs_unified(gs, mat, robin_h=h_s)
        out = g.adi_step_gpu_coeff(cp.asarray(T0_s), gs, mat, prm, ps, Tinf=TINF)
        Tg = cp.asnumpy(out)
        num = float(np.sqrt(np.sum((Tg[m_s] - Tref[m_s]) ** 2)))
        den = float(np.sqrt(np.sum(Tref[m_s] ** 2)))
        parity = {"rel_l2_vs_oracle": num / den, "void_bit_equal": bool(np.array_equal(Tg[~m_s], T0_s[~m_s])),
                  "sample": f"{n}x{ny_s}x{n}", "tol": 1e-12}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(args, world),
        "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 8 * cells, "d2h_bytes_per_step": 8 * cells,
                "steps": e2e_steps, "api": "adi_cart_step_host_async, two staging slots on two streams (pinned host arrays in/out)",
                "serial_value": e2e_serial, "serial_api": "adi_cart_step_host (one blocking call per step)",
                "slots_agree": e2e_ok},
        "gpu_launches": int(launches), "parity": parity,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def run_ours_slab(args, rank, world, local):
    """N>1: z-slab partitioned plate, one n^3 slab per rank (weak scaling), NCCL exchanges."""
    import torch
    import torch.distributed as dist
    from adi_thermal_fields_b200 import slab

    dev = torch.device("cuda", local)
    comm = slab.TorchDistComm() if world > 1 else slab.LocalComm(1).view(0)
    c5 = args.workload == "c5"

    class Mat:
        rho, cp, k = RHO, CP, K

    class Prm:
        dt, theta = DT, THETA
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
        packs = slab.precompute_coeff_packs_unified(grid, Mat, robin_h={f: 10.0 for f in FACES})
        bpc = 51.0
        wl = {"workload": f"synthetic Cartesian {nx}x{ny}x{nzg} (BASELINE configs[4]): full mask, scalar Robin h=10 x6 "
                          f"(on-the-fly coefficients), theta={THETA}, dt={DT}; strong scaling",
              "grid": [nx, ny, nzg], "cells": nx * ny * nzg, "bytes_per_cell_step": 51,
              "l2": "fields (34 GB) exceed the 126 MB L2; no flush needed",
              "parallelism": f"z-slab x{world}: {nzl} planes per GPU; per step T-plane halos + all-gather of 6 doubles "
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
        packs = slab.precompute_coeff_packs_unified(grid, Mat, robin_h=h)
        del h
        bpc = 75.0
        wl = workload_config(args, world)
        scaling = "weak"
    cells = nx * ny * nzl          # per rank
    for o in args.opt:
        name, _, val = o.partition("=")
        grid.be.set_option(name, int(val))
    torch.cuda.synchronize()
    A, B = T0.clone(), torch.empty_like(T0)
    stream = torch.cuda.current_stream()

    def step(src, dst):
        slab.adi_step_gpu_coeff(src, grid, Mat, Prm, packs, Tinf=TINF, out=dst)

    for _ in range(args.warmup):
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
    for _ in range(args.steps):
        step(A, B); A, B = B, A
    e1.record(stream)
    barrier()
    ms_total = e0.elapsed_time(e1)
    launches = grid.be.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    ms3, nst = grid.be.profile_read()
    grid.be.profile(False)
    ms_total = allmax(ms_total)
    ms_per_step = ms_total / args.steps
    value = world * cells * args.steps / (ms_total * 1e-3)

    # e2e: pinned host slabs in and out every step (H2D + step + D2H inside the timed region)
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    if 8 * cells > 12 * 2 ** 30:   # two pinned slabs of > 12 GB each: not attempted
        e2e_steps = 0
    hin = torch.empty((nx, ny, nzl) if e2e_steps else (1,), dtype=torch.float64, pin_memory=True)
    hout = torch.empty((nx, ny, nzl) if e2e_steps else (1,), dtype=torch.float64, pin_memory=True)
    if e2e_steps:
        hin.copy_(T0)
    torch.cuda.synchronize()

    def host_step():
        A.copy_(hin, non_blocking=True)
        step(A, B)
        hout.copy_(B, non_blocking=True)
        torch.cuda.synchronize()

    if e2e_steps:
        host_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        host_step()
    t_e2e = time.perf_counter() - t0
    t_e2e = allmax(t_e2e)
    e2e_value = world * cells * e2e_steps / t_e2e if e2e_steps else None

    if rank == 0:
        peak, peak_src = peaks()
        per = [ms3[i] / max(1, nst) for i in range(4)]
        solve_first = bool(getattr(grid, "_spikes", None))
        names = ["k_explicit", "k_sweep_strided<x>", "k_sweep_strided<y>",
                 "k_sweep_z solve-first + all-gather + k_spike_apply" if solve_first
                 else "k_sweep_z pass1 + all-gather + pass2"]
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
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": scaling,
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": wl,
            "roofline": roofline, "cpu_baseline": None, "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 8 * cells * world,
                    "d2h_bytes_per_step": 8 * cells * world, "steps": e2e_steps,
                    "api": "slab.adi_step_gpu_coeff on pinned host slabs (H2D + step + D2H per rank)"},
            "gpu_launches": int(launches), "parity": None,
            "exchange": {"halo_bytes_per_step_per_boundary": 2 * 8 * nx * ny,
                         "allgather_bytes_per_rank_per_step": 2 * 8 * nx * ny,
                         "allgather_bytes_per_rank_on_matrix_change": 4 * 8 * nx * ny},
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def run_ours_c4(args, local):
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
    ax = (torch.arange(n, device=dev, dtype=torch.float64) + 0.5) / n - 0.5
    X, Y, Z = ax[:, None, None], ax[None, :, None], ax[None, None, :]
    full = ((X / 0.35) ** 2 + (Y / 0.42) ** 2 + ((Z - 0.05) / 0.45) ** 2 <= 1.0) | \
           ((X * X + Y * Y <= 0.12 ** 2) & (Z < -0.3))
    del X, Y, Z
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

    def run(layer_list):
        nonlocal T
        nsteps = 0
        tb = ts = 0.0
        for ks, ke in layer_list:
            e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            e0.record()
            packs = birth(ks, ke)
            e1.record()
            for _ in range(steps_per_layer):
                T = g.adi_step_gpu_coeff(T, grid, mat, prm, packs, Tinf=TINF)
                nsteps += 1
            e2.record()
            torch.cuda.synchronize()
            tb += e0.elapsed_time(e1)
            ts += e1.elapsed_time(e2)
        return nsteps, tb, ts

    nwarm = max(1, args.warmup // steps_per_layer)
    run(layers[:nwarm])
    nlay = max(1, args.steps // steps_per_layer)
    mid = layers[len(layers) // 2: len(layers) // 2 + nlay]   # mid-build: half of the head is active
    for ks, ke in layers[nwarm:len(layers) // 2]:             # fast-forward the activation (untimed)
        born = full[:, :, ks:ke]
        T._t[:, :, ks:ke][born] = 1000.0
        act[:, :, ks:ke] |= born
    sampler = ClockSampler(local)
    sampler.start()
    l0 = g.launch_count()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    nsteps, tb, ts = run(mid)
    wall = time.perf_counter() - t0
    launches = g.launch_count() - l0
    clocks = sampler.stop()
    cells = n ** 3
    peak, peak_src = peaks()
    total_ms = tb + ts
    line = {
        "metric": METRIC, "value": cells * nsteps / (total_ms * 1e-3), "unit": UNIT, "n_gpus": 1, "steps": nsteps,
        "warmup": nwarm * steps_per_layer, "ms_per_step": total_ms / nsteps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"waam_from_stl_v7_mm {n}^3 (BASELINE configs[3]), synthetic head (ellipsoid + neck; the STL is "
                               f"not in the tree), {n_per_layer} z planes per birth, {steps_per_layer} steps per layer, "
                               f"per-face dense h fields, theta={THETA}, cfl=2000; births + device pack rebuilds inside the timed region",
                   "grid": [n, n, n], "cells": cells, "bytes_per_cell_step": 75,
                   "active_fraction_mid_build": float(act.sum().item()) / cells, "parallelism": "single GPU"},
        "roofline": {"bound": "hbm", "kernel": "whole step (explicit + x + y + z), steady state between births",
                     "achieved": 75.0 * cells / (ts / nsteps * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                     "frac": 75.0 * cells / (ts / nsteps * 1e-3) / 1e9 / peak, "traffic": None, "peak_source": peak_src,
                     "steady_ms_per_step": ts / nsteps, "birth_ms": tb / len(mid),
                     "birth_note": "mask update + k_build_packs (6 dense h fields -> 3 coeff fields) + neighbour code rebuild",
                     "note": "achieved/frac restate the metric (all cells of the box, void included, at SURVEY 8(d)'s "
                             "75 B/cell-step) in GB/s; they are not DRAM utilisation: sweep tiles without an active "
                             "cell are skipped and coefficient fields are read at exposed cells only, so the bytes "
                             "moved per step are well below 75 B x cells while the part is being built"},
        "cpu_baseline": None, "clocks": clocks,
        "e2e": {"value": cells * nsteps / wall, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                "api": "adi3d_gpu_coeff.precompute_coeff_packs_unified + adi_step_gpu_coeff (device arrays, host wall clock)"},
        "gpu_launches": int(launches), "parity": None,
    }
    print(json.dumps(line), flush=True)
    return 0


def run_ours_cyl(args, local):
    """BASELINE configs[2]: cylindrical r-phi-z grid 256 x 1024 x 512 (at --size 512), periodic phi,
    RobinR(500, 20), ZBC('neumann0', 'robin'), scheme 'be', dt = min(dr^2, dz^2, (R dphi)^2) / alpha
    (quick_compare_layer_birth_robin_cyl_v3.py:115,127); layer births grow nz 448 -> 512 by 16 planes on a
    buffer pre-pitched at 512 before the timed region; timed at nz = 512.  48 B/cell-step."""
    import math
    import torch
    from adi_thermal_fields_b200 import _capi, adi3d_cyl_phi_v3 as gc

    nr, nphi, nzf = args.size // 2, 2 * args.size, args.size
    dev = torch.device("cuda", local)
    R = 0.02
    dr = R / nr
    dz, dphi = dr, 2.0 * math.pi / nphi
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
        g = gc.GridCyl(nr, nphi, nz, dr, dphi, dz, R)
        for _ in range(2):
            B.copy_(A)
            gc.adi_step_device(A, g, mat, prm, rob, zbc, out=B, nz_pitch=nzf)
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
    # e2e: the reference's own calling convention -- host NumPy in, host NumPy out (adi_cyl_step_host)
    host = A.cpu().numpy()
    gc.adi_step(host, grid, mat, prm, rob, zbc)
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        out = gc.adi_step(host, grid, mat, prm, rob, zbc)
    t_e2e = time.perf_counter() - t0
    peak, peak_src = peaks()
    per = [ms4[i] / max(1, nst.value) for i in range(1, 4)]
    names = ["k_cyl_strided<r>", "k_cyl_strided<phi>", "k_cyl_z"]
    dom = int(np.argmax(per))
    achieved = 16.0 * cells / (per[dom] * 1e-3) / 1e9
    # parity of this very state against the oracle on a sub-block is not possible (the solve couples the whole
    # grid); the golden / oracle parity of the cylindrical path is in tests/test_gpu_cyl.py
    line = {
        "metric": METRIC, "value": cells * args.steps / (ms_total * 1e-3), "unit": UNIT, "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"cylindrical r-phi-z {nr}x{nphi}x{nzf} (BASELINE configs[2]): periodic phi, RobinR(500,20), "
                               f"ZBC(neumann0, robin), scheme be, cfl 1; nz grown {nzf - 4 * (nzf // 32)}->{nzf} by births on a pitched "
                               f"buffer before timing", "grid": [nr, nphi, nzf], "cells": cells, "bytes_per_cell_step": 48,
                   "l2": "fields (1.07 GB) exceed the 126 MB L2; no flush needed", "parallelism": "single GPU"},
        "roofline": {"bound": "hbm", "kernel": names[dom], "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": 16.0 * cells,
                     "kernel_ms": {"r": per[0], "phi": per[1], "z": per[2]},
                     "step_achieved_GBs": 48.0 * cells / (ms_per_step * 1e-3) / 1e9,
                     "step_frac": 48.0 * cells / (ms_per_step * 1e-3) / 1e9 / peak},
        "cpu_baseline": None, "clocks": clocks,
        "e2e": {"value": cells * e2e_steps / t_e2e, "unit": UNIT, "h2d_bytes_per_step": 8 * cells, "d2h_bytes_per_step": 8 * cells,
                "steps": e2e_steps, "api": "adi3d_cyl_phi_v3.adi_step (host NumPy in / out, the reference's calling convention)"},
        "gpu_launches": int(launches), "parity": None,
    }
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            line["roofline"]["traffic"] = json.load(open(tp)).get(names[dom].split("<")[0], {}).get("dram_bytes_per_launch")
        except Exception:
            pass
    print(json.dumps(line), flush=True)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--workload", default="plate", choices=["plate", "c5", "c4", "cyl"],
                    help="plate: BASELINE configs[1] (N>1: one size^3 slab per GPU, weak scaling); "
                         "c5: BASELINE configs[4], 2048x2048x1024 scalar-Robin strong scaling, z-slab over N GPUs; "
                         "cyl: BASELINE configs[2], cylindrical 256x1024x512 backward-Euler step (1 GPU); "
                         "c4: BASELINE configs[3], waam 1024^3 synthetic head, layer births with device pack rebuilds (1 GPU)")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--cpu-ny", type=int, default=128, help="y extent of the CPU-baseline sample slab")
    ap.add_argument("--ref-ny", type=int, default=64, help="y extent of the --impl reference sample slab")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--c4-steps-per-layer", type=int, default=4)
    ap.add_argument("--opt", action="append", default=[], metavar="NAME=VALUE",
                    help="engine tuning option (adi_set_option), e.g. --opt m=32 --opt kt=16")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
