#!/usr/bin/env python
# -*- coding: utf-8 -*-
"""Generate tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference) on the seeded inputs of tests/cases.py.

Runs only in the build container (the reference tree does not travel to the GPU
box); the vectors it writes are committed.  Nothing here is imported by the
product.  Usage:  python tools/gen_golden.py [--ref /root/reference]

What is pinned:
  cart_<case>.npz   adi3d_numba_coeff.adi_step_numba_coeff  (adi3d_numba_coeff.py:290)
                    after `nsteps` steps (+ per-step probe trace for multi-step cases)
  cartgpu_<case>.npz the reference CuPy algorithm adi3d_gpu_coeff.adi_step_gpu_coeff
                    (adi3d_gpu_coeff.py:213) executed on NumPy (sys.modules['cupy']=numpy)
                    for a few cases -- a second, independent pin
  packs_<case>.npz  precompute_coeff_packs_unified outputs (adi3d_numba_coeff.py:57)
  cyl_<case>.npz    adi3d_cyl_phi_v3.adi_step scheme "be" (adi3d_cyl_phi_v3.py:332),
                    or quick_spiral_deposition_gif_v5.adi_step_masked (:31) when masked
  spiral_sim.npz    tests/test_spiral_vs_analytic.py:_run_numeric_simulation snapshots,
                    with GridCyl accepting (and ignoring) R_in -- see SURVEY.md F2
  vtk_text.npz      the files written by vtk_writer.write_vtk_structured_points (vtk_writer.py:12) and
                    waam_from_stl_v7_mm.write_vtk_structured_points (:186) for small seeded fields
  cyl_birth.npz     the nz-growth event loop of quick_compare_layer_birth_robin_cyl_v3.py:171-204
"""
from __future__ import annotations

import argparse
import math
import os
import sys
import tempfile
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(ROOT, "tests"))

os.environ.setdefault("NUMBA_CACHE_DIR", tempfile.mkdtemp(prefix="numba_cache_"))
os.environ["PYTHONDONTWRITEBYTECODE"] = "1"
sys.dont_write_bytecode = True

import numpy as np  # noqa: E402

import cases  # noqa: E402


def _stub_matplotlib():
    """quick_spiral_deposition_gif_v5.py:24-26 imports matplotlib at module top; it is absent."""
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.animation", "matplotlib.colors"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            sys.modules[name] = m
    sys.modules["matplotlib.colors"].LogNorm = object
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["matplotlib"].animation = sys.modules["matplotlib.animation"]
    sys.modules["matplotlib"].colors = sys.modules["matplotlib.colors"]


def gen_cart(ref, outdir):
    import adi3d_numba_coeff as adi
    for name in cases.CART_CASES:
        c = cases.build_cart_case(name)
        nx, ny, nz = c["shape"]
        grid = adi.Grid3D(nx, ny, nz, c["dx"], c["mask"])
        mat = adi.Material(c["rho"], c["cp"], c["k"])
        prm = adi.Params(c["dt"], c["theta"])
        packs = adi.precompute_coeff_packs_unified(grid, mat, robin_Tinf=c["Tinf"], **c["bcs"])
        T = c["T0"].copy()
        trace = []
        for _ in range(c["nsteps"]):
            T = adi.adi_step_numba_coeff(T, grid, mat, prm, packs, Tinf=c["Tinf"])
            trace.append(T[nx // 2, ny // 2, :].copy())
        np.savez_compressed(os.path.join(outdir, f"cart_{name}.npz"), T_out=T,
                            trace=np.array(trace))
        if name in ("holes_combined", "track_mixed", "random_neumann_fields"):
            np.savez_compressed(
                os.path.join(outdir, f"packs_{name}.npz"),
                coeff_x=packs[0].coeff, coeff_y=packs[1].coeff, coeff_z=packs[2].coeff,
                q_x=packs[0].qflux, q_y=packs[1].qflux, q_z=packs[2].qflux,
                dir_mask=packs[0].dir_mask, dir_val=packs[0].dir_val)
        print(f"[cart] {name}: T in [{np.nanmin(T) if T.size else 0:.6g}, "
              f"{np.nanmax(T) if T.size else 0:.6g}]")


def gen_cart_gpu_algo(ref, outdir):
    """adi3d_gpu_coeff.py executed with cupy:=numpy."""
    saved = sys.modules.get("cupy")
    sys.modules["cupy"] = np
    try:
        import importlib
        g = importlib.import_module("adi3d_gpu_coeff")
        for name in ("cyl_robin6", "cyl_dirtop", "B_track_dict3d", "full_dict3d_cfl3000"):
            c = cases.build_cart_case(name)
            nx, ny, nz = c["shape"]
            grid = g.Grid3D(nx, ny, nz, c["dx"], c["mask"])
            mat = g.Material(c["rho"], c["cp"], c["k"])
            prm = g.Params(c["dt"], c["theta"])
            packs = g.precompute_coeff_packs_unified(grid, mat, robin_Tinf=c["Tinf"], **c["bcs"])
            T = c["T0"].copy()
            for _ in range(c["nsteps"]):
                T = g.adi_step_gpu_coeff(T, grid, mat, prm, packs, Tinf=c["Tinf"])
            np.savez_compressed(os.path.join(outdir, f"cartgpu_{name}.npz"), T_out=T)
            print(f"[cartgpu] {name}")
    finally:
        if saved is None:
            sys.modules.pop("cupy", None)
        else:
            sys.modules["cupy"] = saved
        sys.modules.pop("adi3d_gpu_coeff", None)


def gen_cyl(ref, outdir):
    import adi3d_cyl_phi_v3 as cyl
    _stub_matplotlib()
    import quick_spiral_deposition_gif_v5 as spiral
    for name in cases.ALL_CYL_CASES + ["c3_slice"]:
        c = cases.build_cyl_case(name)
        grid = cyl.GridCyl(c["nr"], c["nphi"], c["nz"], c["dr"], c["dphi"], c["dz"], c["R"])
        mat = cyl.Material(c["rho"], c["cp"], c["k"])
        prm = cyl.Params(c["dt"], 1.0, "be")
        rob = cyl.RobinR(c["h_r"], c["Tinf_r"])
        zbc = cyl.ZBC(**c["zbc"])
        if c["active"] is not None:
            T = spiral.adi_step_masked(c["T0"], grid, mat, prm, rob, zbc, c["active"],
                                       robin_inner=cyl.RobinR(c["h_r"], c["T_inner"]),
                                       robin_void=cyl.RobinR(c["h_r"], c["T_void"]))
        else:
            T = cyl.adi_step(c["T0"], grid, mat, prm, rob, zbc, S=c["S"])
        if name == "c3_slice":
            # 33 MB as it stands: every 32nd phi row + the SHA-256 of the whole array pin it bit for bit
            import hashlib
            np.savez_compressed(os.path.join(outdir, f"cyl_{name}.npz"), T_sub=T[:, ::32, :].copy(),
                                sha256=np.frombuffer(hashlib.sha256(np.ascontiguousarray(T).tobytes()).digest(), dtype=np.uint8))
        else:
            np.savez_compressed(os.path.join(outdir, f"cyl_{name}.npz"), T_out=T)
        print(f"[cyl] {name}: T in [{T.min():.6g}, {T.max():.6g}]")


def gen_spiral_sim(ref, outdir):
    """Run tests/test_spiral_vs_analytic.py:_run_numeric_simulation with a GridCyl that
    accepts R_in (ignored; r stays (i+1/2)dr as in adi3d_cyl_phi_v3.py:39)."""
    import adi3d_cyl_phi_v3 as cyl
    _stub_matplotlib()
    import quick_spiral_deposition_gif_v5 as spiral

    class GridCylRin(cyl.GridCyl):
        def __init__(self, nr, nphi, nz, dr, dphi, dz, R, R_in=0.0):
            super().__init__(nr, nphi, nz, dr, dphi, dz, R)
            self.R_in = float(R_in)

    spiral.GridCyl = GridCylRin
    sys.path.insert(0, os.path.join(ref, "tests"))
    import test_spiral_vs_analytic as t
    # configuration of tests/test_spiral_vs_analytic.py:124-162
    R_in, wall = 0.03, 0.002
    nphi, tau_dep, n_layers, layer_h = 36, 2.0, 2, 0.004
    cfg = {"R_out": R_in + wall, "wall_thickness": wall, "height": layer_h * n_layers,
           "z_back": 0.02, "nr": 6, "nphi": nphi, "dz_override": layer_h,
           "rho": 7800.0, "cp": 490.0, "k": 54.0, "h_side": 400.0, "h_end": 500.0,
           "T_inf": 20.0, "T_deposit": 900.0, "h_void": 400.0, "layer_cells": 1,
           "n_layers": n_layers, "loops_per_layer": 1, "dt": tau_dep / nphi,
           "omega": 2.0 * math.pi / tau_dep}
    times = np.linspace(0.0, tau_dep * n_layers, 5)
    grid, snaps, act = t._run_numeric_simulation(times, cfg)
    np.savez_compressed(os.path.join(outdir, "spiral_sim.npz"), times=times,
                        snapshots=np.array(snaps), active=np.array(act),
                        shape=np.array([grid.nr, grid.nphi, grid.nz]),
                        dr=grid.dr, dphi=grid.dphi, dz=grid.dz)
    print(f"[spiral] grid {grid.nr}x{grid.nphi}x{grid.nz}, {len(snaps)} snapshots")


def gen_cyl_birth(ref, outdir):
    """Event loop of quick_compare_layer_birth_robin_cyl_v3.py:171-204 at nr=8,nphi=16."""
    import adi3d_cyl_phi_v3 as cyl
    R, z_back, d, t_step, N_total, t_tail = 0.02, 0.02, 0.005, 0.5, 3, 0.5
    nr, nphi = 8, 16
    k_, rho, cp = 54.0, 7800.0, 490.0
    h_side, h_end, T_inf, Ts, cfl = 500.0, 500.0, 20.0, 1000.0, 1.0
    dr = R / nr
    dz = dr
    dphi = (2.0 * np.pi) / max(nphi, 1)
    alpha = k_ / (rho * cp)
    dt0 = cfl * min(dr * dr, dz * dz, (R * dphi) ** 2) / alpha
    nz0 = int(round((z_back + d) / dz))
    grid = cyl.GridCyl(nr, nphi, nz0, dr, dphi, dz, R)
    mat = cyl.Material(rho, cp, k_)
    rob = cyl.RobinR(h_side, T_inf)
    zbc = cyl.ZBC("neumann0", "robin", h_top=h_end, T_inf_top=T_inf)
    T = np.full((nr, nphi, nz0), T_inf, dtype=float)
    nz_extra = int(round(d / dz))
    T[:, :, -nz_extra:] = Ts
    nz_final = int(round((z_back + N_total * d) / dz))
    times = np.linspace(0.0, (N_total - 1) * t_step + t_tail, 7)
    t = 0.0
    next_birth = t_step
    eps = 1e-12
    frames = []
    for t_target in times[1:]:
        while t < t_target - eps:
            dt_step = min(dt0, t_target - t, max(eps, next_birth - t))
            T = cyl.adi_step(T, grid, mat, cyl.Params(dt_step, 1.0, "be"), rob, zbc, S=None)
            t += dt_step
            if abs(t - next_birth) <= eps:
                if grid.nz + nz_extra <= nz_final:
                    old = T
                    nz_new = grid.nz + nz_extra
                    T = np.full((nr, nphi, nz_new), T_inf, float)
                    T[:, :, : old.shape[2]] = old
                    T[:, :, -nz_extra:] = Ts
                    grid = cyl.GridCyl(nr, nphi, nz_new, dr, dphi, dz, R)
                next_birth += t_step
        t = t_target
        y = np.full(nz_final, np.nan)
        y[: grid.nz] = T[0, 0, :]
        frames.append(y)
    np.savez_compressed(os.path.join(outdir, "cyl_birth.npz"), frames=np.array(frames),
                        T_final=T, times=times)
    print(f"[cyl_birth] final nz={grid.nz}, {len(frames)} frames")


def gen_voxel_bc(ref, outdir):
    """voxel_bc_correction.build_corrected_robin_fields on a duck-typed ellipsoid mesh."""
    import voxel_bc_correction as vb
    for name in ("ellipsoid", "coarse"):
        c = cases.build_voxel_bc_case(name)
        robin, scale = vb.build_corrected_robin_fields(c["mesh"], c["mask"], c["origin"], c["dx"], c["base_h"],
                                                       fallback_to_base=True, max_subdiv=c["max_subdiv"])
        robin_nf, _ = vb.build_corrected_robin_fields(c["mesh"], c["mask"], c["origin"], c["dx"], c["base_h"],
                                                      fallback_to_base=False, max_subdiv=c["max_subdiv"])
        out = {}
        for f in robin:
            out["robin_" + f] = robin[f]
            out["scale_" + f] = scale[f]
            out["robin_nofallback_" + f] = robin_nf[f]
        np.savez_compressed(os.path.join(outdir, f"voxel_bc_{name}.npz"), **out)
        print(f"[voxel_bc] {name}: {len(c['mesh'].triangles)} triangles, "
              f"{int(sum((robin[f] > 0).sum() for f in robin))} corrected face entries")


def gen_vtk_text(ref, outdir):
    """Files written by the reference's two ASCII VTK writers (vtk_writer.py:12,
    waam_from_stl_v7_mm.py:186) for tests/cases.py:vtk_text_cases -> vtk_text.npz (raw bytes)."""
    import vtk_writer
    _stub_matplotlib()
    import waam_from_stl_v7_mm as waam
    out = {}
    tmp = tempfile.mkdtemp(prefix="vtk_golden_")
    for name, c in cases.vtk_text_cases().items():
        for fl, fn in ((0, vtk_writer.write_vtk_structured_points), (1, waam.write_vtk_structured_points)):
            path = os.path.join(tmp, f"{name}_{fl}.vtk")
            fn(path, c["T"], c["dx"], c["origin"], c["field_name"], c["mask"])
            with open(path, "rb") as f:
                out[f"{name}__{fl}"] = np.frombuffer(f.read(), dtype=np.uint8)
            print(f"[vtk] {name} flavour {fl}: {out[f'{name}__{fl}'].size} bytes")
    import hashlib
    c = cases.vtk_text_big_case()
    for fl, fn in ((0, vtk_writer.write_vtk_structured_points), (1, waam.write_vtk_structured_points)):
        path = os.path.join(tmp, f"big_{fl}.vtk")
        fn(path, c["T"], c["dx"], c["origin"], c["field_name"], c["mask"])
        with open(path, "rb") as f:
            data = f.read()
        out[f"big_sha256__{fl}"] = np.frombuffer(hashlib.sha256(data).digest(), dtype=np.uint8)
        out[f"big_size__{fl}"] = np.array([len(data)], dtype=np.int64)
        print(f"[vtk] big flavour {fl}: {len(data)} bytes, sha256 {hashlib.sha256(data).hexdigest()[:16]}...")
    np.savez_compressed(os.path.join(outdir, "vtk_text.npz"), **out)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="", help="generate one family only (e.g. vtk)")
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden"))
    a = ap.parse_args()
    sys.path.insert(0, a.ref)
    os.makedirs(a.out, exist_ok=True)
    if a.only == "vtk":
        gen_vtk_text(a.ref, a.out)
        return
    if a.only == "cyl":
        gen_cyl(a.ref, a.out)
        return
    gen_cart(a.ref, a.out)
    gen_cart_gpu_algo(a.ref, a.out)
    gen_cyl(a.ref, a.out)
    gen_spiral_sim(a.ref, a.out)
    gen_cyl_birth(a.ref, a.out)
    gen_voxel_bc(a.ref, a.out)
    gen_vtk_text(a.ref, a.out)


if __name__ == "__main__":
    main()
