# -*- coding: utf-8 -*-
"""Deposition event loop of the reference's only test
(tests/test_spiral_vs_analytic.py:17-120, `_run_numeric_simulation`, configuration :124-162),
restated over any module `m` that offers the reference's cylindrical interface
(build_grid_annular, Material, Params, RobinR, ZBC, adi_step_masked).  Test infrastructure:
used with the oracle (CPU) and with the CUDA backend (GPU) against tests/golden/spiral_sim.npz."""
import math

import numpy as np

CFG = dict(R_in=0.03, wall=0.002, nphi=36, tau_dep=2.0, n_layers=2, layer_h=0.004, z_back=0.02, nr=6,
           rho=7800.0, cp=490.0, k=54.0, h_side=400.0, h_end=500.0, T_inf=20.0, T_deposit=900.0,
           h_void=400.0, layer_cells=1, loops=1)


def run(m, times, cfg=CFG):
    R_out = cfg["R_in"] + cfg["wall"]
    grid, _, _, _ = m.build_grid_annular(R_out, cfg["wall"], cfg["layer_h"] * cfg["n_layers"], cfg["z_back"],
                                         cfg["nr"], cfg["nphi"], dz_override=cfg["layer_h"])
    mat = m.Material(cfg["rho"], cfg["cp"], cfg["k"])
    outer = m.RobinR(cfg["h_side"], cfg["T_inf"])
    inner = m.RobinR(cfg["h_side"], cfg["T_inf"])
    void = m.RobinR(cfg["h_void"], cfg["T_inf"])
    zbc = m.ZBC(kind_bot="neumann0", kind_top="robin", h_top=cfg["h_end"], T_inf_top=cfg["T_inf"])
    dt = cfg["tau_dep"] / cfg["nphi"]
    omega = 2.0 * math.pi / cfg["tau_dep"]
    iz0 = int(round(cfg["z_back"] / grid.dz))
    T = np.full((grid.nr, grid.nphi, grid.nz), cfg["T_inf"], dtype=float)
    active = np.zeros(T.shape, dtype=bool)
    active[:, :, :iz0] = True
    layer = loop = 0
    angle = 0.0
    iz = iz0
    prm = m.Params(dt, 1.0, "be")

    def deposit(iz, a0, a1):   # :59-76
        if iz < 0 or iz > grid.nz - 1 or a1 <= a0:
            return
        i0 = int(math.floor(a0 / grid.dphi))
        i1 = max(i0, int(math.floor((a1 - 1e-12) / grid.dphi)))
        for i in range(i0, i1 + 1):
            j = i % grid.nphi
            if not active[0, j, iz]:
                active[:, j, iz] = True
                T[:, j, iz] = cfg["T_deposit"]

    snaps, acts = [], []
    t, eps = 0.0, 1e-12
    for t_target in times:
        while t < t_target - eps:
            t_next = min(t + dt, t_target)
            left = omega * (t_next - t)
            while left > 0.0 and layer < cfg["n_layers"]:   # :84-103
                seg = min(left, 2.0 * math.pi - angle)
                if seg > 0.0:
                    deposit(iz, angle, angle + seg)
                    angle += seg
                    left -= seg
                if angle >= 2.0 * math.pi - 1e-15:
                    angle = 0.0
                    loop += 1
                    if loop >= cfg["loops"]:
                        loop = 0
                        layer += 1
                        iz = iz0 + layer * cfg["layer_cells"]
                        if iz > grid.nz - 1:
                            layer = cfg["n_layers"]
                            break
            prm.dt = t_next - t
            T[:] = m.adi_step_masked(T, grid, mat, prm, outer, zbc, active, robin_inner=inner, robin_void=void)
            t = t_next
        snaps.append(T.copy())
        acts.append(active.copy())
    return snaps, acts
