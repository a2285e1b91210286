# -*- coding: utf-8 -*-
"""The table-driven partitioned solve of csrc/adi_tab_core.h (host-tabulated elimination
factors, PCR on the chunk separators, wrap-around PCR for the periodic phi lines), compiled
for the host and run against golden outputs of the UNMODIFIED reference
(adi3d_cyl_phi_v3.adi_step scheme "be" / quick_spiral_deposition_gif_v5.adi_step_masked).
Tolerance: relative L2 <= 1e-12 per step (north_star)."""
import os

import numpy as np
import pytest

import cases
import emu

TOL = 1e-12


@pytest.mark.parametrize("M", [4, 8, 16, 32])
@pytest.mark.parametrize("name", cases.ALL_CYL_CASES)
def test_emulated_cyl_step_matches_reference(name, M, golden_dir):
    c = cases.build_cyl_case(name)
    g = np.load(os.path.join(golden_dir, f"cyl_{name}.npz"))
    out = emu.cyl_step(c, M=M)
    assert cases.rel_l2(out, g["T_out"]) <= TOL
    if c["active"] is not None:   # clamped cells are exact
        assert np.array_equal(out[~c["active"]], g["T_out"][~c["active"]])


def test_phi_rings_worst_case():
    """Periodic solve at the conditioning of BASELINE config 3 (fac up to ~1.2e4 on ring 1):
    every ring against a dense solve."""
    import math
    from oracle import cyl
    nr, nphi, nz = 24, 256, 2
    R = 0.02
    dr = R / 256          # radial spacing of the 256 x 1024 x 512 grid
    dphi = 2 * math.pi / 1024
    grid = cyl.GridCyl(nr, nphi, nz, dr, dphi, dr, R)
    mat = cyl.Material(cases.C_RHO, cases.C_CP, cases.C_K)
    dt = dr * dr / mat.alpha
    fac = cyl.phi_fac(grid, mat, dt)
    assert fac[1] > 1.0e4
    c = dict(nr=nr, nphi=nphi, nz=nz, dr=dr, dphi=dphi, dz=dr, rho=mat.rho, cp=mat.cp, k=mat.k, dt=dt,
             h_r=0.0, Tinf_r=20.0, T_void=0.0, T_inner=0.0, active=None, S=None,
             zbc=dict(kind_bot="neumann0", kind_top="neumann0", h_bot=0.0, h_top=0.0, T_inf_bot=0.0,
                      T_inf_top=0.0, T_bot=0.0, T_top=0.0),
             T0=20.0 + 980.0 * cases.splitmix_uniform(99, (nr, nphi, nz)))
    out = emu.cyl_step(c, M=16)
    ref = cyl.adi_step(c["T0"], grid, mat, cyl.Params(dt, 1.0, "be"), cyl.RobinR(0.0, 20.0), cyl.ZBC(**c["zbc"]))
    for ir in range(nr):
        assert cases.rel_l2(out[ir], ref[ir]) <= TOL, ir


@pytest.mark.parametrize("nslab", [2, 3, 5])
@pytest.mark.parametrize("name", ["mid_default", "mid_dd", "mid_rr", "mid_rd", "odd_sizes", "big", "big_cfl50", "masked_mid", "mid_source"])
def test_emulated_cyl_z_slabs_match_reference(name, nslab, golden_dir):
    """The z sweep cut into nslab segments per line (ghost couplings, tabulated ghost responses,
    inter-segment solve): the multi-GPU z-slab algorithm of the cylindrical path, against the reference."""
    c = cases.build_cyl_case(name)
    g = np.load(os.path.join(golden_dir, f"cyl_{name}.npz"))
    for M in (4, 16):
        out = emu.cyl_step(c, M=M, nslab=nslab)
        assert cases.rel_l2(out, g["T_out"]) <= TOL
