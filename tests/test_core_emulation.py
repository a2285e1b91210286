# -*- coding: utf-8 -*-
"""The partitioned line solve of csrc/adi_core.h (chunk elimination + PCR on the separators),
compiled for the host and run against the golden vectors of the reference.
Tolerance: relative L2 <= 1e-12 on active cells (north_star), void cells bit-identical."""
import os

import numpy as np
import pytest

import cases
import emu
from oracle import cart

TOL = 1e-12


def _packs(c):
    nx, ny, nz = c["shape"]
    grid = cart.Grid3D(nx, ny, nz, c["dx"], c["mask"])
    mat = cart.Material(c["rho"], c["cp"], c["k"])
    return cart.precompute_coeff_packs_unified(grid, mat, robin_Tinf=c["Tinf"], **c["bcs"])


@pytest.mark.parametrize("variant", [0, 1], ids=["M16_two_factors", "M32_one_factor"])
@pytest.mark.parametrize("name", sorted(cases.CART_CASES))
def test_emulated_kernel_matches_reference(name, variant, golden_dir):
    c = cases.build_cart_case(name)
    g = np.load(os.path.join(golden_dir, f"cart_{name}.npz"))
    packs = _packs(c)
    kappa = c["k"] / (c["rho"] * c["cp"])
    has_dir = packs[0].dir_mask.any()
    T = c["T0"]
    for _ in range(c["nsteps"]):
        T = emu.cart_step(T, c["mask"], c["dx"], c["dt"], c["theta"], kappa, c["Tinf"],
                          coeff=[p.coeff for p in packs],
                          dirm=[p.dir_mask if has_dir else None for p in packs],
                          dirv=[p.dir_val if has_dir else None for p in packs],
                          q=[p.qflux if p.qflux.any() else None for p in packs], variant=variant)
    m = c["mask"]
    assert cases.rel_l2(T, g["T_out"], m) <= TOL
    assert np.array_equal(T[~m], c["T0"][~m], equal_nan=True)


@pytest.mark.parametrize("name", ["full_robin6", "cyl_robin6", "track_robin6", "B_full_robin6"])
def test_emulated_scalar_robin_mode(name, golden_dir):
    """On-the-fly Robin coefficients from the neighbour code (no dense coeff array)."""
    c = cases.build_cart_case(name)
    g = np.load(os.path.join(golden_dir, f"cart_{name}.npz"))
    kappa = c["k"] / (c["rho"] * c["cp"])
    A, V = c["dx"] * c["dx"], c["dx"] ** 3
    Ccell = c["rho"] * c["cp"] * V
    fc = [c["bcs"]["robin_h"][f] * A / Ccell for f in cart.FACES]
    T = emu.cart_step(c["T0"], c["mask"], c["dx"], c["dt"], c["theta"], kappa, c["Tinf"], face_coeff=fc)
    assert cases.rel_l2(T, g["T_out"], c["mask"]) <= TOL
