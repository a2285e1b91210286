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


@pytest.mark.parametrize("variant", [0, 1], ids=["M16", "M32"])
@pytest.mark.parametrize("cfl,theta", [(0.128, 0.5), (40.0, 1.0), (3000.0, 0.5)])
@pytest.mark.parametrize("shape,mask_kind,bk", [((70, 6, 37), "full", "robin_dict3d"), ((40, 70, 35), "plate_track", "robin_dict3d"),
                                                ((6, 5, 133), "full", "robin6"), ((48, 51, 53), "cyl_holes", "robin_dict3d")])
def test_uniform_chunk_paths_against_the_oracle(shape, mask_kind, bk, cfl, theta, variant):
    """Runs of uniform cells (adi_core.h UniConst: tabulated elimination factors) with a general separator and,
    at the start of a line, a hand-eliminated first cell -- against the oracle directly, from cfl 0.128 to 3000."""
    mask = cases.make_mask(mask_kind, shape, 11)
    bcs = cases.make_bcs(bk, shape, mask, 11, 20.0)
    T0 = 20.0 + 1380.0 * cases.splitmix_uniform(12, shape)
    T0[~mask] = np.nan
    kappa = cases.K / (cases.RHO * cases.CP)
    dt = cfl * cases.DX ** 2 / kappa
    nx, ny, nz = shape
    hg, hm = cart.Grid3D(nx, ny, nz, cases.DX, mask), cart.Material(cases.RHO, cases.CP, cases.K)
    packs = cart.precompute_coeff_packs_unified(hg, hm, **bcs)
    ref = cart.adi_step_numba_coeff(T0, hg, hm, cart.Params(dt, theta), packs, Tinf=20.0)
    out = {}
    for on in (1, 0):
        emu.lib().emu_set_uniform(on)
        try:
            out[on] = emu.cart_step(T0, mask, cases.DX, dt, theta, kappa, 20.0, coeff=[p.coeff for p in packs], variant=variant)
        finally:
            emu.lib().emu_set_uniform(1)
        assert cases.rel_l2(out[on], ref, mask) <= TOL
        assert np.array_equal(out[on][~mask], T0[~mask], equal_nan=True)
    assert not np.array_equal(out[0], out[1], equal_nan=True) or min(shape) < 4   # the tabulated path really ran
