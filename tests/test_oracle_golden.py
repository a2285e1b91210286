# -*- coding: utf-8 -*-
"""Pins the oracle (oracle/adi_oracle.c, oracle/cyl.py) to outputs of the UNMODIFIED
reference (tests/golden/*.npz, written by tools/gen_golden.py in the build container).
Bit-exact for the Cartesian C restatement and for the NumPy cylindrical restatement."""
import os

import numpy as np
import pytest

import cases
from oracle import cart, cyl


def _cart_run(c, threads=1):
    cart.set_threads(threads)
    nx, ny, nz = c["shape"]
    grid = cart.Grid3D(nx, ny, nz, c["dx"], c["mask"])
    mat = cart.Material(c["rho"], c["cp"], c["k"])
    prm = cart.Params(c["dt"], c["theta"])
    packs = cart.precompute_coeff_packs_unified(grid, mat, robin_Tinf=c["Tinf"], **c["bcs"])
    T = c["T0"].copy()
    trace = []
    for _ in range(c["nsteps"]):
        T = cart.adi_step_numba_coeff(T, grid, mat, prm, packs, Tinf=c["Tinf"])
        trace.append(T[nx // 2, ny // 2, :].copy())
    return T, np.array(trace), packs


@pytest.mark.parametrize("name", sorted(cases.CART_CASES))
def test_cart_oracle_bit_exact(name, golden_dir):
    c = cases.build_cart_case(name)
    g = np.load(os.path.join(golden_dir, f"cart_{name}.npz"))
    T, trace, _ = _cart_run(c)
    assert np.array_equal(T, g["T_out"], equal_nan=True)
    assert np.array_equal(trace, g["trace"], equal_nan=True)
    # void cells pass through bit-identically (adi3d_numba_coeff.py:135 out = R0.copy())
    v = ~c["mask"]
    assert np.array_equal(T[v], c["T0"][v], equal_nan=True)


@pytest.mark.parametrize("name", ["holes_combined", "B_holes_combined"])
def test_cart_oracle_threads_do_not_change_bits(name, golden_dir):
    c = cases.build_cart_case(name)
    g = np.load(os.path.join(golden_dir, f"cart_{name}.npz"))
    T, _, _ = _cart_run(c, threads=max(2, min(4, cart.max_threads())))
    cart.set_threads(1)
    assert np.array_equal(T, g["T_out"], equal_nan=True)


@pytest.mark.parametrize("name", ["holes_combined", "track_mixed", "random_neumann_fields"])
def test_pack_builder_bit_exact(name, golden_dir):
    c = cases.build_cart_case(name)
    g = np.load(os.path.join(golden_dir, f"packs_{name}.npz"))
    _, _, packs = _cart_run(c)
    for a, ax in enumerate("xyz"):
        assert np.array_equal(packs[a].coeff, g[f"coeff_{ax}"])
        assert np.array_equal(packs[a].qflux, g[f"q_{ax}"])
    assert np.array_equal(packs[0].dir_mask, g["dir_mask"])
    assert np.array_equal(packs[0].dir_val, g["dir_val"])


@pytest.mark.parametrize("name", ["cyl_robin6", "cyl_dirtop", "B_track_dict3d", "full_dict3d_cfl3000"])
def test_reference_gpu_algorithm_agrees(name, golden_dir):
    """Second pin: adi3d_gpu_coeff.py's un-compressed-line algorithm (run on NumPy) vs the oracle."""
    c = cases.build_cart_case(name)
    g = np.load(os.path.join(golden_dir, f"cartgpu_{name}.npz"))
    T, _, _ = _cart_run(c)
    assert cases.rel_l2(T, g["T_out"], c["mask"]) < 1e-13
    v = ~c["mask"]
    assert np.array_equal(T[v], g["T_out"][v], equal_nan=True)


def test_exposed_mask_faces_and_error():
    m = cases.make_mask("thin", cases.SHAPE_A, 3)
    e = cart.exposed_mask(m, "z+")
    ref = np.zeros_like(m)
    ref[:, :, :-1] = m[:, :, :-1] & ~m[:, :, 1:]
    ref[:, :, -1] = m[:, :, -1]
    assert np.array_equal(e, ref)
    e = cart.exposed_mask(m, "x-")
    ref = np.zeros_like(m)
    ref[1:] = m[1:] & ~m[:-1]
    ref[0] = m[0]
    assert np.array_equal(e, ref)
    with pytest.raises(ValueError):
        cart.exposed_mask(m, "w+")


def _cyl_objs(c):
    grid = cyl.GridCyl(c["nr"], c["nphi"], c["nz"], c["dr"], c["dphi"], c["dz"], c["R"])
    mat = cyl.Material(c["rho"], c["cp"], c["k"])
    prm = cyl.Params(c["dt"], 1.0, "be")
    rob = cyl.RobinR(c["h_r"], c["Tinf_r"])
    zbc = cyl.ZBC(**c["zbc"])
    return grid, mat, prm, rob, zbc


def _cyl_run(c, solver):
    grid, mat, prm, rob, zbc = _cyl_objs(c)
    if c["active"] is not None:
        return cyl.adi_step_masked(c["T0"], grid, mat, prm, rob, zbc, c["active"],
                                   robin_inner=cyl.RobinR(c["h_r"], c["T_inner"]),
                                   robin_void=cyl.RobinR(c["h_r"], c["T_void"]), phi_solver=solver)
    return cyl.adi_step(c["T0"], grid, mat, prm, rob, zbc, S=c["S"], phi_solver=solver)


@pytest.mark.parametrize("name", cases.ALL_CYL_CASES)
def test_cyl_oracle_bit_exact(name, golden_dir):
    c = cases.build_cyl_case(name)
    g = np.load(os.path.join(golden_dir, f"cyl_{name}.npz"))
    assert np.array_equal(_cyl_run(c, cyl.phi_solve_spectral), g["T_out"])


@pytest.mark.parametrize("name", cases.ALL_CYL_CASES)
def test_cyclic_sherman_morrison_matches_spectral(name, golden_dir):
    """The direct cyclic solve the CUDA kernel implements vs the reference's FFT solve."""
    c = cases.build_cyl_case(name)
    g = np.load(os.path.join(golden_dir, f"cyl_{name}.npz"))
    assert cases.rel_l2(_cyl_run(c, cyl.phi_solve_cyclic), g["T_out"]) < 1e-13


def test_cyl_c3_slice_bit_exact(golden_dir):
    """A 256 x 1024 x 16 slab of BASELINE configs[2] itself (r lines of 256 cells, phi rings of 1024): the golden
    file holds every 32nd phi row and the SHA-256 of the reference's whole output."""
    import hashlib
    c = cases.build_cyl_case("c3_slice")
    g = np.load(os.path.join(golden_dir, "cyl_c3_slice.npz"))
    out = _cyl_run(c, cyl.phi_solve_spectral)
    assert np.array_equal(out[:, ::32, :], g["T_sub"])
    assert hashlib.sha256(np.ascontiguousarray(out).tobytes()).digest() == g["sha256"].tobytes()


def test_cyl_bad_kind_raises():
    c = cases.build_cyl_case("mid_default")
    grid, mat, prm, rob, _ = _cyl_objs(c)
    with pytest.raises(ValueError):
        cyl.adi_step(c["T0"], grid, mat, prm, rob, cyl.ZBC(kind_bot="bogus"))
    with pytest.raises(ValueError):
        cyl.adi_step(c["T0"], grid, mat, prm, rob, cyl.ZBC(kind_top="bogus"))


def test_cyl_birth_sequence(golden_dir):
    """nz-growth event loop of quick_compare_layer_birth_robin_cyl_v3.py:171-204 (see gen_golden)."""
    g = np.load(os.path.join(golden_dir, "cyl_birth.npz"))
    R, z_back, d, t_step, N_total, t_tail = 0.02, 0.02, 0.005, 0.5, 3, 0.5
    nr, nphi = 8, 16
    dr = R / nr
    dz = dr
    dphi = (2.0 * np.pi) / nphi
    mat = cyl.Material(7800.0, 490.0, 54.0)
    dt0 = 1.0 * min(dr * dr, dz * dz, (R * dphi) ** 2) / mat.alpha
    nz0 = int(round((z_back + d) / dz))
    grid = cyl.GridCyl(nr, nphi, nz0, dr, dphi, dz, R)
    rob = cyl.RobinR(500.0, 20.0)
    zbc = cyl.ZBC("neumann0", "robin", h_top=500.0, T_inf_top=20.0)
    T = np.full((nr, nphi, nz0), 20.0)
    nz_extra = int(round(d / dz))
    T[:, :, -nz_extra:] = 1000.0
    nz_final = int(round((z_back + N_total * d) / dz))
    t, next_birth, eps = 0.0, t_step, 1e-12
    frames = []
    for t_target in g["times"][1:]:
        while t < t_target - eps:
            dt_step = min(dt0, t_target - t, max(eps, next_birth - t))
            T = cyl.adi_step(T, grid, mat, cyl.Params(dt_step, 1.0, "be"), rob, zbc)
            t += dt_step
            if abs(t - next_birth) <= eps:
                if grid.nz + nz_extra <= nz_final:
                    old = T
                    T = np.full((nr, nphi, grid.nz + nz_extra), 20.0)
                    T[:, :, : old.shape[2]] = old
                    T[:, :, -nz_extra:] = 1000.0
                    grid = cyl.GridCyl(nr, nphi, T.shape[2], dr, dphi, dz, R)
                next_birth += t_step
        t = t_target
        y = np.full(nz_final, np.nan)
        y[: grid.nz] = T[0, 0, :]
        frames.append(y)
    assert np.array_equal(np.array(frames), g["frames"], equal_nan=True)
    assert np.array_equal(T, g["T_final"])


def test_spiral_simulation_oracle_bit_exact(golden_dir):
    """The deposition event loop of the reference's only test (tests/test_spiral_vs_analytic.py:
    17-120) run on the oracle reproduces the unmodified reference's snapshots bit for bit."""
    import spiral_loop
    g = np.load(os.path.join(golden_dir, "spiral_sim.npz"))
    snaps, acts = spiral_loop.run(cyl, g["times"])
    assert np.array_equal(np.array(acts), g["active"])
    assert np.array_equal(np.array(snaps), g["snapshots"])


@pytest.mark.parametrize("name", ["ellipsoid", "coarse"])
def test_voxel_bc_oracle_bit_exact(name, golden_dir):
    """oracle/voxel_bc.py against voxel_bc_correction.build_corrected_robin_fields of the unmodified
    reference on a duck-typed ellipsoid mesh (with and without the fallback to the base coefficient)."""
    from oracle import voxel_bc
    c = cases.build_voxel_bc_case(name)
    g = np.load(os.path.join(golden_dir, f"voxel_bc_{name}.npz"))
    robin, scale = voxel_bc.build_corrected_robin_fields(c["mesh"], c["mask"], c["origin"], c["dx"], c["base_h"],
                                                         True, c["max_subdiv"])
    nofb, _ = voxel_bc.build_corrected_robin_fields(c["mesh"], c["mask"], c["origin"], c["dx"], c["base_h"],
                                                    False, c["max_subdiv"])
    assert list(robin) == list(c["base_h"])
    for f in robin:
        assert np.array_equal(robin[f], g["robin_" + f])
        assert np.array_equal(scale[f], g["scale_" + f])
        assert np.array_equal(nofb[f], g["robin_nofallback_" + f])


# ---- ASCII VTK writers (oracle/vtk_text.py vs the reference's files) --------------------------
@pytest.mark.parametrize("fmt", [0, 1])
@pytest.mark.parametrize("name", sorted(cases.vtk_text_cases()))
def test_vtk_text_oracle_byte_exact(name, fmt, golden_dir):
    from oracle import vtk_text
    g = np.load(os.path.join(golden_dir, "vtk_text.npz"))
    c = cases.vtk_text_cases()[name]
    got = vtk_text.vtk_bytes(fmt, c["T"], c["dx"], c["origin"], c["field_name"], c["mask"])
    assert got == g[f"{name}__{fmt}"].tobytes()


@pytest.mark.parametrize("fmt", [0, 1])
def test_vtk_text_big_case_digest(fmt, golden_dir):
    """Rows longer than one 256-value piece: oracle and the host build of the device formatter against the
    digest of the reference's own file."""
    import hashlib
    import emu
    from oracle import vtk_text
    g = np.load(os.path.join(golden_dir, "vtk_text.npz"))
    c = cases.vtk_text_big_case()
    data = vtk_text.vtk_bytes(fmt, c["T"], c["dx"], c["origin"], c["field_name"], c["mask"])
    assert len(data) == int(g[f"big_size__{fmt}"][0])
    assert hashlib.sha256(data).digest() == g[f"big_sha256__{fmt}"].tobytes()
    head = vtk_text.header(fmt, c["T"].shape, c["dx"], c["origin"], c["field_name"]).encode()
    sec = vtk_text.section_header("mask" if fmt == 0 else "Mask").encode()
    dev = head + emu.text_field(c["T"], fmt) + sec + emu.text_field(c["mask"], fmt)
    assert dev == data
