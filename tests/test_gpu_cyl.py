# -*- coding: utf-8 -*-
"""GPU parity of the cylindrical path through the C ABI (adi_cyl_step / adi_cyl_step_host)
against golden outputs of the UNMODIFIED reference (adi3d_cyl_phi_v3.adi_step scheme "be",
quick_spiral_deposition_gif_v5.adi_step_masked) and against the oracle.
Tolerance: relative L2 <= 1e-12 per step (north_star)."""
import math
import os

import numpy as np
import pytest

import cases

pytestmark = pytest.mark.gpu
TOL = 1e-12


@pytest.fixture(scope="module")
def gc():
    from adi_thermal_fields_b200 import adi3d_cyl_phi_v3 as m
    return m


def _objs(m, c):
    grid = m.GridCyl(c["nr"], c["nphi"], c["nz"], c["dr"], c["dphi"], c["dz"], c["R"])
    return grid, m.Material(c["rho"], c["cp"], c["k"]), m.Params(c["dt"], 1.0, "be"), \
        m.RobinR(c["h_r"], c["Tinf_r"]), m.ZBC(**c["zbc"])


def _run(m, c):
    grid, mat, prm, rob, zbc = _objs(m, c)
    if c["active"] is not None:
        return m.adi_step_masked(c["T0"], grid, mat, prm, rob, zbc, c["active"],
                                 robin_inner=m.RobinR(c["h_r"], c["T_inner"]),
                                 robin_void=m.RobinR(c["h_r"], c["T_void"]))
    return m.adi_step(c["T0"], grid, mat, prm, rob, zbc, S=c["S"])


@pytest.mark.parametrize("name", cases.ALL_CYL_CASES)
def test_cyl_step_matches_reference(name, gc, golden_dir):
    c = cases.build_cyl_case(name)
    g = np.load(os.path.join(golden_dir, f"cyl_{name}.npz"))
    T0 = c["T0"].copy()
    n0 = gc.launch_count()
    out = _run(gc, c)
    assert gc.launch_count() - n0 == (3 if c["nphi"] > 1 else 2)   # the CUDA kernels ran
    assert isinstance(out, np.ndarray) and out.shape == T0.shape
    assert np.array_equal(c["T0"], T0)                             # input not modified
    assert cases.rel_l2(out, g["T_out"]) <= TOL
    if c["active"] is not None:
        assert np.array_equal(out[~c["active"]], g["T_out"][~c["active"]])


@pytest.mark.parametrize("name", ["mid_default", "mid_source", "masked_mid", "odd_sizes", "big"])
def test_device_resident_step(name, gc, golden_dir):
    """adi_step_device: torch CUDA tensors in/out, no host transfer."""
    import torch
    c = cases.build_cyl_case(name)
    g = np.load(os.path.join(golden_dir, f"cyl_{name}.npz"))
    grid, mat, prm, rob, zbc = _objs(gc, c)
    dev = torch.device("cuda", 0)
    T = torch.from_numpy(c["T0"]).to(dev)
    kw = {}
    if c["S"] is not None:
        kw["S"] = torch.from_numpy(c["S"]).to(dev)
    if c["active"] is not None:
        kw.update(active=torch.from_numpy(c["active"]).to(dev), robin_inner=gc.RobinR(0.0, c["T_inner"]),
                  robin_void=gc.RobinR(0.0, c["T_void"]))
    out = gc.adi_step_device(T, grid, mat, prm, rob, zbc, **kw)
    assert out.data_ptr() != T.data_ptr()
    assert np.array_equal(T.cpu().numpy(), c["T0"])
    assert cases.rel_l2(out.cpu().numpy(), g["T_out"]) <= TOL


def test_pitched_buffers_grow_in_place(gc, golden_dir):
    """Layer births on a pre-pitched device buffer: nz grows, nothing is reallocated
    (quick_compare_layer_birth_robin_cyl_v3.py:195-204), cells beyond nz stay untouched."""
    import torch
    from oracle import cyl
    c = cases.build_cyl_case("mid_default")
    nr, nphi, pitch = c["nr"], c["nphi"], 48
    dev = torch.device("cuda", 0)
    buf = torch.full((nr, nphi, pitch), -7.0, dtype=torch.float64, device=dev)
    host = c["T0"][:, :, :24].copy()
    buf[:, :, :24] = torch.from_numpy(host).to(dev)
    mat, rob, zbc = gc.Material(c["rho"], c["cp"], c["k"]), gc.RobinR(c["h_r"], c["Tinf_r"]), gc.ZBC(**c["zbc"])
    omat, orob, ozbc = cyl.Material(c["rho"], c["cp"], c["k"]), cyl.RobinR(c["h_r"], c["Tinf_r"]), cyl.ZBC(**c["zbc"])
    for nz in (24, 32, 40):
        if nz > host.shape[2]:
            grown = np.full((nr, nphi, nz), 20.0)
            grown[:, :, : host.shape[2]] = host
            grown[:, :, host.shape[2]:] = 1000.0
            buf[:, :, host.shape[2]:nz] = 1000.0
            host = grown
        args = (nr, nphi, nz, c["dr"], c["dphi"], c["dz"], c["R"])
        for _ in range(2):
            buf = gc.adi_step_device(buf, gc.GridCyl(*args), mat, gc.Params(c["dt"], 1.0, "be"), rob, zbc,
                                     nz_pitch=pitch)
            host = cyl.adi_step(host, cyl.GridCyl(*args), omat, cyl.Params(c["dt"], 1.0, "be"), orob, ozbc)
        got = buf.cpu().numpy()
        assert cases.rel_l2(got[:, :, :nz], host) <= 10 * TOL       # six steps accumulated
        assert np.all(got[:, :, nz:] == -7.0)


def test_birth_sequence_matches_reference(gc, golden_dir):
    """The nz-growth event loop of quick_compare_layer_birth_robin_cyl_v3.py:171-204 through the
    host-array entry point, against frames produced by the unmodified reference."""
    g = np.load(os.path.join(golden_dir, "cyl_birth.npz"))
    R, z_back, d, t_step, N_total = 0.02, 0.02, 0.005, 0.5, 3
    nr, nphi = 8, 16
    dr = R / nr
    dz = dr
    dphi = (2.0 * np.pi) / nphi
    mat = gc.Material(7800.0, 490.0, 54.0)
    dt0 = 1.0 * min(dr * dr, dz * dz, (R * dphi) ** 2) / mat.alpha
    nz0 = int(round((z_back + d) / dz))
    grid = gc.GridCyl(nr, nphi, nz0, dr, dphi, dz, R)
    rob = gc.RobinR(500.0, 20.0)
    zbc = gc.ZBC("neumann0", "robin", h_top=500.0, T_inf_top=20.0)
    T = np.full((nr, nphi, nz0), 20.0)
    nz_extra = int(round(d / dz))
    T[:, :, -nz_extra:] = 1000.0
    nz_final = int(round((z_back + N_total * d) / dz))
    t, next_birth, eps = 0.0, t_step, 1e-12
    frames = []
    for t_target in g["times"][1:]:
        while t < t_target - eps:
            dt_step = min(dt0, t_target - t, max(eps, next_birth - t))
            T = gc.adi_step(T, grid, mat, gc.Params(dt_step, 1.0, "be"), rob, zbc)
            t += dt_step
            if abs(t - next_birth) <= eps:
                if grid.nz + nz_extra <= nz_final:
                    old = T
                    T = np.full((nr, nphi, grid.nz + nz_extra), 20.0)
                    T[:, :, : old.shape[2]] = old
                    T[:, :, -nz_extra:] = 1000.0
                    grid = gc.GridCyl(nr, nphi, T.shape[2], dr, dphi, dz, R)
                next_birth += t_step
        t = t_target
        y = np.full(nz_final, np.nan)
        y[: grid.nz] = T[0, 0, :]
        frames.append(y)
    ref = g["frames"]
    ok = ~np.isnan(ref)
    assert np.array_equal(np.isnan(np.array(frames)), ~ok)
    assert cases.rel_l2(np.array(frames)[ok], ref[ok]) <= 1e-10       # ~100 steps accumulated
    assert cases.rel_l2(T, g["T_final"]) <= 1e-10


def test_bad_kinds_and_scheme_raise(gc):
    c = cases.build_cyl_case("mid_default")
    grid, mat, prm, rob, _ = _objs(gc, c)
    with pytest.raises(ValueError):
        gc.adi_step(c["T0"], grid, mat, prm, rob, gc.ZBC(kind_bot="bogus"))
    with pytest.raises(ValueError):
        gc.adi_step(c["T0"], grid, mat, prm, rob, gc.ZBC(kind_top="bogus"))
    with pytest.raises(NotImplementedError):
        gc.adi_step(c["T0"], grid, mat, gc.Params(c["dt"], 0.5, "douglas"), rob, gc.ZBC())
    with pytest.raises(ValueError):
        gc.adi_step(c["T0"][:, :, :-1], grid, mat, prm, rob, gc.ZBC())


def test_long_lines_M32(gc):
    """Lines longer than 512 cells use the 32-cell chunks (phi with 1024 cells as in BASELINE
    config 3, ring conditioning fac ~ 1e4); checked against the oracle."""
    from oracle import cyl
    nr, nphi, nz = 5, 1024, 6
    R = 0.02
    dr, dphi = R / 256, 2 * math.pi / nphi
    mat = gc.Material(cases.C_RHO, cases.C_CP, cases.C_K)
    dt = dr * dr / mat.alpha
    T0 = 20.0 + 980.0 * cases.splitmix_uniform(4242, (nr, nphi, nz))
    zk = dict(kind_bot="neumann0", kind_top="robin", h_top=500.0, T_inf_top=20.0)
    out = gc.adi_step(T0, gc.GridCyl(nr, nphi, nz, dr, dphi, dr, R), mat, gc.Params(dt, 1.0, "be"),
                      gc.RobinR(500.0, 20.0), gc.ZBC(**zk))
    ref = cyl.adi_step(T0, cyl.GridCyl(nr, nphi, nz, dr, dphi, dr, R), cyl.Material(mat.rho, mat.cp, mat.k),
                       cyl.Params(dt, 1.0, "be"), cyl.RobinR(500.0, 20.0), cyl.ZBC(**zk))
    assert cases.rel_l2(out, ref) <= TOL
    # z lines of 700 cells
    nr, nphi, nz = 3, 4, 700
    T0 = 20.0 + 980.0 * cases.splitmix_uniform(4243, (nr, nphi, nz))
    dphi = 2 * math.pi / nphi
    out = gc.adi_step(T0, gc.GridCyl(nr, nphi, nz, dr, dphi, dr, R), mat, gc.Params(dt, 1.0, "be"),
                      gc.RobinR(500.0, 20.0), gc.ZBC(**zk))
    ref = cyl.adi_step(T0, cyl.GridCyl(nr, nphi, nz, dr, dphi, dr, R), cyl.Material(mat.rho, mat.cp, mat.k),
                       cyl.Params(dt, 1.0, "be"), cyl.RobinR(500.0, 20.0), cyl.ZBC(**zk))
    assert cases.rel_l2(out, ref) <= TOL


def test_spiral_simulation_snapshots(gc, golden_dir):
    """tests/test_spiral_vs_analytic.py:_run_numeric_simulation (the deposition event loop of the
    reference's only test) driven through adi_step_masked on the GPU, against snapshots the
    unmodified reference produced (GridCyl accepting R_in on both sides)."""
    g = np.load(os.path.join(golden_dir, "spiral_sim.npz"))
    import spiral_loop
    snaps, act = spiral_loop.run(gc, g["times"])
    assert np.array_equal(np.array(act), g["active"])
    for s, r in zip(snaps, g["snapshots"]):
        assert cases.rel_l2(s, r) <= 1e-11      # up to 72 steps accumulated
    # per step the bar is 1e-12 (north_star): snapshot 0 is the initial state, snapshot 1 (t = 1 s) has 18 steps behind
    # it and still meets the one-step bar
    assert cases.rel_l2(snaps[1], g["snapshots"][1]) <= 1e-12


def test_c3_slice_matches_reference_and_oracle(gc, golden_dir):
    """256 x 1024 x 16: lines of BASELINE configs[2]'s own length along r (256 cells, Robin row at the far end) and phi
    (1024-cell rings) -- against the reference's golden sub-sample and, cell for cell, against the oracle (which
    reproduces the reference's whole array bit for bit, tests/test_oracle_golden.py)."""
    from oracle import cyl
    c = cases.build_cyl_case("c3_slice")
    g = np.load(os.path.join(golden_dir, "cyl_c3_slice.npz"))
    out = _run(gc, c)
    assert cases.rel_l2(out[:, ::32, :], g["T_sub"]) <= TOL
    ref = cyl.adi_step(c["T0"], cyl.GridCyl(c["nr"], c["nphi"], c["nz"], c["dr"], c["dphi"], c["dz"], c["R"]),
                       cyl.Material(c["rho"], c["cp"], c["k"]), cyl.Params(c["dt"], 1.0, "be"),
                       cyl.RobinR(c["h_r"], c["Tinf_r"]), cyl.ZBC(**c["zbc"]))
    assert cases.rel_l2(out, ref) <= TOL
    # worst single r line / phi ring, not only the global norm
    num = np.sqrt(((out - ref) ** 2).sum(axis=0)); den = np.sqrt((ref ** 2).sum(axis=0))
    assert float((num / den).max()) <= 10 * TOL


@pytest.mark.parametrize("zt", [1, 0], ids=["k_sweep_zt", "k_cyl_z"])
@pytest.mark.parametrize("shape", [(6, 8, 64), (5, 12, 96), (4, 6, 130), (3, 5, 512), (7, 3, 2048)], ids=lambda s: "x".join(map(str, s)))
@pytest.mark.parametrize("zk", [dict(kind_bot="neumann0", kind_top="robin", h_top=500.0, T_inf_top=20.0),
                                dict(kind_bot="robin", kind_top="neumann0", h_bot=300.0, T_inf_bot=35.0),
                                dict(kind_bot="robin", kind_top="robin", h_bot=80.0, h_top=500.0, T_inf_bot=20.0, T_inf_top=20.0),
                                dict(kind_bot="robin", kind_top="robin", h_bot=80.0, h_top=500.0, T_inf_bot=15.0, T_inf_top=25.0),
                                dict(kind_bot="neumann0", kind_top="neumann0"),
                                dict(kind_bot="dirichlet", kind_top="robin", T_bot=400.0, h_top=0.0, T_inf_top=20.0)],
                         ids=["n0-robin", "robin-n0", "robin-robin", "robin-robin-2Tinf", "n0-n0", "dirichlet-robin"])
def test_z_sweep_on_the_cartesian_kernel(shape, zk, zt, gc):
    """Unmasked z lines of 64..2048 cells run on the Cartesian z kernel (k_sweep_zt: the z rows of build_coeff_z are
    those of a full line with a Robin term at its ends); Dirichlet ends and two different ambient temperatures keep
    k_cyl_z.  cfl 1 and 30 (slowly decaying couplings), against the oracle; option cylzt=0 for comparison."""
    from oracle import cyl
    nr, nphi, nz = shape
    R = 0.02
    dr, dphi = R / 64, 2 * math.pi / nphi
    mat = gc.Material(cases.C_RHO, cases.C_CP, cases.C_K)
    gc.set_option("cylzt", zt)
    try:
        for cfl in (1.0, 30.0):
            dt = cfl * dr * dr / mat.alpha
            T0 = 20.0 + 980.0 * cases.splitmix_uniform(5100 + nz, shape)
            out = gc.adi_step(T0, gc.GridCyl(nr, nphi, nz, dr, dphi, dr, R), mat, gc.Params(dt, 1.0, "be"),
                              gc.RobinR(500.0, 20.0), gc.ZBC(**zk))
            ref = cyl.adi_step(T0, cyl.GridCyl(nr, nphi, nz, dr, dphi, dr, R), cyl.Material(mat.rho, mat.cp, mat.k),
                               cyl.Params(dt, 1.0, "be"), cyl.RobinR(500.0, 20.0), cyl.ZBC(**zk))
            assert cases.rel_l2(out, ref) <= TOL
    finally:
        gc.set_option("cylzt", 1)
