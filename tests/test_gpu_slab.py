# -*- coding: utf-8 -*-
"""z-slab decomposition on the GPU: R virtual ranks (threads, one engine context each) share
cuda:0 and run the CUDA kernels of the N>1 path -- halo-aware neighbour code and packs, explicit
stage with T halo planes, z sweep pass 1 (interface relations) / inter-rank solve / pass 2 --
against the oracle on the undivided grid.  rel-L2 <= 1e-12 per step, void cells bit-identical.
(The NCCL transport itself is exercised by tests/dist_check.py under torchrun.)"""
import numpy as np
import pytest

import cases
from slab_cases import CASES, assemble, make_problem, oracle_steps, rank_run

pytestmark = pytest.mark.gpu
TOL = 1e-12


@pytest.mark.parametrize("world,shape,mk,bk,theta,cfl,nsteps", CASES)
def test_slab_cuda_matches_oracle(world, shape, mk, bk, theta, cfl, nsteps):
    from adi_thermal_fields_b200 import slab
    pb = make_problem(shape, mk, bk, theta, cfl)
    ref = oracle_steps(pb, nsteps)
    parts = slab.LocalComm(world).run(lambda v: rank_run(v, pb, nsteps, None))
    out = assemble(shape, parts)
    m = pb["mask"]
    assert all(p[3] >= 5 * nsteps for p in parts)     # pack, x, y, z pass 1, iface solve, z pass 2 ran
    assert cases.rel_l2(out, ref, m) <= TOL * nsteps
    assert np.array_equal(out[~m], pb["T0"][~m], equal_nan=True)


@pytest.mark.parametrize("world", [2, 4, 8])
def test_slab_cuda_long_lines(world):
    """512-cell lines split over up to 8 ranks (64-cell segments), scalar and dense Robin."""
    from adi_thermal_fields_b200 import slab
    shape = (20, 33, 512)
    for bk in ("robin6", "robin_dict3d"):
        pb = make_problem(shape, "cyl_holes", bk, 0.5, 2.0, seed=11)
        ref = oracle_steps(pb, 1)
        parts = slab.LocalComm(world).run(lambda v: rank_run(v, pb, 1, None))
        out = assemble(shape, parts)
        assert cases.rel_l2(out, ref, pb["mask"]) <= TOL
        assert np.array_equal(out[~pb["mask"]], pb["T0"][~pb["mask"]], equal_nan=True)


def test_slab_rejects_bad_extent():
    from adi_thermal_fields_b200 import slab
    pb = make_problem((6, 6, 40), "full", "robin6", 0.5, 2.0)   # 20 planes per rank: not a multiple of 16
    with pytest.raises(ValueError):
        slab.LocalComm(2).run(lambda v: rank_run(v, pb, 1, None))


@pytest.mark.parametrize("world", [2, 3, 4])
@pytest.mark.parametrize("name", ["mid_default", "mid_dd", "mid_rr", "odd_sizes", "big", "big_cfl50", "masked_mid", "mid_source"])
def test_cyl_slab_cuda_matches_reference(name, world, golden_dir):
    """Cylindrical path, z-slab decomposition: R virtual ranks on cuda:0 (r / phi sweeps local, z sweep as
    pass 1 -> all-gather of (yf, yl) -> ghost map -> pass 2) against golden outputs of the reference."""
    import os
    import torch
    from adi_thermal_fields_b200 import adi3d_cyl_phi_v3 as gc, slab
    c = cases.build_cyl_case(name)
    g = np.load(os.path.join(golden_dir, f"cyl_{name}.npz"))
    ext = slab.split_z(c["nz"], world)
    nzs = [b - a for a, b in ext]
    mat, prm = gc.Material(c["rho"], c["cp"], c["k"]), gc.Params(c["dt"], 1.0, "be")
    rob, zbc = gc.RobinR(c["h_r"], c["Tinf_r"]), gc.ZBC(**c["zbc"])

    def rank_fn(v):
        z0, z1 = ext[v.rank]
        dev = torch.device("cuda", 0)
        grid = slab.SlabGridCyl(c["nr"], c["nphi"], z1 - z0, c["dr"], c["dphi"], c["dz"], c["R"], v, nz_per_rank=nzs)
        T = torch.from_numpy(np.ascontiguousarray(c["T0"][:, :, z0:z1])).to(dev)
        kw = {}
        if c["S"] is not None:
            kw["S"] = torch.from_numpy(np.ascontiguousarray(c["S"][:, :, z0:z1])).to(dev)
        if c["active"] is not None:
            kw.update(active=torch.from_numpy(np.ascontiguousarray(c["active"][:, :, z0:z1])).to(dev),
                      robin_inner=gc.RobinR(0.0, c["T_inner"]), robin_void=gc.RobinR(0.0, c["T_void"]))
        n0 = grid.launch_count()
        out = slab.adi_step_cyl(T, grid, mat, prm, rob, zbc, **kw)
        return z0, z1, out.cpu().numpy(), grid.launch_count() - n0

    parts = slab.LocalComm(world).run(rank_fn)
    out = np.empty_like(c["T0"])
    for z0, z1, t, nl in parts:
        out[:, :, z0:z1] = t
        assert nl == 5      # r, phi, z pass 1, ghost map, z pass 2
    assert cases.rel_l2(out, g["T_out"]) <= TOL


# ---- steady stepping: solve-first z pass with unit-ghost response corrections ---------------------
@pytest.mark.parametrize("world,shape,mk,bk,theta,cfl,kmax,expect", [
    (2, (9, 11, 64),  "full",        "robin6",       0.5, 0.128, 32, True),
    (3, (10, 7, 96),  "cyl_holes",   "combined",     0.5, 0.128, 32, True),    # voids and Dirichlet cells at slab faces
    (4, (6, 9, 128),  "random",      "robin_dict3d", 0.5, 0.3,   32, True),
    (2, (12, 8, 32),  "plate_track", "robin_mixed",  1.0, 0.128, 32, True),    # 16-cell segments: both ends reach every cell
    (2, (8, 8, 256),  "cyl_holes",   "robin_dict3d", 0.5, 2.0,   64, True),    # slower decay, longer reach
    (2, (8, 8, 128),  "full",        "robin6",       0.5, 50.0,  32, False),   # reach > kmax: stays with the two-pass form
])
def test_slab_solve_first_z_pass(world, shape, mk, bk, theta, cfl, kmax, expect):
    """After `spike_after` steps with unchanged (mask, packs, dt, theta) the z sweep becomes: solve the local
    segment with zero ghosts in place -> all-gather (yf, yl) -> ghosts -> add L*v + R*w on the cells within
    reach of the slab faces.  Same answer as the undivided grid, void cells untouched."""
    import torch
    from adi_thermal_fields_b200 import slab
    from slab_cases import Mat, slice_bcs
    nsteps = 5
    pb = make_problem(shape, mk, bk, theta, cfl, seed=31)
    ref = oracle_steps(pb, nsteps)

    def rank_fn(comm):
        class Prm:
            dt, theta = pb["dt"], pb["theta"]
        nx, ny, nz = pb["shape"]
        z0, z1 = slab.split_z(nz, comm.world)[comm.rank]
        grid = slab.SlabGrid3D(nx, ny, z1 - z0, cases.DX, pb["mask"][:, :, z0:z1], comm)
        grid.spike_after, grid.spike_kmax = 1, kmax
        packs = slab.precompute_coeff_packs_unified(grid, Mat, **slice_bcs(pb["bcs"], z0, z1))
        T = grid.be.asarray(pb["T0"][:, :, z0:z1], torch.float64)
        for _ in range(nsteps):
            T = slab.adi_step_gpu_coeff(T, grid, Mat, Prm, packs, Tinf=pb["Tinf"])
        return (z0, z1, T.cpu().numpy(), bool(grid._spikes))

    parts = slab.LocalComm(world).run(rank_fn)
    out = assemble(shape, parts)
    m = pb["mask"]
    assert [p[3] for p in parts] == [expect] * world
    assert cases.rel_l2(out, ref, m) <= TOL * nsteps
    assert np.array_equal(out[~m], pb["T0"][~m], equal_nan=True)
