# -*- coding: utf-8 -*-
"""Shared pieces of the z-slab tests (host backend on CPU, CUDA backend on the GPU box)."""
import numpy as np
import torch

import cases
from adi_thermal_fields_b200 import slab
from oracle import cart


def make_problem(shape, mask_kind, bc_kind, theta, cfl, seed=7, nan_void=True):
    mask = cases.make_mask(mask_kind, shape, seed)
    bcs = cases.make_bcs(bc_kind, shape, mask, seed, 20.0)
    T0 = 20.0 + 1380.0 * cases.splitmix_uniform(seed + 1, shape)
    if nan_void:
        T0[~mask] = np.nan
    kappa = cases.K / (cases.RHO * cases.CP)
    return dict(shape=shape, mask=mask, bcs=bcs, T0=T0, theta=theta, dt=cfl * cases.DX ** 2 / kappa, Tinf=20.0)


def oracle_steps(pb, nsteps):
    nx, ny, nz = pb["shape"]
    g = cart.Grid3D(nx, ny, nz, cases.DX, pb["mask"])
    m = cart.Material(cases.RHO, cases.CP, cases.K)
    packs = cart.precompute_coeff_packs_unified(g, m, **pb["bcs"])
    T = pb["T0"]
    for _ in range(nsteps):
        T = cart.adi_step_numba_coeff(T, g, m, cart.Params(pb["dt"], pb["theta"]), packs, Tinf=pb["Tinf"])
    return T


def slice_bcs(bcs, z0, z1):
    def cut(v):
        if isinstance(v, np.ndarray):
            return np.ascontiguousarray(v[:, :, z0:z1])
        if isinstance(v, dict):
            return {k: cut(x) for k, x in v.items()}
        return v
    return {k: cut(v) for k, v in bcs.items()}


class Mat:
    rho, cp, k = cases.RHO, cases.CP, cases.K


def rank_run(comm, pb, nsteps, backend, options=None):
    """What one rank does: the calls a user of the reference API would make, on its slab."""
    class Prm:
        dt, theta = pb["dt"], pb["theta"]
    nx, ny, nz = pb["shape"]
    z0, z1 = slab.split_z(nz, comm.world)[comm.rank]
    grid = slab.SlabGrid3D(nx, ny, z1 - z0, cases.DX, pb["mask"][:, :, z0:z1], comm, backend=backend)
    for k, v in (options or {}).items():
        if getattr(grid.be, "dist", False) or k not in ("batches", "batch_min_lines", "spike_after", "spike_kmax", "overlap_halo", "spike_thr_log2"):
            grid.be.set_option(k, v)
    packs = slab.precompute_coeff_packs_unified(grid, Mat, **slice_bcs(pb["bcs"], z0, z1))
    T = grid.be.asarray(pb["T0"][:, :, z0:z1], torch.float64)
    n0 = grid.be.launch_count()
    for _ in range(nsteps):
        T = slab.adi_step_gpu_coeff(T, grid, Mat, Prm, packs, Tinf=pb["Tinf"])
    return (z0, z1, T.cpu().numpy(), grid.be.launch_count() - n0)


def assemble(shape, parts):
    out = np.empty(shape)
    for z0, z1, t, _ in parts:
        out[:, :, z0:z1] = t
    return out


CASES = [
    # world, shape,          mask,          bcs,            theta, cfl,  steps
    (2, (9, 11, 32),  "full",        "robin6",        0.5, 0.128, 2),
    (3, (10, 7, 48),  "cyl_holes",   "combined",      0.5, 2.0,   2),
    (4, (6, 9, 64),   "random",      "robin_dict3d",  0.5, 3000., 1),
    (2, (12, 8, 32),  "plate_track", "robin_mixed",   1.0, 2.0,   2),
    (3, (7, 6, 48),   "thin",        "dir_scalar_interior", 0.5, 2.0, 2),
    (2, (8, 8, 64),   "random",      "neumann_fields", 0.5, 0.128, 3),
]
