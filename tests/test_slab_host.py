# -*- coding: utf-8 -*-
"""N>1 path on the CPU box: the z-slab decomposition of adi_thermal_fields_b200.slab (halo
exchange, interface all-gather, inter-rank solve) with the host build of the kernels' code as
backend -- (a) R in-process ranks (LocalComm), (b) two gloo processes (TorchDistComm).
Checked against the oracle on the undivided grid: rel-L2 <= 1e-12 per step, void cells
bit-identical."""
import os
import sys

import numpy as np
import pytest

import cases
from adi_thermal_fields_b200 import slab
from slab_cases import CASES, assemble, make_problem, oracle_steps, rank_run

TOL = 1e-12
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("world,shape,mk,bk,theta,cfl,nsteps", CASES)
def test_slab_local_ranks_match_oracle(world, shape, mk, bk, theta, cfl, nsteps):
    from host_backend import HostBackend
    pb = make_problem(shape, mk, bk, theta, cfl)
    ref = oracle_steps(pb, nsteps)
    comm = slab.LocalComm(world)
    parts = comm.run(lambda v: rank_run(v, pb, nsteps, HostBackend()))
    out = assemble(shape, parts)
    m = pb["mask"]
    assert cases.rel_l2(out, ref, m) <= TOL * nsteps
    assert np.array_equal(out[~m], pb["T0"][~m], equal_nan=True)


def test_slab_one_rank_is_the_plain_step():
    from host_backend import HostBackend
    pb = make_problem((9, 11, 32), "cyl_holes", "combined", 0.5, 2.0)
    ref = oracle_steps(pb, 1)
    parts = slab.LocalComm(1).run(lambda v: rank_run(v, pb, 1, HostBackend()))
    assert cases.rel_l2(assemble(pb["shape"], parts), ref, pb["mask"]) <= TOL


def _gloo_worker(rank, world, port, q):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    import torch.distributed as dist
    from host_backend import HostBackend
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        pb = make_problem((10, 7, 32), "cyl_holes", "combined", 0.5, 2.0)
        # 4 steps: the last two take the solve-first z pass (slab.py, spike_after = 2)
        q.put(rank_run(slab.TorchDistComm(), pb, 4, HostBackend())[:3] + (0,))
    finally:
        dist.destroy_process_group()


def test_slab_two_gloo_processes_match_oracle():
    import socket
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    parts = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    pb = make_problem((10, 7, 32), "cyl_holes", "combined", 0.5, 2.0)
    ref = oracle_steps(pb, 4)
    out = assemble(pb["shape"], parts)
    assert cases.rel_l2(out, ref, pb["mask"]) <= 4 * TOL
    assert np.array_equal(out[~pb["mask"]], pb["T0"][~pb["mask"]], equal_nan=True)


# ---- steady stepping: solve-first z pass (slab.py switches to it after `spike_after` repeats) ------
@pytest.mark.parametrize("world,shape,mk,bk,theta,cfl,kmax,expect", [
    (2, (9, 11, 64),  "full",        "robin6",       0.5, 0.128, 32, True),
    (3, (10, 7, 96),  "cyl_holes",   "combined",     0.5, 0.128, 32, True),
    (2, (12, 8, 32),  "plate_track", "robin_mixed",  1.0, 0.128, 32, True),    # 16-cell segments
    (2, (6, 5, 128),  "random",      "robin_dict3d", 0.5, 2.0,   64, True),
    (2, (6, 6, 128),  "full",        "robin6",       0.5, 50.0,  32, False),   # slow decay: two-pass form stays
])
def test_slab_solve_first_z_pass_host(world, shape, mk, bk, theta, cfl, kmax, expect):
    import torch
    from host_backend import HostBackend
    from slab_cases import Mat, slice_bcs
    nsteps = 4
    pb = make_problem(shape, mk, bk, theta, cfl, seed=31)
    ref = oracle_steps(pb, nsteps)

    def rank_fn(comm):
        class Prm:
            dt, theta = pb["dt"], pb["theta"]
        nx, ny, nz = pb["shape"]
        z0, z1 = slab.split_z(nz, comm.world)[comm.rank]
        grid = slab.SlabGrid3D(nx, ny, z1 - z0, cases.DX, pb["mask"][:, :, z0:z1], comm, backend=HostBackend())
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
