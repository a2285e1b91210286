#!/usr/bin/env python
"""Multi-GPU parity check of the z-slab path over NCCL (run under torchrun on the GPU box):
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
      --master-port 29511 tests/dist_check.py
Every rank steps its slab of a seeded problem; rank 0 gathers the slabs and compares with the
oracle on the undivided grid (rel-L2 <= 1e-12 per step, void cells bit-identical).
A parity checker: it lives in tests/ because it imports oracle/ (test infrastructure); it is not collected by pytest
(no test_ prefix) -- tools/gpu_scale.sh launches it."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402
from adi_thermal_fields_b200 import slab  # noqa: E402
from slab_cases import make_problem, oracle_steps, rank_run  # noqa: E402

if "--python-seq" in sys.argv:    # the call-by-call sequencing of slab.py instead of adi_cart_slab_step
    slab.USE_LIBRARY_SEQUENCING = False
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ok = True
for shape, mk, bk, theta, cfl, nsteps, opts in [
        ((24, 20, 16 * world * 2), "cyl_holes", "combined", 0.5, 2.0, 3, None),
        # the explicit stage running while the T planes travel (option overlap_halo: face cells on the communication stream)
        ((24, 20, 16 * world * 2), "cyl_holes", "combined", 0.5, 2.0, 3, dict(overlap_halo=1)),
        ((20, 24, 16 * world), "random", "robin_dict3d", 0.5, 0.7, 3, dict(overlap_halo=1)),
        ((16, 40, 64 * world), "random", "robin_dict3d", 0.5, 3000.0, 2, None),
        ((32, 16, 16 * world), "plate_track", "robin6", 1.0, 0.128, 2, None),
        # steady stepping: the solve-first z form, its all-gathers in three overlapped line batches
        ((40, 37, 32 * world), "cyl_holes", "robin_dict3d", 0.5, 0.128, 6, dict(batches=3, batch_min_lines=64)),
        ((48, 48, 16 * world), "full", "robin6", 0.5, 0.3, 5, dict(batches=4, batch_min_lines=32))]:
    pb = make_problem(shape, mk, bk, theta, cfl)
    z0, z1, t, nl = rank_run(slab.TorchDistComm(), pb, nsteps, None, opts)
    parts = [None] * world
    dist.all_gather_object(parts, (z0, z1, t, nl))
    if rank == 0:
        out = np.empty(shape)
        for a, b, tt, _ in parts:
            out[:, :, a:b] = tt
        ref = oracle_steps(pb, nsteps)
        err = cases.rel_l2(out, ref, pb["mask"])
        void = bool(np.array_equal(out[~pb["mask"]], pb["T0"][~pb["mask"]], equal_nan=True))
        good = err <= 1e-12 * nsteps and void
        ok &= good
        print(f"[dist_check] world={world} seq={'library' if slab.USE_LIBRARY_SEQUENCING else 'python'} shape={shape} {mk}/{bk} theta={theta}: rel_l2={err:.2e} "
              f"void_bit_equal={void} launches/rank={parts[0][3]} {'OK' if good else 'FAIL'}", flush=True)
# cylindrical path: z-slab decomposition against the oracle on the undivided grid
from adi_thermal_fields_b200 import adi3d_cyl_phi_v3 as gc  # noqa: E402
from oracle import cyl  # noqa: E402
for nr, nphi, nz, kb, kt in [(16, 32, 24 * world, "neumann0", "robin"), (12, 40, 7 * world + 3, "robin", "dirichlet")]:
    R = 0.02
    dr = R / nr
    dphi = 2 * np.pi / nphi
    mat = gc.Material(cases.C_RHO, cases.C_CP, cases.C_K)
    dt = 2.0 * dr * dr / mat.alpha
    T0 = 20.0 + 980.0 * cases.splitmix_uniform(31, (nr, nphi, nz))
    zk = dict(kind_bot=kb, kind_top=kt, h_bot=180.0, h_top=500.0, T_inf_bot=35.0, T_inf_top=20.0, T_bot=250.0, T_top=60.0)
    ext = slab.split_z(nz, world)
    z0, z1 = ext[rank]
    grid = slab.SlabGridCyl(nr, nphi, z1 - z0, dr, dphi, dr, R, slab.TorchDistComm(), nz_per_rank=[b - a for a, b in ext])
    T = torch.from_numpy(np.ascontiguousarray(T0[:, :, z0:z1])).cuda()
    for _ in range(3):
        T = slab.adi_step_cyl(T, grid, mat, gc.Params(dt, 1.0, "be"), gc.RobinR(500.0, 20.0), gc.ZBC(**zk))
    parts = [None] * world
    dist.all_gather_object(parts, (z0, z1, T.cpu().numpy()))
    if rank == 0:
        out = np.empty((nr, nphi, nz))
        for a, b, tt in parts:
            out[:, :, a:b] = tt
        ref = T0
        og, om = cyl.GridCyl(nr, nphi, nz, dr, dphi, dr, R), cyl.Material(mat.rho, mat.cp, mat.k)
        for _ in range(3):
            ref = cyl.adi_step(ref, og, om, cyl.Params(dt, 1.0, "be"), cyl.RobinR(500.0, 20.0), cyl.ZBC(**zk))
        err = cases.rel_l2(out, ref)
        good = err <= 3e-12
        ok &= good
        print(f"[dist_check] world={world} cylindrical {nr}x{nphi}x{nz} zbc={kb}/{kt}: rel_l2={err:.2e} {'OK' if good else 'FAIL'}", flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
