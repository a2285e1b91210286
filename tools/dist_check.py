#!/usr/bin/env python
"""Multi-GPU parity check of the z-slab path over NCCL (run under torchrun on the GPU box):
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
      --master-port 29511 tools/dist_check.py
Every rank steps its slab of a seeded problem; rank 0 gathers the slabs and compares with the
oracle on the undivided grid (rel-L2 <= 1e-12 per step, void cells bit-identical)."""
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

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ok = True
for shape, mk, bk, theta, cfl, nsteps in [((24, 20, 16 * world * 2), "cyl_holes", "combined", 0.5, 2.0, 3),
                                           ((16, 40, 64 * world), "random", "robin_dict3d", 0.5, 3000.0, 2),
                                           ((32, 16, 16 * world), "plate_track", "robin6", 1.0, 0.128, 2)]:
    pb = make_problem(shape, mk, bk, theta, cfl)
    z0, z1, t, nl = rank_run(slab.TorchDistComm(), pb, nsteps, None)
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
        print(f"[dist_check] world={world} shape={shape} {mk}/{bk} theta={theta}: rel_l2={err:.2e} "
              f"void_bit_equal={void} launches/rank={parts[0][3]} {'OK' if good else 'FAIL'}", flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
