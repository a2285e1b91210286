#!/usr/bin/env python
"""Developer probe (N-GPU box, under torchrun): cylindrical configs[2] grid (256 x 1024 x 512) z-slab sharded over the
ranks (strong scaling of adi3d_cyl_phi_v3.adi_step, scheme 'be'), device-resident, CUDA events, max over ranks.
  python -m torch.distributed.run --nproc-per-node N tools/cyl_slab_probe.py [--steps 20] [--nz 512]"""
import argparse
import json
import math
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from adi_thermal_fields_b200 import adi3d_cyl_phi_v3 as gc, slab  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--nr", type=int, default=256)
ap.add_argument("--nphi", type=int, default=1024)
ap.add_argument("--nz", type=int, default=512)
a = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
comm = slab.TorchDistComm() if world > 1 else slab.LocalComm(1).view(0)
R = 0.02
dr = R / a.nr
dphi = 2.0 * math.pi / a.nphi
mat = gc.Material(7800.0, 490.0, 54.0)
dt = min(dr * dr, (R * dphi) ** 2) / mat.alpha
ext = slab.split_z(a.nz, world)
z0, z1 = ext[rank]
grid = slab.SlabGridCyl(a.nr, a.nphi, z1 - z0, dr, dphi, dr, R, comm, nz_per_rank=[e - b for b, e in ext])
gen = torch.Generator(device="cuda").manual_seed(2 + rank)
A = 20.0 + 5.0 * torch.rand((a.nr, a.nphi, z1 - z0), dtype=torch.float64, device="cuda", generator=gen)
B = torch.empty_like(A)
prm, rob, zbc = gc.Params(dt, 1.0, "be"), gc.RobinR(500.0, 20.0), gc.ZBC("neumann0", "robin", h_top=500.0, T_inf_top=20.0)
for _ in range(3):
    slab.adi_step_cyl(A, grid, mat, prm, rob, zbc, out=B); A, B = B, A
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.steps):
    slab.adi_step_cyl(A, grid, mat, prm, rob, zbc, out=B); A, B = B, A
e1.record()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / a.steps], dtype=torch.float64, device="cuda")
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    cells = a.nr * a.nphi * a.nz
    print(json.dumps({"workload": f"cylindrical {a.nr}x{a.nphi}x{a.nz} z-slab x{world} (strong scaling)", "n_gpus": world,
                      "ms_per_step": float(ms.item()), "cell_steps_per_s": cells / (float(ms.item()) * 1e-3),
                      "z_form": "two-pass (pass 1, all-gather of 2 doubles per line and rank, pass 2)" if world > 1 else "single GPU"}))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
