#!/usr/bin/env python
"""Developer probe (1 GPU): two virtual z-slab ranks (threads) of a 512 x 512 x 512 plate on cuda:0, a few steps --
exercises k_pack_zplanes, k_explicit with halo planes, the two z-sweep passes and k_iface_solve; followed by one
2048-cell-line sweep (cluster kernel).  Meant to be run under ncu for profiles/."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from adi_thermal_fields_b200 import adi3d_gpu_coeff as g, devarray as cp, slab  # noqa: E402

N, half = 512, 256
dev = torch.device("cuda", 0)
mask = torch.ones((N, N, N), dtype=torch.bool, device=dev)
mask[:, :, N - 8:] = False
T0 = 20.0 + 1380.0 * torch.rand((N, N, N), dtype=torch.float64, device=dev)
h = 10.0 * (0.3 + torch.rand((N, N, N), dtype=torch.float64, device=dev))


class Mat:
    rho, cp, k = 7800.0, 500.0, 25.0


class Prm:
    dt, theta = 0.02, 0.5


def rank_fn(v):
    z0, z1 = v.rank * half, (v.rank + 1) * half
    grid = slab.SlabGrid3D(N, N, half, 1e-3, mask[:, :, z0:z1].contiguous(), v)
    hs = h[:, :, z0:z1].contiguous()
    packs = slab.precompute_coeff_packs_unified(grid, Mat, robin_h={f: hs for f in slab.FACES})
    A, B = T0[:, :, z0:z1].contiguous(), torch.empty((N, N, half), dtype=torch.float64, device=dev)
    for _ in range(4):
        slab.adi_step_gpu_coeff(A, grid, Mat, Prm, packs, Tinf=20.0, out=B)
        A, B = B, A
    torch.cuda.synchronize()
    return True


slab.LocalComm(2).run(rank_fn)
del mask, T0, h
torch.cuda.empty_cache()
# long lines: 2048 x 64 x 256, scalar Robin (cluster kernel on the x sweep)
nx, ny, nz = 2048, 64, 256
m2 = torch.ones((nx, ny, nz), dtype=torch.bool, device=dev)
grid = g.Grid3D.__new__(g.Grid3D)
grid.nx, grid.ny, grid.nz, grid.dx, grid.mask = nx, ny, nz, 1e-3, cp.ndarray(m2)
mat = g.Material(7800.0, 500.0, 25.0)
packs = g.precompute_coeff_packs_unified(grid, mat, robin_h={f: 10.0 for f in g.FACES})
T = cp.ndarray(20.0 + torch.rand((nx, ny, nz), dtype=torch.float64, device=dev))
for _ in range(3):
    T = g.adi_step_gpu_coeff(T, grid, mat, g.Params(0.02, 0.5), packs, Tinf=20.0)
torch.cuda.synchronize()
print("slab_probe done")
