#!/usr/bin/env python
"""Developer probe (GPU box): per-sweep device times of the Cartesian step for an arbitrary grid.
  python tools/sweep_probe.py NX NY NZ [--theta 0.5] [--scalar] [--steps 10] [--opt m=32 ...]
Prints one line: ms per sweep (x|y|z), GB/s of each against its algorithmic bytes."""
import argparse
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from adi_thermal_fields_b200 import _capi, adi3d_gpu_coeff as g, devarray as cp  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("shape", type=int, nargs=3)
ap.add_argument("--theta", type=float, default=0.5)
ap.add_argument("--scalar", action="store_true")
ap.add_argument("--full", action="store_true", help="all-active mask instead of plate+track")
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--opt", action="append", default=[])
a = ap.parse_args()
nx, ny, nz = a.shape
dev = torch.device("cuda", 0)
mask = torch.ones((nx, ny, nz), dtype=torch.bool, device=dev)
if not a.full:
    nzp = nz - max(1, nz // 64)
    mask[:, :, nzp:] = False
    mask[: max(1, nx // 32), : ny // 2, nzp:] = True
T0 = 20.0 + 1380.0 * torch.rand((nx, ny, nz), dtype=torch.float64, device=dev)
grid = g.Grid3D.__new__(g.Grid3D)
grid.nx, grid.ny, grid.nz, grid.dx, grid.mask = nx, ny, nz, 1e-3, cp.ndarray(mask)
mat = g.Material(7800.0, 500.0, 25.0)
if a.scalar:
    h = {f: 10.0 for f in g.FACES}
else:
    h = {f: cp.ndarray(10.0 * (0.3 + torch.rand((nx, ny, nz), dtype=torch.float64, device=dev))) for f in g.FACES}
packs = g.precompute_coeff_packs_unified(grid, mat, robin_h=h)
del h
e = g._engine
e.bind(grid); e.set_mask(grid); e.set_packs(packs)
L, ctx = _capi.load(), e.context()
for o in a.opt:
    k, _, v = o.partition("=")
    L.adi_set_option(ctx, k.encode(), int(v))
A, B = T0.clone(), torch.empty_like(T0)
st = torch.cuda.current_stream().cuda_stream
kappa = 25.0 / (7800.0 * 500.0)


def step(s, d):
    _capi.check(L.adi_cart_step(ctx, s.data_ptr(), d.data_ptr(), 0.02, a.theta, kappa, 20.0, st), "step")


for _ in range(3):
    step(A, B); A, B = B, A
torch.cuda.synchronize()
L.adi_set_option(ctx, b"profile", 1)
L.adi_profile_reset(ctx)
for _ in range(a.steps):
    step(A, B); A, B = B, A
torch.cuda.synchronize()
ms = (C.c_double * 4)()
n = C.c_long()
L.adi_profile_read(ctx, ms, C.byref(n))
cells = nx * ny * nz
bpc = 17 if a.scalar else 25
per = [ms[i] / n.value for i in range(4)]
print(f"shape {nx}x{ny}x{nz} theta {a.theta} {'scalar' if a.scalar else 'dense'} opts {a.opt}: "
      + f"expl {per[0]:.3f} ms {17 * cells / max(per[0], 1e-9) / 1e6:.0f} GB/s  "
      + "  ".join(f"{ax} {t:.3f} ms {bpc * cells / t / 1e6:.0f} GB/s" for ax, t in zip("xyz", per[1:]))
      + f"  | step {sum(per):.3f} ms")
