#!/usr/bin/env python
"""Developer probe (GPU box): per-sweep device times of the cylindrical BE step.
  python tools/cyl_probe.py NR NPHI NZ [--masked] [--steps 10] [--opt kt=8 ...]
Prints ms per sweep (r|phi|z) and GB/s of each against 16 B/cell (8 read + 8 written)."""
import argparse
import ctypes as C
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from adi_thermal_fields_b200 import _capi, adi3d_cyl_phi_v3 as gc  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("shape", type=int, nargs=3)
ap.add_argument("--masked", action="store_true")
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--opt", action="append", default=[])
a = ap.parse_args()
nr, nphi, nz = a.shape
dev = torch.device("cuda", 0)
R = 0.02
dr = R / nr
dz = dr
dphi = 2 * math.pi / nphi
mat = gc.Material(7800.0, 490.0, 54.0)
dt = min(dr * dr, dz * dz, (R * dphi) ** 2) / mat.alpha
grid = gc.GridCyl(nr, nphi, nz, dr, dphi, dz, R)
rob, zbc = gc.RobinR(500.0, 20.0), gc.ZBC("neumann0", "robin", h_top=500.0, T_inf_top=20.0)
prm = gc.Params(dt, 1.0, "be")
T = 20.0 + 980.0 * torch.rand((nr, nphi, nz), dtype=torch.float64, device=dev)
act = (torch.rand((nr, nphi, nz), device=dev) < 0.8) if a.masked else None
L = _capi.load()
ctx = gc._engine.context()
for o in a.opt:
    k, _, v = o.partition("=")
    L.adi_set_option(ctx, k.encode(), int(v))
A, B = T, torch.empty_like(T)
for _ in range(3):
    gc.adi_step_device(A, grid, mat, prm, rob, zbc, active=act, out=B); A, B = B, A
torch.cuda.synchronize()
L.adi_set_option(ctx, b"profile", 1)
L.adi_profile_reset(ctx)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.steps):
    gc.adi_step_device(A, grid, mat, prm, rob, zbc, active=act, out=B); A, B = B, A
e1.record()
torch.cuda.synchronize()
ms = (C.c_double * 4)()
n = C.c_long()
L.adi_profile_read(ctx, ms, C.byref(n))
cells = nr * nphi * nz
per = [ms[i] / n.value for i in range(4)]
tot = e0.elapsed_time(e1) / a.steps
print(f"cyl {nr}x{nphi}x{nz} masked={a.masked} opts {a.opt}: "
      + "  ".join(f"{ax} {t:.3f} ms {16 * cells / max(t, 1e-9) / 1e6:.0f} GB/s" for ax, t in zip(("r", "phi", "z"), per[1:]))
      + f"  | step {tot:.3f} ms  {cells / tot / 1e6:.2f} Gcell-steps/s  {48 * cells / tot / 1e6:.0f} GB/s")
