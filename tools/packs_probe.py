# -*- coding: utf-8 -*-
"""B200 probe: precompute_coeff_packs_unified (k_build_packs_v) at n^3 with dense per-face h fields on a half-built
part, against the time a plain memset of the three coefficient fields takes.  usage: python tools/packs_probe.py [n]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from adi_thermal_fields_b200 import adi3d_gpu_coeff as g, devarray as cp   # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 1024
dev = torch.device("cuda", 0)
ax = (torch.arange(n, device=dev, dtype=torch.float64) + 0.5) / n - 0.5
X, Y, Z = ax[:, None, None], ax[None, :, None], ax[None, None, :]
mask = (((X / 0.35) ** 2 + (Y / 0.42) ** 2 + ((Z - 0.05) / 0.45) ** 2 <= 1.0) | ((X * X + Y * Y <= 0.12 ** 2) & (Z < -0.3))) & (Z < 0.0)
h = {f: cp.ndarray(40.0 * (0.3 + torch.rand((n, n, n), dtype=torch.float64, device=dev))) for f in g.FACES}
grid = g.Grid3D(n, n, n, 1e-3, cp.ndarray(mask))
mat = g.Material(7800.0, 490.0, 54.0)


def timed(fn, reps=4):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


bufs = [torch.empty((n, n, n), dtype=torch.float64, device=dev) for _ in range(3)]
gb = 3 * 8 * n ** 3 / 1e9
t = timed(lambda: [b.zero_() for b in bufs])
print(f"memset of 3 fields: {t:.3f} ms = {gb / t:.2f} TB/s")
del bufs
ref = None
one = "--one" in sys.argv
for pkm in ((0,) if one else (0, 0, 1, 4)):
    for pkb in ((0,) if one else (0, 64, 2048)):
        g.set_option("pkm", pkm)
        g.set_option("pkb", pkb)
        keep = []
        t = timed(lambda: keep.append(g.precompute_coeff_packs_unified(grid, mat, robin_h=h)) or keep.__delitem__(slice(0, -1)))
        out = keep[-1][2].coeff._t
        if ref is None:
            ref = out.clone()
        print(f"pkm {pkm} pkb {pkb}: {t:.3f} ms = {gb / t:.2f} TB/s  same {bool(torch.equal(out.view(torch.int64), ref.view(torch.int64)))}")
        del keep, out
if one:
    sys.exit(0)
g.set_option("maskv", 0)
t = timed(lambda: g.precompute_coeff_packs_unified(grid, mat, robin_h=h))
print(f"cell form: {t:.3f} ms = {gb / t:.2f} TB/s")
