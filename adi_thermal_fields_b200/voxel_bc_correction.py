# -*- coding: utf-8 -*-
"""STL geometry -> corrected voxel Robin coefficients on B200 -- the reference's
`voxel_bc_correction` interface (voxel_bc_correction.py:33-226) in front of libadi_b200.so.

  STLBoundaryCorrector(mesh, mask, origin, dx, max_subdiv=6, area_epsilon=1e-16)      :33
      .compute_voxel_projected_areas()  -> dict face -> device array of projected area    :53
            (the reference returns a dict voxel -> per-face areas; the dense per-face form is what
             build_corrected_fields needs and what stays on the device)
      .build_corrected_fields(base_h, fallback_to_base=True) -> (robin_h fields, scale fields)  :110
  build_corrected_robin_fields(mesh, mask, origin, dx, base_h, fallback_to_base=True, max_subdiv=6)   :207

`mesh` is duck-typed like in the reference (triangles, face_normals, area_faces).  The returned
fields are `cupy`-shim device arrays keyed by face name, ready for
precompute_coeff_packs_unified(robin_h=...).  The scatter uses fp64 atomics: sums into one voxel
agree with the reference to rounding (<= 1e-15 relative), voxel membership is identical.
No CPU fallback."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _capi
from . import devarray as cp

FACES = ("x-", "x+", "y-", "y+", "z-", "z+")


def _ctx():
    if not torch.cuda.is_available():
        raise RuntimeError("voxel_bc_correction: no CUDA device (there is no CPU fallback)")
    return _capi.context(torch.cuda.current_device())


def _st():
    return torch.cuda.current_stream().cuda_stream


class STLBoundaryCorrector:
    def __init__(self, mesh, mask, origin, dx, max_subdiv=6, area_epsilon=1e-16):
        self.mesh = mesh
        self.mask = cp.asarray(mask, dtype=cp.bool_)
        self.origin = np.asarray(origin, dtype=float)
        self.dx = float(dx)
        self.shape = self.mask.shape
        self.max_subdiv = max(1, int(max_subdiv))
        self.area_epsilon = float(area_epsilon)

    def compute_voxel_projected_areas(self):
        L, ctx = _capi.load(), _ctx()
        dev = self.mask._t.device
        tri = torch.from_numpy(np.ascontiguousarray(np.asarray(self.mesh.triangles, dtype=np.float64))).to(dev)
        nrm = torch.from_numpy(np.ascontiguousarray(np.asarray(self.mesh.face_normals, dtype=np.float64))).to(dev)
        area = torch.from_numpy(np.ascontiguousarray(np.asarray(self.mesh.area_faces, dtype=np.float64))).to(dev)
        ntri = int(tri.shape[0])
        proj = [torch.zeros(self.shape, dtype=torch.float64, device=dev) for _ in FACES]
        nx, ny, nz = self.shape
        _capi.check(L.adi_voxel_project(ctx, tri.data_ptr(), nrm.data_ptr(), area.data_ptr(), ntri,
                                        (C.c_double * 3)(*[float(v) for v in self.origin]), self.dx, self.max_subdiv,
                                        self.area_epsilon, self.mask._t.data_ptr(), nx, ny, nz,
                                        (C.c_void_p * 6)(*[p.data_ptr() for p in proj]), _st()), "adi_voxel_project")
        torch.cuda.current_stream().synchronize()   # tri / nrm / area are released on return
        return {f: cp.ndarray(p) for f, p in zip(FACES, proj)}

    def build_corrected_fields(self, base_h, fallback_to_base=True):
        for f in base_h:
            if f not in FACES:
                raise ValueError("bad face")
        L, ctx = _capi.load(), _ctx()
        proj = self.compute_voxel_projected_areas()
        dev = self.mask._t.device
        has = [1 if f in base_h else 0 for f in FACES]
        base = [float(base_h.get(f, 0.0)) for f in FACES]
        robin = [torch.empty(self.shape, dtype=torch.float64, device=dev) if h else None for h in has]
        scale = [torch.empty(self.shape, dtype=torch.float64, device=dev) if h else None for h in has]
        nx, ny, nz = self.shape
        vp = C.c_void_p
        _capi.check(L.adi_voxel_correct(ctx, self.mask._t.data_ptr(), nx, ny, nz, self.dx,
                                        (vp * 6)(*[proj[f]._t.data_ptr() for f in FACES]), (C.c_int * 6)(*has),
                                        (C.c_double * 6)(*base), 1 if fallback_to_base else 0,
                                        (vp * 6)(*[None if t is None else t.data_ptr() for t in robin]),
                                        (vp * 6)(*[None if t is None else t.data_ptr() for t in scale]), _st()),
                    "adi_voxel_correct")
        torch.cuda.current_stream().synchronize()
        robin_fields = {f: cp.ndarray(robin[i]) for i, f in enumerate(FACES) if f in base_h}
        scale_fields = {f: cp.ndarray(scale[i]) for i, f in enumerate(FACES) if f in base_h}
        # dict order of the reference: the order of base_h
        return ({f: robin_fields[f] for f in base_h}, {f: scale_fields[f] for f in base_h})


def build_corrected_robin_fields(mesh, mask, origin, dx, base_h, fallback_to_base=True, max_subdiv=6):
    """Helper wrapper (voxel_bc_correction.py:207-226)."""
    corrector = STLBoundaryCorrector(mesh=mesh, mask=mask, origin=origin, dx=dx, max_subdiv=max_subdiv)
    return corrector.build_corrected_fields(base_h=base_h, fallback_to_base=fallback_to_base)
