# -*- coding: utf-8 -*-
"""Cylindrical (r, phi, z) backward-Euler ADI heat step on B200 -- the reference's
`adi3d_cyl_phi_v3` interface (adi3d_cyl_phi_v3.py:33-68, 332-350) in front of libadi_b200.so.

Same names, argument meaning, ownership and errors as the reference module:
  GridCyl(nr,nphi,nz,dr,dphi,dz,R[,R_in])   :33   (R_in is accepted and ignored: the reference's
                                                   build_grid_annular passes it)
  Material(rho,cp,k) (+ .alpha)             :45
  Params(dt,theta=0.5,scheme="be")          :52
  RobinR(h,T_inf)                           :56
  ZBC(kind_bot,kind_top,h_*,T_inf_*,T_*)    :60
  adi_step(Tn,grid,mat,prm,robin_r,zbc,S=None,theta=None) -> ndarray(nr,nphi,nz)   :332
and the activation-mask wrapper of the deposition drivers
  adi_step_masked(Tn,grid,mat,prm,robin_outer,zbc,active,robin_inner=None,robin_void=None)
  build_grid_annular(R_out,wall_thickness,height,z_back,nr,nphi,dz_override=None) -> (grid,R_in,R_out,dz)
(quick_spiral_deposition_gif_v5.py:31-80).

adi_step / adi_step_masked take and return HOST NumPy arrays, as in the reference (H2D, three
sweeps, D2H inside the C ABI); adi_step_device runs the same step on device-resident arrays.
Only scheme "be" exists: the reference's "douglas" branch reads uninitialised memory
(np.empty_like, :149-151) and is not reproducible.
fp64 only.  No CPU fallback: without the built library and a CUDA device every call raises.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _capi
from . import devarray as cp

_KINDS = {"neumann0": 0, "dirichlet": 1, "robin": 2}


class GridCyl:
    def __init__(self, nr, nphi, nz, dr, dphi, dz, R, R_in=0.0):
        self.nr = int(nr); self.nphi = int(nphi); self.nz = int(nz)
        self.dr = float(dr); self.dphi = float(dphi); self.dz = float(dz)
        self.R = float(R)
        self.R_in = float(R_in)
        self.r = (np.arange(self.nr, dtype=np.float64) + 0.5) * self.dr
        self.r_imh = self.r - 0.5 * self.dr
        self.r_iph = self.r + 0.5 * self.dr
        self.r_outer_face = self.r_iph[-1]


class Material:
    def __init__(self, rho, cp, k):
        self.rho = float(rho); self.cp = float(cp); self.k = float(k)

    @property
    def alpha(self):
        return self.k / (self.rho * self.cp)


class Params:
    def __init__(self, dt, theta=0.5, scheme="be"):
        self.dt = float(dt); self.theta = float(theta); self.scheme = str(scheme).lower()


class RobinR:
    def __init__(self, h, T_inf):
        self.h = float(h); self.T_inf = float(T_inf)


class ZBC:
    def __init__(self, kind_bot='neumann0', kind_top='robin', h_bot=0.0, h_top=0.0,
                 T_inf_bot=20.0, T_inf_top=20.0, T_bot=20.0, T_top=20.0):
        self.kind_bot = kind_bot; self.kind_top = kind_top
        self.h_bot = float(h_bot); self.h_top = float(h_top)
        self.T_inf_bot = float(T_inf_bot); self.T_inf_top = float(T_inf_top)
        self.T_bot = float(T_bot); self.T_top = float(T_top)


class _Engine:
    def __init__(self):
        self.ctx = None
        self.bound = None

    def context(self):
        if self.ctx is None:
            if not torch.cuda.is_available():
                raise RuntimeError("adi3d_cyl_phi_v3: no CUDA device (there is no CPU fallback)")
            self.ctx = _capi.context(torch.cuda.current_device())
        return self.ctx

    def bind(self, grid, nz_pitch=None):
        pitch = grid.nz if nz_pitch is None else int(nz_pitch)
        key = (grid.nr, grid.nphi, grid.nz, pitch, grid.dr, grid.dphi, grid.dz)
        if key != self.bound:
            _capi.check(_capi.load().adi_cyl_bind(self.context(), grid.nr, grid.nphi, grid.nz, pitch,
                                                  grid.dr, grid.dphi, grid.dz), "adi_cyl_bind")
            self.bound = key


_engine = _Engine()


def _params(mat, prm, robin_r, zbc, T_void=0.0, T_inner=0.0):
    scheme = prm.scheme if prm.scheme in ("be", "douglas") else "be"  # :335
    if scheme != "be":
        raise NotImplementedError("adi3d_cyl_phi_v3 (B200): only scheme='be' is provided; the reference's "
                                  "'douglas' branch reads uninitialised memory and is not reproducible")
    if zbc.kind_bot not in _KINDS:
        raise ValueError("unknown zbc.kind_bot")  # :283
    if zbc.kind_top not in _KINDS:
        raise ValueError("unknown zbc.kind_top")  # :296
    p = _capi.CylParams()
    p.dt = float(prm.dt)
    p.rho, p.cp, p.k = mat.rho, mat.cp, mat.k
    p.h_r, p.Tinf_r = float(robin_r.h), float(robin_r.T_inf)
    p.kind_bot, p.kind_top = _KINDS[zbc.kind_bot], _KINDS[zbc.kind_top]
    p.h_bot, p.h_top = zbc.h_bot, zbc.h_top
    p.Tinf_bot, p.Tinf_top = zbc.T_inf_bot, zbc.T_inf_top
    p.T_bot, p.T_top = zbc.T_bot, zbc.T_top
    p.T_void, p.T_inner = float(T_void), float(T_inner)
    return p


def _stream_ptr():
    return torch.cuda.current_stream().cuda_stream


def _host_step(Tn, grid, p, active=None, S=None, nsteps=1):
    shape = (grid.nr, grid.nphi, grid.nz)
    T = np.ascontiguousarray(Tn, dtype=np.float64)
    if T.shape != shape:
        raise ValueError("field shape does not match the grid")
    act = None
    if active is not None:
        act = np.ascontiguousarray(active, dtype=np.bool_)
        if act.shape != shape:
            raise ValueError("active shape does not match the grid")
    src = None
    if S is not None:
        src = np.ascontiguousarray(np.broadcast_to(np.asarray(S, dtype=np.float64), shape))
    e = _engine
    e.bind(grid)
    out = cp.pinned_empty(shape, np.float64)     # page-locked: PCIe-speed download, and upload when it comes back as Tn
    _capi.check(_capi.load().adi_cyl_step_host(
        e.context(), T.ctypes.data, out.ctypes.data, int(nsteps), C.byref(p),
        None if act is None else act.ctypes.data, None if src is None else src.ctypes.data,
        _stream_ptr()), "adi_cyl_step_host")
    return out


def adi_step(Tn, grid, mat, prm, robin_r, zbc, S=None, theta=None):
    """One backward-Euler ADI step (adi3d_cyl_phi_v3.py:332-350): r-, phi- (periodic) and
    z-implicit solves.  Host NumPy in, new host NumPy array out; `Tn` is not modified."""
    return _host_step(Tn, grid, _params(mat, prm, robin_r, zbc), S=S)


def adi_step_masked(Tn, grid, mat, prm, robin_outer, zbc, active, robin_inner=None, robin_void=None):
    """quick_spiral_deposition_gif_v5.py:31-70: void cells are held at robin_void.T_inf before
    and after the step, inactive cells of the axis ring at robin_inner.T_inf (the `h` of
    robin_inner / robin_void is unused, as in the reference)."""
    if robin_inner is None:
        robin_inner = robin_outer
    if robin_void is None:
        robin_void = robin_outer
    p = _params(mat, prm, robin_outer, zbc, T_void=float(robin_void.T_inf), T_inner=float(robin_inner.T_inf))
    return _host_step(Tn, grid, p, active=np.asarray(active, dtype=bool))


def build_grid_annular(R_out, wall_thickness, height, z_back, nr, nphi, dz_override=None):
    """quick_spiral_deposition_gif_v5.py:74-80 -> (grid, R_in, R_out, dz)."""
    R_in = max(0.0, R_out - wall_thickness)
    dr = (R_out - R_in) / float(nr)
    dz = dr if (dz_override is None or dz_override <= 0.0) else float(dz_override)
    nz = int(round((z_back + height) / dz))
    dphi = (2.0 * np.pi) / max(1, nphi)
    return GridCyl(nr, nphi, nz, dr, dphi, dz, R_out, R_in=R_in), R_in, R_out, dz


def adi_step_device(Tn, grid, mat, prm, robin_r, zbc, S=None, active=None, robin_inner=None,
                    robin_void=None, out=None, nz_pitch=None):
    """The same step on DEVICE-resident arrays (`cupy`-shim ndarray or torch CUDA tensors):
    no host transfer.  `active` selects the masked variant.  With `nz_pitch` the arrays are
    (nr, nphi, nz_pitch) buffers of which the first grid.nz planes along z are stepped (layer
    births grow nz in place, quick_compare_layer_birth_robin_cyl_v3.py:195-204).
    Returns a device array of the same kind as `Tn`."""
    def tens(x, dtype):
        t = x._t if isinstance(x, cp.ndarray) else x
        if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == dtype and t.is_contiguous()):
            raise TypeError("adi_step_device: contiguous CUDA arrays of the right dtype expected")
        return t
    pitch = grid.nz if nz_pitch is None else int(nz_pitch)
    shape = (grid.nr, grid.nphi, pitch)
    T = tens(Tn, torch.float64)
    if tuple(T.shape) != shape:
        raise ValueError("field shape does not match the grid")
    if active is not None:
        ri = robin_inner or robin_r
        rv = robin_void or robin_r
        p = _params(mat, prm, robin_r, zbc, T_void=float(rv.T_inf), T_inner=float(ri.T_inf))
        a = tens(active, torch.bool)
        if tuple(a.shape) != shape:
            raise ValueError("active shape does not match the grid")
    else:
        p = _params(mat, prm, robin_r, zbc)
        a = None
    s = None if S is None else tens(S, torch.float64)
    if out is None:
        # cells beyond grid.nz (pitched buffers) keep the input values
        o = torch.empty_like(T) if pitch == grid.nz else T.clone()
    else:
        o = tens(out, torch.float64)
    e = _engine
    e.bind(grid, pitch)
    _capi.check(_capi.load().adi_cyl_step(e.context(), T.data_ptr(), o.data_ptr(), C.byref(p),
                                          None if a is None else a.data_ptr(),
                                          None if s is None else s.data_ptr(), _stream_ptr()), "adi_cyl_step")
    return cp.ndarray(o) if isinstance(Tn, cp.ndarray) else o


def launch_count():
    return int(_capi.load().adi_launch_count(_engine.context()))


def set_option(name, value):
    """Engine tuning knob of the cylindrical path (adi_set_option): 'cylsm', 'cylzt', 'm', 'kt', 'lt' ..."""
    _capi.check(_capi.load().adi_set_option(_engine.context(), str(name).encode(), int(value)), "adi_set_option")


def get_option(name):
    """Current value of an engine option (adi_get_option)."""
    return int(_capi.load().adi_get_option(_engine.context(), str(name).encode()))
