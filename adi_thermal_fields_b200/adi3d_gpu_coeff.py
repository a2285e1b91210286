# -*- coding: utf-8 -*-
"""Cartesian ADI theta-scheme heat step on B200 -- the reference's `adi3d_gpu_coeff`
interface (adi3d_gpu_coeff.py:6-230) in front of libadi_b200.so.

Same names, argument meaning and ownership as the reference module:
  Grid3D(nx,ny,nz,dx,mask)            :6     mask is copied to the device
  Material(rho,cp_,k)                 :14
  Params(dt,theta=0.5)                :18
  AxisCoeffPack(coeff,dir_mask,dir_val,qflux=None)   :22
  precompute_coeff_packs_unified(...) :50    -> (packx, packy, packz), built on the device
  adi_step_gpu_coeff(Tn,grid,mat,params,packs,Tinf=0.0) :213  -> new device array

Arrays are `cupy`-shim device arrays (adi_thermal_fields_b200.devarray.ndarray); host
NumPy arrays are accepted wherever the reference calls cp.asarray on its inputs.
fp64 only.  No CPU / CuPy / Numba fallback: without the built library and a CUDA device
every call raises.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _capi
from . import devarray as cp

FACES = ("x-", "x+", "y-", "y+", "z-", "z+")


class Grid3D:
    def __init__(self, nx, ny, nz, dx, mask):
        self.nx, self.ny, self.nz = int(nx), int(ny), int(nz)
        self.dx = float(dx)
        # mask goes to the device straight away (adi3d_gpu_coeff.py:11); copy, like the CPU class
        self.mask = cp.array(mask, dtype=cp.bool_)
        assert self.mask.shape == (self.nx, self.ny, self.nz)


class Material:
    def __init__(self, rho, cp_, k):
        self.rho = float(rho); self.cp = float(cp_); self.k = float(k)


class Params:
    def __init__(self, dt, theta=0.5):
        self.dt = float(dt); self.theta = float(theta)


class AxisCoeffPack:
    """Per-axis operand bundle (adi3d_gpu_coeff.py:22-29).

    Packs made by precompute_coeff_packs_unified may leave operands symbolic
    (all-zero qflux, no Dirichlet cell, scalar Robin): the dense arrays the reference
    would hold are then materialised only if somebody reads the attribute
    (e.g. `packs[2].qflux`, quick_compare_neumann_robin.py:104)."""

    def __init__(self, coeff, dir_mask, dir_val, qflux=None):
        self._shape = tuple(coeff.shape)
        self._coeff = cp.asarray(coeff, dtype=cp.float64)
        self._dir_mask = cp.asarray(dir_mask, dtype=cp.bool_)
        self._dir_val = cp.asarray(dir_val, dtype=cp.float64)
        self._qflux = None if qflux is None else cp.asarray(qflux, dtype=cp.float64)
        for name, arr in (("dir_mask", self._dir_mask), ("dir_val", self._dir_val), ("qflux", self._qflux)):
            if arr is not None and tuple(arr.shape) != self._shape:
                # the kernels index every operand with the grid's strides: a mis-shaped array would be read out of bounds
                raise ValueError(f"AxisCoeffPack: {name} has shape {tuple(arr.shape)}, coeff has {self._shape}")
        self._face_coeff = None     # (lo, hi) scalars when coeff is symbolic
        self._scalar_src = None     # (engine build args) to materialise coeff lazily
        self._dir_any = None        # cached "dir_mask has a True" keyed by tensor version
        self._q_zero = qflux is None

    @classmethod
    def _symbolic(cls, shape, coeff, dir_mask, dir_val, qflux, face_coeff=None, builder=None):
        self = cls.__new__(cls)
        self._shape = tuple(shape)
        self._coeff, self._dir_mask, self._dir_val, self._qflux = coeff, dir_mask, dir_val, qflux
        self._face_coeff = face_coeff
        self._scalar_src = builder
        self._dir_any = None
        self._q_zero = qflux is None
        self._mask_epoch = None     # scalar Robin: the engine's mask state the symbolic coefficients stand for
        return self

    # dense views, as the reference exposes them
    @property
    def coeff(self):
        if self._coeff is None:
            self._coeff = self._scalar_src() if self._scalar_src else cp.zeros(self._shape, cp.float64)
        return self._coeff

    @property
    def dir_mask(self):
        if self._dir_mask is None:
            self._dir_mask = cp.zeros(self._shape, cp.bool_)
        return self._dir_mask

    @property
    def dir_val(self):
        if self._dir_val is None:
            self._dir_val = cp.zeros(self._shape, cp.float64)
        return self._dir_val

    @property
    def qflux(self):
        if self._qflux is None:
            self._qflux = cp.zeros(self._shape, cp.float64)
            self._q_zero = False  # somebody holds it now and may write to it
        return self._qflux

    # what the engine binds
    def _has_dirichlet(self):
        if self._dir_mask is None:
            return False
        # the three packs of one precompute share ONE dir_mask array (adi3d_gpu_coeff.py:108-110): the answer is kept
        # on that array, so a pack rebuild costs one reduction + read-back instead of three
        m = self._dir_mask
        t = m._t
        key = (t.data_ptr(), t._version)
        cached = getattr(m, "_any_true", None)
        if cached is None or cached[0] != key:
            cached = (key, bool(t.any().item()))
            m._any_true = cached
        self._dir_any = cached
        return cached[1]


class _Engine:
    """Keeps the C context in step with the Python objects the driver mutates
    (grid.mask rebinding / in-place edits, params.dt, pack rebuilds)."""

    def __init__(self):
        self.ctx = None
        self.bound = None
        self.mask_key = None
        self.mask_hold = None
        self.pack_keys = [None, None, None]
        self.scalar_key = None
        self.hold = []
        self.trust = 0
        self.mask_epoch = 0          # counts adi_cart_set_mask calls: identifies the bound mask state

    def context(self):
        if self.ctx is None:
            if not torch.cuda.is_available():
                raise RuntimeError("adi3d_gpu_coeff: no CUDA device (there is no CPU fallback)")
            self.ctx = _capi.context(torch.cuda.current_device())
        return self.ctx

    def lib(self):
        return _capi.load()

    def bind(self, grid):
        key = (grid.nx, grid.ny, grid.nz, grid.dx)
        if self.bound != key:
            _capi.check(self.lib().adi_cart_bind(self.context(), grid.nx, grid.ny, grid.nz, grid.dx),
                        "adi_cart_bind")
            self.bound = key
            self.mask_key = None
            self.pack_keys = [None, None, None]
            self.scalar_key = None

    def set_mask(self, grid):
        m = grid.mask
        if not isinstance(m, cp.ndarray):
            # drivers rebind grid.mask to their live NumPy array (waam_from_stl_v7_mm.py:494-495)
            m = cp.asarray(np.asarray(m), dtype=cp.bool_)
            key = None
        else:
            if m.dtype != np.bool_ or not m._t.is_contiguous():
                m = cp.asarray(m, dtype=cp.bool_)
            key = (m._t.data_ptr(), m._t._version)
        if m.shape != (grid.nx, grid.ny, grid.nz):
            raise AssertionError("mask shape does not match the grid")
        if key is None and self.mask_hold is not None and self.mask_key is None \
                and self.mask_hold.shape == m.shape and bool(torch.equal(m._t, self.mask_hold._t)):
            return   # a host mask with the same content as the one bound: nothing to rebuild
        if key is None or key != self.mask_key:
            _capi.check(self.lib().adi_cart_set_mask(self.context(), m._t.data_ptr()), "adi_cart_set_mask")
            self.mask_key = key
            self.mask_hold = m
            self.mask_epoch += 1
            self.trust = -1          # the engine dropped its trust bits: set_packs re-asserts them

    def set_packs(self, packs):
        """Bind the three packs.  The arrays are kept alive in self.hold until replaced, so a
        (pointer, version) key can never be reused by a different tensor while it is cached."""
        L, ctx = self.lib(), self.context()
        scalar = all(p._coeff is None and p._face_coeff is not None for p in packs)
        if scalar:
            # Symbolic scalar-Robin packs stand for coefficients frozen from the mask they were precomputed for
            # (adi3d_gpu_coeff.py:86-92); the kernels derive them from the mask bound NOW.  When the mask has
            # moved on since, freeze them for real: the dense arrays are built from a snapshot the pack keeps.
            for p in packs:
                ep = getattr(p, "_mask_epoch", None)
                if ep is not None and ep != (self.mask_epoch, id(self)):
                    raise RuntimeError("adi_step_gpu_coeff: these scalar-Robin packs were precomputed for an earlier "
                                       "state of grid.mask; call precompute_coeff_packs_unified again after changing "
                                       "the mask (the reference would keep stepping with the stale coefficients)")
        hold = []
        touched = False
        for a, p in enumerate(packs):
            if p._shape != self.bound[:3]:
                raise ValueError("pack shape does not match the grid")
            coeff = None if (scalar or (p._coeff is None and p._face_coeff is None)) else p.coeff
            dirm = p._dir_mask if p._has_dirichlet() else None
            dirv = p.dir_val if dirm is not None else None
            q = None if p._q_zero else p._qflux
            arrs = (coeff, dirm, dirv, q)
            key = tuple(None if x is None else (x._t.data_ptr(), x._t._version) for x in arrs)
            hold.append(arrs)
            if key != self.pack_keys[a] or (not scalar and self.scalar_key is not None):
                ptr = [None if x is None else x._t.data_ptr() for x in arrs]
                _capi.check(L.adi_cart_set_pack(ctx, a, ptr[0], ptr[1], ptr[2], ptr[3]), "adi_cart_set_pack")
                self.pack_keys[a] = key
                touched = True
        # packs straight from the device builder, for the mask that is bound now and untouched since, are
        # surface-only by construction: the engine may skip its examination pass (option sparse_trust)
        trust = 0
        if not scalar:
            for a, p in enumerate(packs):
                b = getattr(p, "_built", None)
                c = hold[a][0]
                if b is not None and c is not None and b == (c._t.data_ptr(), c._t._version, self.mask_epoch):
                    trust |= 1 << a
        if touched or trust != self.trust:
            # set_pack clears the engine's trust bits, so they are re-asserted after every rebind
            _capi.check(L.adi_set_option(ctx, b"sparse_trust", trust), "adi_set_option")
            self.trust = trust
        if scalar:
            fc = tuple(float(v) for p in packs for v in p._face_coeff)
            if touched or fc != self.scalar_key:
                _capi.check(L.adi_cart_set_robin_scalar(ctx, (C.c_double * 6)(*fc)), "adi_cart_set_robin_scalar")
                self.scalar_key = fc
        else:
            self.scalar_key = None
        self.hold = hold


_engine = _Engine()


def _stream_ptr():
    return torch.cuda.current_stream().cuda_stream


def _exposed_mask(mask, face):
    """exposed_mask(mask, face)  adi3d_gpu_coeff.py:31-48."""
    if face not in FACES:
        raise ValueError("bad face")
    m = cp.asarray(mask, dtype=cp.bool_)
    nx, ny, nz = m.shape
    e = _engine
    saved = e.bound
    L, ctx = e.lib(), e.context()
    _capi.check(L.adi_cart_bind(ctx, nx, ny, nz, 1.0), "adi_cart_bind")
    e.bound = None
    _capi.check(L.adi_cart_set_mask(ctx, m._t.data_ptr()), "adi_cart_set_mask")
    out = cp.empty(m.shape, cp.bool_)
    _capi.check(L.adi_cart_exposed_mask(ctx, FACES.index(face), out._t.data_ptr(), _stream_ptr()),
                "adi_cart_exposed_mask")
    torch.cuda.current_stream().synchronize()
    del saved
    return out


exposed_mask = _exposed_mask


def precompute_coeff_packs_unified(grid, mat, dir_mask=None, dir_value=None, neumann=None,
                                   robin_h=None, robin_Tinf=None):
    """adi3d_gpu_coeff.py:50-110 on the device (kernel k_build_packs).
    Dirichlet: dir_mask (bool 3-D), dir_value (scalar or 3-D).  Neumann: dict face -> q''
    (W/m^2, >0 heats the solid) on exposed cells of that face.  Robin: scalar / 3-D / dict
    face -> (scalar | 3-D); the ambient is the step's Tinf (`robin_Tinf` is unused, as in
    the reference)."""
    nx, ny, nz = grid.nx, grid.ny, grid.nz
    shape = (nx, ny, nz)
    e = _engine
    e.bind(grid)
    e.set_mask(grid)
    L, ctx = e.lib(), e.context()

    dm = None if dir_mask is None else cp.asarray(dir_mask, dtype=cp.bool_)
    if dir_value is None:
        dv = None
    elif np.isscalar(dir_value):
        dv = cp.full(shape, float(dir_value), dtype=cp.float64)
    else:
        dv = cp.asarray(dir_value, dtype=cp.float64)
    if dm is not None and dv is None:
        dv = cp.zeros(shape, cp.float64)
    for name, arr in (("dir_mask", dm), ("dir_value", dv)):
        if arr is not None and tuple(arr.shape) != shape:
            raise ValueError(f"{name} has shape {tuple(arr.shape)}, the grid is {shape}")
    # one Dirichlet mask / value array shared by the three packs, as in the reference (adi3d_gpu_coeff.py:73-78,108-110)
    if dm is None and dv is not None:
        dm = cp.zeros(shape, cp.bool_)

    def classify(v):
        if v is None:
            return 0, 0.0, None
        if np.isscalar(v):
            return 1, float(v), None
        a = cp.asarray(v, dtype=cp.float64)
        if a.shape != shape:
            raise ValueError("field shape does not match the grid")
        return 2, 0.0, a

    hk, hs, hf = [0] * 6, [0.0] * 6, [None] * 6
    if robin_h is not None:
        for i, f in enumerate(FACES):
            v = robin_h.get(f, 0.0) if isinstance(robin_h, dict) else robin_h
            hk[i], hs[i], hf[i] = classify(v)
    qk, qs, qf = [0] * 6, [0.0] * 6, [None] * 6
    if neumann is not None:
        for f, v in neumann.items():
            if f not in FACES:
                raise ValueError("bad face")
            i = FACES.index(f)
            qk[i], qs[i], qf[i] = classify(v)

    dense_h = any(k == 2 for k in hk)
    have_q = any(k != 0 for k in qk)
    A = grid.dx * grid.dx
    V = grid.dx ** 3
    Ccell = mat.rho * mat.cp * V
    face_coeff = [(hs[i] * A / Ccell) if hk[i] == 1 else 0.0 for i in range(6)]

    def run_build(want_coeff, want_q):
        qk_use = qk if want_q else [0] * 6
        coeffs = [cp.empty(shape, cp.float64) if want_coeff else None for _ in range(3)]
        qs_out = [cp.empty(shape, cp.float64) if (want_q and (qk[2 * a] or qk[2 * a + 1])) else None
                  for a in range(3)]
        vp = C.c_void_p
        hfp = (vp * 6)(*[None if a is None else a._t.data_ptr() for a in hf])
        qfp = (vp * 6)(*[None if a is None else a._t.data_ptr() for a in qf])
        ptr = [None if a is None else a._t.data_ptr() for a in coeffs + qs_out]
        _capi.check(L.adi_cart_build_packs(ctx, mat.rho, mat.cp, (C.c_int * 6)(*hk), (C.c_double * 6)(*hs),
                                           hfp, (C.c_int * 6)(*qk_use), (C.c_double * 6)(*qs), qfp,
                                           *ptr, _stream_ptr()), "adi_cart_build_packs")
        return coeffs, qs_out

    coeffs, qouts = run_build(dense_h, have_q) if (dense_h or have_q) else ([None] * 3, [None] * 3)

    packs = []
    for a in range(3):
        builder = None
        fc = None
        if coeffs[a] is None:
            # scalar Robin: the kernel derives h*A/Ccell from the mask; a dense array is only
            # produced if somebody reads pack.coeff (from the mask bound at that time)
            fc = (face_coeff[2 * a], face_coeff[2 * a + 1])

            def builder(a=a):
                e.bind(grid)
                e.set_mask(grid)
                return run_build(True, False)[0][a]
        pk = AxisCoeffPack._symbolic(shape, coeffs[a], dm, dv, qouts[a], face_coeff=fc, builder=builder)
        if fc is not None:
            pk._mask_epoch = (e.mask_epoch, id(e))
        if coeffs[a] is not None:
            pk._built = (coeffs[a]._t.data_ptr(), coeffs[a]._t._version, e.mask_epoch)
        packs.append(pk)
    return tuple(packs)


def adi_step_gpu_coeff(Tn, grid, mat, params, packs, Tinf=0.0):
    """One ADI theta-step (adi3d_gpu_coeff.py:213-230): explicit part + x, y, z implicit
    sweeps; cells outside the mask are returned unchanged.  Returns a new device array;
    `Tn` is not modified."""
    theta = params.theta; dt = params.dt
    kappa = mat.k / (mat.rho * mat.cp)
    T = Tn if isinstance(Tn, cp.ndarray) else cp.asarray(Tn, dtype=cp.float64)
    if T.dtype != np.float64:
        # the reference promotes a float32 field on its first step (waam_from_stl_v7_mm.py --precision float32
        # builds T with cp.full(..., dtype=float32)); the arithmetic and the result are float64
        T = cp.asarray(T, dtype=cp.float64)
    if not T._t.is_contiguous():
        T = cp.asarray(T)
    if T.shape != (grid.nx, grid.ny, grid.nz):
        raise ValueError("field shape does not match the grid")
    e = _engine
    e.bind(grid)
    e.set_mask(grid)
    e.set_packs(packs)
    out = cp.empty(T.shape, cp.float64)
    _capi.check(e.lib().adi_cart_step(e.context(), T._t.data_ptr(), out._t.data_ptr(), dt, theta, kappa,
                                      float(Tinf), _stream_ptr()), "adi_cart_step")
    return out


def adi_step_host(Tn, grid, mat, params, packs, Tinf=0.0, nsteps=1):
    """The same step for HOST NumPy fields (the CPU module's calling convention,
    adi3d_numba_coeff.py:290): H2D, `nsteps` steps, D2H.  Returns a NumPy array."""
    T = np.ascontiguousarray(Tn, dtype=np.float64)
    if T.shape != (grid.nx, grid.ny, grid.nz):
        raise ValueError("field shape does not match the grid")
    kappa = mat.k / (mat.rho * mat.cp)
    e = _engine
    e.bind(grid)
    e.set_mask(grid)
    e.set_packs(packs)
    out = cp.pinned_empty(T.shape, np.float64)   # page-locked: PCIe-speed download, and upload when it comes back as Tn
    _capi.check(e.lib().adi_cart_step_host(e.context(), T.ctypes.data, out.ctypes.data, int(nsteps),
                                           params.dt, params.theta, kappa, float(Tinf), _stream_ptr()),
                "adi_cart_step_host")
    return out


def adi_step_host_pipelined(fields, grid, mat, params, packs, Tinf=0.0):
    """One ADI step of each of several independent HOST fields (e.g. an ensemble, or frames that
    are streamed out while the next input streams in): two staging slots on two CUDA streams keep
    the upload of one field, the compute of another and the download of a third in flight
    (adi_cart_step_host_async).  `fields`: sequence of (nx,ny,nz) float64 arrays; returns the list of
    stepped arrays.  Page-locked inputs/outputs (torch pinned tensors' numpy views) transfer fastest."""
    shape = (grid.nx, grid.ny, grid.nz)
    kappa = mat.k / (mat.rho * mat.cp)
    e = _engine
    e.bind(grid)
    e.set_mask(grid)
    e.set_packs(packs)
    L, ctx = e.lib(), e.context()
    ins = [np.ascontiguousarray(f, dtype=np.float64) for f in fields]
    for f in ins:
        if f.shape != shape:
            raise ValueError("field shape does not match the grid")
    outs = [np.empty(shape, dtype=np.float64) for _ in ins]
    if ins:  # the first step runs alone: it (re)builds the neighbour code before two streams share it
        _capi.check(L.adi_cart_step_host(ctx, ins[0].ctypes.data, outs[0].ctypes.data, 1, params.dt, params.theta,
                                         kappa, float(Tinf), _stream_ptr()), "adi_cart_step_host")
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    for i in range(1, len(ins)):
        s = i & 1
        if i >= 3:
            streams[s].synchronize()   # the slot's previous occupant is done
        _capi.check(L.adi_cart_step_host_async(ctx, s, ins[i].ctypes.data, outs[i].ctypes.data, params.dt,
                                               params.theta, kappa, float(Tinf), streams[s].cuda_stream),
                    "adi_cart_step_host_async")
    for st in streams:
        st.synchronize()
    return outs


def set_option(name, value):
    """Engine tuning knob (adi_set_option): 'm' chunk length (16|32), 'kt' / 'lt' lines per
    block of the strided / z sweeps; 0 restores the default."""
    _capi.check(_engine.lib().adi_set_option(_engine.context(), str(name).encode(), int(value)),
                "adi_set_option")


def get_option(name):
    """Current value of an engine option (adi_get_option)."""
    return int(_engine.lib().adi_get_option(_engine.context(), str(name).encode()))


def launch_count():
    return int(_engine.lib().adi_launch_count(_engine.context()))
