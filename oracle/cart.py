# -*- coding: utf-8 -*-
"""ctypes front-end of oracle/adi_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Mirrors the call signatures of the reference's CPU module (adi3d_numba_coeff.py)
closely enough for the parity tests to read like the reference's drivers.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  Parity status: pinned (tests/test_oracle_golden.py).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
FACES = ("x-", "x+", "y-", "y+", "z-", "z+")


def build(force: bool = False) -> str:
    """Compile liboracle.so next to its source (gcc; OpenMP when libgomp is usable)."""
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "adi_oracle.c")
    if not force and os.path.exists(so) and os.path.getmtime(so) >= os.path.getmtime(src):
        return so
    base = ["-O2", "-fPIC", "-shared", "-ffp-contract=off", "-std=c11", "-o", so, src]
    last = None
    for cc in ("/usr/bin/gcc", "gcc", "cc"):
        for omp in (["-fopenmp"], []):
            try:
                subprocess.run([cc] + omp + base, check=True, capture_output=True)
                return so
            except (OSError, subprocess.CalledProcessError) as e:  # try the next recipe
                last = e
    raise RuntimeError(f"could not build the oracle: {last}")


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        dp, bp, ip = C.POINTER(C.c_double), C.POINTER(C.c_uint8), C.POINTER(C.c_int)
        L.oracle_set_threads.argtypes = [C.c_int]
        L.oracle_max_threads.restype = C.c_int
        L.oracle_exposed_mask.argtypes = [bp, C.c_int, C.c_int, C.c_int, C.c_int, bp]
        L.oracle_precompute_packs.argtypes = (
            [bp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double,
             ip, dp, C.POINTER(dp), C.c_int, ip, ip, dp, C.POINTER(dp)] + [dp] * 6)
        L.oracle_lap1d.argtypes = [dp, bp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, dp]
        L.oracle_sweep_axis.argtypes = ([C.c_int, dp, bp, dp, bp, dp, dp, C.c_int, C.c_int, C.c_int]
                                        + [C.c_double] * 4 + [dp])
        L.oracle_adi_step.argtypes = ([dp, bp, C.c_int, C.c_int, C.c_int] + [C.c_double] * 7
                                      + [dp, bp, dp, dp] * 3 + [dp, dp])
        _LIB = L
    return _LIB


def set_threads(n: int) -> None:
    lib().oracle_set_threads(int(n))


def max_threads() -> int:
    return int(lib().oracle_max_threads())


def _d(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _b(a):
    return a.view(np.uint8).ctypes.data_as(C.POINTER(C.c_uint8))


class Grid3D:  # adi3d_numba_coeff.py:14-19
    def __init__(self, nx, ny, nz, dx, mask):
        self.nx, self.ny, self.nz = int(nx), int(ny), int(nz)
        self.dx = float(dx)
        self.mask = np.array(mask, dtype=np.bool_, order="C", copy=True)
        assert self.mask.shape == (self.nx, self.ny, self.nz)


class Material:  # :21-23
    def __init__(self, rho, cp, k):
        self.rho, self.cp, self.k = float(rho), float(cp), float(k)


class Params:  # :25-27
    def __init__(self, dt, theta=0.5):
        self.dt, self.theta = float(dt), float(theta)


class AxisCoeffPack:  # :29-36 (constructor copies)
    def __init__(self, coeff, dir_mask, dir_val, qflux=None):
        self.coeff = np.array(coeff, dtype=np.float64, order="C", copy=True)
        self.dir_mask = np.array(dir_mask, dtype=np.bool_, order="C", copy=True)
        self.dir_val = np.array(dir_val, dtype=np.float64, order="C", copy=True)
        if qflux is None:
            qflux = np.zeros_like(self.coeff)
        self.qflux = np.array(qflux, dtype=np.float64, order="C", copy=True)


def exposed_mask(mask, face):  # :38-55
    if face not in FACES:
        raise ValueError("bad face")
    m = np.ascontiguousarray(mask, dtype=np.bool_)
    nx, ny, nz = m.shape
    out = np.zeros(m.shape, dtype=np.bool_)
    lib().oracle_exposed_mask(_b(m), nx, ny, nz, FACES.index(face), _b(out))
    return out


def precompute_coeff_packs_unified(grid, mat, dir_mask=None, dir_value=None, neumann=None,
                                   robin_h=None, robin_Tinf=None):  # :57-118
    nx, ny, nz = grid.nx, grid.ny, grid.nz
    shape = (nx, ny, nz)
    if dir_mask is None:
        dir_mask = np.zeros(shape, dtype=np.bool_)
    if dir_value is None:
        dir_value = np.zeros(shape, dtype=np.float64)
    elif np.isscalar(dir_value):
        dir_value = np.full(shape, float(dir_value), dtype=np.float64)

    def classify(v):
        if v is None:
            return 0, 0.0, None
        if np.isscalar(v):
            return 1, float(v), None
        return 2, 0.0, np.ascontiguousarray(v, dtype=np.float64)

    hk, hs, hf = [0] * 6, [0.0] * 6, [None] * 6
    if robin_h is not None:
        for i, f in enumerate(FACES):
            v = robin_h.get(f, 0.0) if isinstance(robin_h, dict) else robin_h
            hk[i], hs[i], hf[i] = classify(v)
    qk, qs, qf, order = [0] * 6, [0.0] * 6, [None] * 6, []
    if neumann is not None:
        for f, v in neumann.items():
            i = FACES.index(f)
            order.append(i)
            qk[i], qs[i], qf[i] = classify(v)
    dp = C.POINTER(C.c_double)
    null = C.cast(None, dp)
    hfp = (dp * 6)(*[(_d(a) if a is not None else null) for a in hf])
    qfp = (dp * 6)(*[(_d(a) if a is not None else null) for a in qf])
    outs = [np.empty(shape, dtype=np.float64) for _ in range(6)]
    mask = np.ascontiguousarray(grid.mask, dtype=np.bool_)
    rc = lib().oracle_precompute_packs(
        _b(mask), nx, ny, nz, grid.dx, mat.rho, mat.cp,
        (C.c_int * 6)(*hk), (C.c_double * 6)(*hs), hfp,
        len(order), (C.c_int * max(1, len(order)))(*(order or [0])),
        (C.c_int * 6)(*qk), (C.c_double * 6)(*qs), qfp, *[_d(o) for o in outs])
    if rc:
        raise RuntimeError(f"oracle_precompute_packs rc={rc}")
    return tuple(AxisCoeffPack(outs[a], dir_mask, dir_value, outs[3 + a]) for a in range(3))


def lap1D(T, mask, dx, axis):  # :240-288
    T = np.ascontiguousarray(T, dtype=np.float64)
    m = np.ascontiguousarray(mask, dtype=np.bool_)
    out = np.empty_like(T)
    lib().oracle_lap1d(_d(T), _b(m), *T.shape, float(dx), int(axis), _d(out))
    return out


def sweep_axis(axis, prev, mask, pack, theta, gam, dt, Tinf):  # :133-237
    prev = np.ascontiguousarray(prev, dtype=np.float64)
    m = np.ascontiguousarray(mask, dtype=np.bool_)
    out = np.empty_like(prev)
    rc = lib().oracle_sweep_axis(int(axis), _d(prev), _b(m), _d(pack.coeff), _b(pack.dir_mask),
                                 _d(pack.dir_val), _d(pack.qflux), *prev.shape,
                                 float(theta), float(gam), float(dt), float(Tinf), _d(out))
    if rc:
        raise RuntimeError(f"oracle_sweep_axis rc={rc}")
    return out


def adi_step_numba_coeff(Tn, grid, mat, params, packs, Tinf=0.0, work=None):  # :290-302
    Tn = np.ascontiguousarray(Tn, dtype=np.float64)
    mask = np.ascontiguousarray(grid.mask, dtype=np.bool_)
    assert Tn.shape == (grid.nx, grid.ny, grid.nz) == mask.shape
    out = np.empty_like(Tn)
    args = []
    for p in packs:
        args += [_d(p.coeff), _b(p.dir_mask), _d(p.dir_val), _d(p.qflux)]
    wp = _d(work) if work is not None else C.cast(None, C.POINTER(C.c_double))
    rc = lib().oracle_adi_step(_d(Tn), _b(mask), grid.nx, grid.ny, grid.nz, grid.dx,
                               mat.rho, mat.cp, mat.k, params.dt, params.theta, float(Tinf),
                               *args, _d(out), wp)
    if rc:
        raise RuntimeError(f"oracle_adi_step rc={rc}")
    return out
