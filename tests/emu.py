# -*- coding: utf-8 -*-
"""Builds and binds csrc/host_emulation.cpp (the kernel's per-thread code run on the CPU).
Test-only helper."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "adi_thermal_fields_b200", "csrc")
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(ROOT, "tests", "_emu.so")
        srcs = [os.path.join(CSRC, f) for f in ("host_emulation.cpp", "adi_core.h", "adi_mask_core.h", "adi_tab_core.h", "adi_fmt_core.h", "adi_pow10_tab.h")]
        if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
            subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off",
                            "-o", so, srcs[0]], check=True)
        _LIB = C.CDLL(so)
    return _LIB


def _ptr(a, ty):
    return a.ctypes.data_as(C.POINTER(ty)) if a is not None else C.cast(None, C.POINTER(ty))


def cart_step(T, mask, dx, dt, theta, kappa, Tinf, coeff=(None,) * 3, dirm=(None,) * 3,
              dirv=(None,) * 3, q=(None,) * 3, face_coeff=None, variant=0):
    L = lib()
    T = np.ascontiguousarray(T, dtype=np.float64)
    nx, ny, nz = T.shape
    out = np.empty_like(T)
    m8 = np.ascontiguousarray(mask, dtype=np.bool_).view(np.uint8)
    dp, bp = C.POINTER(C.c_double), C.POINTER(C.c_uint8)
    keep = []

    def arr3(xs, ty, dtype):
        ps = []
        for x in xs:
            if x is None:
                ps.append(C.cast(None, C.POINTER(ty)))
            else:
                a = np.ascontiguousarray(x, dtype=dtype)
                if dtype == np.bool_:
                    a = a.view(np.uint8)
                keep.append(a)
                ps.append(_ptr(a, ty))
        return (C.POINTER(ty) * 3)(*ps)

    fc = None
    if face_coeff is not None:
        fc = np.ascontiguousarray(face_coeff, dtype=np.float64)
    L.emu_cart_step.argtypes = [dp, dp, bp, C.c_int, C.c_int, C.c_int] + [C.c_double] * 5 + \
        [C.POINTER(dp), C.POINTER(bp), C.POINTER(dp), C.POINTER(dp), dp, C.c_int]
    rc = L.emu_cart_step(_ptr(T, C.c_double), _ptr(out, C.c_double), _ptr(m8, C.c_uint8), nx, ny, nz,
                         dx, dt, theta, kappa, Tinf, arr3(coeff, C.c_double, np.float64),
                         arr3(dirm, C.c_uint8, np.bool_), arr3(dirv, C.c_double, np.float64),
                         arr3(q, C.c_double, np.float64), _ptr(fc, C.c_double), int(variant))
    assert rc == 0
    return out


KINDS = {"neumann0": 0, "dirichlet": 1, "robin": 2}


def cyl_step(c, M=16, nslab=1):
    """One cylindrical BE step of case dict `c` (tests/cases.build_cyl_case) through the
    host-compiled table-driven solve (csrc/adi_tab_core.h).  nslab > 1: the z sweep as nslab segments per
    line (the z-slab algorithm of the multi-GPU path)."""
    L = lib()
    T = np.ascontiguousarray(c["T0"], dtype=np.float64)
    out = np.empty_like(T)
    z = c["zbc"]
    prm = np.array([c["dt"], c["rho"], c["cp"], c["k"], c["h_r"], c["Tinf_r"], z["h_bot"], z["h_top"],
                    z["T_inf_bot"], z["T_inf_top"], z["T_bot"], z["T_top"], c["T_void"], c["T_inner"]],
                   dtype=np.float64)
    act = None if c.get("active") is None else np.ascontiguousarray(c["active"], dtype=np.bool_).view(np.uint8)
    S = None if c.get("S") is None else np.ascontiguousarray(c["S"], dtype=np.float64)
    dp, bp = C.POINTER(C.c_double), C.POINTER(C.c_uint8)
    L.emu_cyl_step_slab.argtypes = [dp, dp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, dp,
                                    C.c_int, C.c_int, bp, dp, C.c_int, C.c_int]
    rc = L.emu_cyl_step_slab(_ptr(T, C.c_double), _ptr(out, C.c_double), c["nr"], c["nphi"], c["nz"], c["dr"],
                             c["dphi"], c["dz"], _ptr(prm, C.c_double), KINDS[z["kind_bot"]], KINDS[z["kind_top"]],
                             _ptr(act, C.c_uint8), _ptr(S, C.c_double), int(M), int(nslab))
    assert rc == 0
    return out


def text_field(T, fmt):
    """Value lines of a whole (nx,ny,nz) field, formatted by csrc/adi_fmt_core.h on the CPU."""
    L = lib()
    a = np.ascontiguousarray(T)
    if a.dtype == np.bool_:
        a = a.view(np.uint8)
    code = {np.dtype(np.float64): 0, np.dtype(np.float32): 1, np.dtype(np.uint8): 2}[a.dtype]
    nx, ny, nz = a.shape
    out = np.zeros(a.size * 15 + 16, dtype=np.uint8)
    L.emu_text_field.restype = C.c_long
    n = L.emu_text_field(a.ctypes.data_as(C.c_void_p), C.c_int(code), C.c_int(nx), C.c_int(ny), C.c_int(nz),
                         C.c_int(fmt), out.ctypes.data_as(C.c_void_p))
    return out[:n].tobytes()
