# -*- coding: utf-8 -*-
"""ctypes binding of libadi_b200.so (include/adi_b200.h).  Fails loudly: no fallback."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_lib", "libadi_b200.so")

OK, EINVAL, ECUDA, ESTATE, ENOMEM = 0, -1, -2, -3, -4

# every symbol include/adi_b200.h declares (checked by tests/test_capi_symbols.py)
SYMBOLS = (
    "adi_ctx_create", "adi_ctx_destroy", "adi_last_error", "adi_version", "adi_sync",
    "adi_malloc", "adi_free", "adi_h2d", "adi_d2h",
    "adi_cart_bind", "adi_cart_set_mask", "adi_cart_set_pack", "adi_cart_set_robin_scalar",
    "adi_cart_step", "adi_cart_step_host", "adi_cart_build_packs", "adi_cart_exposed_mask",
    "adi_set_option", "adi_get_option", "adi_launch_count", "adi_profile_reset", "adi_profile_read",
    "adi_cyl_bind", "adi_cyl_step", "adi_cyl_step_host",
    "adi_cart_set_slab", "adi_cart_set_mask_halo", "adi_cart_pack_zplanes", "adi_cart_step_xy",
    "adi_cart_zsweep_reduce", "adi_cart_zsweep_finish",
    "adi_cart_zsweep_spike", "adi_cart_zsweep_solve0", "adi_cart_zsweep_apply",
    "adi_voxel_project", "adi_voxel_correct", "adi_cart_step_host_async",
    "adi_cyl_set_slab", "adi_cyl_step_rphi", "adi_cyl_zsweep_reduce", "adi_cyl_zsweep_finish",
    "adi_text_capacity", "adi_text_format", "adi_text_write",
    "adi_probe_open", "adi_probe_record", "adi_probe_fetch",
    "adi_dist_unique_id", "adi_dist_init", "adi_dist_init_comm", "adi_dist_comm", "adi_dist_destroy", "adi_dist_set_option",
    "adi_dist_info", "adi_cart_slab_sync_mask", "adi_cart_slab_step",
)


class CylParams(C.Structure):
    _fields_ = [("dt", C.c_double), ("rho", C.c_double), ("cp", C.c_double), ("k", C.c_double),
                ("h_r", C.c_double), ("Tinf_r", C.c_double),
                ("kind_bot", C.c_int), ("kind_top", C.c_int),
                ("h_bot", C.c_double), ("h_top", C.c_double),
                ("Tinf_bot", C.c_double), ("Tinf_top", C.c_double),
                ("T_bot", C.c_double), ("T_top", C.c_double),
                ("T_void", C.c_double), ("T_inner", C.c_double)]


class AdiError(RuntimeError):
    pass


_lib = None


def load():
    """Load the shared library; ImportError when it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m adi_thermal_fields_b200._build` "
            "(or __graft_entry__.build()); there is no CPU fallback")
    L = C.CDLL(LIB_PATH)
    vp, dp, bp, ip = C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int)
    dbl = C.c_double
    L.adi_last_error.restype = C.c_char_p
    L.adi_version.restype = C.c_char_p
    L.adi_ctx_create.argtypes = [C.c_int, C.POINTER(vp)]
    L.adi_ctx_destroy.argtypes = [vp]
    L.adi_sync.argtypes = [vp, vp]
    L.adi_malloc.argtypes = [vp, C.c_size_t, C.POINTER(vp)]
    L.adi_free.argtypes = [vp, vp]
    L.adi_h2d.argtypes = [vp, vp, vp, C.c_size_t, vp]
    L.adi_d2h.argtypes = [vp, vp, vp, C.c_size_t, vp]
    L.adi_cart_bind.argtypes = [vp, C.c_int, C.c_int, C.c_int, dbl]
    L.adi_cart_set_mask.argtypes = [vp, bp]
    L.adi_cart_set_pack.argtypes = [vp, C.c_int, dp, bp, dp, dp]
    L.adi_cart_set_robin_scalar.argtypes = [vp, C.POINTER(dbl)]
    L.adi_cart_step.argtypes = [vp, dp, dp, dbl, dbl, dbl, dbl, vp]
    L.adi_cart_step_host.argtypes = [vp, vp, vp, C.c_int, dbl, dbl, dbl, dbl, vp]
    L.adi_cart_step_host_async.argtypes = [vp, C.c_int, vp, vp, dbl, dbl, dbl, dbl, vp]
    L.adi_cart_build_packs.argtypes = [vp, dbl, dbl, ip, C.POINTER(dbl), C.POINTER(vp),
                                       ip, C.POINTER(dbl), C.POINTER(vp)] + [dp] * 6 + [vp]
    L.adi_cart_exposed_mask.argtypes = [vp, C.c_int, bp, vp]
    L.adi_set_option.argtypes = [vp, C.c_char_p, C.c_long]
    L.adi_get_option.argtypes = [vp, C.c_char_p]
    L.adi_get_option.restype = C.c_long
    L.adi_launch_count.argtypes = [vp]
    L.adi_launch_count.restype = C.c_long
    L.adi_profile_reset.argtypes = [vp]
    L.adi_profile_read.argtypes = [vp, C.POINTER(dbl), C.POINTER(C.c_long)]
    L.adi_cart_set_slab.argtypes = [vp, C.c_int, C.c_int]
    L.adi_cart_set_mask_halo.argtypes = [vp, bp, bp]
    L.adi_cart_pack_zplanes.argtypes = [vp, vp, C.c_int, vp, vp, vp]
    L.adi_cart_step_xy.argtypes = [vp, dp, dp, dp, dp, dbl, dbl, dbl, dbl, vp]
    L.adi_cart_zsweep_reduce.argtypes = [vp, dp, dp, dp, dbl, dbl, dbl, dbl, vp]
    L.adi_cart_zsweep_finish.argtypes = [vp, dp, dp, dp, dbl, dbl, dbl, dbl, vp]
    L.adi_cart_zsweep_spike.argtypes = [vp, dp, C.c_int, C.c_int, dbl, dp, vp, ip, dbl, dbl, dbl, vp]
    L.adi_cart_zsweep_solve0.argtypes = [vp, dp, dp, dbl, dbl, dbl, dbl, vp]
    L.adi_cart_zsweep_apply.argtypes = [vp, dp, dp, dp, dp, dp, vp, vp, C.c_int, vp]
    L.adi_voxel_project.argtypes = [vp, dp, dp, dp, C.c_int, C.POINTER(dbl), dbl, C.c_int, dbl, bp, C.c_int, C.c_int,
                                    C.c_int, C.POINTER(vp), vp]
    L.adi_voxel_correct.argtypes = [vp, bp, C.c_int, C.c_int, C.c_int, dbl, C.POINTER(vp), ip, C.POINTER(dbl), C.c_int,
                                    C.POINTER(vp), C.POINTER(vp), vp]
    L.adi_cyl_bind.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, dbl, dbl, dbl]
    L.adi_cyl_step.argtypes = [vp, dp, dp, C.POINTER(CylParams), bp, dp, vp]
    L.adi_cyl_set_slab.argtypes = [vp, C.c_int, C.c_int, ip]
    L.adi_cyl_step_rphi.argtypes = [vp, dp, dp, C.POINTER(CylParams), bp, dp, vp]
    L.adi_cyl_zsweep_reduce.argtypes = [vp, dp, C.POINTER(CylParams), dp, vp]
    L.adi_cyl_zsweep_finish.argtypes = [vp, dp, C.POINTER(CylParams), dp, bp, vp]
    L.adi_cyl_step_host.argtypes = [vp, vp, vp, C.c_int, C.POINTER(CylParams), vp, vp, vp]
    ull = C.c_ulonglong
    L.adi_text_capacity.argtypes = [C.c_size_t]
    L.adi_text_capacity.restype = C.c_size_t
    L.adi_text_format.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp,
                                  C.c_size_t, C.POINTER(ull), vp]
    L.adi_text_write.argtypes = [vp, C.c_char_p, C.c_int, C.c_char_p, C.c_size_t, vp, C.c_int, C.c_int, C.c_int,
                                 C.c_int, C.c_int, C.POINTER(ull), vp]
    L.adi_probe_open.argtypes = [vp, C.c_int, C.c_size_t]
    L.adi_probe_record.argtypes = [vp, C.c_int, vp, C.c_int, C.c_int, C.c_int, C.c_int, ip, ip, vp]
    L.adi_probe_fetch.argtypes = [vp, C.c_int, vp, C.c_size_t, C.POINTER(C.c_size_t), C.c_int]
    L.adi_dist_unique_id.argtypes = [vp]
    L.adi_dist_init.argtypes = [vp, vp, C.c_int, C.c_int]
    L.adi_dist_init_comm.argtypes = [vp, vp, C.c_int, C.c_int]
    L.adi_dist_comm.argtypes = [vp, C.POINTER(vp)]
    L.adi_dist_destroy.argtypes = [vp]
    L.adi_dist_set_option.argtypes = [vp, C.c_char_p, C.c_long]
    L.adi_dist_info.argtypes = [vp, ip, ip, C.POINTER(C.c_long), C.POINTER(C.c_long)]
    L.adi_cart_slab_sync_mask.argtypes = [vp, vp]
    L.adi_cart_slab_step.argtypes = [vp, dp, dp, dbl, dbl, dbl, dbl, vp]
    _lib = L
    return L


def check(rc: int, what: str = "") -> None:
    """Map a C-ABI status to the Python exception the reference's interface would raise."""
    if rc == OK:
        return
    msg = load().adi_last_error().decode("utf-8", "replace")
    if rc == EINVAL:
        raise ValueError(msg or what)
    if rc == ENOMEM:
        raise MemoryError(msg or what)
    raise AdiError(f"{what}: {msg} (rc={rc})")


_ctx = {}


def context(device: int = 0):
    """The per-process, per-device engine context (SURVEY.md 8b "Threading")."""
    if device not in _ctx:
        L = load()
        h = C.c_void_p()
        check(L.adi_ctx_create(int(device), C.byref(h)), "adi_ctx_create")
        _ctx[device] = h
    return _ctx[device]
