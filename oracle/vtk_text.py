# -*- coding: utf-8 -*-
"""Restatement of the reference's two ASCII VTK writers -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

  flavour 0  vtk_writer.py:4-30        write_vtk_structured_points(path,T,dx,origin,field_name,mask)
             values "{float(v):.6e}", Fortran-order flattening, nine per line; ORIGIN at the cell
             centre, "%.9e"; mask section named "mask"
  flavour 1  waam_from_stl_v7_mm.py:186-215  write_vtk_structured_points(path,T,dx_mm,origin_mm,...)
             values "{float(T[i,j,k]):.6g}", one line per (k, j) row of nx values; ORIGIN / SPACING
             "%.9g"; mask section named "Mask"

Parity status: PINNED against the unmodified reference functions (tools/gen_golden.py --only vtk
-> tests/golden/vtk_text.npz holds the reference's files for the seeded cases of tests/cases.py;
tests/test_oracle_golden.py compares byte for byte).  Pure-Python loops: small cases only."""
from __future__ import annotations

import numpy as np

SPEC = {0: ".6e", 1: ".6g"}


def data_section(T, fmt):
    """The value lines of one SCALARS section, as bytes."""
    T = np.asarray(T)
    nx, ny, nz = T.shape
    if fmt == 0:                                     # vtk_writer.py:4-9,17
        flat = T.reshape(-1, order="F")
        lines = []
        for i in range(0, flat.size, 9):
            lines.append(" ".join(format(float(v), ".6e") for v in flat[i:i + 9]) + "\n")
        return "".join(lines).encode()
    lines = []                                       # waam_from_stl_v7_mm.py:203-206
    for k in range(nz):
        for j in range(ny):
            lines.append(" ".join(format(float(T[i, j, k]), ".6g") for i in range(nx)) + "\n")
    return "".join(lines).encode()


def header(fmt, shape, dx, origin, field_name):
    nx, ny, nz = shape
    ox, oy, oz = map(float, origin)
    dx = float(dx)
    if fmt == 0:                                     # vtk_writer.py:15-28
        c = (ox + dx * 0.5, oy + dx * 0.5, oz + dx * 0.5)
        s = ("# vtk DataFile Version 3.0\n" "Uniform grid with Temperature and mask\n" "ASCII\n"
             "DATASET STRUCTURED_POINTS\n" f"DIMENSIONS {nx} {ny} {nz}\n"
             f"ORIGIN {c[0]:.9e} {c[1]:.9e} {c[2]:.9e}\n" f"SPACING {dx:.9e} {dx:.9e} {dx:.9e}\n"
             f"POINT_DATA {nx*ny*nz}\n")
    else:                                            # waam_from_stl_v7_mm.py:192-200
        s = ("# vtk DataFile Version 3.0\n" "WAAM Structured Points (mm)\n" "ASCII\n"
             "DATASET STRUCTURED_POINTS\n" f"DIMENSIONS {nx} {ny} {nz}\n"
             f"ORIGIN {ox:.9g} {oy:.9g} {oz:.9g}\n" f"SPACING {dx:.9g} {dx:.9g} {dx:.9g}\n"
             f"POINT_DATA {nx*ny*nz}\n")
    return s + section_header(field_name)


def section_header(name):
    return f"SCALARS {name} float 1\n" "LOOKUP_TABLE default\n"


def vtk_bytes(fmt, T, dx, origin=(0.0, 0.0, 0.0), field_name="Temperature", mask=None):
    """The whole file the reference writer of flavour `fmt` produces, as bytes."""
    T = np.asarray(T)
    out = header(fmt, T.shape, dx, origin, field_name).encode("utf-8") + data_section(T, fmt)
    if mask is not None:
        M = np.asarray(mask, dtype=np.float32)       # vtk_writer.py:29, waam...:208
        out += section_header("mask" if fmt == 0 else "Mask").encode() + data_section(M, fmt)
    return out
