# -*- coding: utf-8 -*-
"""Restatement of the reference's STL -> voxel-face projected-area correction
(voxel_bc_correction.py:53-108 compute_voxel_projected_areas, :110-168 build_corrected_fields,
:170-204 helpers) -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Parity status: PINNED against the unmodified reference (tools/gen_golden.py -> tests/golden/
voxel_bc_*.npz; tests/test_oracle_golden.py, bit-exact).

The reference walks the triangles in order and accumulates per voxel in a dict; here the same
per-triangle arithmetic (same operation order) scatters into dense per-face arrays."""
from __future__ import annotations

import math

import numpy as np

FACES = ("x-", "x+", "y-", "y+", "z-", "z+")


def exposed_mask(mask, face):   # adi3d_numba_coeff.py:38-55
    m = np.asarray(mask, dtype=bool)
    nb = np.zeros_like(m)
    ax, sgn = "xyz".index(face[0]), face[1]
    src = [slice(None)] * 3
    dst = [slice(None)] * 3
    if sgn == "-":
        dst[ax], src[ax] = slice(1, None), slice(None, -1)
    else:
        dst[ax], src[ax] = slice(None, -1), slice(1, None)
    nb[tuple(dst)] = m[tuple(src)]
    return m & ~nb


def projected_areas(mesh, mask, origin, dx, max_subdiv=6, area_epsilon=1e-16):
    """-> dict face -> dense array of projected area per voxel (:53-108, :170-183)."""
    mask = np.asarray(mask, dtype=bool)
    origin = np.asarray(origin, dtype=float)
    shape = mask.shape
    out = {f: np.zeros(shape) for f in FACES}
    normals = np.asarray(mesh.face_normals, dtype=float)
    areas = np.asarray(mesh.area_faces, dtype=float)
    verts = np.asarray(mesh.triangles, dtype=float)
    max_subdiv = max(1, int(max_subdiv))
    for t in range(len(verts)):
        area = float(areas[t])
        if area <= area_epsilon:
            continue
        v0, v1, v2 = verts[t]
        span = (verts[t].max(axis=0) - verts[t].min(axis=0)) / dx
        smax = float(np.max(span))
        n = int(math.ceil(smax)) if smax > 1.0 else 1
        n = max(1, min(n, max_subdiv))
        sub_area = area if n == 1 else area / (n * n)

        def bary(i, j):
            a, b = i / float(n), j / float(n)
            c = 1.0 - a - b
            return c * v0 + a * v1 + b * v2

        subs = []
        if n == 1:
            subs.append((v0, v1, v2))
        else:
            for i in range(n):
                for j in range(n - i):
                    p0, p1, p2 = bary(i, j), bary(i + 1, j), bary(i, j + 1)
                    subs.append((p0, p1, p2))
                    if i + j < n - 1:
                        subs.append((p1, bary(i + 1, j + 1), p2))
        for p0, p1, p2 in subs:
            cen = ((p0 + p1) + p2) / 3.0        # np.mean over the three vertices
            idx = np.floor((cen - origin) / dx).astype(int)
            if np.any(idx < 0) or np.any(idx >= shape):
                continue
            key = (int(idx[0]), int(idx[1]), int(idx[2]))
            if not mask[key]:
                continue
            for comp, fneg, fpos in ((normals[t, 0], "x-", "x+"), (normals[t, 1], "y-", "y+"), (normals[t, 2], "z-", "z+")):
                if comp > 1e-12:
                    a = sub_area * comp
                    if a > 0.0:
                        out[fpos][key] += a
                elif comp < -1e-12:
                    a = sub_area * (-comp)
                    if a > 0.0:
                        out[fneg][key] += a
    return out


def build_corrected_robin_fields(mesh, mask, origin, dx, base_h, fallback_to_base=True, max_subdiv=6):
    """:110-168 / :207-226 -> (robin_h fields, area-scale fields), dicts keyed by the faces of base_h."""
    proj = projected_areas(mesh, mask, origin, dx, max_subdiv)
    face_area = dx * dx
    robin, scale = {}, {}
    for face, base in base_h.items():
        base = float(base)
        arr = np.zeros(np.shape(mask))
        scl = np.zeros(np.shape(mask))
        if base != 0.0:
            hit = proj[face] > 0.0
            s = proj[face][hit] / face_area
            arr[hit] += base * s
            scl[hit] += s
            if fallback_to_base:
                missing = exposed_mask(mask, face) & (arr <= 0.0)
                arr[missing] = base
                scl[missing] = 1.0
        robin[face], scale[face] = arr, scl
    return robin, scale
