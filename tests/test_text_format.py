# -*- coding: utf-8 -*-
"""ASCII output path, CPU side: the per-value formatter of csrc/adi_fmt_core.h (the code the
device runs, compiled for the host by tests/emu.py) against Python's own float formatting --
which is what the reference writers call (vtk_writer.py:8 "{float(v):.6e}",
waam_from_stl_v7_mm.py:205 "{float(T[i,j,k]):.6g}").  Bit-exact text is the bar."""
import ctypes as C
import math
from fractions import Fraction

import numpy as np
import pytest

import emu

SPEC = {0: ".6e", 1: ".6g"}


def fmt_values(vals, fmt):
    L = emu.lib()
    v = np.ascontiguousarray(vals, dtype=np.float64)
    n = v.size
    out = np.zeros(16 * n, dtype=np.uint8)
    lens = np.zeros(n, dtype=np.int32)
    L.emu_format_values(v.ctypes.data_as(C.c_void_p), C.c_long(n), C.c_int(fmt),
                        out.ctypes.data_as(C.c_void_p), lens.ctypes.data_as(C.c_void_p))
    o = out.reshape(n, 16)
    return [bytes(o[i, :lens[i]]).decode() for i in range(n)]


def check(vals, fmt):
    vals = np.asarray(vals, dtype=np.float64)
    got = fmt_values(vals, fmt)
    bad = [(repr(v), g, format(v, SPEC[fmt])) for v, g in zip(vals.tolist(), got) if g != format(v, SPEC[fmt])]
    assert not bad, bad[:5]


def neighbours(vals):
    v = np.asarray(vals, dtype=np.float64)
    return np.concatenate([v, np.nextafter(v, np.inf), np.nextafter(v, -np.inf), -v])


@pytest.mark.parametrize("fmt", [0, 1])
def test_special_values(fmt):
    check([0.0, -0.0, math.inf, -math.inf, math.nan, -math.nan, 5e-324, -5e-324, 2.2250738585072014e-308,
           2.225073858507201e-308, 1.7976931348623157e308, 1.0, -1.0, 20.0, 1400.0, 0.1, 0.3, 2 / 3,
           9.9999995, 9.9999994999, 99999.95, 999999.5, 9999995.0, 9.9999999e99, 9.9999999e-101,
           1e-5, 1e-4, 0.0001, 0.00012345678, 123456.5, 1234565.0, 100000.0, 999999.4], fmt)


@pytest.mark.parametrize("fmt", [0, 1])
def test_powers_of_ten_and_neighbours(fmt):
    check(neighbours([float("1e%d" % j) for j in range(-323, 309)]), fmt)
    check(neighbours([float("9.999995e%d" % j) for j in range(-320, 308)]), fmt)
    check(neighbours([float("9.99999949999e%d" % j) for j in range(-320, 308)]), fmt)


@pytest.mark.parametrize("fmt", [0, 1])
def test_exact_ties_round_half_even(fmt):
    """(n + 1/2) * 10^q that are exactly representable, and their neighbours in the last bit."""
    P = 7 if fmt == 0 else 6
    rng = np.random.default_rng(11)
    ties = []
    for q in range(0, 12):                       # (2n+1) * 5^q * 2^(q-1): exact while it fits 53 bits
        for n in rng.integers(10 ** (P - 1), 10 ** P, size=40):
            fr = Fraction(2 * int(n) + 1, 2) * Fraction(10) ** q
            if Fraction(float(fr)) == fr:
                ties.append(float(fr))
    for q in range(1, 9):                        # (2n+1) / (2 * 10^q): needs 5^q | 2n+1
        step = 5 ** q
        lo = (2 * 10 ** (P - 1) + 1 + step - 1) // step
        hi = (2 * 10 ** P - 1) // step
        for m in rng.integers(lo, hi, size=60):
            t = int(m) * step
            if t % 2 == 0:
                t += step
            if t > 2 * 10 ** P - 1:
                continue
            fr = Fraction(t, 2) / Fraction(10) ** q
            if Fraction(float(fr)) == fr:
                ties.append(float(fr))
    assert len(ties) > 300
    check(neighbours(ties), fmt)


@pytest.mark.parametrize("fmt", [0, 1])
def test_random_bit_patterns(fmt):
    rng = np.random.default_rng(5 + fmt)
    check(rng.integers(0, 2 ** 64, size=300_000, dtype=np.uint64).view(np.float64), fmt)


@pytest.mark.parametrize("fmt", [0, 1])
def test_temperature_like_fields(fmt):
    rng = np.random.default_rng(7)
    v = np.concatenate([20.0 + 1380.0 * rng.random(200_000), rng.normal(0, 1e-3, 50_000),
                        np.round(rng.random(50_000) * 1e4) / 8.0,         # short dyadic values: tie candidates
                        rng.integers(0, 2, 1000).astype(np.float64),
                        rng.random(20_000).astype(np.float32).astype(np.float64)])
    check(v, fmt)


def test_exact_comparison_against_rationals():
    """exact_cmp_half(a, n, q) = sign(a - (n + 1/2) 10^q) over the whole exponent range."""
    L = emu.lib()
    L.emu_exact_cmp_half.argtypes = [C.c_double, C.c_uint, C.c_int]
    rng = np.random.default_rng(3)
    for _ in range(3000):
        k = int(rng.integers(-323, 309))
        q = k - 6
        n = int(rng.integers(10 ** 6, 10 ** 7))
        half = Fraction(2 * n + 1, 2) * Fraction(10) ** q
        try:
            a = float(half)
        except OverflowError:
            continue
        if a == 0.0 or math.isinf(a):
            continue
        for cand in (a, math.nextafter(a, math.inf), math.nextafter(a, 0.0)):
            if cand == 0.0 or math.isinf(cand):
                continue
            want = (Fraction(cand) > half) - (Fraction(cand) < half)
            assert L.emu_exact_cmp_half(cand, n, q) == want, (cand, n, q)


@pytest.mark.parametrize("fmt", [0, 1])
@pytest.mark.parametrize("shape", [(5, 4, 3), (9, 1, 1), (1, 1, 1), (7, 3, 11), (300, 2, 3)])
def test_field_layout_matches_writer_loops(fmt, shape):
    """emu_text_field (the kernel's ordering and separators) against the reference writers' loops."""
    import oracle.vtk_text as ov
    rng = np.random.default_rng(sum(shape))
    T = (rng.random(shape) - 0.3) * 10.0 ** rng.integers(-8, 9, size=shape)
    out = np.zeros(T.size * 16 + 16, dtype=np.uint8)
    L = emu.lib()
    L.emu_text_field.restype = C.c_long
    n = L.emu_text_field(T.ctypes.data_as(C.c_void_p), 0, *[C.c_int(s) for s in shape], C.c_int(fmt),
                         out.ctypes.data_as(C.c_void_p))
    want = ov.data_section(T, fmt)
    assert bytes(out[:n]) == want


@pytest.mark.parametrize("fmt", [0, 1])
def test_writer_headers_match_the_reference_files(fmt, golden_dir):
    """Host side of the writers (adi_thermal_fields_b200/vtk_writer.py): the header and section lines it
    hands to adi_text_write as `prefix` are the reference's, byte for byte (no device needed)."""
    import os
    import cases
    from adi_thermal_fields_b200 import vtk_writer as vw
    g = np.load(os.path.join(golden_dir, "vtk_text.npz"))
    for name, c in cases.vtk_text_cases().items():
        want = g[f"{name}__{fmt}"].tobytes()
        head = vw._header(fmt, c["T"].shape, c["dx"], c["origin"], c["field_name"]).encode("utf-8")
        assert want.startswith(head)
        if c["mask"] is not None:
            assert vw._section("mask" if fmt == 0 else "Mask").encode() in want[len(head):]
        # data section by the host build of the device formatter completes the file
        body = emu.text_field(np.asarray(c["T"]), fmt)
        assert want[len(head):len(head) + len(body)] == body
