# -*- coding: utf-8 -*-
"""GPU parity of the output path (csrc/adi_text.cu through the C ABI and the reference's
write_vtk_structured_points interfaces): the files must equal, BYTE FOR BYTE, what the reference's
writers produce (tests/golden/vtk_text.npz holds the reference's own files) and, at sizes the
Python loops cannot reach, what the same formatting code produces on the CPU (tests/emu.py, itself
checked against Python's float formatting in tests/test_text_format.py)."""
import os

import numpy as np
import pytest

import cases
import emu

pytestmark = pytest.mark.gpu

FLAVOURS = [0, 1]


def _writer(fmt):
    from adi_thermal_fields_b200 import vtk_writer as vw
    return vw.write_vtk_structured_points if fmt == 0 else vw.write_vtk_structured_points_mm


@pytest.mark.parametrize("fmt", FLAVOURS)
@pytest.mark.parametrize("name", sorted(cases.vtk_text_cases()))
@pytest.mark.parametrize("resident", ["host", "device"])
def test_files_equal_the_reference_files(name, fmt, resident, golden_dir, tmp_path):
    import torch
    from adi_thermal_fields_b200 import _capi, devarray as cp
    g = np.load(os.path.join(golden_dir, "vtk_text.npz"))
    c = cases.vtk_text_cases()[name]
    T, m = c["T"], c["mask"]
    if resident == "device":
        T = cp.asarray(T)
        m = None if m is None else cp.asarray(m)
    n0 = _capi.load().adi_launch_count(_capi.context(0))
    path = tmp_path / "out.vtk"
    nbytes = _writer(fmt)(str(path), T, c["dx"], c["origin"], c["field_name"], m)
    assert _capi.load().adi_launch_count(_capi.context(0)) - n0 == (3 if m is None else 6)
    got = path.read_bytes()
    assert got == g[f"{name}__{fmt}"].tobytes()
    assert nbytes == len(got)
    torch.cuda.synchronize()


@pytest.mark.parametrize("fmt", FLAVOURS)
@pytest.mark.parametrize("shape", [(300, 37, 45), (513, 9, 17), (31, 2, 100), (1, 1, 1), (256, 3, 16), (257, 5, 33)])
def test_value_lines_match_cpu_formatter(fmt, shape):
    """Ragged x tiles (nx not a multiple of 32 / 256), plane counts off the 16-plane block, wide
    dynamic range, negative values, ties."""
    from adi_thermal_fields_b200 import vtk_writer as vw
    rng = np.random.default_rng(sum(shape) + fmt)
    T = (rng.random(shape) - 0.3) * 10.0 ** rng.integers(-9, 10, size=shape)
    T.flat[:: 7] = np.round(T.flat[:: 7] * 64.0) / 64.0
    T.flat[:: 11] = 20.0
    T.flat[:: 13] = 0.0
    assert vw.format_planes(T, fmt) == emu.text_field(T, fmt)


@pytest.mark.parametrize("fmt", FLAVOURS)
def test_random_bit_patterns_whole_exponent_range(fmt):
    from adi_thermal_fields_b200 import vtk_writer as vw
    rng = np.random.default_rng(77 + fmt)
    T = rng.integers(0, 2 ** 64, size=(130, 40, 50), dtype=np.uint64).view(np.float64)
    assert vw.format_planes(T, fmt) == emu.text_field(T, fmt)


@pytest.mark.parametrize("fmt", FLAVOURS)
def test_plane_ranges_concatenate(fmt):
    """Chunks are independent: text(planes a..b) + text(planes b..c) == text(planes a..c)."""
    from adi_thermal_fields_b200 import vtk_writer as vw
    rng = np.random.default_rng(5)
    T = 20.0 + 1380.0 * rng.random((70, 11, 41))
    whole = vw.format_planes(T, fmt)
    parts = b"".join(vw.format_planes(T, fmt, k0, kc) for k0, kc in ((0, 16), (16, 7), (23, 1), (24, 17)))
    assert parts == whole == emu.text_field(T, fmt)


@pytest.mark.parametrize("dtype", [np.float32, np.bool_, np.uint8, np.int32])
@pytest.mark.parametrize("fmt", FLAVOURS)
def test_other_dtypes_follow_float_of_value(dtype, fmt):
    from adi_thermal_fields_b200 import vtk_writer as vw
    from oracle import vtk_text
    rng = np.random.default_rng(3)
    if dtype == np.float32:
        A = (rng.random((19, 6, 5)) * 2000.0 - 500.0).astype(np.float32)
    elif dtype == np.bool_:
        A = rng.random((19, 6, 5)) < 0.5
    else:
        A = rng.integers(0, 200, size=(19, 6, 5)).astype(dtype)
    assert vw.format_planes(A, fmt) == vtk_text.data_section(A, fmt)


@pytest.mark.parametrize("fmt", FLAVOURS)
def test_multi_chunk_file_with_mask(fmt, tmp_path):
    """512 x 512 x 40: two plane chunks through the pinned double buffers, then the mask section."""
    import torch
    from adi_thermal_fields_b200 import vtk_writer as vw
    g = torch.Generator(device="cuda").manual_seed(9)
    T = 20.0 + 1380.0 * torch.rand((512, 512, 40), dtype=torch.float64, device="cuda", generator=g)
    T[:, :, 5] = -T[:, :, 5]
    T[::3, ::5, 7] = 0.0
    mask = torch.rand((512, 512, 40), device="cuda", generator=g) < 0.8
    path = tmp_path / "big.vtk"
    nbytes = _writer(fmt)(str(path), T, 1e-3, (0.0, 0.0, 0.0), "Temperature", mask)
    data = path.read_bytes()
    assert len(data) == nbytes
    Th, mh = T.cpu().numpy(), mask.cpu().numpy()
    t_txt, m_txt = emu.text_field(Th, fmt), emu.text_field(mh, fmt)
    sec = b"SCALARS %s float 1\nLOOKUP_TABLE default\n" % (b"mask" if fmt == 0 else b"Mask")
    head_end = data.index(b"LOOKUP_TABLE default\n") + len(b"LOOKUP_TABLE default\n")
    assert data[head_end:head_end + len(t_txt)] == t_txt
    assert data[head_end + len(t_txt):] == sec + m_txt
    assert b"DIMENSIONS 512 512 40\n" in data[:head_end] and b"POINT_DATA 10485760\n" in data[:head_end]


def test_async_writer_snapshots_and_overlaps(tmp_path):
    """Frames submitted to the worker are snapshots: in-place changes after submit() (births do
    T[born] = Ts) must not leak into the file; results equal the synchronous writer's."""
    import torch
    from adi_thermal_fields_b200 import vtk_writer as vw, devarray as cp
    rng = np.random.default_rng(21)
    T = cp.asarray(20.0 + 100.0 * rng.random((40, 33, 20)))
    mask = cp.asarray(rng.random((40, 33, 20)) < 0.6)
    w = vw.AsyncVTKWriter("waam")
    frames = []
    for f in range(3):
        p = tmp_path / f"a{f}.vtk"
        vw.write_vtk_structured_points_mm(str(tmp_path / f"s{f}.vtk"), T, 0.5, (1.0, 2.0, 3.0), "Temperature", mask)
        w.submit(str(p), T, 0.5, (1.0, 2.0, 3.0), "Temperature", mask)
        frames.append(p)
        T[:, :, f] = 1400.0                      # in-place change right after the submit
        mask[:, :, f] = True
    w.wait()
    w.close()
    for f, p in enumerate(frames):
        assert p.read_bytes() == (tmp_path / f"s{f}.vtk").read_bytes()
    assert w.bytes_written == sum(p.stat().st_size for p in frames)
    torch.cuda.synchronize()


def test_probe_recorder_lines_slices_boxes():
    import torch
    from adi_thermal_fields_b200 import vtk_writer as vw, devarray as cp
    rng = np.random.default_rng(8)
    Th = rng.random((23, 17, 40))
    T = cp.asarray(Th)
    rec = vw.ProbeRecorder(nslots=4, slot_bytes=23 * 40 * 8)
    picks = [(3, 5, slice(None)), (slice(None), 9, slice(None)), (slice(2, 9), slice(4, 6), slice(10, 30)), (-1, -1, -1)]
    tickets = [rec.record(T, ix) for ix in picks]
    T[:] = 0.0                                    # later writes do not change recorded probes
    for tk, ix in zip(tickets, picks):
        got = rec.fetch(tk)
        assert got.shape == np.asarray(Th[ix]).shape
        assert np.array_equal(got, Th[ix])
    # masks too, and polling
    m = torch.from_numpy(Th > 0.5).cuda()
    tk = rec.record(m, (slice(None), 0, slice(None)))
    got = None
    while got is None:
        got = rec.fetch(tk, wait=False)
    assert np.array_equal(got, (Th > 0.5)[:, 0, :])
    # a slot keeps its record until fetched; a box beyond the slot is refused
    tks = [rec.record(T, (0, 0, slice(None))) for _ in range(4)]
    with pytest.raises(Exception):
        rec.record(T, (0, 0, slice(None)))
    for tk in tks:
        rec.fetch(tk)
    with pytest.raises(MemoryError):
        rec.record(T, (slice(None), slice(None), slice(None)))
    with pytest.raises(IndexError):
        rec.record(T, (slice(0, 10, 2), 0, 0))


def test_bad_arguments_raise(tmp_path):
    from adi_thermal_fields_b200 import vtk_writer as vw
    with pytest.raises(AssertionError):
        vw.write_vtk_structured_points(str(tmp_path / "x.vtk"), np.zeros((3, 3)), 1.0)
    with pytest.raises(AssertionError):
        vw.write_vtk_structured_points_mm(str(tmp_path / "x.vtk"), np.zeros((3, 3, 3)), 1.0, mask=np.ones((3, 3, 2), bool))
    with pytest.raises(ValueError):
        vw.write_vtk_structured_points(str(tmp_path / "no_such_dir" / "x.vtk"), np.zeros((3, 3, 3)), 1.0)
