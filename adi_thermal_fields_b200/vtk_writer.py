# -*- coding: utf-8 -*-
"""Output path of the deposition drivers on B200: the reference's ASCII VTK writers and probe
downloads, with the value formatting done by the GPU (libadi_b200.so, csrc/adi_text.cu).

Same names, arguments and FILE BYTES as the reference:
  write_vtk_structured_points(path, T, dx, origin=(0,0,0), field_name="Temperature", mask=None)
      vtk_writer.py:12-30 -- "%.6e", Fortran-order flattening, nine values per line, ORIGIN at the
      cell centre
  write_vtk_structured_points_mm(path, T, dx_mm, origin_mm=(0,0,0), field_name="Temperature", mask=None)
      waam_from_stl_v7_mm.py:186-215 (there it is also called write_vtk_structured_points) --
      "%.6g", one line per (k, j) row of nx values
`T` / `mask` may be device arrays (the cupy shim's ndarray, torch CUDA tensors) -- no host copy of
the field is made, the text is what crosses PCIe -- or host NumPy arrays (uploaded first).

Beyond the reference ("async D2H of probe lines / slices and VTK frames"):
  AsyncVTKWriter   queues snapshots and writes them from a worker thread while stepping continues
  ProbeRecorder    records T[i0, j0, :] / T[:, j0, :] / boxes without stalling the stepping stream
No CPU fallback: without the built library and a CUDA device every call raises."""
from __future__ import annotations

import ctypes as C
import queue
import threading

import numpy as np
import torch

from . import _capi
from . import devarray as cp

FMT_E6, FMT_G6 = 0, 1
_DTYPE_CODE = {torch.float64: 0, torch.float32: 1, torch.bool: 2, torch.uint8: 2}


def _context():
    if not torch.cuda.is_available():
        raise RuntimeError("vtk_writer: no CUDA device (there is no CPU fallback)")
    return _capi.context(torch.cuda.current_device())


def _device_field(x):
    """-> contiguous CUDA tensor of a dtype the formatter reads (fp64 / fp32 / 1-byte)."""
    if isinstance(x, cp.ndarray):
        t = x._t
    elif torch.is_tensor(x):
        t = x
    else:
        a = np.asarray(x)
        if a.dtype not in (np.float64, np.float32, np.bool_, np.uint8):
            a = a.astype(np.float64)             # float(v) of any other dtype
        t = torch.from_numpy(np.ascontiguousarray(a))
    if not t.is_cuda:
        t = t.cuda()
    if t.dtype not in _DTYPE_CODE:
        t = t.to(torch.float64)
    if t.dim() != 3:
        raise AssertionError("T.ndim == 3")      # waam_from_stl_v7_mm.py:188
    return t.contiguous()


def _snapshot(x):
    """A private device copy of `x` (drivers mutate their arrays in place at births)."""
    t = _device_field(x)
    src = x._t if isinstance(x, cp.ndarray) else x
    return t.clone() if torch.is_tensor(src) and src.is_cuda and t.data_ptr() == src.data_ptr() else t


def _header(fmt, shape, dx, origin, field_name):
    nx, ny, nz = shape
    ox, oy, oz = (float(o) for o in origin)
    dx = float(dx)
    if fmt == FMT_E6:                            # vtk_writer.py:15-26
        cx, cy, cz = ox + dx * 0.5, oy + dx * 0.5, oz + dx * 0.5
        title = "Uniform grid with Temperature and mask"
        geo = f"ORIGIN {cx:.9e} {cy:.9e} {cz:.9e}\nSPACING {dx:.9e} {dx:.9e} {dx:.9e}\n"
    else:                                        # waam_from_stl_v7_mm.py:192-200
        title = "WAAM Structured Points (mm)"
        geo = f"ORIGIN {ox:.9g} {oy:.9g} {oz:.9g}\nSPACING {dx:.9g} {dx:.9g} {dx:.9g}\n"
    return (f"# vtk DataFile Version 3.0\n{title}\nASCII\nDATASET STRUCTURED_POINTS\n"
            f"DIMENSIONS {nx} {ny} {nz}\n{geo}POINT_DATA {nx*ny*nz}\n") + _section(field_name)


def _section(name):
    return f"SCALARS {name} float 1\nLOOKUP_TABLE default\n"


def _write_field(ctx, path, append, prefix, t, fmt, stream):
    nx, ny, nz = t.shape
    pre = prefix.encode("utf-8")
    n = C.c_ulonglong(0)
    _capi.check(_capi.load().adi_text_write(ctx, str(path).encode(), int(append), pre, len(pre), t.data_ptr(),
                                            _DTYPE_CODE[t.dtype], nx, ny, nz, fmt, C.byref(n), stream),
                "adi_text_write")
    return int(n.value)


def _write(fmt, path, T, dx, origin, field_name, mask, ctx=None, stream=None):
    t = _device_field(T)
    m = None
    if mask is not None:
        m = _device_field(mask)
        if tuple(m.shape) != tuple(t.shape):
            raise AssertionError("mask.shape == T.shape")   # waam_from_stl_v7_mm.py:209
    ctx = _context() if ctx is None else ctx
    stream = torch.cuda.current_stream().cuda_stream if stream is None else stream
    nbytes = _write_field(ctx, path, False, _header(fmt, tuple(t.shape), dx, origin, field_name), t, fmt, stream)
    if m is not None:
        nbytes += _write_field(ctx, path, True, _section("mask" if fmt == FMT_E6 else "Mask"), m, fmt, stream)
    return nbytes


def write_vtk_structured_points(path, T, dx, origin=(0.0, 0.0, 0.0), field_name="Temperature", mask=None):
    """vtk_writer.py:12-30, byte for byte.  Returns the number of bytes written."""
    return _write(FMT_E6, path, T, dx, origin, field_name, mask)


def write_vtk_structured_points_mm(path, T, dx_mm, origin_mm=(0.0, 0.0, 0.0), field_name="Temperature", mask=None):
    """waam_from_stl_v7_mm.py:186-215, byte for byte.  Returns the number of bytes written."""
    return _write(FMT_G6, path, T, dx_mm, origin_mm, field_name, mask)


def format_planes(T, fmt, k0=0, kc=None):
    """The value lines of planes [k0, k0+kc) of `T` as bytes (device-formatted; tests and tools)."""
    t = _device_field(T)
    nx, ny, nz = t.shape
    kc = nz - k0 if kc is None else int(kc)
    L = _capi.load()
    cap = int(L.adi_text_capacity(nx * ny * kc))
    buf = torch.empty(cap, dtype=torch.uint8, device=t.device)
    n = C.c_ulonglong(0)
    _capi.check(L.adi_text_format(_context(), t.data_ptr(), _DTYPE_CODE[t.dtype], nx, ny, nz, int(k0), kc, int(fmt),
                                  buf.data_ptr(), cap, C.byref(n), torch.cuda.current_stream().cuda_stream),
                "adi_text_format")
    return buf[:int(n.value)].cpu().numpy().tobytes()


class AsyncVTKWriter:
    """Frames are snapshotted on the stepping stream (one device-to-device copy) and written by a
    worker thread through the library's own output stream, so `T = step(T)` keeps running.
    Errors surface at wait()/close()."""

    def __init__(self, flavour="vtk_writer", max_pending=2):
        self.fmt = {"vtk_writer": FMT_E6, "waam": FMT_G6}[flavour]
        self.ctx = _context()
        self.device = torch.cuda.current_device()
        self.q = queue.Queue(maxsize=max_pending)   # bounds the snapshots held in HBM
        self.errors = []
        self.bytes_written = 0
        self.th = threading.Thread(target=self._run, daemon=True)
        self.th.start()

    def _run(self):
        torch.cuda.set_device(self.device)
        while True:
            job = self.q.get()
            try:
                if job is None:
                    return
                path, t, dx, origin, name, m, ev = job
                ev.synchronize()                     # the snapshot copies are complete
                self.bytes_written += _write(self.fmt, path, t, dx, origin, name, m, ctx=self.ctx, stream=None)
            except Exception as e:                   # noqa: BLE001 -- reported by wait()
                self.errors.append(e)
            finally:
                self.q.task_done()

    def submit(self, path, T, dx, origin=(0.0, 0.0, 0.0), field_name="Temperature", mask=None):
        t = _snapshot(T)
        m = None if mask is None else _snapshot(mask)
        ev = torch.cuda.Event()
        ev.record()
        self.q.put((path, t, dx, origin, field_name, m, ev))

    def wait(self):
        self.q.join()
        if self.errors:
            raise self.errors.pop(0)

    def close(self):
        self.q.put(None)
        self.th.join()
        if self.errors:
            raise self.errors.pop(0)


class ProbeRecorder:
    """Asynchronous T[index] downloads (basic indexing: ints and unit-step slices).

        rec = ProbeRecorder(nslots=8, slot_bytes=nz * 8)
        tk = rec.record(T, (i0, j0, slice(None)))      # returns at once; PCIe copy on a copy stream
        ...                                            # more steps
        line = rec.fetch(tk)                           # NumPy array, shape of T[i0, j0, :]
    """

    def __init__(self, nslots=8, slot_bytes=1 << 20):
        self.ctx = _context()
        self.nslots = int(nslots)
        self.slot_bytes = int(slot_bytes)
        _capi.check(_capi.load().adi_probe_open(self.ctx, self.nslots, self.slot_bytes), "adi_probe_open")
        self.next = 0

    def record(self, T, index):
        t = T._t if isinstance(T, cp.ndarray) else T
        if not (torch.is_tensor(t) and t.is_cuda and t.is_contiguous() and t.dim() == 3):
            raise TypeError("ProbeRecorder.record: contiguous 3-D CUDA array expected")
        if not isinstance(index, tuple):
            index = (index,)
        index = tuple(index) + (slice(None),) * (3 - len(index))
        lo, hi, keep = [], [], []
        for ax, (ix, n) in enumerate(zip(index, t.shape)):
            if isinstance(ix, slice):
                a, b, st = ix.indices(n)
                if st != 1:
                    raise IndexError("ProbeRecorder.record: unit-step slices only")
                lo.append(a); hi.append(b); keep.append(b - a)
            else:
                i = int(ix) + (n if int(ix) < 0 else 0)
                if not 0 <= i < n:
                    raise IndexError("index out of range")
                lo.append(i); hi.append(i + 1)
        slot = self.next
        self.next = (self.next + 1) % self.nslots
        I3 = C.c_int * 3
        _capi.check(_capi.load().adi_probe_record(self.ctx, slot, t.data_ptr(), t.element_size(), *t.shape,
                                                  I3(*lo), I3(*hi), torch.cuda.current_stream().cuda_stream),
                    "adi_probe_record")
        dtype = {torch.float64: np.float64, torch.float32: np.float32, torch.bool: np.bool_,
                 torch.uint8: np.uint8}[t.dtype]
        return (slot, tuple(keep), dtype)

    def fetch(self, ticket, wait=True):
        """-> the recorded array, or None when wait=False and the copy is still in flight."""
        slot, shape, dtype = ticket
        out = np.empty(shape, dtype=dtype)
        n = C.c_size_t(0)
        rc = _capi.load().adi_probe_fetch(self.ctx, slot, out.ctypes.data, out.nbytes, C.byref(n), int(bool(wait)))
        if rc == 1:
            return None
        _capi.check(rc, "adi_probe_fetch")
        if n.value != out.nbytes:
            raise RuntimeError("ProbeRecorder.fetch: record size does not match the ticket")
        return out
