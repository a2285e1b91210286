# -*- coding: utf-8 -*-
"""Device array type behind the `cupy` shim (dropin/cupy).

The reference's drivers talk to their GPU backend through `import cupy as cp`
(quick_compare_neumann_robin_backend.py:140-141,163-164,184;
quick_compare_layer_birth_robin_v3.py:138,149,272-277,300,314,326;
waam_from_stl_v7_mm.py:410-411,491-493,500).  CuPy is not a dependency here (and is not
allowed as a compute fallback), so this module provides the small surface those drivers
use -- creation, slicing, boolean / index-tuple assignment, host transfer, stream sync --
on top of torch CUDA tensors.  torch is plumbing only: the ADI step itself runs in
libadi_b200.so on the raw device pointers of these arrays.
"""
from __future__ import annotations

import numpy as np
import torch

float64 = np.float64
float32 = np.float32
int64 = np.int64
int32 = np.int32
uint8 = np.uint8
bool_ = np.bool_

_NP2T = {np.dtype(np.float64): torch.float64, np.dtype(np.float32): torch.float32,
         np.dtype(np.int64): torch.int64, np.dtype(np.int32): torch.int32,
         np.dtype(np.uint8): torch.uint8, np.dtype(np.bool_): torch.bool,
         np.dtype(np.int8): torch.int8, np.dtype(np.int16): torch.int16}
_T2NP = {v: k for k, v in _NP2T.items()}


def _tdtype(dtype):
    if dtype is None:
        return None
    if isinstance(dtype, torch.dtype):
        return dtype
    return _NP2T[np.dtype(dtype)]


_FORCE_DEVICE = None  # tests of the array semantics set this to torch.device("cpu"); the
                      # engine itself (adi3d_gpu_coeff.adi_step_gpu_coeff) still requires CUDA


def _device():
    if _FORCE_DEVICE is not None:
        return _FORCE_DEVICE
    if not torch.cuda.is_available():
        raise RuntimeError("adi_thermal_fields_b200: no CUDA device; there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _unwrap(x):
    if isinstance(x, ndarray):
        return x._t
    if isinstance(x, np.ndarray):
        return torch.from_numpy(np.ascontiguousarray(x)).to(_device())
    return x


def _index(idx):
    if isinstance(idx, tuple):
        return tuple(_unwrap(i) for i in idx)
    return _unwrap(idx)


class ndarray:
    """C-contiguous-or-view device array.  `_t` is the backing torch tensor."""
    __array_priority__ = 100.0

    def __init__(self, t: torch.Tensor):
        self._t = t

    # -- metadata ----------------------------------------------------------------
    @property
    def shape(self):
        return tuple(self._t.shape)

    @property
    def ndim(self):
        return self._t.dim()

    @property
    def size(self):
        return self._t.numel()

    @property
    def dtype(self):
        return _T2NP[self._t.dtype]

    @property
    def nbytes(self):
        return self._t.numel() * self._t.element_size()

    @property
    def T(self):
        return ndarray(self._t.permute(*reversed(range(self._t.dim()))))

    @property
    def data_ptr(self):
        return self._t.data_ptr()

    def __len__(self):
        return self._t.shape[0]

    def __repr__(self):
        return f"devarray({self.get()!r})"

    # -- transfer ------------------------------------------------------------------
    def get(self):
        return self._t.detach().cpu().numpy()

    def item(self):
        return self._t.item()

    def __float__(self):
        return float(self._t.item())

    def __int__(self):
        return int(self._t.item())

    def __bool__(self):
        return bool(self._t.item())

    # -- indexing ------------------------------------------------------------------
    def __getitem__(self, idx):
        return ndarray(self._t[_index(idx)])

    def __setitem__(self, idx, value):
        v = _unwrap(value)
        if isinstance(v, torch.Tensor) and v.dtype != self._t.dtype:
            v = v.to(self._t.dtype)
        if isinstance(idx, type(Ellipsis)) or (isinstance(idx, slice) and idx == slice(None)):
            if isinstance(v, torch.Tensor):
                self._t.copy_(v)
            else:
                self._t.fill_(v)
            return
        self._t[_index(idx)] = v

    # -- conversion ----------------------------------------------------------------
    def astype(self, dtype, copy=True, order="C"):
        t = self._t.to(_tdtype(dtype))
        if copy and t.data_ptr() == self._t.data_ptr():
            t = t.clone()
        return ndarray(t.contiguous())

    def copy(self, order="C"):
        return ndarray(self._t.clone(memory_format=torch.contiguous_format))

    def reshape(self, *shape):
        if len(shape) == 1 and isinstance(shape[0], (tuple, list)):
            shape = tuple(shape[0])
        return ndarray(self._t.reshape(shape))

    def ravel(self):
        return ndarray(self._t.reshape(-1))

    def transpose(self, *axes):
        if len(axes) == 1 and isinstance(axes[0], (tuple, list)):
            axes = tuple(axes[0])
        if not axes:
            return self.T
        return ndarray(self._t.permute(*axes))

    def fill(self, v):
        self._t.fill_(v)

    # -- reductions ----------------------------------------------------------------
    def any(self, axis=None):
        return ndarray(self._t.any() if axis is None else self._t.any(dim=axis))

    def all(self, axis=None):
        return ndarray(self._t.all() if axis is None else self._t.all(dim=axis))

    def sum(self, axis=None):
        t = self._t if self._t.dtype != torch.bool else self._t.to(torch.int64)
        return ndarray(t.sum() if axis is None else t.sum(dim=axis))

    def min(self, axis=None):
        return ndarray(self._t.min() if axis is None else self._t.amin(dim=axis))

    def max(self, axis=None):
        return ndarray(self._t.max() if axis is None else self._t.amax(dim=axis))

    def mean(self, axis=None):
        return ndarray(self._t.mean() if axis is None else self._t.mean(dim=axis))

    # -- elementwise ---------------------------------------------------------------
    @staticmethod
    def _numpy_promote(a, b, op):
        """NumPy/CuPy promotion where torch differs: a Python float (or true division) meeting a
        bool / integer array gives float64, not torch's default float32
        (e.g. `(-theta*gam) * m` with a bool `m`, adi3d_gpu_coeff.py:175-176)."""
        def nonfloat(t):
            return isinstance(t, torch.Tensor) and not (t.is_floating_point() or t.is_complex())
        def pyfloat(x):
            return isinstance(x, (float, np.floating)) and not isinstance(x, torch.Tensor)
        if op is torch.true_divide:
            if nonfloat(a) and not (isinstance(b, torch.Tensor) and b.is_floating_point()):
                a = a.to(torch.float64)
            elif nonfloat(b) and not (isinstance(a, torch.Tensor) and a.is_floating_point()):
                b = b.to(torch.float64)
        if nonfloat(a) and pyfloat(b):
            a = a.to(torch.float64)
        if nonfloat(b) and pyfloat(a):
            b = b.to(torch.float64)
        if isinstance(a, np.floating):
            a = float(a)
        if isinstance(b, np.floating):
            b = float(b)
        return a, b

    def _bin(self, other, op):
        a, b = self._numpy_promote(self._t, _unwrap(other), op)
        return ndarray(op(a, b))

    def _rbin(self, other, op):
        a, b = self._numpy_promote(_unwrap(other), self._t, op)
        if not isinstance(a, torch.Tensor):
            a = torch.as_tensor(a, device=self._t.device, dtype=b.dtype if b.is_floating_point() else None)
        return ndarray(op(a, b))

    def __add__(self, o): return self._bin(o, torch.add)
    def __radd__(self, o): return self._bin(o, torch.add)
    def __sub__(self, o): return self._bin(o, torch.sub)
    def __rsub__(self, o): return self._rbin(o, torch.sub)
    def __mul__(self, o): return self._bin(o, torch.mul)
    def __rmul__(self, o): return self._bin(o, torch.mul)
    def __truediv__(self, o): return self._bin(o, torch.true_divide)
    def __rtruediv__(self, o): return self._rbin(o, torch.true_divide)
    def __neg__(self): return ndarray(-self._t)
    def __abs__(self): return ndarray(self._t.abs())
    def __invert__(self): return ndarray(~self._t)
    def __and__(self, o): return self._bin(o, torch.bitwise_and)
    def __rand__(self, o): return self._bin(o, torch.bitwise_and)
    def __or__(self, o): return self._bin(o, torch.bitwise_or)
    def __ror__(self, o): return self._bin(o, torch.bitwise_or)
    def __xor__(self, o): return self._bin(o, torch.bitwise_xor)
    def __lt__(self, o): return self._bin(o, torch.lt)
    def __le__(self, o): return self._bin(o, torch.le)
    def __gt__(self, o): return self._bin(o, torch.gt)
    def __ge__(self, o): return self._bin(o, torch.ge)
    def __eq__(self, o): return self._bin(o, torch.eq)
    def __ne__(self, o): return self._bin(o, torch.ne)
    __hash__ = None

    def _ibin(self, o, op):
        op(_unwrap(o))
        return self

    def __iadd__(self, o): return self._ibin(o, self._t.add_)
    def __isub__(self, o): return self._ibin(o, self._t.sub_)
    def __imul__(self, o): return self._ibin(o, self._t.mul_)
    def __itruediv__(self, o): return self._ibin(o, self._t.div_)
    def __iand__(self, o): return self._ibin(o, self._t.bitwise_and_)
    def __ior__(self, o): return self._ibin(o, self._t.bitwise_or_)


# -- creation -----------------------------------------------------------------------
def asarray(a, dtype=None, order=None):
    """cp.asarray: no copy for a device array that already has the dtype."""
    td = _tdtype(dtype)
    if isinstance(a, ndarray):
        t = a._t if td is None or a._t.dtype == td else a._t.to(td)
        return ndarray(t if t.is_contiguous() else t.contiguous())
    if isinstance(a, torch.Tensor):
        t = a.to(_device())
        if td is not None:
            t = t.to(td)
        return ndarray(t.contiguous())
    h = np.ascontiguousarray(a) if dtype is None else np.ascontiguousarray(a, dtype=np.dtype(dtype))
    t = torch.from_numpy(h.copy() if not h.flags.writeable else h).to(_device())
    if t.device.type == "cpu" and h.size and t.data_ptr() == h.ctypes.data:
        t = t.clone()  # (test mode) never alias the caller's host array
    return ndarray(t)


def array(a, dtype=None, copy=True, order=None):
    out = asarray(a, dtype=dtype)
    if copy and isinstance(a, ndarray) and out._t.data_ptr() == a._t.data_ptr():
        out = out.copy()
    return out


def asnumpy(a):
    if isinstance(a, ndarray):
        return a.get()
    return np.asarray(a)


def _shape(shape):
    return (int(shape),) if np.isscalar(shape) else tuple(int(s) for s in shape)


def empty(shape, dtype=float64, order="C"):
    return ndarray(torch.empty(_shape(shape), dtype=_tdtype(dtype), device=_device()))


def zeros(shape, dtype=float64, order="C"):
    return ndarray(torch.zeros(_shape(shape), dtype=_tdtype(dtype), device=_device()))


def ones(shape, dtype=float64, order="C"):
    return ndarray(torch.ones(_shape(shape), dtype=_tdtype(dtype), device=_device()))


def full(shape, fill_value, dtype=None, order="C"):
    if dtype is None:
        dtype = np.asarray(fill_value).dtype
    return ndarray(torch.full(_shape(shape), fill_value, dtype=_tdtype(dtype), device=_device()))


def _like(a, dtype):
    shape = a.shape
    dt = dtype if dtype is not None else a.dtype
    return shape, dt


def empty_like(a, dtype=None):
    return empty(*_like(a, dtype))


def zeros_like(a, dtype=None):
    return zeros(*_like(a, dtype))


def ones_like(a, dtype=None):
    return ones(*_like(a, dtype))


def full_like(a, fill_value, dtype=None):
    s, d = _like(a, dtype)
    return full(s, fill_value, dtype=d)


def where(cond, x=None, y=None):
    c = _unwrap(cond)
    if x is None and y is None:
        return tuple(ndarray(t) for t in torch.where(c))
    xt, yt = _unwrap(x), _unwrap(y)
    if not isinstance(xt, torch.Tensor):
        xt = torch.as_tensor(xt, device=c.device, dtype=yt.dtype if isinstance(yt, torch.Tensor) else None)
    if not isinstance(yt, torch.Tensor):
        yt = torch.as_tensor(yt, device=c.device, dtype=xt.dtype)
    return ndarray(torch.where(c, xt, yt))


def transpose(a, axes=None):
    return a.transpose(*(axes or ()))


def isnan(a):
    return ndarray(torch.isnan(_unwrap(a)))


def isfinite(a):
    return ndarray(torch.isfinite(_unwrap(a)))


def logical_not(a):
    return ndarray(torch.logical_not(_unwrap(a)))


def logical_and(a, b):
    return ndarray(torch.logical_and(_unwrap(a), _unwrap(b)))


def logical_or(a, b):
    return ndarray(torch.logical_or(_unwrap(a), _unwrap(b)))


def any(a, axis=None):  # noqa: A001
    return asarray(a).any(axis)


def all(a, axis=None):  # noqa: A001
    return asarray(a).all(axis)


def sum(a, axis=None):  # noqa: A001
    return asarray(a).sum(axis)


def abs(a):  # noqa: A001
    return ndarray(_unwrap(a).abs())


def sqrt(a):
    return ndarray(torch.sqrt(_unwrap(a)))


def nanmin(a):
    t = _unwrap(a)
    return ndarray(torch.where(torch.isnan(t), torch.full_like(t, float("inf")), t).min())


def nanmax(a):
    t = _unwrap(a)
    return ndarray(torch.where(torch.isnan(t), torch.full_like(t, float("-inf")), t).max())


def isscalar(x):
    return np.isscalar(x)


# -- cp.cuda --------------------------------------------------------------------------
class _Stream:
    def __init__(self, null=False):
        self._null = null

    def synchronize(self):
        torch.cuda.current_stream().synchronize() if not self._null else torch.cuda.synchronize()


class _StreamNS(_Stream):
    null = _Stream(null=True)


class _DeviceCtx:
    def __init__(self, idx=0):
        self.id = int(idx)

    def use(self):
        torch.cuda.set_device(self.id)

    def synchronize(self):
        torch.cuda.synchronize(self.id)

    def __enter__(self):
        self._prev = torch.cuda.current_device()
        torch.cuda.set_device(self.id)
        return self

    def __exit__(self, *a):
        torch.cuda.set_device(self._prev)


class _Runtime:
    @staticmethod
    def getDeviceCount():
        return torch.cuda.device_count()

    @staticmethod
    def deviceSynchronize():
        torch.cuda.synchronize()


class cuda:  # namespace
    Stream = _StreamNS
    Device = _DeviceCtx
    runtime = _Runtime

    @staticmethod
    def is_available():
        return torch.cuda.is_available()


def pinned_empty(shape, dtype=np.float64):
    """A host NumPy array in page-locked memory (from torch's caching host allocator: reuse after the first allocation
    is free).  The host-array entry points return their results in such arrays: the download runs at PCIe speed, and
    in a time loop T = step(T) the next upload does too.  Small arrays come from the ordinary heap."""
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    if n < (1 << 20) or not torch.cuda.is_available():
        return np.empty(shape, dtype=dtype)
    tdt = {np.dtype(np.float64): torch.float64, np.dtype(np.float32): torch.float32, np.dtype(np.bool_): torch.bool,
           np.dtype(np.uint8): torch.uint8}[np.dtype(dtype)]
    return torch.empty(tuple(int(v) for v in shape), dtype=tdt, pin_memory=True).numpy()
