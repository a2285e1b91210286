# -*- coding: utf-8 -*-
"""`cupy`-named shim for the reference's unchanged drivers (`import cupy as cp`).
Not CuPy: a thin device-array surface on torch CUDA tensors (see
adi_thermal_fields_b200/devarray.py); the ADI step runs in libadi_b200.so."""
import os as _os
import sys as _sys

_root = _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))))
if _root not in _sys.path:
    _sys.path.insert(0, _root)

from adi_thermal_fields_b200.devarray import *  # noqa: F401,F403,E402
from adi_thermal_fields_b200.devarray import (  # noqa: F401,E402
    ndarray, cuda, asarray, array, asnumpy, any, all, sum, abs, bool_, float64, float32, int64, int32, uint8)

__version__ = "0.0-adi_b200-shim"
