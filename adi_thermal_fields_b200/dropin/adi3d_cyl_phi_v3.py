# -*- coding: utf-8 -*-
"""Drop-in for the reference's adi3d_cyl_phi_v3.py: same module name, same symbols
(GridCyl, Material, Params, RobinR, ZBC, adi_step), B200 kernels underneath.
Put this directory on sys.path ahead of the reference tree."""
import os as _os
import sys as _sys

_root = _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
if _root not in _sys.path:
    _sys.path.insert(0, _root)

from adi_thermal_fields_b200.adi3d_cyl_phi_v3 import (  # noqa: F401,E402
    GridCyl, Material, Params, RobinR, ZBC, adi_step, adi_step_device)
