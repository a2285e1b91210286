# -*- coding: utf-8 -*-
"""Drop-in for the reference's voxel_bc_correction.py (STLBoundaryCorrector,
build_corrected_robin_fields): the triangle scatter and the field correction run on the GPU."""
import os as _os
import sys as _sys

_root = _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
if _root not in _sys.path:
    _sys.path.insert(0, _root)

from adi_thermal_fields_b200.voxel_bc_correction import (  # noqa: F401,E402
    STLBoundaryCorrector, build_corrected_robin_fields)
