# -*- coding: utf-8 -*-
"""Drop-in for the two HOT-PATH functions of the reference's quick_spiral_deposition_gif_v5.py
(adi_step_masked :31-70, build_grid_annular :74-80), which tests/test_spiral_vs_analytic.py:9
imports.  The rest of that file is a matplotlib GIF command line (out of scope); shadowing the
module name therefore hides that CLI -- run the CLI from the reference tree without this
directory on the path."""
import os as _os
import sys as _sys

_root = _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
if _root not in _sys.path:
    _sys.path.insert(0, _root)

from adi_thermal_fields_b200.adi3d_cyl_phi_v3 import (  # noqa: F401,E402
    GridCyl, Material, Params, RobinR, ZBC, adi_step, adi_step_masked, build_grid_annular)
