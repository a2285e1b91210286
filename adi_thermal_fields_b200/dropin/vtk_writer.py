# -*- coding: utf-8 -*-
"""Drop-in for the reference's `vtk_writer` module (vtk_writer.py:12): same function, same file
bytes, values formatted on the GPU.  Put adi_thermal_fields_b200/dropin first on sys.path."""
from adi_thermal_fields_b200.vtk_writer import write_vtk_structured_points  # noqa: F401
