# -*- coding: utf-8 -*-
"""adi_thermal_fields_b200 -- B200 (sm_100a) engine for the ADI heat step of
Matemusi/ADI_thermal_fields.

The package holds only the hot path: CUDA kernels + C ABI (csrc/, include/adi_b200.h)
and host-side mirrors of the reference's module interfaces:

  adi3d_gpu_coeff     Cartesian step, same names as the reference's adi3d_gpu_coeff.py
  adi3d_cyl_phi_v3    cylindrical (r, phi, z) backward-Euler step
  vtk_writer          the reference's ASCII VTK writers (same bytes, formatted on the GPU), async frame
                      writer and probe recorder
  dropin/             directory to put on sys.path in front of the reference tree:
                      `cupy` shim + the two module names above

There is no CPU fallback: importing the compute modules without the built library, or
calling them without a CUDA device, raises.
"""
__version__ = "0.1.0"
