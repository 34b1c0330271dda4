"""emme_b200 -- B200-native (sm_100a) implementation of EMME's eigenmatrix assembly and
Newton/secant dense step, behind a C ABI (include/emme_b200.h).

Python here is a thin host-side mirror of the reference's solver entry points
(EigenSolver, solve_once_eigen, the scan generator) over that C ABI; the compute path is the
CUDA library and fails loudly without it.
"""
from .capi import EmmeError, EmmeParams, EmmePicParams, EmmeStats, load  # noqa: F401
from .solver import EigenSolver, Input, scan_values, solve_once_eigen  # noqa: F401
from .pic import PIC_State, Integrator, calculate_omega, solve_once_pic  # noqa: F401

__all__ = ["EigenSolver", "Input", "solve_once_eigen", "scan_values", "PIC_State", "Integrator",
           "calculate_omega", "solve_once_pic", "EmmePicParams", "EmmeError", "EmmeParams",
           "EmmeStats", "load"]
