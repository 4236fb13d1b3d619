"""pyefvlib_b200 -- B200-native EbFVM assembly + sparse solve behind PyEFVLib's API.

Host-side names mirror the reference package (PyEFVLib/__init__.py:1-39): Grid, GridData, MSHReader, read,
ProblemData, NumericalSettings, PropertyData, BoundaryConditions, Solver, LinearSystem(CSR), the
Neumann/Dirichlet/Constant/Variable constants.  All arithmetic of the hot path runs in hand-written CUDA
kernels (pyefvlib_b200/csrc) through the C-ABI of include/efv_b200.h; there is no CPU fallback.
"""
from .grid import GridData, MSHReader, Grid, Region, Boundary, read, structured_grid, structured_grid_data  # noqa: F401
from .shapes import SHAPES, TRIANGLE, QUADRILATERAL, TETRAHEDRON, HEXAHEDRON  # noqa: F401

Neumann = "NEUMANN_BOUNDARY_CONDITION"          # PyEFVLib/__init__.py:28-31
Dirichlet = "DIRICHLET_BOUNDARY_CONDITION"
Constant = "BOUNDARY_CONDITION_CONSTANT_VALUE"
Variable = "BOUNDARY_CONDITION_VARIABLE_VALUE"
