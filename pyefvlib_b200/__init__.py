"""pyefvlib_b200 -- B200-native EbFVM assembly + sparse solve behind PyEFVLib's API.

Host-side names mirror the reference package (PyEFVLib/__init__.py:1-39): Grid, GridData, MSHReader, read,
ProblemData, NumericalSettings, PropertyData, BoundaryConditions, Solver, LinearSystem(CSR), the
Neumann/Dirichlet/Constant/Variable constants.  All arithmetic of the hot path runs in hand-written CUDA
kernels (pyefvlib_b200/csrc) through the C-ABI of include/efv_b200.h; there is no CPU fallback.
"""
from .grid import GridData, MSHReader, Grid, Region, Boundary, read, structured_grid, structured_grid_data  # noqa: F401
from .shapes import SHAPES, TRIANGLE, QUADRILATERAL, TETRAHEDRON, HEXAHEDRON  # noqa: F401

from .problem import (ProblemData, NumericalSettings, PropertyData, BoundaryConditions, DirichletBoundaryCondition,  # noqa: F401
                      NeumannBoundaryCondition, Neumann, Dirichlet, Constant, Variable)
from .solver import Solver, Saver, CsvSaver, NullSaver  # noqa: F401
from .savers import VtuSaver, VtuSeriesSaver  # noqa: F401
from .linear_system import LinearSystem, LinearSystemCSR  # noqa: F401
