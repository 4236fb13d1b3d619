"""ProblemData / NumericalSettings / PropertyData / BoundaryConditions with the reference's constructor
arguments, attributes and error behaviour (PyEFVLib/ProblemData.py:5-156, PyEFVLib/BoundaryConditions.py:4-31).
Host-side bookkeeping only: values end up in small arrays handed to the C-ABI."""
import os

import numpy as np

from . import grid as _grid

Neumann = "NEUMANN_BOUNDARY_CONDITION"
Dirichlet = "DIRICHLET_BOUNDARY_CONDITION"
Constant = "BOUNDARY_CONDITION_CONSTANT_VALUE"
Variable = "BOUNDARY_CONDITION_VARIABLE_VALUE"


def _get_function(expr, with_time):
    """Expression BCs are eval'd after textual substitution of x, y, z (, t) exactly like
    BoundaryConditions.py:4-7 / ProblemData.py:75-79 (including the naive str.replace)."""
    from numpy import (pi, sin, cos, tan, arcsin, arccos, arctan, sinh, cosh, tanh, arcsinh, arccosh, arctanh,  # noqa: F401
                       sqrt, e, log, exp, inf, mod, floor)
    if with_time:
        def function(x, y, z, t):
            return eval(expr.replace("x", str(x)).replace("y", str(y)).replace("z", str(z)).replace("t", str(t)))
    else:
        def function(x, y, z):
            return eval(expr.replace("x", str(x)).replace("y", str(y)).replace("z", str(z)))
    return function


class _BoundaryConditionBase:
    def __init__(self, grid, boundary, value, handle, expression=False):
        self.grid, self.boundary, self.value, self.handle, self.expression = grid, boundary, value, handle, expression
        if self.expression:
            self.function = _get_function(value, True)

    def getValue(self, index, time=0.0):
        if self.expression:
            x, y, z = self.grid.vertices[index].getCoordinates()
            return self.function(x, y, z, time)
        return self.value

    def values_on_vertices(self, time=0.0):
        """Value at every vertex of the boundary, in the boundary's vertex order."""
        idx = self.boundary.verticesIndexes
        if not self.expression:
            return np.full(len(idx), float(self.value))
        return np.array([self.getValue(int(i), time) for i in idx], dtype=np.float64)


class DirichletBoundaryCondition(_BoundaryConditionBase):
    __type__ = "DIRICHLET"


class NeumannBoundaryCondition(_BoundaryConditionBase):
    __type__ = "NEUMANN"


class NumericalSettings:
    def __init__(self, timeStep, finalTime=np.inf, tolerance=1e-4, maxNumberOfIterations=1000):
        self.timeStep, self.finalTime, self.tolerance, self.maxNumberOfIterations = timeStep, finalTime, tolerance, maxNumberOfIterations

    def init(self, problemData):
        self.problemData = problemData
        problemData.timeStep = self.timeStep
        problemData.finalTime = self.finalTime
        problemData.tolerance = self.tolerance
        problemData.maxNumberOfIterations = self.maxNumberOfIterations


class PropertyData:
    def __init__(self, propertiesDict):
        self.propertiesDict = propertiesDict
        self.properties = list(list(self.propertiesDict.values())[0].keys())

    def init(self, problemData):
        self.problemData = problemData
        self.regions = problemData.grid.gridData.regionsNames
        for regionName in self.propertiesDict.keys():
            if regionName not in self.regions:
                print(f"Warning! Region {regionName} not in the grid.")
        for regionName in self.regions:
            if regionName not in self.propertiesDict.keys():
                raise Exception(f"Error, must specify {regionName} properties.")

    def get(self, handle, propertyName):
        if handle in range(len(self.regions)):
            if propertyName in self.properties:
                return self.propertiesDict[self.regions[handle]][propertyName]
            raise Exception(f"Property {propertyName} not defined")
        raise Exception(f"Invalid region handle {handle}")

    def set(self, handle, propertyName, value):
        if handle in range(len(self.regions)):
            if propertyName in self.properties:
                self.propertiesDict[self.regions[handle]][propertyName] = value
            else:
                raise Exception(f"Property {propertyName} not defined")
        else:
            raise Exception(f"Invalid region handle {handle}")

    def table(self, names):
        """[nRegions, len(names)] array for the C-ABI."""
        return np.array([[float(self.get(h, n)) for n in names] for h in range(len(self.regions))], dtype=np.float64)


class BoundaryConditions:
    def __init__(self, boundaryConditionsDict):
        self.boundaryConditionsDict = boundaryConditionsDict

    def init(self, problemData):
        self.problemData = problemData
        grid = problemData.grid
        variables = self.boundaryConditionsDict.keys()
        boundaryNames = grid.gridData.boundariesNames
        first = list(self.boundaryConditionsDict.values())[0]
        for boundaryName in first.keys():
            if boundaryName != "InitialValue" and boundaryName not in boundaryNames:
                print(f"Warning! Boundary {boundaryName} not in the grid.")
        for boundaryName in boundaryNames:
            if boundaryName not in first.keys():
                raise Exception(f"Error, must specify {boundaryName} condition.")
        self.neumannBoundaries = {v: [] for v in variables}
        self.dirichletBoundaries = {v: [] for v in variables}
        self.boundaryConditions = [dict() for _ in grid.boundaries]
        self.initialValues = dict()
        nV = grid.numberOfVertices
        for variable in variables:
            iv = self.boundaryConditionsDict[variable]["InitialValue"]
            if np.isscalar(iv) and not isinstance(iv, str):
                self.initialValues[variable] = np.repeat(iv, nV)
            elif isinstance(iv, (list, np.ndarray)):
                if len(iv) == nV:
                    self.initialValues[variable] = np.array(iv)
                else:
                    raise Exception(f"Length of \"{variable}\" InitialValue must be equal to the number of vertices ({nV})")
            elif isinstance(iv, str):
                fn = _get_function(iv, False)
                c = grid.coords
                self.initialValues[variable] = np.array([fn(c[i, 0], c[i, 1], c[i, 2]) for i in range(nV)])
        bcHandle = 0
        for idx, boundary in enumerate(grid.boundaries):
            for variable in variables:
                spec = self.boundaryConditionsDict[variable][boundary.name]
                if spec["type"] == Constant:
                    expression = False
                elif spec["type"] == Variable:
                    expression = True
                else:
                    raise Exception(f"Invalid Boundary Condition Type {spec['type']}")
                if spec["condition"] == Neumann:
                    bc = NeumannBoundaryCondition(grid, boundary, spec["value"], bcHandle, expression=expression)
                    self.neumannBoundaries[variable].append(bc)
                elif spec["condition"] == Dirichlet:
                    bc = DirichletBoundaryCondition(grid, boundary, spec["value"], bcHandle, expression=expression)
                    self.dirichletBoundaries[variable].append(bc)
                else:
                    raise Exception("Invalid boundary condition!")
                self.boundaryConditions[idx][variable] = bc
                bcHandle += 1
        problemData.dirichletBoundaries = self.dirichletBoundaries
        problemData.neumannBoundaries = self.neumannBoundaries
        problemData.boundaryConditions = self.boundaryConditions
        problemData.initialValues = self.initialValues


class ProblemData:
    """Same keyword arguments as the reference; `meshFilePath` may also be a Grid / GridData instance so that
    in-memory (synthetic) meshes need no file (the GridData seam of README.md:372)."""

    def __init__(self, meshFilePath, propertyData, boundaryConditions=None, numericalSettings=None, outputFilePath=None):
        self.meshFilePath = meshFilePath
        self.outputFilePath = outputFilePath
        self.propertyData = propertyData
        self.parsePaths()
        self.createGrid()
        (numericalSettings or NumericalSettings(0.0)).init(self)
        propertyData.init(self)
        if boundaryConditions:
            boundaryConditions.init(self)

    def parsePaths(self):
        here = os.path.dirname(os.path.abspath(__file__))
        self.libraryPath = os.path.realpath(os.path.join(here, os.pardir))
        meshes = os.environ.get("PYEFVLIB_MESHES", os.path.join(self.libraryPath, "meshes"))
        results = os.path.join(self.libraryPath, "results")
        if isinstance(self.meshFilePath, str):
            self.meshFilePath = self.meshFilePath.replace("{MESHES}", meshes)
        if isinstance(self.outputFilePath, str):
            self.outputFilePath = self.outputFilePath.replace("{RESULTS}", results)

    def createGrid(self):
        if isinstance(self.meshFilePath, _grid.Grid):
            self.grid = self.meshFilePath
        elif isinstance(self.meshFilePath, _grid.GridData):
            self.grid = _grid.Grid(self.meshFilePath)
        else:
            self.grid = _grid.read(self.meshFilePath)
