"""StressEquilibriumSolver on the GPU: same class, hooks and result attributes as apps/stress_equilibrium.py:9-185.

assembleLinearSystem = CUDA kernels (efv_elasticity_assemble + BC kernels); solveLinearSystem = Jacobi-PCG on the
negated, symmetrically eliminated system (the reference's operator has a negative diagonal and +1 Dirichlet rows,
SURVEY.md fact 7 / appendix B.7) instead of the sparse inverse of stress_equilibrium.py:170-173."""
import numpy as np

from ..device import DeviceSystem, DIRICHLET_ROWS, DIRICHLET_SYMMETRIC, VEC_X, VEC_B
from ..solver import Solver

PROPERTY_NAMES = ["ShearModulus", "PoissonsRatio", "Density", "Gravity"]


class StressEquilibriumSolver(Solver):
    def __init__(self, workspaceDirectory, gravity=False, rtol=1e-10, maxLinearIterations=50000, **kwargs):
        Solver.__init__(self, workspaceDirectory, **kwargs)
        self.gravity, self.rtol, self.maxLinearIterations = gravity, rtol, maxLinearIterations

    def init(self):
        self.system = DeviceSystem(self.grid, self.dimension)
        self.displacements = np.repeat(0.0, self.dimension * self.numberOfVertices)

    def mainloop(self):
        self.assembleLinearSystem()
        self.solveLinearSystem()
        self.saveIterationResults()

    def _send_boundary_conditions(self, system):
        nV = self.numberOfVertices
        names = ["u", "v", "w"][:self.dimension]
        system.bc_clear()
        for bc in self.problemData.boundaryConditions:                      # stress_equilibrium.py:119-163
            boundary = bc["u"].boundary
            for i, var in enumerate(names):
                if bc[var].__type__ == "NEUMANN":
                    if bc[var].expression:
                        raise Exception("Variable Neumann conditions are not supported on the GPU backend")
                    system.bc_neumann(boundary.handle, i, bc[var].value, -1.0)   # `-=` in the reference (:131-141)
            for i, var in enumerate(names):
                if bc[var].__type__ == "DIRICHLET":
                    system.bc_dirichlet(boundary.verticesIndexes + i * nV, bc[var].values_on_vertices())

    def _assemble(self, system, mode):
        self._send_boundary_conditions(system)
        system.assemble_dirichlet_mode(mode)
        system.elasticity_assemble(self.propertyData.table(PROPERTY_NAMES), self.gravity)
        system.build_rhs(False)

    def assembleLinearSystem(self):
        self._assemble(self.system, DIRICHLET_SYMMETRIC)

    @property
    def matrix(self):
        tmp = DeviceSystem(self.grid, self.dimension)
        self._assemble(tmp, DIRICHLET_ROWS)
        A = tmp.to_scipy()
        A.eliminate_zeros()
        return A

    @property
    def independent(self):
        tmp = DeviceSystem(self.grid, self.dimension)
        self._assemble(tmp, DIRICHLET_ROWS)
        return tmp.download(VEC_B)

    def solveLinearSystem(self):
        self.system.upload(VEC_X, np.zeros(self.system.rows))
        iters, relres = self.system.solve_pcg(self.rtol, self.maxLinearIterations, negate=True)
        if not relres <= self.rtol:            # non-symmetric operator (e.g. distorted hexahedra): BiCGStab on the same system
            iters2, relres = self.system.solve_bicgstab(self.rtol, self.maxLinearIterations)
            iters += iters2
        self.linearIterations, self.linearResidual = iters, relres
        self.displacements = self.system.download(VEC_X)

    def saveIterationResults(self):
        nV = self.numberOfVertices
        for i, name in enumerate(["u", "v", "w"][:self.dimension]):
            self.saver.save(name, self.displacements[i * nV:(i + 1) * nV], self.currentTime)


def stressEquilibrium(workspaceDirectory, solve=True, extension="csv", gravity=False, verbosity=False, **kwargs):
    solver = StressEquilibriumSolver(workspaceDirectory, extension=extension, gravity=gravity, verbosity=verbosity, **kwargs)
    if solve:
        solver.solve()
    return solver
