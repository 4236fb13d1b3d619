"""HeatTransferSolver on the GPU: same class, hooks and result attributes as apps/heat_transfer.py:10-158.

What changes is where the work happens: the element loops of assembleMatrix / addToIndependentVector are CUDA
kernels (efv_heat_assemble, efv_bc_*, efv_build_rhs) and the sparse inverse + dense mat-vec of
invertMatrix / solveLinearSystem (heat_transfer.py:78-81,134-135) is a Jacobi-PCG (or BiCGStab on
non-symmetric operators) on the device; fields stay in HBM between time steps."""
import numpy as np

from ..device import DeviceSystem, DIRICHLET_ROWS, DIRICHLET_SYMMETRIC, VEC_X, VEC_B, VEC_XOLD
from ..solver import Solver

PROPERTY_NAMES = ["Conductivity", "Density", "HeatCapacity", "HeatGeneration"]


class HeatTransferSolver(Solver):
    def __init__(self, workspaceDirectory, rtol=1e-10, maxLinearIterations=20000, linearSolver="auto", preconditioner="auto", **kwargs):
        Solver.__init__(self, workspaceDirectory, **kwargs)
        self.rtol, self.maxLinearIterations, self.linearSolver = rtol, maxLinearIterations, linearSolver
        self.preconditioner = preconditioner          # of the CG solve: "jacobi", "amg", or "auto" (multigrid from 50 000 rows up)
        self.linearIterations = []

    def init(self):
        self.system = DeviceSystem(self.grid, 1)
        pc = self.preconditioner
        if pc == "auto":
            pc = "amg" if self.system.rows >= 50000 else "jacobi"
        self.system.set_preconditioner(pc)
        self._T = None                                                # host copy of the current field (downloaded on demand)
        self._Tprev = np.asarray(self.problemData.initialValues["temperature"], dtype=np.float64)
        self.system.upload(VEC_XOLD, self._Tprev)
        self.system.upload(VEC_X, self._Tprev)                        # initial guess of the first solve
        if self.saver.wants_fields:
            self.saver.save("temperature", self._Tprev, self.currentTime)
        self.difference = 0.0
        self._matrix_mode = None
        self._symmetric = True
        self._pending_save = None                                     # time stamp of the field being downloaded asynchronously

    # Fields live in HBM.  The attributes of the reference (heat_transfer.py:17-19,135) are
    # host arrays; here they are downloaded when somebody reads them.
    @property
    def temperatureField(self):
        if self._T is None:
            self._T = self.system.download(VEC_X)
        return self._T

    @temperatureField.setter
    def temperatureField(self, value):
        self._T = value

    @property
    def prevTemperatureField(self):
        if self._Tprev is None:
            self._Tprev = self.system.download(VEC_XOLD)
        return self._Tprev

    @prevTemperatureField.setter
    def prevTemperatureField(self, value):
        self._Tprev = value

    def mainloop(self):
        self.assembleMatrix()
        while not self.converged:
            self.addToIndependentVector()
            self.solveLinearSystem()
            self.printIterationData()
            self.currentTime += self.timeStep
            self.checkConvergence()              # needs X and XOLD: before the advance (the reference checks after saving)
            self.system.advance()
            self._Tprev = self._T                # host copies, if anybody downloaded them
            self.saveIterationResults()
            self.iteration += 1
        self._drain_save()
        if not self.saver.wants_fields:          # savers that keep no history get the final field, once
            self.saver.save("temperature", self.temperatureField, self.currentTime)

    # -- boundary conditions in the reference's visiting order ----------------------------------------------------
    def _send_neumann(self, system):
        for bc in self.problemData.neumannBoundaries["temperature"]:         # heat_transfer.py:115-120
            if bc.expression:
                system.bc_neumann_expression(bc, 0, 0.0, 1.0)         # getValue(handle) without a time: t = 0 (heat_transfer.py:120)
            else:
                system.bc_neumann(bc.boundary.handle, 0, bc.value, 1.0)

    def _send_boundary_conditions(self):
        s = self.system
        s.bc_clear()
        self._send_neumann(s)
        for bc in self.problemData.dirichletBoundaries["temperature"]:       # heat_transfer.py:70-76, 122-126
            s.bc_dirichlet(bc.boundary.verticesIndexes, bc.values_on_vertices(self.currentTime))

    def assembleMatrix(self, mode=DIRICHLET_SYMMETRIC):
        props = self.propertyData.table(PROPERTY_NAMES)
        self._send_boundary_conditions()
        self.system.assemble_dirichlet_mode(mode)          # Dirichlet rows/columns handled inside the assembly kernel
        self.system.heat_assemble(props, self.timeStep, self.transient)
        self._matrix_mode = mode
        # CG or BiCGStab is decided here, from the assembled operator (general quadrilaterals / hexahedra give a
        # non-symmetric EbFVM matrix), not after a failed CG run
        self._symmetric = mode == DIRICHLET_SYMMETRIC and (self.linearSolver != "auto" or self.system.symmetry_defect() <= 1e-12)

    @property
    def matrix(self):
        """The reference's `solver.matrix` (row-replaced Dirichlet form, heat_transfer.py:70-80) as scipy CSR,
        re-assembled on demand into a second system so that the solver's symmetric form stays intact."""
        tmp = DeviceSystem(self.grid, 1)
        tmp.heat_assemble(self.propertyData.table(PROPERTY_NAMES), self.timeStep, self.transient)
        for bc in self.problemData.dirichletBoundaries["temperature"]:
            tmp.bc_dirichlet(bc.boundary.verticesIndexes, bc.values_on_vertices(self.currentTime))
        tmp.apply_dirichlet(DIRICHLET_ROWS)
        A = tmp.to_scipy()
        A.eliminate_zeros()
        return A

    def addToIndependentVector(self):
        self.system.build_rhs(self.transient)

    @property
    def independent(self):
        """RHS of the reference's system (heat_transfer.py:89-132): in symmetric mode the eliminated-column term is
        removed again so that the vector matches `solver.independent` of the reference."""
        if self._matrix_mode == DIRICHLET_ROWS:
            return self.system.download(VEC_B)
        tmp = DeviceSystem(self.grid, 1)
        tmp.heat_assemble(self.propertyData.table(PROPERTY_NAMES), self.timeStep, self.transient)
        self._send_neumann(tmp)
        for bc in self.problemData.dirichletBoundaries["temperature"]:
            tmp.bc_dirichlet(bc.boundary.verticesIndexes, bc.values_on_vertices(self.currentTime))
        tmp.apply_dirichlet(DIRICHLET_ROWS)
        tmp.upload(VEC_XOLD, self.prevTemperatureField)
        tmp.build_rhs(self.transient)
        return tmp.download(VEC_B)

    def solveLinearSystem(self):
        solver = self.linearSolver
        if solver == "auto":
            solver = "pcg" if self._symmetric else "bicgstab"
        if solver == "pcg":
            iters, relres = self.system.solve_pcg(self.rtol, self.maxLinearIterations)
            if not relres <= self.rtol:          # CG broke down (x is its last good iterate): finish with BiCGStab
                iters2, relres = self.system.solve_bicgstab(self.rtol, self.maxLinearIterations)
                iters += iters2
        else:
            iters, relres = self.system.solve_bicgstab(self.rtol, self.maxLinearIterations)
        if not relres <= self.rtol:
            raise Exception(f"linear solver stalled at relative residual {relres:.2e} after {iters} iterations")
        self.linearIterations.append(iters)
        self.linearResidual = relres
        self._T = None                           # the new field stays on the device until somebody asks for it

    def _drain_save(self):
        if self._pending_save is not None:
            self.saver.save("temperature", self.system.download_xold_end(), self._pending_save)
            self._pending_save = None

    def saveIterationResults(self):
        """Called right after the advance: XOLD holds this step's field and stays constant during the next step, so it is
        downloaded on the copy stream while the next step runs; the saver receives it one step later (same order, same
        time stamps as heat_transfer.py:137-138)."""
        if not self.saver.wants_fields:
            return
        self._drain_save()
        self.system.download_xold_begin()
        self._pending_save = self.currentTime

    def checkConvergence(self):
        self.converged = False
        if self.tolerance is not None or self.verbosity:                           # one scalar read back; skipped when nobody uses it
            self.difference = self.system.max_abs_diff()                           # heat_transfer.py:142
        if self.problemData.finalTime is not None and self.currentTime > self.problemData.finalTime:
            self.converged = True
            return
        if self.iteration > 0 and self.tolerance is not None and self.difference < self.tolerance:
            self.converged = True
            return
        if self.problemData.maxNumberOfIterations and self.iteration >= self.problemData.maxNumberOfIterations:
            self.converged = True
            return


def heatTransfer(problemData, solve=True, extension="csv", saverType="default", transient=True, verbosity=True, **kwargs):
    solver = HeatTransferSolver(problemData, extension=extension, saverType=saverType, transient=transient, verbosity=verbosity, **kwargs)
    if solve:
        solver.solve()
    return solver
