"""Fully coupled Biot poroelasticity on the GPU: the time loop of apps/geomechanics.py:243-285 with the dense
(1+d)N x (1+d)N matrix, its np.linalg.inv and the per-step dense mat-vec (geomechanics.py:24,140,254) replaced by a
BSR system assembled by efv_biot_assemble and a BiCGStab solve per time step.  Fields stay on the device between steps."""
import numpy as np

from ..device import DeviceSystem, DIRICHLET_ROWS, VEC_X, VEC_XOLD
from ..solver import Saver

PROPERTY_NAMES = ["nu", "G", "cs", "phi", "k", "cf", "mu", "rhos", "rhof"]


def send_boundary_conditions(system, problemData):
    grid = problemData.grid
    nV, d = grid.numberOfVertices, grid.dimension
    names = ["u_x", "u_y", "u_z"][:d] + ["p"]
    system.bc_clear()
    for i, var in enumerate(names):                                         # geomechanics.py:202-221
        for bc in problemData.neumannBoundaries[var]:
            if bc.expression:
                raise Exception("Variable Neumann conditions are not supported on the GPU backend")
            system.bc_neumann(bc.boundary.handle, i, bc.value, 1.0)
    for i, var in enumerate(names):                                         # geomechanics.py:117-137, 224-239
        for bc in problemData.dirichletBoundaries[var]:
            system.bc_dirichlet(bc.boundary.verticesIndexes + i * nV, bc.values_on_vertices())


def geomechanics(problemData, rtol=1e-10, maxLinearIterations=20000, verbosity=True, saver=None, preconditioner="auto"):
    """Returns dict(u=..., p=..., iterations=[...], saver=...) after the reference's stopping rule
    (geomechanics.py:279-282)."""
    grid = problemData.grid
    nV, d = grid.numberOfVertices, grid.dimension
    timeStep = problemData.timeStep
    saver = saver or Saver(grid, problemData.outputFilePath, problemData.libraryPath)
    system = DeviceSystem(grid, d + 1)
    if preconditioner == "auto":          # block multigrid for BiCGStab from 50 000 rows up, the diagonal below
        preconditioner = "amg" if system.rows >= 50000 else "jacobi"
    system.set_preconditioner(preconditioner)
    names = ["u_x", "u_y", "u_z"][:d]
    old = np.concatenate([np.asarray(problemData.initialValues[v], dtype=np.float64) for v in names] +
                         [np.asarray(problemData.initialValues["p"], dtype=np.float64)])
    send_boundary_conditions(system, problemData)
    system.assemble_dirichlet_mode(DIRICHLET_ROWS)
    system.biot_assemble(problemData.propertyData.table(PROPERTY_NAMES), timeStep)
    system.upload(VEC_XOLD, old)
    system.upload(VEC_X, old)
    tolerance = problemData.tolerance
    iteration, currentTime, converged = 0, 0.0, False
    lin_iters = []
    results = old
    while not converged:
        system.build_rhs(True)
        iters, relres = system.solve_bicgstab(rtol, maxLinearIterations)
        if not relres <= rtol:
            raise Exception(f"BiCGStab stalled at relative residual {relres:.2e} after {iters} iterations")
        lin_iters.append(iters)
        difference = system.max_abs_diff()                                  # geomechanics.py:258
        system.advance()
        currentTime += timeStep
        iteration += 1
        results = system.download(VEC_X)
        for i, var in enumerate(names):
            saver.save(var, results[i * nV:(i + 1) * nV], currentTime)
        saver.save("p", results[d * nV:], currentTime)
        if verbosity:
            print("{:>9}\t{:>14.2e}\t{:>14.2e}\t{:>14.2e}".format(iteration, currentTime, timeStep, difference))
        converged = (difference <= tolerance) or (currentTime >= problemData.finalTime) or (iteration >= problemData.maxNumberOfIterations)
        if iteration >= problemData.maxNumberOfIterations:
            break
    saver.finalize()
    return dict(u=results[:d * nV], p=results[d * nV:], iterations=lin_iters, saver=saver, system=system)
