"""Fixed-stress split of the Biot problem on the GPU: the loops of apps/fixed_stress.py:284-329 with the two dense
matrices and their inverses (fixed_stress.py:40-41,115-116) replaced by two sparse systems -- a heat-type `mass` system
for the pressure and an elasticity `geo` system for the displacements, both assembled by the kernels the heat and
stress apps use -- and the coupling terms of the right-hand sides taken from the blocks of the coupled operator
(efv_fs_mass_rhs / efv_fs_geo_rhs).  All fields stay on the device; the host only sees the convergence scalar."""
import ctypes as C

import numpy as np

from .. import _lib
from ..device import DeviceSystem, DIRICHLET_SYMMETRIC, VEC_X, VEC_XOLD
from ..solver import Saver
from .geomechanics import PROPERTY_NAMES


def _derived(p):
    """fixed_stress.py:56-62 including the precedence quirk of K."""
    K = 2 * p["G"] * (1 + p["nu"]) / 3 * (1 - 2 * p["nu"])
    cb = 1 / K
    alpha = 1 - p["cs"] / cb
    S = p["phi"] * p["cf"] + (alpha - p["phi"]) * p["cs"]
    return cb, alpha, S


def fixedStress(problemData, rtol=1e-12, maxLinearIterations=20000, verbosity=True, saver=None):
    grid = problemData.grid
    nV, d = grid.numberOfVertices, grid.dimension
    dt = problemData.timeStep
    L = _lib.load()
    saver = saver or Saver(grid, problemData.outputFilePath, problemData.libraryPath)
    pdta = problemData.propertyData
    names_u = ["u_x", "u_y", "u_z"][:d]
    # coupled operator (unmodified): source of the coupling blocks and holder of the iterate / previous step
    biot = DeviceSystem(grid, d + 1)
    biot.bc_clear()
    biot.assemble_dirichlet_mode(-1)
    biot.biot_assemble(pdta.table(PROPERTY_NAMES), dt)
    # mass system: -k mu grad p flux + (S + cb alpha^2)/dt vol on the diagonal (fixed_stress.py:64-77)
    mass = DeviceSystem(grid, 1)
    mprops = []
    for h in range(len(grid.regions)):
        p = {k: float(pdta.get(h, k)) for k in PROPERTY_NAMES}
        cb, alpha, S = _derived(p)
        mprops.append([p["k"] * p["mu"], S + cb * alpha ** 2, 1.0, 0.0])
    mass.bc_clear()
    for bc in problemData.neumannBoundaries["p"]:
        mass.bc_neumann(bc.boundary.handle, 0, bc.value, 1.0)
    for bc in problemData.dirichletBoundaries["p"]:
        mass.bc_dirichlet(bc.boundary.verticesIndexes, bc.values_on_vertices())
    mass.assemble_dirichlet_mode(DIRICHLET_SYMMETRIC)
    mass.heat_assemble(np.array(mprops), dt, True)
    # geo system: At Ce Bv (fixed_stress.py:80-92)
    geo = DeviceSystem(grid, d)
    gprops = [[float(pdta.get(h, "G")), float(pdta.get(h, "nu")), 0.0, 0.0] for h in range(len(grid.regions))]
    geo.bc_clear()
    for i, var in enumerate(names_u):
        for bc in problemData.neumannBoundaries[var]:
            geo.bc_neumann(bc.boundary.handle, i, bc.value, 1.0)           # `+=` here (fixed_stress.py:251-265)
    for i, var in enumerate(names_u):
        for bc in problemData.dirichletBoundaries[var]:
            geo.bc_dirichlet(bc.boundary.verticesIndexes + i * nV, bc.values_on_vertices())
    geo.assemble_dirichlet_mode(DIRICHLET_SYMMETRIC)
    geo.elasticity_assemble(np.array(gprops), False)
    prev = np.concatenate([np.asarray(problemData.initialValues[v], dtype=np.float64) for v in names_u] +
                          [np.asarray(problemData.initialValues["p"], dtype=np.float64)])
    biot.upload(VEC_XOLD, prev)
    biot.upload(VEC_X, prev)
    mass.upload(VEC_X, prev[d * nV:])
    geo.upload(VEC_X, prev[:d * nV])
    tolerance = problemData.tolerance
    currentTime, converged = 0.0, False
    inner_log, lin_log = [], []
    results = prev
    diff = C.c_double(0.0)
    while not converged:
        difference, iteration = 2 * tolerance, 0
        while difference >= tolerance or iteration <= 1:                     # fixed_stress.py:292
            _lib.check(L.efv_fs_mass_rhs(biot.handle, mass.handle))
            it_m, rr = mass.solve_pcg(rtol, maxLinearIterations)
            if not rr <= rtol:
                it_m, rr = mass.solve_bicgstab(rtol, maxLinearIterations)
            _lib.check(L.efv_fs_geo_rhs(biot.handle, mass.handle, geo.handle))
            it_g, rr = geo.solve_pcg(rtol, maxLinearIterations, negate=True)
            if not rr <= rtol:
                it_g, rr = geo.solve_bicgstab(rtol, maxLinearIterations)
            _lib.check(L.efv_fs_update(biot.handle, mass.handle, geo.handle, C.byref(diff)))
            difference = diff.value
            lin_log.append((it_m, it_g))
            iteration += 1
        difference = biot.max_abs_diff()                                     # fixed_stress.py:306
        biot.advance()
        currentTime += dt
        results = biot.download(VEC_X)
        for i, var in enumerate(names_u):
            saver.save(var, results[i * nV:(i + 1) * nV], currentTime)
        saver.save("p", results[d * nV:], currentTime)
        inner_log.append(iteration)
        if verbosity:
            print("{:>9}\t{:>14.2e}\t{:>14.2e}\t{:>14.2e}".format(iteration, currentTime, dt, difference))
        converged = (difference <= tolerance) or (currentTime >= problemData.finalTime) or (iteration >= problemData.maxNumberOfIterations)
        if iteration >= problemData.maxNumberOfIterations:
            break
    saver.finalize()
    return dict(u=results[:d * nV], p=results[d * nV:], inner_iterations=inner_log, linear_iterations=lin_log, saver=saver)
