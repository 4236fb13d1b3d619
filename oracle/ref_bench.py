"""TEST / BENCH INFRASTRUCTURE ONLY -- times the UNMODIFIED reference (through oracle/ref_harness.py) and the numpy
oracle port on the synthetic heat problem of bench.py.  Only bench.py's CPU legs (`cpu_baseline`, `--impl reference`)
call this; the product package never does.

What one "step" is on the CPU side (same definition as the GPU arm: assemble + BCs + RHS + solve of one time step):
  reference: HeatTransferSolver.assembleMatrix() [element loops + Dirichlet filter + sparse inverse,
             apps/heat_transfer.py:40-87] + addToIndependentVector() [:89-132] + solveLinearSystem() [:134-135];
             Grid() (PyEFVLib/Grid.py:18-65) is the one-time setup and is reported separately, like the GPU arm's setup_ms
  port:      oracle geometry + vectorised assembly + SuperLU factor/solve (what `sparse.linalg.inv` computes, without
             forming the dense inverse)
"""
import os
import tempfile
import time

import numpy as np

from . import efv_oracle as O
from . import ref_harness as R

BOUNDARY_ORDER = ["West", "East", "South", "North", "Bottom", "Top"]


def write_msh(path, gd):
    """Gmsh 2.2 ASCII file of an OracleGridData (hexahedra + quadrilateral boundary facets), laid out like the bundled
    meshes/msh/3D/Hexas.msh: boundaries first, then the region."""
    nb = len(gd.boundariesNames)
    with open(path, "w") as f:
        f.write("$MeshFormat\n2.2 0 8\n$EndMeshFormat\n$PhysicalNames\n%d\n" % (nb + len(gd.regionsNames)))
        for k, name in enumerate(gd.boundariesNames):
            f.write('2 %d "%s"\n' % (k + 1, name))
        for k, name in enumerate(gd.regionsNames):
            f.write('3 %d "%s"\n' % (nb + k + 1, name))
        f.write("$EndPhysicalNames\n$Nodes\n%d\n" % len(gd.verticesCoordinates))
        for i, (x, y, z) in enumerate(gd.verticesCoordinates):
            f.write("%d %.17g %.17g %.17g\n" % (i + 1, x, y, z))
        nfac = sum(len(b) for b in gd.boundariesIndexes)
        f.write("$EndNodes\n$Elements\n%d\n" % (nfac + len(gd.elementsConnectivities)))
        eid = 1
        for k, idx in enumerate(gd.boundariesIndexes):
            for q in idx:
                c = gd.boundariesConnectivities[q]
                f.write("%d %d 2 %d %d %s\n" % (eid, 3 if len(c) == 4 else 2, k + 1, k + 1, " ".join(str(v + 1) for v in c)))
                eid += 1
        for k, idx in enumerate(gd.regionsElementsIndexes):
            for q in idx:
                c = gd.elementsConnectivities[q]
                f.write("%d %d 2 %d %d %s\n" % (eid, 5 if len(c) == 8 else 4, nb + k + 1, nb + k + 1, " ".join(str(v + 1) for v in c)))
                eid += 1
        f.write("$EndElements\n")


class ReferenceHeatCase:
    """The reference's own HeatTransferSolver on the n^3-vertex structured hexahedral mesh with bench.py's data."""

    def __init__(self, n, props, bcs, dt):
        ns = R.load()
        P = ns["PyEFVLib"]
        self.n = n
        gd = O.structured_grid_data(n, "hex")
        tmp = tempfile.mkdtemp(prefix="efv_ref_")
        path = os.path.join(tmp, "hex%d.msh" % n)
        write_msh(path, gd)
        bc = {"InitialValue": 0.0}
        for name, (kind, val) in bcs.items():
            bc[name] = {"condition": P.Dirichlet if kind == "dirichlet" else P.Neumann, "type": P.Constant, "value": val}
        t0 = time.perf_counter()
        # ProblemData.py:152-153 drops the leading "/" of absolute paths: ref_harness.load() has chdir'ed to "/"
        self.pd = P.ProblemData(meshFilePath=path, outputFilePath="tmp/efv_ref_bench",
                                numericalSettings=P.NumericalSettings(timeStep=dt, finalTime=None, tolerance=None, maxNumberOfIterations=1),
                                propertyData=P.PropertyData(props), boundaryConditions=P.BoundaryConditions({"temperature": bc}))
        self.grid_s = time.perf_counter() - t0                      # MSH parse + Grid() (Grid.py:18-65)
        self.solver = ns["heat_transfer"].HeatTransferSolver(self.pd, extension="csv", transient=True, verbosity=False)
        self.solver.settings()
        self.solver.init()
        self.vertices = self.pd.grid.numberOfVertices

    def step(self):
        s = self.solver
        t0 = time.perf_counter()
        s.coords, s.matrixVals = [], []                               # assembleMatrix appends to these lists (heat_transfer.py:36-38)
        s.assembleMatrix()
        t1 = time.perf_counter()
        s.addToIndependentVector()
        s.solveLinearSystem()
        t2 = time.perf_counter()
        s.prevTemperatureField = s.temperatureField
        return dict(total_s=t2 - t0, assemble_s=t1 - t0, rhs_solve_s=t2 - t1)


def port_step(n, props, bcs, dt):
    """One pass of the oracle port on an n^3-vertex sample (geometry, assembly, SuperLU factor + solve)."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    t0 = time.perf_counter()
    og = O.OracleGrid(O.structured_grid_data(n, "hex"))
    t1 = time.perf_counter()
    A = O.heat_matrix(og, props, bcs, transient=True, dt=dt)
    lu = spla.splu(sp.csc_matrix(A))
    b = O.heat_rhs(og, props, bcs, transient=True, dt=dt, Told=np.zeros(og.numberOfVertices))
    T = lu.solve(b)
    t2 = time.perf_counter()
    assert np.isfinite(T).all()
    return dict(total_s=t2 - t1, grid_s=t1 - t0, vertices=og.numberOfVertices, field=T)
