"""TEST INFRASTRUCTURE ONLY -- golden vector for VARIABLE (expression) Neumann conditions, produced by the UNMODIFIED
reference (apps/heat_transfer.py:115-120 through oracle/ref_harness.py).  Run in the build container:
    python -m oracle.make_golden_variable_neumann
The reference evaluates the expression at the vertex whose number equals the outer face's global counter
(bCondition.getValue(outerFace.handle), Boundary.py:74) -- the fixture pins exactly that behaviour."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_harness as R  # noqa: E402

ns = R.load()
P = ns["PyEFVLib"]
EXPR = {"North": "10*x+3*y+1", "South": "x*x"}


def main():
    path = R.mesh_path("2D/Square.msh")
    names = P.MSHReader(path).getData().boundariesNames
    bc = {"InitialValue": 0.0}
    for n in names:
        if n in EXPR:
            bc[n] = {"condition": P.Neumann, "type": P.Variable, "value": EXPR[n]}
        elif n == "West":
            bc[n] = {"condition": P.Dirichlet, "type": P.Constant, "value": 300.0}
        else:
            bc[n] = {"condition": P.Neumann, "type": P.Constant, "value": 0.0}
    pd = P.ProblemData(meshFilePath=path, outputFilePath="tmp/efv_golden",
                       numericalSettings=P.NumericalSettings(timeStep=0.01, finalTime=None, tolerance=1e-6, maxNumberOfIterations=5),
                       propertyData=P.PropertyData({"Body": {"HeatCapacity": 1.0, "Conductivity": 1.0, "Density": 1.0, "HeatGeneration": 100.0}}),
                       boundaryConditions=P.BoundaryConditions({"temperature": bc}))
    s = ns["heat_transfer"].HeatTransferSolver(pd, extension="csv", transient=False, verbosity=False)
    s.settings(); s.init(); s.assembleMatrix(); s.addToIndependentVector(); s.solveLinearSystem()
    out = os.path.join(ROOT, "tests", "golden", "heat_Square_variable_neumann.npz")
    np.savez_compressed(out, rhs=s.independent.copy(), T=s.temperatureField.copy(), expr_names=np.array(list(EXPR)), expr=np.array(list(EXPR.values())))
    print("wrote", out, "rhs sum", s.independent.sum(), "T range", s.temperatureField.min(), s.temperatureField.max())


if __name__ == "__main__":
    main()
