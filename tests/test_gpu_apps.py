"""The reference-facing Python API on the GPU: app classes / functions with the reference's names run end to end and
reproduce the reference's results (tests/golden) -- the 'apps run with a backend flag' contract of BASELINE.json."""
import numpy as np
import pytest
import scipy.sparse as sp

import pyefvlib_b200 as P
from tests import util
from tests.cases import HEAT_CASES, STRESS_CASES, BIOT_CASES

pytestmark = pytest.mark.gpu


def bc_dict(spec, names, variables):
    out = {}
    for var in variables:
        d = {"InitialValue": 0.0}
        for n in names:
            kind, val = spec[var][n]
            d[n] = {"condition": P.Dirichlet if kind == "dirichlet" else P.Neumann, "type": P.Constant, "value": val}
        out[var] = d
    return out


@pytest.mark.parametrize("tag", ["Square", "test_mixed", "Hexas3x3x3_3d", "eight_sphere"])
def test_heat_transfer_app(tag):
    from pyefvlib_b200.apps.heat_transfer import heatTransfer
    z = util.load("heat_" + tag)
    c = HEAT_CASES[tag]
    grid = util.product_grid(z)
    names = [b.name for b in grid.boundaries]
    steps = len(z["heatT_fields"])
    pd = P.ProblemData(grid, P.PropertyData(c["props"]), P.BoundaryConditions(bc_dict({"temperature": c["bcs"]}, names, ["temperature"])),
                       P.NumericalSettings(timeStep=float(z["heatT_dt"]), finalTime=None, tolerance=None, maxNumberOfIterations=steps - 1))
    s = heatTransfer(pd, extension="csv", transient=True, verbosity=False,
                     linearSolver="bicgstab" if tag == "test_mixed" else "auto")
    fields = s.saver.fields["temperature"]
    assert len(fields) == steps + 1                                   # initial field + one per step (maxIter=k => k+1 solves)
    for k in range(steps):
        assert util.rel_l2(fields[k + 1], z["heatT_fields"][k]) <= 1e-8
    assert util.row_rel_matrix_error(s.matrix, util.csr_from(z, "heatT_")) <= 1e-12
    assert np.isclose(s.currentTime, steps * float(z["heatT_dt"]))
    # steady mode runs the loop twice and stops with difference 0
    pd2 = P.ProblemData(grid, P.PropertyData(c["props"]), P.BoundaryConditions(bc_dict({"temperature": c["bcs"]}, names, ["temperature"])),
                        P.NumericalSettings(timeStep=0.01, finalTime=None, tolerance=1e-6, maxNumberOfIterations=50))
    s2 = heatTransfer(pd2, transient=False, verbosity=False, linearSolver="bicgstab" if tag == "test_mixed" else "auto")
    assert s2.iteration == 2 and s2.difference < 1e-6
    assert util.rel_l2(s2.temperatureField, z["heat_T"]) <= 1e-8
    assert util.rel_max(s2.independent, z["heat_rhs"]) <= 1e-12


def test_all_neumann_mean_temperature():
    """Known answer of the reference's default configuration (apps/heat_transfer.py:165-197): with q = 100,
    rho cp = 1 and insulated walls the mean temperature after k steps is 100 k dt."""
    from pyefvlib_b200.apps.heat_transfer import heatTransfer
    z = util.load("heat_Square")
    grid = util.product_grid(z)
    bcs = {"temperature": dict({"InitialValue": 0.0}, **{b.name: {"condition": P.Neumann, "type": P.Constant, "value": 0.0} for b in grid.boundaries})}
    pd = P.ProblemData(grid, P.PropertyData(HEAT_CASES["Square"]["props"]), P.BoundaryConditions(bcs),
                       P.NumericalSettings(timeStep=1e-2, finalTime=None, tolerance=None, maxNumberOfIterations=4))
    s = heatTransfer(pd, transient=True, verbosity=False)
    for k, T in enumerate(s.saver.fields["temperature"]):
        assert abs(T.mean() - k) < 1e-8 and np.ptp(T) < 1e-7


@pytest.mark.parametrize("tag", ["Square", "Hexas4x4x4"])
def test_stress_equilibrium_app(tag):
    from pyefvlib_b200.apps.stress_equilibrium import stressEquilibrium
    z = util.load("stress_" + tag)
    c = STRESS_CASES[tag]
    grid = util.product_grid(z)
    names = [b.name for b in grid.boundaries]
    variables = ["u", "v", "w"][:grid.dimension]
    pd = P.ProblemData(grid, P.PropertyData(c["props"]), P.BoundaryConditions(bc_dict(c["bcs"], names, variables)))
    s = stressEquilibrium(pd, gravity=c["gravity"], verbosity=False)
    assert util.rel_l2(s.displacements, z["x_equilibrated"]) <= 1e-8
    n = len(z["rhs"])
    assert util.row_rel_matrix_error(s.matrix, sp.csr_matrix((z["data"], z["indices"], z["indptr"]), shape=(n, n))) <= 1e-12
    assert util.rel_max(s.independent, z["rhs"]) <= 1e-12
    assert set(s.saver.fields) == set(variables)


@pytest.mark.parametrize("tag", ["Square", "Hexas3x3x3"])
def test_geomechanics_app(tag):
    from pyefvlib_b200.apps.geomechanics import geomechanics
    z = util.load("biot_" + tag)
    c = BIOT_CASES[tag]
    grid = util.product_grid(z)
    d, nV = grid.dimension, grid.numberOfVertices
    names = [b.name for b in grid.boundaries]
    variables = ["u_x", "u_y", "u_z"][:d] + ["p"]
    pd = P.ProblemData(grid, P.PropertyData(c["props"]), P.BoundaryConditions(bc_dict(c["bcs"], names, variables)),
                       P.NumericalSettings(timeStep=c["dt"], finalTime=1e30, tolerance=0.0, maxNumberOfIterations=c["steps"]))
    out = geomechanics(pd, verbosity=False)
    ref = z["x"][c["steps"] - 1]                                       # the reference's fields after the same steps
    assert util.rel_l2(out["u"], ref[:d * nV]) <= 1e-6                 # the reference's dense inverse is ~1e-8 accurate itself
    assert util.rel_l2(out["p"], ref[d * nV:]) <= 1e-6
    assert len(out["iterations"]) == c["steps"]


def test_linear_system_csr_gpu_backend():
    """The reference's self-demo (PyEFVLib/LinearSystem.py:186-205) and a solve through the 'gpu' backend."""
    stencil = [[0, 1], [0, 2, 3], [], [0, 4], [], [2, 3, 5]]
    ls = P.LinearSystemCSR(stencil, 1)
    ls.initialize()
    for i, st in enumerate(stencil):
        for j in st:
            ls.addValueToMatrix(i, j, -3)
    ls.addValueToMatrix(3, 0, -12); ls.addValueToMatrix(3, 0, -12); ls.setValueToMatrix(3, 4, -5)
    D = np.asarray(ls.getDense())
    assert D[3].tolist() == [-27, 0, 0, 0, -5, 0] and D[5].tolist() == [0, 0, -3, -3, 0, -3]
    ls.matZeroRow(5, -20)
    assert np.asarray(ls.getDense())[5].tolist() == [0, 0, 0, 0, 0, -20]
    with pytest.raises(Exception, match="New entries not allowed"):
        ls.addValueToMatrix(2, 2, 1.0)
    # a real system: 1-D Laplacian with Dirichlet ends through add/set/matZeroRow, solved on the device
    n = 50
    st = [[j for j in (i - 1, i, i + 1) if 0 <= j < n] for i in range(n)]
    ls = P.LinearSystemCSR(st, 1)
    ls.initialize()
    for i in range(n):
        for j in st[i]:
            ls.addValueToMatrix(i, j, 2.0 if i == j else -1.0)
        ls.addValueToRHS(i, 0.01)
    ls.matZeroRow(0, 1.0); ls.setValueToRHS(0, 1.0)
    ls.matZeroRow(n - 1, 1.0); ls.setValueToRHS(n - 1, 2.0)
    x, iters, relres = ls.solve("bicgstab", 1e-12)
    import scipy.sparse.linalg as spla
    ref = spla.spsolve(sp.csc_matrix(ls.matrix), ls.rhs)
    assert relres <= 1e-12 and util.rel_l2(x, ref) <= 1e-9


@pytest.mark.parametrize("tag", ["Square", "Hexas3x3x3"])
def test_fixed_stress_app(tag):
    """apps/fixed_stress.py: same inner fixed-point iteration counts and the same fields as the reference."""
    from pyefvlib_b200.apps.fixed_stress import fixedStress
    z = util.load("fixed_stress_" + tag)
    c = BIOT_CASES[tag]
    grid = util.product_grid(z)
    d, nV = grid.dimension, grid.numberOfVertices
    names = [b.name for b in grid.boundaries]
    variables = ["u_x", "u_y", "u_z"][:d] + ["p"]
    steps = len(z["p"])
    pd = P.ProblemData(grid, P.PropertyData(c["props"]), P.BoundaryConditions(bc_dict(c["bcs"], names, variables)),
                       P.NumericalSettings(timeStep=float(z["dt"]), finalTime=float(z["dt"]) * steps - 1e-9,
                                           tolerance=float(z["tolerance"]), maxNumberOfIterations=10 ** 9))
    out = fixedStress(pd, verbosity=False)
    assert out["inner_iterations"] == z["inner_iterations"].tolist()
    for k in range(steps):
        assert util.rel_l2(out["saver"].fields["p"][k], z["p"][k]) <= 1e-7
        u = np.concatenate([out["saver"].fields[v][k] for v in variables[:d]])
        assert util.rel_l2(u, z["u"][k]) <= 1e-7


@pytest.mark.parametrize("tag,saverType", [("Prisms_3d", "default"), ("Pyrams_3d", "stream")])
def test_heat_app_writes_vtu_results(tag, saverType, tmp_path):
    """App run with an output path on the prism / pyramid meshes: `extension="vtu"` writes the last field like the
    reference's VtuSaver, `saverType="stream"` one file per saved step; the written numbers are the reference's fields."""
    import os
    import xml.etree.ElementTree as ET
    from pyefvlib_b200.apps.heat_transfer import heatTransfer
    z = util.load("heat_" + tag)
    c = HEAT_CASES[tag]
    grid = util.product_grid(z)
    names = [b.name for b in grid.boundaries]
    steps = len(z["heatT_fields"])
    pd = P.ProblemData(meshFilePath=grid, outputFilePath=str(tmp_path),
                       numericalSettings=P.NumericalSettings(timeStep=float(z["heatT_dt"]), finalTime=None, tolerance=None,
                                                             maxNumberOfIterations=steps - 1),
                       propertyData=P.PropertyData(c["props"]),
                       boundaryConditions=P.BoundaryConditions(bc_dict({"temperature": c["bcs"]}, names, ["temperature"])))
    s = heatTransfer(pd, extension="vtu", saverType=saverType, transient=True, verbosity=False)
    assert s.saver.finalized

    def point_data(path):
        piece = ET.parse(path).getroot().find("UnstructuredGrid").find("Piece")
        assert int(piece.get("NumberOfCells")) == grid.numberOfElements
        types = np.array(piece.find("Cells").findall("DataArray")[2].text.split(), dtype=int)
        assert (types == (13 if tag.startswith("Prisms") else 14)).all()
        return np.array(piece.find("PointData").find("DataArray").text.split(), dtype=np.float64)

    if saverType == "default":
        assert util.rel_l2(point_data(s.saver.outputPath), z["heatT_fields"][-1]) <= 1e-8
    else:
        sets = ET.parse(s.saver.outputPath).getroot().find("Collection").findall("DataSet")
        assert len(sets) == steps + 1                                 # initial field + one per step
        for k in range(steps):
            data = point_data(os.path.join(str(tmp_path), sets[k + 1].get("file")))
            assert util.rel_l2(data, z["heatT_fields"][k]) <= 1e-8


def test_heat_variable_neumann_matches_reference():
    """Expression-valued Neumann conditions: RHS and field of the UNMODIFIED reference (oracle/make_golden_variable_neumann.py),
    including its evaluation point (the vertex numbered like the outer face's global counter, apps/heat_transfer.py:120)."""
    from pyefvlib_b200.apps.heat_transfer import heatTransfer
    z = util.load("heat_Square")
    g = np.load(util.GOLDEN + "/heat_Square_variable_neumann.npz")
    grid = util.product_grid(z)
    expr = dict(zip([str(s) for s in g["expr_names"]], [str(s) for s in g["expr"]]))
    bc = {"InitialValue": 0.0}
    for b in grid.boundaries:
        if b.name in expr:
            bc[b.name] = {"condition": P.Neumann, "type": P.Variable, "value": expr[b.name]}
        elif b.name == "West":
            bc[b.name] = {"condition": P.Dirichlet, "type": P.Constant, "value": 300.0}
        else:
            bc[b.name] = {"condition": P.Neumann, "type": P.Constant, "value": 0.0}
    pd = P.ProblemData(grid, P.PropertyData(HEAT_CASES["Square"]["props"]), P.BoundaryConditions({"temperature": bc}),
                       P.NumericalSettings(timeStep=0.01, finalTime=None, tolerance=1e-6, maxNumberOfIterations=5))
    s = heatTransfer(pd, transient=False, verbosity=False)
    assert util.rel_max(s.independent, g["rhs"]) <= 1e-12
    assert util.rel_l2(s.temperatureField, g["T"]) <= 1e-8
