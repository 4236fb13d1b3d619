"""The reference-side binding documented in INTEGRATION.md section 3 is executed as written: the code block is
extracted from the document, pointed at the in-tree shared library, and driven with objects that carry the reference's
attribute names (GridData lists, BC objects with .boundary/.value/.getValue).  Its field must equal the reference's
first transient step (golden fixture)."""
import os
import re
import types

import numpy as np
import pytest

from tests import util
from tests.cases import HEAT_CASES

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def stub_namespace():
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    block = re.search(r"## 3\..*?```python\n(.*?)```", text, re.S).group(1)
    lib_path = os.path.join(ROOT, "pyefvlib_b200", "libefv_b200.so")
    block = block.replace('C.CDLL("libefv_b200.so")', f'C.CDLL("{lib_path}")')
    ns = {}
    exec(compile(block, "INTEGRATION.md#3", "exec"), ns)
    return ns


@pytest.mark.parametrize("tag", ["Square", "Hexas3x3x3_3d"])
def test_documented_binding_reproduces_the_reference_step(tag):
    import pyefvlib_b200 as P
    z = util.load("heat_" + tag)
    c = HEAT_CASES[tag]
    grid = util.product_grid(z)
    names = [b.name for b in grid.boundaries]
    bcd = {"temperature": dict({"InitialValue": 0.0},
                               **{n: {"condition": P.Dirichlet if c["bcs"][n][0] == "dirichlet" else P.Neumann, "type": P.Constant,
                                      "value": c["bcs"][n][1]} for n in names})}
    pd = P.ProblemData(meshFilePath=grid, outputFilePath=None,
                       numericalSettings=P.NumericalSettings(timeStep=float(z["heatT_dt"]), finalTime=None, tolerance=1e-6,
                                                             maxNumberOfIterations=1),
                       propertyData=P.PropertyData(c["props"]), boundaryConditions=P.BoundaryConditions(bcd))
    ns = stub_namespace()
    gpu = ns["GpuHeat"](grid)                     # only uses grid.gridData.* lists, like the reference's Grid
    solver = types.SimpleNamespace(problemData=pd, grid=grid, timeStep=float(z["heatT_dt"]), transient=True,
                                   prevTemperatureField=np.zeros(grid.numberOfVertices))
    gpu.assemble(solver)
    T1 = gpu.step(solver)
    assert util.rel_l2(T1, z["heatT_fields"][0]) <= 1e-8
    solver.prevTemperatureField = T1
    T2 = gpu.step(solver)
    assert util.rel_l2(T2, z["heatT_fields"][1]) <= 1e-8
