"""Host-side logic (no GPU): C-ABI surface, MSH reader, flat Grid, ProblemData / BoundaryConditions mirrors."""
import os
import re

import numpy as np
import pytest

import pyefvlib_b200 as P
from pyefvlib_b200 import _lib
from oracle import efv_oracle as O
from tests import util

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "efv_b200.h")).read()
    declared = set(re.findall(r"\b(efv_[a-z0-9_]+)\s*\(", header))
    declared -= {"efv_mesh", "efv_system"}
    lib = _lib.load()
    missing = [name for name in sorted(declared) if not hasattr(lib, name)]
    assert not missing, missing
    assert declared == set(_lib.PROTOTYPES), declared ^ set(_lib.PROTOTYPES)


def test_no_cpu_fallback():
    """Without a device the compute entry points fail loudly instead of falling back."""
    if _lib.has_gpu():
        pytest.skip("a GPU is present")
    grid = P.structured_grid(3, "hex")
    with pytest.raises(_lib.EfvError):
        grid.device


def write_msh(path, z):
    d = util.lists_from_fixture(z)
    dim = d["dimension"]
    lines = ["$MeshFormat", "2.2 0 8", "$EndMeshFormat", "$PhysicalNames", str(len(d["region_names"]) + len(d["boundary_names"]))]
    tag = 1
    btags, rtags = [], []
    for name in d["boundary_names"]:
        lines.append(f'{dim - 1} {tag} "{name}"'); btags.append(tag); tag += 1
    for name in d["region_names"]:
        lines.append(f'{dim} {tag} "{name}"'); rtags.append(tag); tag += 1
    lines += ["$EndPhysicalNames", "$Nodes", str(len(d["coords"]))]
    for i, c in enumerate(d["coords"]):
        lines.append(f"{i + 1} {float(c[0])!r} {float(c[1])!r} {float(c[2])!r}")
    lines += ["$EndNodes", "$Elements"]
    code = {2: 1, 3: 2, 4: 3 if dim == 2 else 4, 8: 5, 6: 6, 5: 7}          # gmsh: 6 = prism, 7 = pyramid
    elems = []
    for bt, idx in zip(btags, d["bidx"]):
        for i in idx:
            elems.append((code[len(d["bconn"][i])] if len(d["bconn"][i]) != 4 else 3, bt, d["bconn"][i]))
    for rt, idx in zip(rtags, d["regions"]):
        for i in idx:
            elems.append((code[len(d["conn"][i])], rt, d["conn"][i]))
    lines.append(str(len(elems)))
    for k, (c, t, conn) in enumerate(elems):
        lines.append(f"{k + 1} {c} 2 {t} {t} " + " ".join(str(v + 1) for v in conn))
    lines += ["$EndElements", ""]
    with open(path, "w") as f:
        f.write("\n".join(lines))


@pytest.mark.parametrize("tag", ["Square", "test_mixed", "Hexas3x3x3_we", "eight_sphere", "Prisms_3d", "Pyrams_3d"])
def test_msh_reader_matches_oracle_reader(tag, tmp_path):
    z = util.load("heat_" + tag)
    path = str(tmp_path / "m.msh")
    write_msh(path, z)
    mine = P.MSHReader(path).getData()
    ref = O.read_msh(path)
    assert mine.dimension == ref.dimension == int(z["dimension"])
    assert np.array_equal(np.asarray(mine.verticesCoordinates), np.asarray(ref.verticesCoordinates))
    assert np.array_equal(np.asarray(mine.verticesCoordinates), z["coords"])
    assert [list(c) for c in mine.elementsConnectivities] == ref.elementsConnectivities
    assert mine.regionsNames == ref.regionsNames and mine.boundariesNames == ref.boundariesNames
    assert [list(r) for r in mine.regionsElementsIndexes] == ref.regionsElementsIndexes
    assert [list(r) for r in mine.boundariesIndexes] == ref.boundariesIndexes
    assert [list(c) for c in mine.boundariesConnectivities] == ref.boundariesConnectivities
    grid = P.Grid(mine)
    og = O.OracleGrid(ref)
    for b, ob in zip(grid.boundaries, og.boundaries):
        assert b.name == ob.name and np.array_equal(b.verticesIndexes, ob.vertices)      # Boundary.py:19-25 order


def test_msh_version_check(tmp_path):
    p = tmp_path / "bad.msh"
    p.write_text("$MeshFormat\n4.1 0 8\n$EndMeshFormat\n$PhysicalNames\n0\n$EndPhysicalNames\n$Nodes\n0\n$EndNodes\n$Elements\n0\n$EndElements\n")
    with pytest.raises(Exception, match="MSH version must be 2.2"):
        P.MSHReader(str(p))


def test_grid_views_and_errors():
    grid = P.structured_grid(4, "tet")
    assert grid.numberOfVertices == 64 and grid.dimension == 3 and len(grid.elements) == 6 * 27
    v = grid.vertices[5]
    assert (v.x, v.y, v.z) == tuple(grid.coords[5]) and v.handle == 5
    e = grid.elements[7]
    assert [w.handle for w in e.vertices] == grid.conn[28:32].tolist() and e.shape.name == "Tetrahedron"
    assert [b.name for b in grid.boundaries] == ["West", "East", "South", "North", "Bottom", "Top"]
    assert grid.regions[0].name == "Body" and grid.regions[0].handle == 0
    with pytest.raises(Exception, match="Grid argument must be of class GridData"):
        P.Grid("nope")
    gd = P.GridData("bad")
    gd.setDimension(3); gd.setVertices(np.zeros((7, 3))); gd.setElementConnectivity([[0, 1, 2, 3, 4, 5, 6]])
    gd.setRegions(["Body"], [[0]]); gd.setBoundaries([], [], [])
    with pytest.raises(Exception, match="not been registered"):
        P.Grid(gd)                                                                       # Element.py:27


def test_problem_data_mirrors_reference_checks():
    grid = P.structured_grid(3, "hex")
    props = P.PropertyData({"Body": {"Conductivity": 2.0, "Density": 1.0, "HeatCapacity": 3.0, "HeatGeneration": 0.0}})
    bc = {"temperature": dict({"InitialValue": 1.5}, **{n: {"condition": P.Neumann, "type": P.Constant, "value": 0.0}
                                                           for n in ["West", "East", "South", "North", "Bottom", "Top"]})}
    bc["temperature"]["West"] = {"condition": P.Dirichlet, "type": P.Variable, "value": "300+10*y"}
    pd = P.ProblemData(grid, props, P.BoundaryConditions(bc), P.NumericalSettings(timeStep=0.1, finalTime=1.0))
    assert pd.timeStep == 0.1 and pd.finalTime == 1.0 and pd.tolerance == 1e-4 and pd.maxNumberOfIterations == 1000
    assert props.get(0, "Conductivity") == 2.0 and np.array_equal(props.table(["Conductivity", "HeatCapacity"]), [[2.0, 3.0]])
    assert np.array_equal(pd.initialValues["temperature"], np.repeat(1.5, 27))
    assert len(pd.dirichletBoundaries["temperature"]) == 1 and len(pd.neumannBoundaries["temperature"]) == 5
    d = pd.dirichletBoundaries["temperature"][0]
    assert d.__type__ == "DIRICHLET" and d.boundary.name == "West"
    idx = d.boundary.verticesIndexes
    assert np.allclose(d.values_on_vertices(), 300 + 10 * grid.coords[idx, 1])
    with pytest.raises(Exception, match="Property Nope not defined"):
        props.get(0, "Nope")
    with pytest.raises(Exception, match="Invalid region handle 3"):
        props.get(3, "Density")
    with pytest.raises(Exception, match="must specify Body properties"):
        P.ProblemData(grid, P.PropertyData({"Other": {"Conductivity": 1.0}}))
    del bc["temperature"]["Top"]
    with pytest.raises(Exception, match="must specify Top condition"):
        P.ProblemData(grid, props, P.BoundaryConditions(bc))
    with pytest.raises(Exception, match="only ships the CUDA backend"):
        P.Solver(pd, backend="numpy/scipy")


def test_structured_generator_matches_oracle_generator():
    for kind in ("hex", "tet"):
        gd, og = P.structured_grid_data(6, kind), O.structured_grid_data(6, kind)
        assert np.array_equal(gd.verticesCoordinates, np.asarray(og.verticesCoordinates))
        assert np.array_equal(gd.elementsConnectivities, np.asarray(og.elementsConnectivities))
        assert np.array_equal(gd.boundariesConnectivities, np.asarray(og.boundariesConnectivities))
        assert gd.boundariesNames == og.boundariesNames


def test_bench_clock_sampler_parses_nvidia_smi_lines():
    """bench.py's clock record: median SM clock, maximum, throttle reasons from `nvidia-smi --query-gpu ... --format=csv`."""
    import bench
    s = bench.ClockSampler(0)
    s.proc = type("P", (), {"terminate": lambda self: None, "wait": lambda self, timeout=None: 0, "kill": lambda self: None})()
    s.lines = ["0, 1965, 1965, 612.3, 0x0000000000000004, Not Active, Not Active, Not Active, Active",
               "0, 1890, 1965, 998.1, 0x0000000000000004, Not Active, Not Active, Not Active, Active",
               "0, 1920, 1965, 950.0, 0x0000000000000000, Not Active, Not Active, Not Active, Not Active",
               "garbage line"]
    out = s.stop()
    assert out["sm_mhz"] == 1920.0 and out["sm_max_mhz"] == 1965.0 and out["samples"] == 3
    assert out["reasons"] == ["sw_power_cap"]


def test_weak_scaling_sizes_are_odd_and_keep_the_per_gpu_size():
    """bench.py --gpus N (weak): the C2 cube refined to round(216 N^(1/3)) | 1 vertices per side (pyefvlib_b200/distributed.py)."""
    for world, expect in ((2, 273), (4, 343), (8, 433)):
        n = int(round(216 * world ** (1.0 / 3.0))) | 1
        assert n == expect and abs(n ** 3 / world / 216 ** 3 - 1.0) < 0.012
