"""Array-based result writers vs the text the reference's VtuSaver / CsvSaver produce (golden files written by
oracle/make_golden_savers.py from the unmodified reference)."""
import os
import xml.etree.ElementTree as ET

import numpy as np
import pytest

from tests import util

CASES = [("heat_Square", "Square"), ("heat_10x10", "10x10"), ("heat_Hexas3x3x3_we", "Hexas3x3x3"), ("heat_Tetras_we", "Tetras"),
         ("heat_Prisms_3d", "Prisms"), ("heat_Pyrams_3d", "Pyrams")]


def field(coords, step):          # same closed form as oracle/make_golden_savers.py
    x, y, z = coords[:, 0], coords[:, 1], coords[:, 2]
    return 300.0 + 100.0 * x + 7.0 * y - 3.0 * z + (step + 1) * np.sin(3.0 * x + 2.0 * y + z)


def golden_text(name):
    with open(os.path.join(util.GOLDEN, name)) as f:
        return f.read()


@pytest.mark.parametrize("fixture,tag", CASES)
def test_vtu_matches_reference_bytes(fixture, tag, tmp_path):
    from pyefvlib_b200.savers import VtuSaver
    grid = util.product_grid(util.load(fixture))
    saver = VtuSaver(grid, str(tmp_path), "")
    for step in range(2):
        saver.save("temperature", field(grid.coords, step), 0.01 * (step + 1))
    saver.finalize()
    assert saver.finalized and saver.fields["temperature"].shape == (2, grid.numberOfVertices)
    with open(saver.outputPath) as f:
        assert f.read() == golden_text(f"saver_{tag}.vtu")


@pytest.mark.parametrize("fixture,tag", CASES)
def test_csv_matches_reference_bytes(fixture, tag, tmp_path):
    from pyefvlib_b200.solver import CsvSaver
    grid = util.product_grid(util.load(fixture))
    saver = CsvSaver(grid, str(tmp_path), "", extension="csv")
    for step in range(2):
        saver.save("temperature", field(grid.coords, step), 0.01 * (step + 1))
    saver.finalize()
    with open(saver.outputPath) as f:
        assert f.read() == golden_text(f"saver_{tag}.csv")


def test_vtu_series_streams_one_file_per_step(tmp_path):
    from pyefvlib_b200.savers import VtuSeriesSaver
    grid = util.product_grid(util.load("heat_Hexas3x3x3_we"))
    saver = VtuSeriesSaver(grid, str(tmp_path), "")
    for step in range(3):
        saver.save("temperature", field(grid.coords, step), 0.5 * (step + 1))
        saver.save("pressure", -field(grid.coords, step), 0.5 * (step + 1))
        assert saver.fields["temperature"].shape == (1, grid.numberOfVertices)      # only the latest step is kept
    saver.finalize()
    pvd = ET.parse(saver.outputPath).getroot()
    sets = pvd.find("Collection").findall("DataSet")
    assert [float(d.get("timestep")) for d in sets] == [0.5, 1.0, 1.5]
    for step, d in enumerate(sets):
        piece = ET.parse(os.path.join(str(tmp_path), d.get("file"))).getroot().find("UnstructuredGrid").find("Piece")
        assert int(piece.get("NumberOfPoints")) == grid.numberOfVertices
        assert int(piece.get("NumberOfCells")) == grid.numberOfElements
        arrays = {a.get("Name"): np.array(a.text.split(), dtype=np.float64) for a in piece.find("PointData")}
        np.testing.assert_allclose(arrays["temperature"], field(grid.coords, step), rtol=0, atol=1e-12)
        np.testing.assert_allclose(arrays["pressure"], -field(grid.coords, step), rtol=0, atol=1e-12)
        types = np.array(piece.find("Cells").findall("DataArray")[2].text.split(), dtype=int)
        assert (types == 12).all()
