"""The reference-element tables derived in pyefvlib_b200/shapes.py equal the reference's literal constants
(PyEFVLib/Shape.py:7-139, dumped into tests/golden/shape_tables.npz) bit for bit."""
import os

import numpy as np
import pytest

from pyefvlib_b200 import shapes as S
from tests import util

Z = np.load(os.path.join(util.GOLDEN, "shape_tables.npz"))
PAIRS = [("ifD", "innerFaceShapeFunctionDerivatives"), ("ifN", "innerFaceShapeFunctionValues"),
         ("subD", "subelementShapeFunctionDerivatives"), ("subN", "subelementShapeFunctionValues"),
         ("subT", "subelementTransformedVolumes"), ("facetV", "facetVerticesIndexes"), ("ofN", "outerFaceShapeFunctionValues")]


@pytest.mark.parametrize("code", list(S.SHAPES))
def test_tables_bit_exact(code):
    t = S.SHAPES[code]
    for key, attr in PAIRS:
        assert np.array_equal(Z[f"{t.name}_{key}"], np.asarray(getattr(t, attr))), (t.name, attr)


@pytest.mark.parametrize("code", list(S.SHAPES))
def test_neighbour_table(code):
    """(b, f) and the face numbering are identical; the auxiliary vertices of the two adjacent facets are equal as
    sets per facet (their order inside a facet does not enter the area formula, Shape.py:119-139)."""
    t = S.SHAPES[code]
    ref, mine = Z[f"{t.name}_ifNbr"], t.innerFaceNeighborVertices
    assert np.array_equal(ref[:, :2], mine[:, :2])
    if t.dimension == 3 and t.name == "Hexahedron":
        for r, m in zip(ref, mine):
            assert set(r[2:4]) == set(m[2:4]) and set(r[4:6]) == set(m[4:6])
    elif t.dimension == 3:
        assert np.array_equal(ref, mine)


def test_generated_cuda_tables_are_current():
    path = os.path.join(os.path.dirname(S.__file__), "csrc", "shape_tables.inc")
    assert open(path).read() == S.generate_cuda_tables()


def test_partition_of_unity_and_volume():
    for t in S.SHAPES.values():
        assert np.allclose(t.innerFaceShapeFunctionValues.sum(axis=1), 1.0, atol=1e-15)
        assert np.allclose(t.innerFaceShapeFunctionDerivatives.sum(axis=1), 0.0, atol=1e-15)
        assert np.allclose(t.outerFaceShapeFunctionValues.sum(axis=2), 1.0, atol=1e-15)
