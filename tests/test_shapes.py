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
    if t.name == "Pyramid":
        # Shape.py:192 lists only (b, f) for the base edges (their triangular face uses the apex implicitly,
        # Shape.py:222-225); the apex edges list the third vertex of each adjacent triangle
        assert np.array_equal(ref[4:], mine[4:]) and (mine[:4, 2] == 4).all()
        assert t.innerFaceAdjacentCounts.tolist() == [[1, 0]] * 4 + [[1, 1]] * 4
    elif t.dimension == 3 and t.name in ("Hexahedron", "Prism"):
        for r, m, (na, nb) in zip(ref, mine, t.innerFaceAdjacentCounts):
            assert set(r[2:2 + na]) == set(m[2:2 + na]) and set(r[2 + na:2 + na + nb]) == set(m[2 + na:2 + na + nb])
            assert (m[2 + na + nb:] == 0).all() and (r[2 + na + nb:] == 0).all()
    elif t.dimension == 3:
        assert np.array_equal(ref, mine)


def test_generated_cuda_tables_are_current():
    path = os.path.join(os.path.dirname(S.__file__), "csrc", "shape_tables.inc")
    assert open(path).read() == S.generate_cuda_tables()


def test_partition_of_unity_and_volume():
    for t in S.SHAPES.values():
        assert np.allclose(t.innerFaceShapeFunctionValues.sum(axis=1), 1.0, atol=1e-15)
        assert np.allclose(t.innerFaceShapeFunctionDerivatives.sum(axis=1), 0.0, atol=1e-15)
        for lf, n in enumerate(t.facetSizes):         # rows beyond the facet's vertex count are padding
            assert np.allclose(t.outerFaceShapeFunctionValues[lf, :n].sum(axis=1), 1.0, atol=1e-15)
            assert (t.outerFaceShapeFunctionValues[lf, n:] == 0).all()
