"""Aggregation multigrid preconditioner (csrc/amg.cu): the hierarchy against scipy (P^T A P rebuilt from the exported
aggregates), and the preconditioned CG against the reference's fields (tests/golden) and the oracle."""
import os

import numpy as np
import pytest
import scipy.sparse as sp

import pyefvlib_b200 as P
import pyefvlib_b200.device as D
from tests import util
from tests.cases import BC3D, HEAT_CASES, HEAT_PROPS, STRESS_PROPS, _s3
from tests.test_gpu_heat import apply_bcs, props_array

pytestmark = pytest.mark.gpu
PROPS1 = np.array([[1.0, 1.0, 1.0, 100.0]])
DT = 1e-2


def heat_system(grid, transient=True, bcs=BC3D, props=PROPS1):
    s = D.DeviceSystem(grid, 1)
    apply_bcs(s, grid, bcs)
    s.assemble_dirichlet_mode(D.DIRICHLET_SYMMETRIC)
    s.heat_assemble(props, DT, transient)
    s.upload(D.VEC_XOLD, np.zeros(s.rows)); s.upload(D.VEC_X, np.zeros(s.rows))
    s.build_rhs(transient)
    return s


def check_hierarchy(s, expect_lattice=None):
    """Every coarse operator equals P^T A P of the level above for the exported piecewise-constant P, bit for bit up to
    the summation order (<= 1e-14 relative); aggregates hold <= 8 vertices."""
    info = s.amg_info()
    nd = s.ndof
    assert info["levels"] >= 2 and info["rows"][-1] <= 256
    for l in range(info["levels"] - 1):
        A, agg = s.amg_level(l)
        Ac, _ = s.amg_level(l + 1, with_aggregates=False)
        n, nc = len(agg), info["rows"][l + 1] // nd
        assert agg.min() == 0 and agg.max() == nc - 1 and np.bincount(agg).max() <= 8
        rows = (np.arange(n)[:, None] * nd + np.arange(nd)[None, :]).ravel()
        cols = (agg[:, None].astype(np.int64) * nd + np.arange(nd)[None, :]).ravel()
        Pt = sp.csr_matrix((np.ones(n * nd), (rows, cols)), shape=(n * nd, nc * nd))
        ref = (Pt.T @ A @ Pt).tocsr()
        D_ = abs(ref - Ac)
        assert (D_.max() if D_.nnz else 0.0) <= 1e-13 * abs(ref).max(), l
        # same pattern (explicit zeros kept on both sides)
        ref.sort_indices(); Ac.sort_indices()
        pat = sp.csr_matrix((np.ones(A.nnz), A.indices, A.indptr), shape=A.shape)
        refp = (Pt.T @ pat @ Pt).tocsr()
        assert refp.nnz == Ac.nnz
        if expect_lattice and l == 0:
            m = [(k + 1) // 2 for k in expect_lattice]
            assert nc == m[0] * m[1] * m[2]
    return info


@pytest.mark.parametrize("n,transient", [(20, True), (33, False)])
def test_amg_lattice_heat(n, transient):
    grid = P.Grid(P.structured_grid_data(n, "hex"))
    s = heat_system(grid, transient)
    itj, rrj = s.solve_pcg(1e-10, 5000)
    xj = s.download(D.VEC_X)
    s.upload(D.VEC_X, np.zeros(s.rows))
    s.set_preconditioner("amg")
    it, rr = s.solve_pcg(1e-10, 200)
    x = s.download(D.VEC_X)
    info = check_hierarchy(s, (n, n, n))
    assert rr <= 1e-10 and s.true_residual() <= 2e-10, (it, rr)
    assert it <= 25 and it < itj, (it, itj)
    assert util.rel_l2(x, xj) <= 1e-8
    # the hierarchy follows the values: re-assemble with another conductivity and solve again
    s.heat_assemble(np.array([[3.0, 1.0, 1.0, 100.0]]), DT, transient)
    s.build_rhs(transient)
    s.upload(D.VEC_X, np.zeros(s.rows))
    it2, rr2 = s.solve_pcg(1e-10, 200)
    assert rr2 <= 1e-10 and s.true_residual() <= 2e-10
    check_hierarchy(s)
    # bit-reproducible: same system, same iterate
    x1 = s.download(D.VEC_X)
    s.upload(D.VEC_X, np.zeros(s.rows))
    s.solve_pcg(1e-10, 200)
    assert np.array_equal(x1, s.download(D.VEC_X))


@pytest.mark.parametrize("tag", ["Hexas6x6x6_we", "Tetras_we", "eight_sphere", "Square", "Prisms_3d"])
def test_amg_general_meshes_match_reference_field(tag):
    """Unstructured / mixed meshes of the reference: aggregates from pairwise matching; the field of the reference's own
    solve (tests/golden) to 1e-8."""
    z = util.load("heat_" + tag)
    c = HEAT_CASES[tag]
    grid = util.product_grid(z)
    s = D.DeviceSystem(grid, 1)
    apply_bcs(s, grid, c["bcs"])
    s.assemble_dirichlet_mode(D.DIRICHLET_SYMMETRIC)
    s.heat_assemble(props_array(grid, c["props"]), 0.0, False)
    s.build_rhs(False)
    s.upload(D.VEC_X, np.zeros(s.rows))
    s.set_preconditioner("amg")
    it, rr = s.solve_pcg(1e-10, 300)
    assert rr <= 1e-10, (it, rr)
    assert util.rel_l2(s.download(D.VEC_X), z["heat_T"]) <= 1e-8
    if s.rows > 256:
        check_hierarchy(s)


def test_amg_elasticity_tets():
    from oracle import efv_oracle as O
    n = 12
    grid = P.Grid(P.structured_grid_data(n, "tet"))
    og = O.OracleGrid(O.structured_grid_data(n, "tet"))
    A, b = O.elasticity_system(og, STRESS_PROPS, _s3, gravity=False)
    xe = O.solve_equilibrated(A, b)
    s = D.DeviceSystem(grid, 3)
    nV = grid.numberOfVertices
    s.bc_clear()
    for bd in grid.boundaries:
        for i, var in enumerate("uvw"):
            kind, val = _s3[var][bd.name]
            if kind == "neumann":
                s.bc_neumann(bd.handle, i, val, -1.0)
        for i, var in enumerate("uvw"):
            kind, val = _s3[var][bd.name]
            if kind == "dirichlet":
                s.bc_dirichlet(bd.verticesIndexes + i * nV, np.full(len(bd.verticesIndexes), val))
    s.assemble_dirichlet_mode(D.DIRICHLET_SYMMETRIC)
    s.elasticity_assemble(np.array([[STRESS_PROPS["Body"][k] for k in ("ShearModulus", "PoissonsRatio", "Density", "Gravity")]]), False)
    s.build_rhs(False)
    s.upload(D.VEC_X, np.zeros(s.rows))
    itj, rrj = s.solve_pcg(1e-10, 50000, negate=True)
    s.upload(D.VEC_X, np.zeros(s.rows))
    s.set_preconditioner("amg")
    it, rr = s.solve_pcg(1e-10, 2000, negate=True)
    assert rr <= 1e-10 and it < itj / 3, (it, itj, rr)
    assert util.rel_l2(s.download(D.VEC_X), xe) <= 1e-8
    check_hierarchy(s)
