"""GPU parity of the coupled Biot path (apps/geomechanics.py) against the reference's captured matrix, right-hand
sides and fields (tests/golden/biot_*.npz).  The system is badly scaled (u rows ~1e9, p rows ~1e-13): matrix and RHS
are compared row-relative, the solution per field against the equilibrated + refined direct solve."""
import numpy as np
import pytest
import scipy.sparse as sp

from tests import util
from tests.cases import BIOT_CASES

pytestmark = pytest.mark.gpu
NAMES = ["nu", "G", "cs", "phi", "k", "cf", "mu", "rhos", "rhof"]


def props_array(grid, props):
    return np.array([[props[r.name][k] for k in NAMES] for r in grid.regions])


def send_bcs(system, grid, bcs):
    nV, d = grid.numberOfVertices, grid.dimension
    names = ["u_x", "u_y", "u_z"][:d] + ["p"]
    system.bc_clear()
    for i, var in enumerate(names):                            # geomechanics.py:202-221 (`+=`)
        for b in grid.boundaries:
            kind, val = bcs[var][b.name]
            if kind == "neumann":
                system.bc_neumann(b.handle, i, val, 1.0)
    for i, var in enumerate(names):                            # geomechanics.py:117-137, 224-239
        for b in grid.boundaries:
            kind, val = bcs[var][b.name]
            if kind == "dirichlet":
                system.bc_dirichlet(b.verticesIndexes + i * nV, np.full(len(b.verticesIndexes), val))


@pytest.mark.parametrize("tag", list(BIOT_CASES))
def test_matrix_rhs_solution(tag):
    import pyefvlib_b200.device as D
    z = util.load("biot_" + tag)
    c = BIOT_CASES[tag]
    grid = util.product_grid(z)
    d, nV = grid.dimension, grid.numberOfVertices
    s = D.DeviceSystem(grid, d + 1)
    n = s.rows
    Aref = sp.csr_matrix((z["data"], z["indices"], z["indptr"]), shape=(n, n))
    send_bcs(s, grid, c["bcs"])
    s.assemble_dirichlet_mode(D.DIRICHLET_ROWS)
    s.biot_assemble(props_array(grid, c["props"]), c["dt"])
    A = s.to_scipy()
    assert util.row_rel_matrix_error(A, Aref) <= 1e-12
    xold = np.zeros(n)
    for k in range(len(z["rhs"])):
        s.upload(D.VEC_XOLD, xold)
        s.build_rhs(True)
        b = s.download(D.VEC_B)
        scale = np.maximum(np.abs(z["rhs"][k]), abs(Aref) @ np.abs(z["x"][k]))
        assert (np.abs(b - z["rhs"][k]) / np.maximum(scale, 1e-300)).max() <= 1e-12
        s.upload(D.VEC_X, xold)
        iters, relres = s.solve_bicgstab(1e-10, 5000)
        assert relres <= 1e-10, (iters, relres)
        x = s.download(D.VEC_X)
        xe = z["x_equilibrated"][k]
        assert util.rel_l2(x[:d * nV], xe[:d * nV]) <= 1e-8, (iters, relres)      # displacements
        assert util.rel_l2(x[d * nV:], xe[d * nV:]) <= 1e-8, (iters, relres)      # pressure
        xold = z["x"][k]                                       # feed the reference's own fields forward


@pytest.mark.parametrize("tag", ["Hexas3x3x3", "eight_sphere", "Square"])
def test_biot_block_multigrid_preconditioner(tag):
    """BiCGStab right-preconditioned by multigrid V-cycles on the displacement and the pressure block (csrc/amg.cu: BlockPrec):
    same fields as the equilibrated direct solve, in fewer iterations than with the diagonal."""
    import pyefvlib_b200.device as D
    z = util.load("biot_" + tag)
    c = BIOT_CASES[tag]
    grid = util.product_grid(z)
    d, nV = grid.dimension, grid.numberOfVertices
    s = D.DeviceSystem(grid, d + 1)
    send_bcs(s, grid, c["bcs"])
    s.assemble_dirichlet_mode(D.DIRICHLET_ROWS)
    s.biot_assemble(props_array(grid, c["props"]), c["dt"])
    xold = np.zeros(s.rows)
    s.upload(D.VEC_XOLD, xold)
    s.build_rhs(True)
    s.upload(D.VEC_X, xold)
    itj, rrj = s.solve_bicgstab(1e-10, 5000)
    s.set_preconditioner("amg")
    s.upload(D.VEC_X, xold)
    it, rr = s.solve_bicgstab(1e-10, 500)
    assert rr <= 1e-10 and it <= itj, (it, itj, rr)
    x = s.download(D.VEC_X)
    xe = z["x_equilibrated"][0]
    assert util.rel_l2(x[:d * nV], xe[:d * nV]) <= 1e-8, (it, rr)
    assert util.rel_l2(x[d * nV:], xe[d * nV:]) <= 1e-8, (it, rr)
