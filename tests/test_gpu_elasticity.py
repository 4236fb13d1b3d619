"""GPU parity of the linear-elasticity path (apps/stress_equilibrium.py) against the reference outputs in
tests/golden and against the oracle on synthetic meshes.  Matrix values <= 1e-12 (row-relative), solution
<= 1e-8 relative L2 against the equilibrated + refined direct solve (the reference's own
sparse inverse is only 1e-9 accurate on these systems)."""
import numpy as np
import pytest
import scipy.sparse as sp

from tests import util
from tests.cases import STRESS_CASES, STRESS_PROPS

pytestmark = pytest.mark.gpu


def props_array(grid, props):
    return np.array([[props[r.name]["ShearModulus"], props[r.name]["PoissonsRatio"], props[r.name]["Density"],
                      props[r.name]["Gravity"]] for r in grid.regions])


def send_bcs(system, grid, bcs):
    nV = grid.numberOfVertices
    names = ["u", "v", "w"][:grid.dimension]
    system.bc_clear()
    for b in grid.boundaries:                                  # stress_equilibrium.py:119-163
        for i, var in enumerate(names):
            kind, val = bcs[var][b.name]
            if kind == "neumann":
                system.bc_neumann(b.handle, i, val, -1.0)       # note the `-=` (:131-141)
        for i, var in enumerate(names):
            kind, val = bcs[var][b.name]
            if kind == "dirichlet":
                system.bc_dirichlet(b.verticesIndexes + i * nV, np.full(len(b.verticesIndexes), val))


@pytest.mark.parametrize("tag", list(STRESS_CASES))
def test_matrix_rhs_solution(tag):
    import pyefvlib_b200.device as D
    z = util.load("stress_" + tag)
    c = STRESS_CASES[tag]
    grid = util.product_grid(z)
    s = D.DeviceSystem(grid, grid.dimension)
    n = s.rows
    Aref = sp.csr_matrix((z["data"], z["indices"], z["indptr"]), shape=(n, n))
    send_bcs(s, grid, c["bcs"])
    s.assemble_dirichlet_mode(D.DIRICHLET_ROWS)
    s.elasticity_assemble(props_array(grid, c["props"]), c["gravity"])
    s.build_rhs(False)
    A = s.to_scipy()
    assert util.row_rel_matrix_error(A, Aref) <= 1e-12
    assert util.rel_max(s.download(D.VEC_B), z["rhs"]) <= 1e-12
    # separate Dirichlet kernel gives the same matrix as the fused path
    s.assemble_dirichlet_mode(-1)
    s.elasticity_assemble(props_array(grid, c["props"]), c["gravity"])
    s.apply_dirichlet(D.DIRICHLET_ROWS)
    assert abs(s.to_scipy() - A).max() == 0.0
    # symmetric form, negated, PCG
    s.assemble_dirichlet_mode(D.DIRICHLET_SYMMETRIC)
    s.elasticity_assemble(props_array(grid, c["props"]), c["gravity"])
    s.build_rhs(False)
    As = s.to_scipy()
    assert abs(As - As.T).max() <= 1e-10 * abs(As).max()
    s.upload(D.VEC_X, np.zeros(n))
    iters, relres = s.solve_pcg(1e-10, 20000, negate=True)
    assert relres <= 1e-10, (iters, relres)
    x = s.download(D.VEC_X)
    assert util.rel_l2(x, z["x_equilibrated"]) <= 1e-8, (iters, relres)
    assert util.rel_l2(x, z["x"]) <= 1e-6                      # the reference's inv(A) @ b itself


@pytest.mark.parametrize("n,kind", [(7, "tet"), (6, "hex")])
def test_structured_against_oracle(n, kind):
    import pyefvlib_b200 as P
    import pyefvlib_b200.device as D
    from oracle import efv_oracle as O
    N0 = ("neumann", 0.0)
    bcs = {"u": {"West": ("dirichlet", 0.0), "East": ("dirichlet", 0.0), "South": N0, "North": N0, "Bottom": N0, "Top": N0},
           "v": {"West": N0, "East": N0, "South": ("dirichlet", 0.0), "North": ("neumann", -1e4), "Bottom": N0, "Top": N0},
           "w": {"West": N0, "East": N0, "South": N0, "North": N0, "Bottom": ("dirichlet", 0.0), "Top": N0}}
    gd = P.structured_grid_data(n, kind)
    grid = P.Grid(gd)
    og = O.OracleGrid(O.OracleGridData(3, gd.verticesCoordinates.tolist(), gd.elementsConnectivities.tolist(), gd.regionsNames,
                                       [list(r) for r in gd.regionsElementsIndexes], gd.boundariesNames,
                                       [list(r) for r in gd.boundariesIndexes], gd.boundariesConnectivities.tolist()))
    s = D.DeviceSystem(grid, 3)
    ip, ix = s.csr_structure()
    rp, rx = O.csr_pattern(og, 3)
    assert np.array_equal(ip, rp) and np.array_equal(ix, rx)       # LinearSystemCSR layout (LinearSystem.py:114-128)
    send_bcs(s, grid, bcs)
    s.assemble_dirichlet_mode(D.DIRICHLET_ROWS)
    s.elasticity_assemble(props_array(grid, STRESS_PROPS), True)
    s.build_rhs(False)
    Aref, bref = O.elasticity_system(og, STRESS_PROPS, bcs, gravity=True)
    assert util.row_rel_matrix_error(s.to_scipy(), Aref) <= 1e-12
    assert util.rel_max(s.download(D.VEC_B), bref) <= 1e-12
    s.assemble_dirichlet_mode(D.DIRICHLET_SYMMETRIC)
    s.elasticity_assemble(props_array(grid, STRESS_PROPS), True)
    s.build_rhs(False)
    s.upload(D.VEC_X, np.zeros(s.rows))
    iters, relres = s.solve_pcg(1e-10, 20000, negate=True)
    assert relres <= 1e-10
    assert util.rel_l2(s.download(D.VEC_X), O.solve_equilibrated(Aref, bref)) <= 1e-8
