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


NONSYMMETRIC = ("Prisms", "Pyrams")


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
    s.upload(D.VEC_X, np.zeros(n))
    if tag in NONSYMMETRIC:
        # the EbFVM elasticity operator of prisms / pyramids is not symmetric (nor is the reference's): BiCGStab, which is
        # also what the app falls over to (pyefvlib_b200/apps/stress_equilibrium.py:68-70)
        assert abs(As - As.T).max() > 1e-6 * abs(As).max()
        iters, relres = s.solve_bicgstab(1e-10, 20000)
    else:
        assert abs(As - As.T).max() <= 1e-10 * abs(As).max()
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


def test_mixed_mesh_elasticity_and_biot_against_oracle():
    """Mixed triangle + quadrilateral mesh with two regions (meshes/msh/2D/test.msh, taken from the heat fixture):
    the generic (non-uniform) kernels for elasticity and Biot against the oracle."""
    import pyefvlib_b200.device as D
    from oracle import efv_oracle as O
    z = util.load("heat_test_mixed")
    grid = util.product_grid(z)
    og = util.oracle_grid(z)
    nV = grid.numberOfVertices
    N0 = ("neumann", 0.0)
    names = [b.name for b in grid.boundaries]
    eb = {"u": {n: N0 for n in names}, "v": {n: N0 for n in names}}
    eb["u"]["W"] = ("dirichlet", 0.0); eb["v"]["W"] = ("dirichlet", 0.0); eb["u"]["E"] = ("dirichlet", 1e-3); eb["v"]["NE"] = ("neumann", -5e3)
    sp_props = {"WEST": {"Density": 1800.0, "PoissonsRatio": 0.4, "ShearModulus": 6.0e6, "Gravity": 10.0},
                "EAST": {"Density": 2500.0, "PoissonsRatio": 0.25, "ShearModulus": 2.0e7, "Gravity": 10.0}}
    s = D.DeviceSystem(grid, 2)
    send_bcs(s, grid, eb)
    s.assemble_dirichlet_mode(D.DIRICHLET_ROWS)
    s.elasticity_assemble(props_array(grid, sp_props), True)
    s.build_rhs(False)
    Aref, bref = O.elasticity_system(og, sp_props, eb, gravity=True)
    assert util.row_rel_matrix_error(s.to_scipy(), Aref) <= 1e-12
    assert util.rel_max(s.download(D.VEC_B), bref) <= 1e-12
    s.upload(D.VEC_X, np.zeros(s.rows))
    iters, relres = s.solve_bicgstab(1e-12, 20000)          # general quadrilaterals: non-symmetric operator
    assert relres <= 1e-12
    assert util.rel_l2(s.download(D.VEC_X), O.solve_equilibrated(Aref, bref)) <= 1e-8
    # Biot on the same mesh
    from tests.test_gpu_biot import props_array as bprops, send_bcs as bsend
    bp = {"WEST": dict(nu=0.2, G=6.0e9, cs=0.0, phi=0.19, k=1.9e-15, cf=3.0303e-10, mu=1000.0, rhos=2700.0, rhof=1000.0),
          "EAST": dict(nu=0.3, G=2.0e9, cs=1e-11, phi=0.3, k=5e-15, cf=4.0e-10, mu=1000.0, rhos=2500.0, rhof=1000.0)}
    bb = {v: {n: N0 for n in names} for v in ("u_x", "u_y", "p")}
    bb["u_x"]["W"] = ("dirichlet", 0.0); bb["u_y"]["W"] = ("dirichlet", 0.0); bb["u_y"]["NE"] = ("neumann", 1e5); bb["p"]["E"] = ("dirichlet", 0.0)
    sb = D.DeviceSystem(grid, 3)
    bsend(sb, grid, bb)
    sb.assemble_dirichlet_mode(D.DIRICHLET_ROWS)
    sb.biot_assemble(bprops(grid, bp), 20.0)
    u_old = 1e-6 * np.sin(np.arange(2 * nV)); p_old = 1e3 * np.cos(np.arange(nV))
    sb.upload(D.VEC_XOLD, np.concatenate([u_old, p_old]))
    sb.build_rhs(True)
    Ab, bbv = O.biot_system(og, bp, bb, 20.0, u_old, p_old)
    assert util.row_rel_matrix_error(sb.to_scipy(), Ab) <= 1e-12
    scale = np.maximum(np.abs(bbv), abs(Ab) @ np.abs(np.concatenate([u_old, p_old])))
    assert (np.abs(sb.download(D.VEC_B) - bbv) / np.maximum(scale, 1e-300)).max() <= 1e-12


def test_bad_connectivity_is_rejected():
    import pyefvlib_b200 as P
    from pyefvlib_b200 import _lib
    gd = P.structured_grid_data(3, "hex")
    gd.elementsConnectivities = gd.elementsConnectivities.copy()
    gd.elementsConnectivities[0, 0] = 10 ** 6
    with pytest.raises(_lib.EfvError, match="outside"):
        P.Grid(gd).device


@pytest.mark.parametrize("tag", ["Hexas4x4x4", "Prisms", "Pyrams", "eight_sphere", "10x10_gravity"])
@pytest.mark.parametrize("mode", ["rows", "symmetric"])
def test_assembly_kernels_agree(tag, mode, monkeypatch):
    """Lane kernel (default; with a capped grid: many tiles per warp through the cp.async ring) against the round-1 row-group
    kernel (EFV_ASSEMBLY=rows): block values, gravity vector and elimination vector, both fused Dirichlet modes."""
    import pyefvlib_b200.device as D
    z = util.load("stress_" + tag)
    c = STRESS_CASES[tag]
    grid = util.product_grid(z)
    out = {}
    for label, kernel, cap in (("lanes", "lanes", None), ("capped", "lanes", "2"), ("rows", "rows", None)):
        monkeypatch.setenv("EFV_ASSEMBLY", kernel)
        if cap: monkeypatch.setenv("EFV_LANE_GRID", cap)
        else: monkeypatch.delenv("EFV_LANE_GRID", raising=False)
        s = D.DeviceSystem(grid, grid.dimension)
        send_bcs(s, grid, c["bcs"])
        s.assemble_dirichlet_mode(D.DIRICHLET_ROWS if mode == "rows" else D.DIRICHLET_SYMMETRIC)
        s.elasticity_assemble(props_array(grid, c["props"]), c["gravity"])
        s.build_rhs(False)
        out[label] = (s.values(), s.download(D.VEC_B))
    assert np.array_equal(out["lanes"][0], out["capped"][0]) and np.array_equal(out["lanes"][1], out["capped"][1])
    scale = np.abs(out["rows"][0]).max()
    assert np.abs(out["lanes"][0] - out["rows"][0]).max() <= 1e-13 * scale
    assert util.rel_max(out["lanes"][1], out["rows"][1]) <= 1e-12
