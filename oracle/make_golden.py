"""TEST INFRASTRUCTURE ONLY -- generate tests/golden/* from the UNMODIFIED reference and pin the oracle.

Run in the build container (needs /root/reference):   python -m oracle.make_golden

For every case it (1) runs the reference objects / apps through oracle/ref_harness.py, (2) asserts that
oracle/efv_oracle.py reproduces them, (3) stores mesh + reference outputs as a compressed .npz fixture.
The fixtures are what `tests/` use on machines without the reference (the GPU box).
"""
import os
import sys
import time
import warnings

import numpy as np
import scipy.sparse as sp

warnings.filterwarnings("ignore")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")

from oracle import ref_harness as R  # noqa: E402

ns = R.load()
P = ns["PyEFVLib"]


def _padded(rows, fill, dtype):
    rows = [list(r) for r in rows]
    w = max(len(r) for r in rows)
    return np.array([r + [fill] * (w - len(r)) for r in rows], dtype=dtype)


def dump_shape_tables():
    out = {}
    for name in ("Triangle", "Quadrilateral", "Tetrahedron", "Hexahedron", "Prism", "Pyramid"):
        c = getattr(P, name)
        m = len(c.subelementTransformedVolumes)
        out[f"{name}_ifD"] = np.asarray(c.innerFaceShapeFunctionDerivatives, dtype=np.float64)
        out[f"{name}_ifN"] = np.asarray(c.innerFaceShapeFunctionValues, dtype=np.float64)
        out[f"{name}_ifNbr"] = _padded(c.innerFaceNeighborVertices, 0, np.int64)           # pyramid rows are ragged
        out[f"{name}_subD"] = np.asarray(c.subelementShapeFunctionDerivatives, dtype=np.float64)
        out[f"{name}_subN"] = np.asarray(c.subelementShapeFunctionValues, dtype=np.float64)
        out[f"{name}_subT"] = np.asarray(c.subelementTransformedVolumes, dtype=np.float64)
        out[f"{name}_facetV"] = _padded(c.facetVerticesIndexes, -1, np.int64)              # prism / pyramid: 3- and 4-vertex facets
        of = [[list(r) for r in f] for f in c.outerFaceShapeFunctionValues]
        w = max(len(f) for f in of)
        out[f"{name}_ofN"] = np.array([f + [[0.0] * m] * (w - len(f)) for f in of], dtype=np.float64)
    np.savez_compressed(os.path.join(GOLDEN, "shape_tables.npz"), **out)


dump_shape_tables()
from oracle import efv_oracle as O  # noqa: E402  (needs shape_tables.npz)


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)) if a.size else 0.0


def griddata_arrays(gd):
    conn = gd.elementsConnectivities
    bconn = gd.boundariesConnectivities
    return dict(
        dimension=np.int64(gd.dimension),
        coords=np.asarray(gd.verticesCoordinates, dtype=np.float64),
        conn_flat=np.concatenate([np.asarray(c, dtype=np.int64) for c in conn]),
        conn_ptr=np.concatenate([[0], np.cumsum([len(c) for c in conn])]).astype(np.int64),
        region_names=np.array(gd.regionsNames),
        region_ptr=np.concatenate([[0], np.cumsum([len(r) for r in gd.regionsElementsIndexes])]).astype(np.int64),
        region_elems=np.concatenate([np.asarray(r, dtype=np.int64) for r in gd.regionsElementsIndexes]),
        boundary_names=np.array(gd.boundariesNames),
        boundary_ptr=np.concatenate([[0], np.cumsum([len(r) for r in gd.boundariesIndexes])]).astype(np.int64),
        boundary_facets=np.concatenate([np.asarray(r, dtype=np.int64) for r in gd.boundariesIndexes]),
        bconn_flat=np.concatenate([np.asarray(c, dtype=np.int64) for c in bconn]),
        bconn_ptr=np.concatenate([[0], np.cumsum([len(c) for c in bconn])]).astype(np.int64),
    )


def ref_geometry(grid):
    """Flatten the reference object graph."""
    G, S, vol = [], [], []
    for e in grid.elements:
        for f in e.innerFaces:
            G.append(np.asarray(f.globalDerivatives).ravel())
            S.append(f.area.getCoordinates())
        vol.append(np.asarray(e.subelementVolumes, dtype=np.float64))
    out = dict(G_flat=np.concatenate(G), S_flat=np.concatenate(S), vol_flat=np.concatenate(vol),
               vertex_volume=np.array([v.volume for v in grid.vertices]))
    for k, b in enumerate(grid.boundaries):
        out[f"b{k}_vertices"] = np.array([v.handle for v in b.vertices], dtype=np.int64)
        out[f"b{k}_facet_element"] = np.array([f.element.handle for f in b.facets], dtype=np.int64)
        out[f"b{k}_facet_local"] = np.array([f.elementLocalIndex for f in b.facets], dtype=np.int64)
        out[f"b{k}_facet_area"] = np.array([f.area.getCoordinates() for f in b.facets]).reshape(-1, 3)
        out[f"b{k}_outer_vertex"] = np.array([of.vertex.handle for f in b.facets for of in f.outerFaces], dtype=np.int64)
        out[f"b{k}_outer_area"] = np.array([of.area.getCoordinates() for f in b.facets for of in f.outerFaces]).reshape(-1, 3)
    return out


def check_geometry(og, ref):
    G = _flat_by_element(og, "G")
    S = _flat_by_element(og, "S")
    vol = _flat_by_element(og, "vol")
    errs = dict(G=rel(G, ref["G_flat"]), S=rel(S, ref["S_flat"]), vol=rel(vol, ref["vol_flat"]),
                vv=rel(og.vertex_volume, ref["vertex_volume"]))
    for k, b in enumerate(og.boundaries):
        assert np.array_equal(b.vertices, ref[f"b{k}_vertices"]), "boundary vertex order"
        assert np.array_equal(b.facet_element, ref[f"b{k}_facet_element"]), "facet element"
        assert np.array_equal(b.facet_local, ref[f"b{k}_facet_local"]), "facet local"
        assert np.array_equal(b.outer_vertex, ref[f"b{k}_outer_vertex"])
        errs[f"fa{k}"] = rel(b.facet_area, ref[f"b{k}_facet_area"])
        errs[f"oa{k}"] = rel(b.outer_area, ref[f"b{k}_outer_area"])
    worst = max(errs.values())
    assert worst < 1e-13, errs
    return worst


def _flat_by_element(og, what):
    parts = [None] * og.numberOfElements
    for g in og.groups:
        arr = getattr(g, what)
        for a, e in enumerate(g.ids):
            parts[e] = arr[a].ravel()
    return np.concatenate(parts)


def bc_dict(spec, names, variable_names):
    """spec: {var: {boundaryName: (kind, value)}} -> reference BoundaryConditions dict."""
    out = {}
    for var in variable_names:
        d = {"InitialValue": 0.0}
        for n in names:
            kind, val = spec[var][n]
            d[n] = {"condition": P.Dirichlet if kind == "dirichlet" else P.Neumann, "type": P.Constant, "value": val}
        out[var] = d
    return out


def csr_parts(A):
    A = sp.csr_matrix(A)
    A.sum_duplicates()
    A.sort_indices()
    return A.indptr.astype(np.int64), A.indices.astype(np.int64), A.data.astype(np.float64)


def heat_case(mesh, props, bcs, tag, transient_steps=0, dt=0.01):
    t0 = time.time()
    path = R.mesh_path(mesh)
    names = P.MSHReader(path).getData().boundariesNames
    pd = P.ProblemData(
        meshFilePath=path, outputFilePath="tmp/efv_golden",
        numericalSettings=P.NumericalSettings(timeStep=dt, finalTime=None, tolerance=1e-6, maxNumberOfIterations=50),
        propertyData=P.PropertyData(props),
        boundaryConditions=P.BoundaryConditions(bc_dict({"temperature": bcs}, names, ["temperature"])))
    ht = ns["heat_transfer"]
    out = griddata_arrays(pd.grid.gridData)
    ref = ref_geometry(pd.grid)
    out.update(ref)
    og = O.OracleGrid(O.read_msh(path))
    gerr = check_geometry(og, ref)
    # steady system
    s = ht.HeatTransferSolver(pd, extension="csv", transient=False, verbosity=False)
    s.settings(); s.init(); s.assembleMatrix(); s.addToIndependentVector(); s.solveLinearSystem()
    ip, ix, dv = csr_parts(s.matrix)
    out.update(heat_indptr=ip, heat_indices=ix, heat_data=dv, heat_rhs=s.independent.copy(), heat_T=s.temperatureField.copy())
    A = O.heat_matrix(og, props, bcs)
    b = O.heat_rhs(og, props, bcs)
    D = sp.csr_matrix(A) - sp.csr_matrix(s.matrix)
    merr = float(np.abs(D.data).max() / np.abs(dv).max()) if D.nnz else 0.0
    berr = rel(b, s.independent)
    terr = rel(O.solve_direct(A, b), s.temperatureField)
    assert merr < 1e-13 and berr < 1e-13 and terr < 1e-10, (merr, berr, terr)
    # interior pattern = union of element cliques with sorted columns
    gptr, gcol = O.csr_pattern(og, 1)
    out.update(pattern_indptr=gptr, pattern_indices=gcol.astype(np.int64), slot_map=O.slot_map(og, gptr, gcol))
    if transient_steps:
        s = ht.HeatTransferSolver(pd, extension="csv", transient=True, verbosity=False)
        s.settings(); s.init(); s.assembleMatrix()
        ip, ix, dv = csr_parts(s.matrix)
        out.update(heatT_indptr=ip, heatT_indices=ix, heatT_data=dv, heatT_dt=np.float64(dt))
        At = O.heat_matrix(og, props, bcs, transient=True, dt=dt)
        D = sp.csr_matrix(At) - sp.csr_matrix(s.matrix)
        assert (float(np.abs(D.data).max() / np.abs(dv).max()) if D.nnz else 0.0) < 1e-13
        fields, rhss = [], []
        Told = np.zeros(og.numberOfVertices)
        for it in range(transient_steps):
            s.addToIndependentVector(); s.solveLinearSystem()
            rhss.append(s.independent.copy()); fields.append(s.temperatureField.copy())
            bo = O.heat_rhs(og, props, bcs, transient=True, dt=dt, Told=Told)
            assert rel(bo, s.independent) < 1e-13
            Told = O.solve_direct(At, bo)
            assert rel(Told, s.temperatureField) < 1e-10
            s.prevTemperatureField = s.temperatureField
        out.update(heatT_rhs=np.array(rhss), heatT_fields=np.array(fields))
    np.savez_compressed(os.path.join(GOLDEN, f"heat_{tag}.npz"), **out)
    print(f"heat_{tag}: nV={og.numberOfVertices} nE={og.numberOfElements} geom {gerr:.1e} A {merr:.1e} b {berr:.1e} T {terr:.1e}"
          f"  trace={s.matrix.diagonal().sum() if not transient_steps else 0:.12e} ({time.time()-t0:.1f}s)")


def stress_case(mesh, props, bcs, tag, gravity=False):
    t0 = time.time()
    path = R.mesh_path(mesh)
    gd = P.MSHReader(path).getData()
    variables = ["u", "v", "w"][:gd.dimension]
    pd = P.ProblemData(meshFilePath=path, outputFilePath="tmp/efv_golden", propertyData=P.PropertyData(props),
                       boundaryConditions=P.BoundaryConditions(bc_dict(bcs, gd.boundariesNames, variables)))
    se = ns["stress_equilibrium"]
    s = se.StressEquilibriumSolver(pd, extension="csv", gravity=gravity, verbosity=False)
    s.settings(); s.init(); s.assembleLinearSystem(); s.solveLinearSystem()
    out = griddata_arrays(pd.grid.gridData)
    ip, ix, dv = csr_parts(s.matrix)
    out.update(indptr=ip, indices=ix, data=dv, rhs=s.independent.copy(), x=np.asarray(s.displacements).copy())
    og = O.OracleGrid(O.read_msh(path))
    A, b = O.elasticity_system(og, props, bcs, gravity=gravity)
    D = sp.csr_matrix(A) - sp.csr_matrix(s.matrix)
    merr = float(np.abs(D.data).max() / np.abs(dv).max()) if D.nnz else 0.0
    berr = rel(b, s.independent)
    xe = O.solve_equilibrated(A, b)
    xerr = rel(xe, s.displacements)
    out.update(x_equilibrated=xe)
    assert merr < 1e-13 and berr < 1e-13 and xerr < 1e-6, (merr, berr, xerr)
    np.savez_compressed(os.path.join(GOLDEN, f"stress_{tag}.npz"), **out)
    print(f"stress_{tag}: n={len(b)} A {merr:.1e} b {berr:.1e} x(vs ref inv) {xerr:.1e} ({time.time()-t0:.1f}s)")


def biot_case(mesh, props, bcs, tag, dt=20.0, steps=2):
    t0 = time.time()
    path = R.mesh_path(mesh)
    gd = P.MSHReader(path).getData()
    d = gd.dimension
    variables = ["u_x", "u_y", "u_z"][:d] + ["p"]
    pd = P.ProblemData(meshFilePath=path, outputFilePath="tmp/efv_golden",
                       numericalSettings=P.NumericalSettings(timeStep=dt, finalTime=1e30, tolerance=0.0, maxNumberOfIterations=steps),
                       propertyData=P.PropertyData(props),
                       boundaryConditions=P.BoundaryConditions(bc_dict(bcs, gd.boundariesNames, variables)))
    geo = ns["geomechanics"]
    captured = {"A": None, "b": [], "x": []}
    real_inv, real_matmul = np.linalg.inv, np.matmul

    def fake_inv(M):
        if M.ndim == 2 and M.shape[0] == (d + 1) * pd.grid.numberOfVertices:
            captured["A"] = M.copy()
        return real_inv(M)

    def fake_matmul(a, b, *args, **kw):
        r = real_matmul(a, b, *args, **kw)
        if getattr(a, "ndim", 0) == 2 and a.shape[0] == (d + 1) * pd.grid.numberOfVertices and getattr(b, "ndim", 0) == 1:
            captured["b"].append(np.array(b)); captured["x"].append(np.array(r))
        return r

    class NoSaver:
        def __init__(self, *a, **k):
            pass

        def save(self, *a, **k):
            pass

        def finalize(self, *a, **k):
            pass
    old = (P.CsvSaver, P.MeshioSaver)
    P.CsvSaver = P.MeshioSaver = NoSaver
    np.linalg.inv, np.matmul = fake_inv, fake_matmul
    import io, contextlib
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            geo.geomechanics(pd)
    finally:
        np.linalg.inv, np.matmul = real_inv, real_matmul
        P.CsvSaver, P.MeshioSaver = old
    out = griddata_arrays(pd.grid.gridData)
    ip, ix, dv = csr_parts(sp.csr_matrix(captured["A"]))
    out.update(indptr=ip, indices=ix, data=dv, rhs=np.array(captured["b"]), x=np.array(captured["x"]), dt=np.float64(dt))
    og = O.OracleGrid(O.read_msh(path))
    nV = og.numberOfVertices
    u_old, p_old = np.zeros(d * nV), np.zeros(nV)
    xs = []
    for it in range(len(captured["b"])):
        A, b = O.biot_system(og, props, bcs, dt, u_old, p_old)
        if it == 0:
            D = sp.csr_matrix(A) - sp.csr_matrix(captured["A"])
            D.eliminate_zeros()
            # compare entry-wise relative to the row scale (u rows ~1e9, p rows ~1e-13)
            scale = np.abs(captured["A"]).max(axis=1)
            merr = float((np.abs(D).max(axis=1).toarray().ravel() / scale).max())
            assert merr < 1e-13, merr
        bscale = np.maximum(np.abs(captured["b"][it]), np.abs(sp.csr_matrix(captured["A"])) @ np.abs(captured["x"][it]))
        berr = float((np.abs(b - captured["b"][it]) / np.maximum(bscale, 1e-300)).max())
        assert berr < 1e-11, berr
        x = O.solve_equilibrated(A, b)
        xs.append(x)
        # feed the REFERENCE's own fields forward so that RHS comparisons stay independent of solver error
        u_old, p_old = captured["x"][it][:d * nV], captured["x"][it][d * nV:]
    out.update(x_equilibrated=np.array(xs))
    uerr = rel(xs[0][:d * nV], captured["x"][0][:d * nV]); perr = rel(xs[0][d * nV:], captured["x"][0][d * nV:])
    np.savez_compressed(os.path.join(GOLDEN, f"biot_{tag}.npz"), **out)
    print(f"biot_{tag}: n={len(b)} A {merr:.1e} b {berr:.1e} u {uerr:.1e} p {perr:.1e} ({time.time()-t0:.1f}s)")


def fixed_stress_case(mesh, props, bcs, tag, dt=20.0, steps=2, tolerance=1e-4):
    t0 = time.time()
    path = R.mesh_path(mesh)
    gd = P.MSHReader(path).getData()
    d = gd.dimension
    variables = ["u_x", "u_y", "u_z"][:d] + ["p"]
    pd = P.ProblemData(meshFilePath=path, outputFilePath="tmp/efv_golden",
                       numericalSettings=P.NumericalSettings(timeStep=dt, finalTime=dt * steps - 1e-9, tolerance=tolerance, maxNumberOfIterations=10 ** 9),
                       propertyData=P.PropertyData(props),
                       boundaryConditions=P.BoundaryConditions(bc_dict(bcs, gd.boundariesNames, variables)))
    fs = ns["fixed_stress"]
    saved = []

    class Rec:
        def __init__(self, *a, **k):
            pass

        def save(self, name, values, t):
            saved.append((name, np.array(values), t))

        def finalize(self, *a, **k):
            pass

    class Nop(Rec):
        def save(self, *a, **k):
            pass
    old = (P.CsvSaver, P.MeshioSaver)
    P.CsvSaver, P.MeshioSaver = Rec, Nop
    import io, contextlib
    buf = io.StringIO()
    try:
        with contextlib.redirect_stdout(buf):
            fs.geomechanics(pd)
    finally:
        P.CsvSaver, P.MeshioSaver = old
    inner = [int(line.split()[0]) for line in buf.getvalue().strip().splitlines()]
    nV = pd.grid.numberOfVertices
    times = sorted(set(t for _, _, t in saved))
    U, Pf = [], []
    for t in times:
        comp = {n: v for n, v, tt in saved if tt == t}
        U.append(np.concatenate([comp[v] for v in variables[:d]])); Pf.append(comp["p"])
    out = griddata_arrays(pd.grid.gridData)
    out.update(u=np.array(U), p=np.array(Pf), inner_iterations=np.array(inner), dt=np.float64(dt), tolerance=np.float64(tolerance))
    og = O.OracleGrid(O.read_msh(path))
    res = O.fixed_stress_run(og, props, bcs, dt, len(times), tolerance)
    uerr = max(rel(r[0], U[k]) for k, r in enumerate(res)); perr = max(rel(r[1], Pf[k]) for k, r in enumerate(res))
    assert [r[2] for r in res] == inner, ([r[2] for r in res], inner)
    assert uerr < 1e-7 and perr < 1e-7, (uerr, perr)
    np.savez_compressed(os.path.join(GOLDEN, f"fixed_stress_{tag}.npz"), **out)
    print(f"fixed_stress_{tag}: steps={len(times)} inner={inner} u {uerr:.1e} p {perr:.1e} ({time.time()-t0:.1f}s)")


from tests.cases import HEAT_CASES, STRESS_CASES, BIOT_CASES  # noqa: E402


def main():
    """No arguments: every fixture.  With arguments: only the named ones (e.g. `heat_Prisms_3d stress_Square`)."""
    os.makedirs(GOLDEN, exist_ok=True)
    only = set(sys.argv[1:])
    want = lambda name: not only or name in only
    for tag, c in HEAT_CASES.items():
        if want(f"heat_{tag}"):
            heat_case(c["mesh"], c["props"], c["bcs"], tag, transient_steps=c["transient_steps"])
    for tag, c in STRESS_CASES.items():
        if want(f"stress_{tag}"):
            stress_case(c["mesh"], c["props"], c["bcs"], tag, gravity=c["gravity"])
    for tag, c in BIOT_CASES.items():
        if want(f"biot_{tag}"):
            biot_case(c["mesh"], c["props"], c["bcs"], tag, dt=c["dt"], steps=c["steps"])
    for tag in ("Square", "Hexas3x3x3"):
        if want(f"fixed_stress_{tag}"):
            c = BIOT_CASES[tag]
            fixed_stress_case(c["mesh"], c["props"], c["bcs"], tag, dt=c["dt"], steps=2)


if __name__ == "__main__":
    main()
