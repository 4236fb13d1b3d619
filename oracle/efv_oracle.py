"""TEST INFRASTRUCTURE ONLY -- CPU oracle: a vectorised numpy restatement of PyEFVLib's hot path.

This file is the *checker* for the CUDA product in `pyefvlib_b200/`.  It may only be imported by `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs.  The product never
imports it and has no CPU fallback.

Pinning: `oracle/make_golden.py` runs the UNMODIFIED reference (via `oracle/ref_harness.py`, build
container only) on the bundled meshes and (a) asserts that every function below agrees with the reference
objects (geometry <= 1e-13 rel, matrices/RHS <= 1e-13 rel), (b) writes the fixtures in `tests/golden/`.
`tests/test_oracle_golden.py` re-checks the oracle against those committed fixtures on every CPU run.
The reference itself has no tests, so those fixtures are outputs of the reference code.

Every function cites the reference lines it restates (paths relative to the reference root).
"""
import os

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

_HERE = os.path.dirname(os.path.abspath(__file__))
_TABLES = os.path.join(_HERE, os.pardir, "tests", "golden", "shape_tables.npz")

SHAPE_OF = {(3, 2): "Triangle", (4, 2): "Quadrilateral", (4, 3): "Tetrahedron", (8, 3): "Hexahedron",
            (6, 3): "Prism", (5, 3): "Pyramid"}


class _Tab:
    pass


_tables = {}


def tables():
    """Reference-element constants exactly as the reference holds them (dumped from PyEFVLib/Shape.py:7-199
    by oracle/make_golden.py into tests/golden/shape_tables.npz; the ragged prism / pyramid tables are padded:
    facet vertices with -1, neighbour vertices and outer-face values with 0)."""
    if not _tables:
        z = np.load(_TABLES)
        for name in ("Triangle", "Quadrilateral", "Tetrahedron", "Hexahedron", "Prism", "Pyramid"):
            t = _Tab()
            t.name = name
            for key in ("ifD", "ifN", "ifNbr", "subD", "subN", "subT", "facetV", "ofN"):
                setattr(t, key, z[f"{name}_{key}"])
            t.m = t.ifD.shape[1]
            t.nf = t.ifD.shape[0]
            t.dim = t.ifD.shape[2]
            t.facetN = (t.facetV >= 0).sum(axis=1)
            _tables[name] = t
    return _tables


# ---------------------------------------------------------------------------------------------------------
# mesh ingestion  (PyEFVLib/MSHReader.py:14-76, PyEFVLib/GridData.py:5-24)
# ---------------------------------------------------------------------------------------------------------
class OracleGridData:
    """Plain-list mesh description with the field names of PyEFVLib/GridData.py:5-24."""

    def __init__(self, dimension, verticesCoordinates, elementsConnectivities, regionsNames,
                 regionsElementsIndexes, boundariesNames, boundariesIndexes, boundariesConnectivities):
        self.dimension = dimension
        self.verticesCoordinates = verticesCoordinates
        self.elementsConnectivities = elementsConnectivities
        self.regionsNames = regionsNames
        self.regionsElementsIndexes = regionsElementsIndexes
        self.boundariesNames = boundariesNames
        self.boundariesIndexes = boundariesIndexes
        self.boundariesConnectivities = boundariesConnectivities


def read_msh(path):
    """Positional Gmsh-2.2 parse: split on '$', take sections 1,3,5,7 (MSHReader.py:14-17); regions are the
    physical groups of maximal dimension, boundaries the others, in file order (MSHReader.py:48-66)."""
    with open(path, "r") as f:
        parts = f.read().split("$")
    data = [[line.split() for line in parts[i].split("\n")[1:-1]] for i in range(1, 8, 2)]
    version = float(data[0][0][0])
    if version < 2.0 or version > 2.2:
        raise Exception("MSH version must be 2.2")
    sections = [(int(d), int(i), n[1:-1]) for d, i, n in data[1][1:]]
    vertices = [(float(x), float(y), float(z)) for _, x, y, z in data[2][1:]]
    conns = [(int(p_id), [int(v) - 1 for v in nodes]) for _, c1, c2, p_id, e_id, *nodes in data[3][1:]]
    section_elems = [[e[1] for e in conns if e[0] == s[1]] for s in sections]
    max_dim = max(s[0] for s in sections)
    ec, rn, ri, bc, bn, bi = [], [], [], [], [], []
    for (dim, _, name), elems in zip(sections, section_elems):
        if dim == max_dim:
            first = len(ec)
            ec += elems
            rn.append(name)
            ri.append(list(range(first, first + len(elems))))
        else:
            first = len(bc)
            bc += elems
            bn.append(name)
            bi.append(list(range(first, first + len(elems))))
    return OracleGridData(max_dim, vertices, ec, rn, ri, bn, bi, bc)


def structured_grid_data(n, kind="hex"):
    """Synthetic unit-cube mesh of SURVEY.md section 8d: n^3 vertices, id = i + n*j + n*n*k, hex local order as
    meshes/msh/3D/Hexas3x3x3.msh, tets = the 6-tet split of meshes/msh/3D/Tetras.msh; boundaries West East
    South North Bottom Top (meshes/msh/3D/Hexas.msh:6-12)."""
    ax = np.linspace(0.0, 1.0, n)
    K, J, I = np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij")
    coords = np.stack([ax[I.ravel()], ax[J.ravel()], ax[K.ravel()]], axis=1)
    vid = lambda i, j, k: i + n * j + n * n * k
    k, j, i = np.meshgrid(np.arange(n - 1), np.arange(n - 1), np.arange(n - 1), indexing="ij")
    i, j, k = i.ravel(), j.ravel(), k.ravel()
    corner = lambda bits: vid(i + bits[0], j + bits[1], k + bits[2])
    hexc = np.stack([corner(b) for b in [(0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 1, 0), (0, 0, 1), (1, 0, 1), (1, 1, 1), (0, 1, 1)]], axis=1)
    if kind == "hex":
        conn = hexc
    else:
        split = [[(0, 0, 0), (1, 0, 0), (0, 1, 0), (0, 0, 1)], [(1, 0, 0), (1, 0, 1), (0, 1, 1), (0, 0, 1)],
                 [(1, 0, 0), (0, 1, 1), (0, 1, 0), (0, 0, 1)], [(1, 0, 0), (1, 1, 0), (0, 1, 0), (0, 1, 1)],
                 [(1, 0, 0), (1, 0, 1), (1, 1, 0), (0, 1, 1)], [(1, 1, 0), (0, 1, 1), (1, 0, 1), (1, 1, 1)]]
        tets = np.stack([np.stack([corner(b) for b in t], axis=1) for t in split], axis=1)  # [cube, 6, 4]
        conn = tets.reshape(-1, 4)
    # boundary quads (hex) or triangles (tet: each quad split along a diagonal consistent with the tets)
    a, b = np.meshgrid(np.arange(n - 1), np.arange(n - 1), indexing="ij")
    a, b = a.ravel(), b.ravel()
    o = np.zeros_like(a)
    L = o + (n - 1)

    def quad(f):  # f(a,b) -> 4 vertex ids
        return np.stack(f, axis=1)
    faces = {
        "West": quad([vid(o, b, a), vid(o, b, a + 1), vid(o, b + 1, a + 1), vid(o, b + 1, a)]),
        "East": quad([vid(L, b, a), vid(L, b + 1, a), vid(L, b + 1, a + 1), vid(L, b, a + 1)]),
        "South": quad([vid(b, o, a), vid(b + 1, o, a), vid(b + 1, o, a + 1), vid(b, o, a + 1)]),
        "North": quad([vid(b, L, a), vid(b, L, a + 1), vid(b + 1, L, a + 1), vid(b + 1, L, a)]),
        "Bottom": quad([vid(b, a, o), vid(b, a + 1, o), vid(b + 1, a + 1, o), vid(b + 1, a, o)]),
        "Top": quad([vid(b, a, L), vid(b + 1, a, L), vid(b + 1, a + 1, L), vid(b, a + 1, L)]),
    }
    bconn, bidx, names = [], [], []
    if kind == "hex":
        for name, q in faces.items():
            first = len(bconn)
            bconn += q.tolist()
            names.append(name)
            bidx.append(list(range(first, first + len(q))))
    else:
        tetfaces = set()
        for t in conn.tolist():
            for tri in ((t[0], t[2], t[1]), (t[0], t[3], t[2]), (t[0], t[1], t[3]), (t[1], t[2], t[3])):
                tetfaces.add(tuple(sorted(tri)))
        for name, q in faces.items():
            first = len(bconn)
            for v in q.tolist():
                if tuple(sorted((v[0], v[1], v[2]))) in tetfaces:
                    tris = [[v[0], v[1], v[2]], [v[0], v[2], v[3]]]
                else:
                    tris = [[v[0], v[1], v[3]], [v[1], v[2], v[3]]]
                for t in tris:
                    assert tuple(sorted(t)) in tetfaces
                bconn += tris
            names.append(name)
            bidx.append(list(range(first, len(bconn))))
    return OracleGridData(3, coords.tolist(), conn.tolist(), ["Body"], [list(range(len(conn)))], names, bidx, bconn)


# ---------------------------------------------------------------------------------------------------------
# geometry  (PyEFVLib/Grid.py:18-65, Element.py:29-48,69-74, InnerFace.py:13-25, Shape.py area formulas)
# ---------------------------------------------------------------------------------------------------------
class _Group:
    """All elements of one shape."""


class _Boundary:
    pass


class OracleGrid:
    def __init__(self, gd):
        self.gridData = gd
        self.dimension = d = gd.dimension
        self.coords = np.asarray(gd.verticesCoordinates, dtype=np.float64).reshape(-1, 3)
        self.numberOfVertices = nV = len(self.coords)
        conn_list = gd.elementsConnectivities
        self.numberOfElements = nE = len(conn_list)
        lens = np.fromiter((len(c) for c in conn_list), dtype=np.int64, count=nE)
        self.region_of_element = np.zeros(nE, dtype=np.int32)
        for r, idx in enumerate(gd.regionsElementsIndexes):
            self.region_of_element[np.asarray(idx, dtype=np.int64)] = r
        self.vertex_volume = np.zeros(nV)
        self.groups = []
        self.group_of_element = np.zeros(nE, dtype=np.int32)
        self.index_in_group = np.zeros(nE, dtype=np.int64)
        T = tables()
        for m in sorted(set(lens.tolist())):
            if (m, d) not in SHAPE_OF:
                raise Exception("This element has not been registered in Grid yet")  # Element.py:27
            g = _Group()
            g.tab = T[SHAPE_OF[(m, d)]]
            g.ids = np.nonzero(lens == m)[0]
            g.conn = np.array([conn_list[e] for e in g.ids], dtype=np.int64).reshape(-1, m)
            self.group_of_element[g.ids] = len(self.groups)
            self.index_in_group[g.ids] = np.arange(len(g.ids))
            self._group_geometry(g)
            self.groups.append(g)
        # Vertex.volume accumulates sub-element volumes in element order (Element.py:46)
        order = []
        for g in self.groups:
            order.append((g.ids, g.conn, g.vol))
        for ids, conn, vol in order:
            np.add.at(self.vertex_volume, conn.ravel(), vol.ravel())
        self._boundaries()

    # -- inner faces and sub-elements ---------------------------------------------------------------------
    def _group_geometry(self, g, chunk=200000):
        t, d = g.tab, self.dimension
        n = len(g.ids)
        g.G = np.empty((n, t.nf, d, t.m))
        g.S = np.empty((n, t.nf, 3))
        g.vol = np.empty((n, t.m))
        for a in range(0, n, chunk):
            sl = slice(a, min(n, a + chunk))
            X3 = self.coords[g.conn[sl]]                       # [e, m, 3]
            X = X3[:, :, :d]                                   # Element.py:70
            # Jt = D^T X (Element.py:71);  G = inv(Jt) D^T (InnerFace.py:25)
            Jt = np.einsum("fma,emk->efak", t.ifD, X)
            g.G[sl] = np.einsum("efab,fmb->efam", np.linalg.inv(Jt), t.ifD)
            # sub-element volumes: T_k det(D_k^T X) (Element.py:44)
            Js = np.einsum("kma,emb->ekab", t.subD, X)
            g.vol[sl] = t.subT[None, :] * np.linalg.det(Js)
            # area vectors
            c = X3.sum(axis=1) / t.m                           # Element.py:59 centroid
            vb = X3[:, t.ifNbr[:, 0], :]
            vf = X3[:, t.ifNbr[:, 1], :]
            if d == 2:                                         # Shape.py:29-33, 57-61
                av = c[:, None, :] - (vb + vf) / 2.0
                g.S[sl] = np.stack([av[..., 1], -av[..., 0], np.zeros_like(av[..., 0])], axis=-1)
            else:
                p0 = c[:, None, :]
                p2 = (vb + vf) / 2.0
                if t.name == "Tetrahedron":                    # Shape.py:84-96
                    p1 = (vb + vf + X3[:, t.ifNbr[:, 2], :]) / 3.0
                    p3 = (vb + vf + X3[:, t.ifNbr[:, 3], :]) / 3.0
                elif t.name == "Hexahedron":                   # Shape.py:119-139
                    p1 = (vb + vf + X3[:, t.ifNbr[:, 2], :] + X3[:, t.ifNbr[:, 3], :]) / 4.0
                    p3 = (vb + vf + X3[:, t.ifNbr[:, 4], :] + X3[:, t.ifNbr[:, 5], :]) / 4.0
                elif t.name == "Prism":                        # Shape.py:162-185: quad facet, then triangle (faces 0-5) or quad
                    p1 = (vb + vf + X3[:, t.ifNbr[:, 2], :] + X3[:, t.ifNbr[:, 3], :]) / 4.0
                    p3 = (vb + vf + X3[:, t.ifNbr[:, 4], :]) / 3.0
                    p3[:, 6:] = (vb[:, 6:] + vf[:, 6:] + X3[:, t.ifNbr[6:, 4], :] + X3[:, t.ifNbr[6:, 5], :]) / 4.0
                else:                                          # Pyramid, Shape.py:208-247: p0 is the centre of the BASE
                    p0 = (X3[:, :4, :].sum(axis=1) / 4.0)[:, None, :]
                    p1 = (vb + vf + X3[:, t.ifNbr[:, 2], :]) / 3.0
                    p1[:, :4] = (vb[:, :4] + vf[:, :4] + X3[:, 4:5, :]) / 3.0
                    p3 = (vb + vf + X3[:, t.ifNbr[:, 3], :]) / 3.0
                S = 0.5 * (np.cross(p1 - p0, p3 - p0) + np.cross(p3 - p2, p1 - p2))
                if t.name == "Pyramid":                        # base edges: the triangle p0-p1-p2 only (Shape.py:222-231)
                    S[:, :4] = 0.5 * np.cross(p1[:, :4] - p0, p2[:, :4] - p0)
                g.S[sl] = S

    # -- boundaries -----------------------------------------------------------------------------------------
    def _boundaries(self):
        """Facet.py:22-36 (area), Boundary.py:19-25 (vertex order), :27-50 (facet->element), :52-60 (outward
        fix), :62-75 + OuterFace.py:18-19 (outer faces)."""
        gd, d = self.gridData, self.dimension
        nV = self.numberOfVertices
        self.boundaries = []
        is_bv = np.zeros(nV, dtype=bool)
        for name, fidx in zip(gd.boundariesNames, gd.boundariesIndexes):
            b = _Boundary()
            b.name = name
            b.facets = [list(gd.boundariesConnectivities[i]) for i in fidx]
            # ordered unique vertices, first appearance (Boundary.py:19-25)
            flat = np.array([v for f in b.facets for v in f], dtype=np.int64)
            _, first = np.unique(flat, return_index=True)
            b.vertices = flat[np.sort(first)]
            is_bv[b.vertices] = True
            self.boundaries.append(b)
        # facet vertex set -> (first element in grid order containing it, local facet)  (Boundary.py:37-47)
        owner = {}
        for g in self.groups:
            for l in range(len(g.tab.facetV)):
                fverts = g.conn[:, g.tab.facetV[l, :g.tab.facetN[l]]]      # [e, mf] (facets of a prism / pyramid differ in size)
                for a in np.nonzero(is_bv[fverts].all(axis=1))[0].tolist():
                    key = tuple(sorted(fverts[a].tolist()))
                    e = int(g.ids[a])
                    if key not in owner or e < owner[key][0]:
                        owner[key] = (e, l)
        for b in self.boundaries:
            nF = len(b.facets)
            b.facet_element = np.zeros(nF, dtype=np.int64)
            b.facet_local = np.zeros(nF, dtype=np.int64)
            b.facet_area = np.zeros((nF, 3))
            b.outer_vertex, b.outer_area, b.outer_element, b.outer_N = [], [], [], []
            for k, fv in enumerate(b.facets):
                e, lf = owner[tuple(sorted(fv))]
                b.facet_element[k], b.facet_local[k] = e, lf
                P = self.coords[np.asarray(fv)]
                if len(fv) == 2:                                # Facet.py:23-24
                    area = np.array([P[0, 1] - P[1, 1], P[1, 0] - P[0, 0], 0.0])
                elif len(fv) == 3:                              # Facet.py:26-28
                    area = np.cross(P[1] - P[0], P[2] - P[0]) / 2
                else:                                           # Facet.py:30-36
                    CM, LR = P[2] - P[0], P[3] - P[1]
                    area = 0.5 * np.array([CM[1] * LR[2] - CM[2] * LR[1], CM[2] * LR[0] - CM[0] * LR[2], CM[0] * LR[1] - CM[1] * LR[0]])
                centroid = P.sum(axis=0) / len(fv)              # Facet.py:38-39
                g = self.groups[self.group_of_element[e]]
                econn = g.conn[self.index_in_group[e]]
                inner = next(v for v in econn if v not in fv)   # Boundary.py:57
                if np.dot(area, centroid - self.coords[inner]) < 0:     # Boundary.py:58-60
                    area = area * -1
                b.facet_area[k] = area
                locals_in_elem = list(g.tab.facetV[lf])
                for v in fv:                                    # Facet.py:41-45, Boundary.py:62-75
                    lv = int(np.nonzero(econn == v)[0][0])
                    ol = locals_in_elem.index(lv)
                    b.outer_vertex.append(v)
                    b.outer_area.append(area / len(fv))         # OuterFace.py:18-19
                    b.outer_element.append(e)
                    b.outer_N.append(g.tab.ofN[lf][ol])
            b.outer_vertex = np.array(b.outer_vertex, dtype=np.int64)
            b.outer_area = np.array(b.outer_area).reshape(-1, 3)
            b.outer_element = np.array(b.outer_element, dtype=np.int64)


# ---------------------------------------------------------------------------------------------------------
# CSR pattern and slot map  (LinearSystem.py:113-128 structure; SURVEY.md appendix B.6 canonical form)
# ---------------------------------------------------------------------------------------------------------
def csr_pattern(grid, ndof=1):
    """Union of element cliques, sorted columns; DOFs field-major: row = v + i*nV and, inside a row, DOF block
    j then stencil vertex (LinearSystem.py:117-124 with each stencil sorted).  int64 indptr, int32 indices."""
    nV = grid.numberOfVertices
    rows, cols = [], []
    for g in grid.groups:
        m = g.tab.m
        rows.append(np.repeat(g.conn, m, axis=1).ravel())
        cols.append(np.tile(g.conn, (1, m)).ravel())
    rows, cols = np.concatenate(rows), np.concatenate(cols)
    P = sp.csr_matrix((np.ones(len(rows), dtype=np.int8), (rows, cols)), shape=(nV, nV))
    P.sum_duplicates()
    P.sort_indices()
    gptr, gcol = P.indptr.astype(np.int64), P.indices.astype(np.int32)
    if ndof == 1:
        return gptr, gcol
    nnzg = len(gcol)
    rowlen = np.diff(gptr)
    indptr = np.zeros(ndof * nV + 1, dtype=np.int64)
    indptr[1:] = np.cumsum(np.tile(rowlen * ndof, ndof))
    indices = np.empty(ndof * ndof * nnzg, dtype=np.int32)
    for i in range(ndof):
        base = ndof * i * nnzg
        # row v of block-row i: [gcol(v) + 0*nV | gcol(v) + 1*nV | ...]
        starts = ndof * gptr[:-1] + base
        for j in range(ndof):
            pos = np.repeat(starts + j * rowlen, rowlen) + (np.arange(nnzg) - np.repeat(gptr[:-1], rowlen))
            indices[pos] = gcol + j * nV
    return indptr, indices


def slot_map(grid, gptr, gcol):
    """Element-to-CSR slot map on the vertex graph: for element e (grid order) and local pair (i, j) the index
    into `gcol` of entry (conn[e][i], conn[e][j]).  Returned as a flat int64 array in element order, each
    element contributing m*m entries, i-major."""
    nV = grid.numberOfVertices
    keys = np.repeat(np.arange(nV, dtype=np.int64), np.diff(gptr)) * nV + gcol
    out = [None] * len(grid.groups)
    sizes = np.zeros(grid.numberOfElements, dtype=np.int64)
    for gi, g in enumerate(grid.groups):
        m = g.tab.m
        q = np.repeat(g.conn, m, axis=1) * nV + np.tile(g.conn, (1, m))
        out[gi] = np.searchsorted(keys, q)
        sizes[g.ids] = m * m
    offs = np.concatenate([[0], np.cumsum(sizes)])
    flat = np.empty(offs[-1], dtype=np.int64)
    for gi, g in enumerate(grid.groups):
        m2 = g.tab.m ** 2
        idx = offs[g.ids][:, None] + np.arange(m2)[None, :]
        flat[idx] = out[gi]
    return flat


# ---------------------------------------------------------------------------------------------------------
# heat transfer  (apps/heat_transfer.py:40-132)
# ---------------------------------------------------------------------------------------------------------
def _props(props, grid, name):
    """Per-element property from a {regionName: {prop: value}} dict (ProblemData.py:36-43)."""
    per_region = np.array([props[r][name] for r in grid.gridData.regionsNames], dtype=np.float64)
    return per_region[grid.region_of_element]


def dirichlet_rows(grid, bcs, ndof_offset=0):
    """Ordered (row, value) pairs exactly as the reference visits them: boundaries in grid order, vertices in
    first-appearance order; later entries overwrite earlier ones (heat_transfer.py:122-126)."""
    rows, vals = [], []
    for b in grid.boundaries:
        bc = bcs.get(b.name)
        if bc is not None and bc[0] == "dirichlet":
            rows.append(b.vertices + ndof_offset)
            vals.append(np.full(len(b.vertices), float(bc[1])))
    if not rows:
        return np.zeros(0, dtype=np.int64), np.zeros(0)
    return np.concatenate(rows), np.concatenate(vals)


def heat_interior_coo(grid, props, transient, dt):
    """COO triplets of diffusionTerm (+ accumulationTerm) before Dirichlet (heat_transfer.py:41-68)."""
    R, C, V = [], [], []
    d = grid.dimension
    for g in grid.groups:
        t = g.tab
        k = _props(props, grid, "Conductivity")[g.ids]
        # diffusiveFlux = k * G^T s[:d]   (heat_transfer.py:47)
        w = k[:, None, None] * np.einsum("efam,efa->efm", g.G, g.S[:, :, :d])
        b = g.conn[:, t.ifNbr[:, 0]]
        f = g.conn[:, t.ifNbr[:, 1]]
        cols = np.broadcast_to(g.conn[:, None, :], w.shape)
        R += [np.broadcast_to(b[:, :, None], w.shape).ravel(), np.broadcast_to(f[:, :, None], w.shape).ravel()]
        C += [cols.ravel(), cols.ravel()]
        V += [(-w).ravel(), w.ravel()]                          # heat_transfer.py:52-54
        if transient:
            acc = (_props(props, grid, "Density") * _props(props, grid, "HeatCapacity"))[g.ids] / dt
            R.append(g.conn.ravel()); C.append(g.conn.ravel())
            V.append((g.vol * acc[:, None]).ravel())            # heat_transfer.py:62-67
    return np.concatenate(R), np.concatenate(C), np.concatenate(V)


def heat_matrix(grid, props, bcs, transient=False, dt=None):
    """`solver.matrix` of the reference: csc(COO) after deleting Dirichlet rows and adding a unit diagonal
    (heat_transfer.py:70-80).  bcs: {boundaryName: ("dirichlet"|"neumann", value)}."""
    nV = grid.numberOfVertices
    R, C, V = heat_interior_coo(grid, props, transient, dt)
    drows, _ = dirichlet_rows(grid, bcs)
    is_d = np.zeros(nV, dtype=bool)
    is_d[drows] = True
    keep = ~is_d[R]
    dr = np.nonzero(is_d)[0]
    R = np.concatenate([R[keep], dr]); C = np.concatenate([C[keep], dr]); V = np.concatenate([V[keep], np.ones(len(dr))])
    return sp.csc_matrix((V, (R, C)), shape=(nV, nV))


def heat_rhs(grid, props, bcs, transient=False, dt=None, Told=None):
    """`solver.independent` (heat_transfer.py:89-132)."""
    nV = grid.numberOfVertices
    b = np.zeros(nV)
    for g in grid.groups:
        q = _props(props, grid, "HeatGeneration")[g.ids]
        np.add.at(b, g.conn.ravel(), (g.vol * q[:, None]).ravel())                      # :92-100
    if transient:
        for g in grid.groups:
            acc = (_props(props, grid, "Density") * _props(props, grid, "HeatCapacity"))[g.ids] / dt
            np.add.at(b, g.conn.ravel(), (g.vol * acc[:, None] * Told[g.conn]).ravel())  # :102-113
    for bd in grid.boundaries:                                                           # :115-120
        bc = bcs.get(bd.name)
        if bc is not None and bc[0] == "neumann":
            np.add.at(b, bd.outer_vertex, float(bc[1]) * np.linalg.norm(bd.outer_area, axis=1))
    rows, vals = dirichlet_rows(grid, bcs)                                               # :122-126
    for r, v in zip(rows, vals):
        b[r] = v
    return b


# ---------------------------------------------------------------------------------------------------------
# linear elasticity  (apps/stress_equilibrium.py:29-163)
# ---------------------------------------------------------------------------------------------------------
def constitutive_matrix(G, nu, d):
    """stress_equilibrium.py:29-48."""
    lame = 2 * G * nu / (1 - 2 * nu)
    if d == 2:
        return np.array([[2 * G + lame, lame, 0], [lame, 2 * G + lame, 0], [0, 0, G]], dtype=np.float64)
    return np.array([[2 * G + lame, lame, lame, 0, 0, 0], [lame, 2 * G + lame, lame, 0, 0, 0], [lame, lame, 2 * G + lame, 0, 0, 0],
                     [0, 0, 0, G, 0, 0], [0, 0, 0, 0, G, 0], [0, 0, 0, 0, 0, G]], dtype=np.float64)


def _voigt_area(S, d):
    """Transposed Voigt area [.., d, nvoigt] (stress_equilibrium.py:50-55)."""
    Sx, Sy, Sz = S[..., 0], S[..., 1], S[..., 2]
    z = np.zeros_like(Sx)
    if d == 2:
        return np.stack([np.stack([Sx, z, Sy], -1), np.stack([z, Sy, Sx], -1)], -2)
    return np.stack([np.stack([Sx, z, z, Sy, z, Sz], -1), np.stack([z, Sy, z, Sx, Sz, z], -1),
                     np.stack([z, z, Sz, z, Sy, Sx], -1)], -2)


def _voigt_grad(G, d):
    """Voigt gradient operator [.., nvoigt, d, m] (stress_equilibrium.py:57-67)."""
    z = np.zeros_like(G[..., 0, :])
    if d == 2:
        Nx, Ny = G[..., 0, :], G[..., 1, :]
        return np.stack([np.stack([Nx, z], -2), np.stack([z, Ny], -2), np.stack([Ny, Nx], -2)], -3)
    Nx, Ny, Nz = G[..., 0, :], G[..., 1, :], G[..., 2, :]
    return np.stack([np.stack([Nx, z, z], -2), np.stack([z, Ny, z], -2), np.stack([z, z, Nz], -2),
                     np.stack([Ny, Nx, z], -2), np.stack([z, Nz, Ny], -2), np.stack([Nz, z, Nx], -2)], -3)


def _region_param(grid, props, fn):
    per_region = [fn(props[r]) for r in grid.gridData.regionsNames]
    return per_region


def elasticity_interior_coo(grid, props):
    """stressTerm triplets (stress_equilibrium.py:92-117): rows of b get +M, rows of f get -M."""
    d, nV = grid.dimension, grid.numberOfVertices
    Ce_r = np.array(_region_param(grid, props, lambda p: constitutive_matrix(p["ShearModulus"], p["PoissonsRatio"], d)))
    R, C, V = [], [], []
    for g in grid.groups:
        t = g.tab
        Ce = Ce_r[grid.region_of_element[g.ids]]                                 # [e, nv, nv]
        At = _voigt_area(g.S, d)                                                  # [e, f, d, nv]
        Bv = _voigt_grad(g.G, d)                                                  # [e, f, nv, d, m]
        M = np.einsum("efia,eab,efbjl->efijl", At, Ce, Bv)                        # :100
        b = g.conn[:, t.ifNbr[:, 0]]
        f = g.conn[:, t.ifNbr[:, 1]]
        for i in range(d):
            for j in range(d):
                Mij = M[:, :, i, j, :]
                cols = np.broadcast_to(g.conn[:, None, :], Mij.shape) + j * nV
                R += [(np.broadcast_to(b[:, :, None], Mij.shape) + i * nV).ravel(), (np.broadcast_to(f[:, :, None], Mij.shape) + i * nV).ravel()]
                C += [cols.ravel(), cols.ravel()]
                V += [Mij.ravel(), (-Mij).ravel()]
    return np.concatenate(R), np.concatenate(C), np.concatenate(V)


def elasticity_system(grid, props, bcs, gravity=False):
    """Matrix and RHS of StressEquilibriumSolver.assembleLinearSystem (stress_equilibrium.py:74-168).
    bcs: {"u": {boundary: (kind, value)}, "v": {...}, ["w": {...}]}."""
    d, nV = grid.dimension, grid.numberOfVertices
    n = d * nV
    b = np.zeros(n)
    if gravity:                                                                   # :81-90 (y component only)
        for g in grid.groups:
            rho = _props(props, grid, "Density")[g.ids]
            grav = _props(props, grid, "Gravity")[g.ids]
            np.add.at(b, g.conn.ravel() + nV, (-(rho * grav)[:, None] * g.vol).ravel())
    R, C, V = elasticity_interior_coo(grid, props)
    names = ["u", "v", "w"][:d]
    is_d = np.zeros(n, dtype=bool)
    for bd in grid.boundaries:                                                    # :119-163
        for i, var in enumerate(names):
            kind, val = bcs[var][bd.name]
            if kind == "neumann":
                np.add.at(b, bd.outer_vertex + i * nV, -float(val) * np.linalg.norm(bd.outer_area, axis=1))
        for i, var in enumerate(names):
            kind, val = bcs[var][bd.name]
            if kind == "dirichlet":
                b[bd.vertices + i * nV] = float(val)
                is_d[bd.vertices + i * nV] = True
    keep = ~is_d[R]
    dr = np.nonzero(is_d)[0]
    A = sp.csc_matrix((np.concatenate([V[keep], np.ones(len(dr))]),
                       (np.concatenate([R[keep], dr]), np.concatenate([C[keep], dr]))), shape=(n, n))
    return A, b


# ---------------------------------------------------------------------------------------------------------
# Biot poroelasticity  (apps/geomechanics.py:39-241)
# ---------------------------------------------------------------------------------------------------------
def biot_params(p, d):
    """geomechanics.py:41-58 including the precedence quirk of K (:55) and k*mu (:81)."""
    nu, G = p["nu"], p["G"]
    lame = 2 * G * nu / (1 - 2 * nu)
    Ce = constitutive_matrix(G, nu, d)
    rho = p["phi"] * p["rhof"] + (1 - p["phi"]) * p["rhos"]
    K = 2 * G * (1 + nu) / 3 * (1 - 2 * nu)
    cb = 1 / K
    alpha = 1 - p["cs"] / cb
    S = p["phi"] * p["cf"] + (alpha - p["phi"]) * p["cs"]
    return dict(Ce=Ce, rho=rho, alpha=alpha, S=S, kmu=p["k"] * p["mu"], lame=lame)


def _outer_faces_by_group(grid):
    """All outer faces (every boundary, grid order): element, vertex, area vector, shape values."""
    el = np.concatenate([b.outer_element for b in grid.boundaries]) if grid.boundaries else np.zeros(0, dtype=np.int64)
    vx = np.concatenate([b.outer_vertex for b in grid.boundaries]) if grid.boundaries else np.zeros(0, dtype=np.int64)
    ar = np.concatenate([b.outer_area for b in grid.boundaries]) if grid.boundaries else np.zeros((0, 3))
    Ns = [n for b in grid.boundaries for n in b.outer_N]
    return el, vx, ar, Ns


def biot_system(grid, props, bcs, dt, u_old, p_old, g_vec=None):
    """Matrix (pre-inverse) and independent vector of geomechanics.py:39-141 / :143-241.  Unknown order
    [u_x | u_y | (u_z) | p].  bcs: {"u_x": {...}, "u_y": {...}, ["u_z"], "p": {...}}.  Returns scipy csr, rhs."""
    d, nV = grid.dimension, grid.numberOfVertices
    n = (d + 1) * nV
    gv = np.zeros(d) if g_vec is None else np.asarray(g_vec, dtype=np.float64)[:d]
    par_r = _region_param(grid, props, lambda p: biot_params(p, d))
    R, C, V = [], [], []
    rhs = np.zeros(n)
    P0 = d * nV

    def add(r, c, v):
        R.append(np.asarray(r).ravel()); C.append(np.asarray(c).ravel()); V.append(np.asarray(v).ravel())
    oel, ovx, oar, oN = _outer_faces_by_group(grid)
    for g in grid.groups:
        t = g.tab
        reg = grid.region_of_element[g.ids]
        alpha = np.array([par_r[r]["alpha"] for r in reg])
        S = np.array([par_r[r]["S"] for r in reg])
        kmu = np.array([par_r[r]["kmu"] for r in reg])
        rho = np.array([par_r[r]["rho"] for r in reg])
        Ce = np.array([par_r[r]["Ce"] for r in reg])
        b = g.conn[:, t.ifNbr[:, 0]]
        f = g.conn[:, t.ifNbr[:, 1]]
        Sd = g.S[:, :, :d]
        shp3 = (len(g.ids), t.nf, t.m)
        cols = np.broadcast_to(g.conn[:, None, :], shp3)
        B3 = np.broadcast_to(b[:, :, None], shp3)
        F3 = np.broadcast_to(f[:, :, None], shp3)
        # S/dt * vol on pp diagonal (:62-63)
        add(g.conn + P0, g.conn + P0, g.vol * (S / dt)[:, None])
        # -alpha * At [N;N;N;0;0;0]  -> u rows x p cols (:66-77): coefficient[coord][local] = -alpha s[coord] N[local]
        for c in range(d):
            co = -alpha[:, None, None] * Sd[:, :, c, None] * t.ifN[None, :, :]
            add(B3 + c * nV, cols + P0, co)
            add(F3 + c * nV, cols + P0, -co)
        # -k*mu * s^T G -> pp (:80-86)
        co = -kmu[:, None, None] * np.einsum("efa,efam->efm", Sd, g.G)
        add(B3 + P0, cols + P0, co)
        add(F3 + P0, cols + P0, -co)
        # At Ce Bv -> uu (:89-100)
        M = np.einsum("efia,eab,efbjl->efijl", _voigt_area(g.S, d), Ce, _voigt_grad(g.G, d))
        for i in range(d):
            for j in range(d):
                add(B3 + i * nV, cols + j * nV, M[:, :, i, j, :])
                add(F3 + i * nV, cols + j * nV, -M[:, :, i, j, :])
        # alpha/dt * N_l * s_c -> p rows x u cols, inner faces (:103-112)
        for c in range(d):
            co = (alpha / dt)[:, None, None] * t.ifN[None, :, :] * Sd[:, :, c, None]
            add(B3 + P0, cols + c * nV, co)
            add(F3 + P0, cols + c * nV, -co)
            # RHS alpha/dt * N s u_old (:185-193)
            uo = u_old[g.conn + c * nV]                                            # [e, m]
            val = np.einsum("efm,em->ef", co, uo)
            np.add.at(rhs, (b + P0).ravel(), val.ravel())
            np.add.at(rhs, (f + P0).ravel(), -val.ravel())
        # body force (:166-169), S/dt p_old (:172-173), gravity flux (:176-183)
        for c in range(d):
            np.add.at(rhs, g.conn.ravel() + c * nV, (g.vol * (-rho * gv[c])[:, None]).ravel())
        np.add.at(rhs, g.conn.ravel() + P0, (g.vol * (S / dt)[:, None] * p_old[g.conn]).ravel())
        co = np.einsum("efa,ea->ef", Sd, (-kmu * rho)[:, None] * gv[None, :])
        np.add.at(rhs, (b + P0).ravel(), co.ravel())
        np.add.at(rhs, (f + P0).ravel(), -co.ravel())
    # outer faces: alpha/dt * N_l * s_c at row p(face.vertex) (:113-115, :194-196)
    for k in range(len(oel)):
        e = int(oel[k])
        g = grid.groups[grid.group_of_element[e]]
        econn = g.conn[grid.index_in_group[e]]
        alpha = par_r[grid.region_of_element[e]]["alpha"]
        for c in range(d):
            co = alpha / dt * oN[k] * oar[k, c]
            add(np.full(len(econn), ovx[k] + P0), econn + c * nV, co)
            rhs[ovx[k] + P0] += np.dot(co, u_old[econn + c * nV])
    names = (["u_x", "u_y", "u_z"][:d]) + ["p"]
    for i, var in enumerate(names):                                                # Neumann (:202-221)
        for bd in grid.boundaries:
            kind, val = bcs[var][bd.name]
            if kind == "neumann":
                np.add.at(rhs, bd.outer_vertex + i * nV, float(val) * np.linalg.norm(bd.outer_area, axis=1))
    is_d = np.zeros(n, dtype=bool)
    for i, var in enumerate(names):                                                # Dirichlet (:224-239, :117-137)
        for bd in grid.boundaries:
            kind, val = bcs[var][bd.name]
            if kind == "dirichlet":
                rhs[bd.vertices + i * nV] = float(val)
                is_d[bd.vertices + i * nV] = True
    R, C, V = np.concatenate(R), np.concatenate(C), np.concatenate(V)
    keep = ~is_d[R]
    dr = np.nonzero(is_d)[0]
    A = sp.csr_matrix((np.concatenate([V[keep], np.ones(len(dr))]),
                       (np.concatenate([R[keep], dr]), np.concatenate([C[keep], dr]))), shape=(n, n))
    return A, rhs


# ---------------------------------------------------------------------------------------------------------
# fixed-stress split  (apps/fixed_stress.py:39-329)
# ---------------------------------------------------------------------------------------------------------
def fixed_stress_operators(grid, props, bcs, dt):
    """Mass (p) and geo (u) matrices of fixed_stress.py:39-117 plus the coupling blocks its right-hand sides apply
    (:163-184 mass, :239-248 geo).  All of them are blocks of the coupled Biot operator (same element terms), except the
    extra  cb alpha^2 / dt  vol  on the mass diagonal (:68)."""
    d, nV = grid.dimension, grid.numberOfVertices
    names = (["u_x", "u_y", "u_z"][:d]) + ["p"]
    free = {v: {b.name: ("neumann", 0.0) for b in grid.boundaries} for v in names}
    A, _ = biot_system(grid, props, free, dt, np.zeros(d * nV), np.zeros(nV))        # no Dirichlet rows: raw operator
    A = sp.csr_matrix(A)
    P0 = d * nV
    Auu, Aup, Apu, App = A[:P0, :P0], A[:P0, P0:], A[P0:, :P0], A[P0:, P0:]
    par_r = _region_param(grid, props, lambda p: biot_params(p, d))
    accS, accC = np.zeros(nV), np.zeros(nV)
    for g in grid.groups:
        reg = grid.region_of_element[g.ids]
        S = np.array([par_r[r]["S"] for r in reg])
        # cb = 1/K with the precedence quirk of K (fixed_stress.py:60-61)
        K = np.array([2 * props[grid.gridData.regionsNames[r]]["G"] * (1 + props[grid.gridData.regionsNames[r]]["nu"]) / 3 *
                      (1 - 2 * props[grid.gridData.regionsNames[r]]["nu"]) for r in reg])
        alpha = np.array([par_r[r]["alpha"] for r in reg])
        np.add.at(accS, g.conn.ravel(), (g.vol * (S / dt)[:, None]).ravel())
        np.add.at(accC, g.conn.ravel(), (g.vol * ((1 / K) * (1 / dt) * alpha ** 2)[:, None]).ravel())
    mass = sp.lil_matrix(App + sp.diags(accC))
    geo = sp.lil_matrix(Auu)
    dir_p = np.zeros(nV, dtype=bool)
    dir_u = np.zeros(P0, dtype=bool)
    for bd in grid.boundaries:
        if bcs["p"][bd.name][0] == "dirichlet":
            dir_p[bd.vertices] = True
        for i, var in enumerate(names[:d]):
            if bcs[var][bd.name][0] == "dirichlet":
                dir_u[bd.vertices + i * nV] = True
    mass = sp.csr_matrix(mass)
    geo = sp.csr_matrix(geo)
    keep_p = sp.diags((~dir_p).astype(float))
    keep_u = sp.diags((~dir_u).astype(float))
    mass = keep_p @ mass + sp.diags(dir_p.astype(float))                               # fixed_stress.py:109-112
    geo = keep_u @ geo + sp.diags(dir_u.astype(float))                                 # fixed_stress.py:94-107
    return dict(mass=sp.csr_matrix(mass), geo=sp.csr_matrix(geo), Aup=Aup, Apu=Apu, accS=accS, accC=accC)


def fixed_stress_mass_rhs(grid, ops, bcs, p_prev, u_prev, p_k, u_k):
    """fixed_stress.py:119-199 (gravity vector is zero in the reference)."""
    b = ops["accS"] * p_prev + ops["accC"] * p_k + ops["Apu"] @ (u_prev - u_k)
    for bd in grid.boundaries:
        kind, val = bcs["p"][bd.name]
        if kind == "neumann":
            np.add.at(b, bd.outer_vertex, float(val) * np.linalg.norm(bd.outer_area, axis=1))
    for bd in grid.boundaries:
        kind, val = bcs["p"][bd.name]
        if kind == "dirichlet":
            b[bd.vertices] = float(val)
    return b


def fixed_stress_geo_rhs(grid, ops, bcs, p_next):
    """fixed_stress.py:201-275."""
    d, nV = grid.dimension, grid.numberOfVertices
    b = -(ops["Aup"] @ p_next)
    names = ["u_x", "u_y", "u_z"][:d]
    for i, var in enumerate(names):
        for bd in grid.boundaries:
            kind, val = bcs[var][bd.name]
            if kind == "neumann":
                np.add.at(b, bd.outer_vertex + i * nV, float(val) * np.linalg.norm(bd.outer_area, axis=1))
    for i, var in enumerate(names):
        for bd in grid.boundaries:
            kind, val = bcs[var][bd.name]
            if kind == "dirichlet":
                b[bd.vertices + i * nV] = float(val)
    return b


def fixed_stress_run(grid, props, bcs, dt, steps, tolerance):
    """Time loop of fixed_stress.py:284-329 with direct solves; returns per time step (u, p, inner iterations)."""
    d, nV = grid.dimension, grid.numberOfVertices
    ops = fixed_stress_operators(grid, props, bcs, dt)
    lu_m, lu_g = spla.splu(sp.csc_matrix(ops["mass"])), spla.splu(sp.csc_matrix(ops["geo"]))
    p_prev, u_prev = np.zeros(nV), np.zeros(d * nV)
    out = []
    for _ in range(steps):
        p_k, u_k = p_prev.copy(), u_prev.copy()
        difference, it = 2 * tolerance, 0
        while difference >= tolerance or it <= 1:
            p_n = lu_m.solve(fixed_stress_mass_rhs(grid, ops, bcs, p_prev, u_prev, p_k, u_k))
            u_n = lu_g.solve(fixed_stress_geo_rhs(grid, ops, bcs, p_n))
            difference = max(np.abs(p_n - p_k).max(), np.abs(u_n - u_k).max())
            p_k, u_k = p_n, u_n
            it += 1
        p_prev, u_prev = p_k.copy(), u_k.copy()
        out.append((u_k.copy(), p_k.copy(), it))
    return out


# ---------------------------------------------------------------------------------------------------------
# solves  (heat_transfer.py:78-81,134-135 ; stress_equilibrium.py:170-173 ; geomechanics.py:140,254)
# ---------------------------------------------------------------------------------------------------------
def solve_direct(A, b):
    """What the reference computes (inv(A) @ b) without forming the inverse: SuperLU spsolve(A, b).
    SURVEY.md section 8c: inv(A)@b and spsolve agree to 2-5e-16 on the bundled heat cases."""
    return spla.spsolve(sp.csc_matrix(A), b)


def solve_equilibrated(A, b, refine=4):
    """Row/column equilibrated LU + long-double iterative refinement: the parity target
    for the badly scaled elasticity / Biot systems where plain spsolve is itself off by 1e-9 .. 3e-5."""
    A = sp.csr_matrix(A)
    r = 1.0 / np.sqrt(np.maximum(np.abs(A).max(axis=1).toarray().ravel(), 1e-300))
    Ar = sp.diags(r) @ A
    c = 1.0 / np.maximum(np.abs(Ar).max(axis=0).toarray().ravel(), 1e-300)
    As = sp.csc_matrix(Ar @ sp.diags(c))
    lu = spla.splu(As)
    bs = r * b
    y = lu.solve(bs)
    Ac = As.tocoo()
    vals = Ac.data.astype(np.longdouble)
    for _ in range(refine):
        res = bs.astype(np.longdouble).copy()
        np.subtract.at(res, Ac.row, vals * y.astype(np.longdouble)[Ac.col])
        y = y + lu.solve(res.astype(np.float64))
    return c * y
