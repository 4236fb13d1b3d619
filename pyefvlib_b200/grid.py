"""Flat-array mesh: GridData / MSHReader / Grid with the reference's attribute names.

The reference builds an object graph (PyEFVLib/Grid.py:18-65: one Python object per vertex, element, inner
face, facet, outer face, ~500 elements/s).  Here the same information is a handful of numpy arrays on the host
and SoA buffers in HBM (see csrc/efv_internal.cuh); `grid.vertices[i]`, `grid.elements[i]`,
`grid.boundaries[i]` ... are thin views created on access so that ProblemData / BoundaryConditions / savers keep
working (SURVEY.md section 8b lists the attributes they touch).
"""
import os

import numpy as np

from . import shapes as _shapes


# ---------------------------------------------------------------------------------------------------------
# GridData  (PyEFVLib/GridData.py:1-24) -- the mesh injection seam
# ---------------------------------------------------------------------------------------------------------
class GridData:
    def __init__(self, path):
        self.path = path

    def setDimension(self, gridDimension):
        self.dimension = gridDimension

    def setVertices(self, verticesCoordinates):
        self.verticesCoordinates = verticesCoordinates

    def setElementConnectivity(self, elementsConnectivities):
        self.elementsConnectivities = elementsConnectivities

    def setRegions(self, regionsNames, regionsElementsIndexes):
        self.regionsNames = regionsNames
        self.regionsElementsIndexes = regionsElementsIndexes

    def setBoundaries(self, boundariesNames, boundariesIndexes, boundariesConnectivities):
        self.boundariesConnectivities = boundariesConnectivities
        self.boundariesNames = boundariesNames
        self.boundariesIndexes = boundariesIndexes

    def setShapes(self, shapes):
        self.lines, self.triangles, self.quadrilaterals, self.tetrahedrons, self.hexahedrons, self.prisms, self.pyramids = shapes


# ---------------------------------------------------------------------------------------------------------
# MSHReader  (PyEFVLib/MSHReader.py:6-76) -- same positional semantics, array based
# ---------------------------------------------------------------------------------------------------------
class MSHReader:
    """Gmsh 2.x ASCII reader.  Like the reference it is positional: the file is split on '$' and sections
    1, 3, 5, 7 are taken as MeshFormat, PhysicalNames, Nodes, Elements whatever their names
    (MSHReader.py:14-17); every element line must carry exactly two tags (MSHReader.py:38)."""

    def __init__(self, path):
        self.path = path
        with open(path, "r") as f:
            parts = f.read().split("$")
        if len(parts) < 8:
            raise Exception("MSH file must have MeshFormat, PhysicalNames, Nodes and Elements sections")
        body = [parts[i].split("\n")[1:-1] for i in range(1, 8, 2)]
        version = float(body[0][0].split()[0])
        if version < 2.0 or version > 2.2:
            raise Exception("MSH version must be 2.2")                                   # MSHReader.py:21-22
        self.sections = []
        for line in body[1][1:]:
            dim, idx, name = line.split()[:3]
            self.sections.append((int(dim), int(idx), name[1:-1]))
        nodes = np.array([ln.split()[1:4] for ln in body[2][1:]], dtype=np.float64).reshape(-1, 3)
        self.verticesCoordinates = nodes
        self.numberOfVertices = len(nodes)
        tags, conns = [], []
        for ln in body[3][1:]:
            t = ln.split()
            tags.append(int(t[3]))
            conns.append([int(v) - 1 for v in t[5:]])
        self.physical = np.array(tags, dtype=np.int64)
        self.connectivities = conns
        self.gridDimension = max(s[0] for s in self.sections)

    def getData(self):
        ec, rn, ri, bc, bn, bi = [], [], [], [], [], []
        for dim, idx, name in self.sections:                                             # MSHReader.py:52-66
            members = [self.connectivities[k] for k in np.nonzero(self.physical == idx)[0]]
            if dim == self.gridDimension:
                first = len(ec)
                ec += members
                rn.append(name)
                ri.append(list(range(first, first + len(members))))
            else:
                first = len(bc)
                bc += members
                bn.append(name)
                bi.append(list(range(first, first + len(members))))
        gd = GridData(self.path)
        gd.setDimension(self.gridDimension)
        gd.setVertices(self.verticesCoordinates)
        gd.setElementConnectivity(ec)
        gd.setRegions(rn, ri)
        gd.setBoundaries(bn, bi, bc)
        gd.setShapes([[]] * 7)
        return gd


# ---------------------------------------------------------------------------------------------------------
# lazy object views
# ---------------------------------------------------------------------------------------------------------
class _VertexView:
    __slots__ = ("_g", "handle")

    def __init__(self, g, handle):
        self._g, self.handle = g, int(handle)

    x = property(lambda s: float(s._g.coords[s.handle, 0]))
    y = property(lambda s: float(s._g.coords[s.handle, 1]))
    z = property(lambda s: float(s._g.coords[s.handle, 2]))

    def getCoordinates(self):
        return self._g.coords[self.handle].copy()

    @property
    def volume(self):
        return float(self._g.device.vertex_volume()[self.handle])

    def __eq__(self, o):
        return isinstance(o, _VertexView) and o.handle == self.handle and o._g is self._g

    def __hash__(self):
        return hash(self.handle)


class _ElementView:
    __slots__ = ("_g", "handle")

    def __init__(self, g, handle):
        self._g, self.handle = g, int(handle)

    @property
    def vertices(self):
        g = self._g
        return [_VertexView(g, v) for v in g.conn[g.elem_ptr[self.handle]:g.elem_ptr[self.handle + 1]]]

    @property
    def shape(self):
        g = self._g
        m = int(g.elem_ptr[self.handle + 1] - g.elem_ptr[self.handle])
        return _shapes.SHAPES[_shapes.shape_code_for(m, g.dimension)]

    @property
    def numberOfVertices(self):
        return int(self._g.elem_ptr[self.handle + 1] - self._g.elem_ptr[self.handle])

    @property
    def region(self):
        return self._g.regions[int(self._g.region_of_element[self.handle])]


class _Seq:
    """Read-only sequence that builds a view per access (grid.vertices, grid.elements)."""

    def __init__(self, n, make):
        self._n, self._make = int(n), make

    def __len__(self):
        return self._n

    size = property(lambda s: s._n)

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self._make(k) for k in range(*i.indices(self._n))]
        i = int(i)
        if i < 0:
            i += self._n
        if not 0 <= i < self._n:
            raise IndexError(i)
        return self._make(i)

    def __iter__(self):
        return (self._make(k) for k in range(self._n))


class Region:
    """PyEFVLib/Region.py:3-9"""

    def __init__(self, grid, elementsIndexes, name, handle):
        self.grid, self.name, self.handle = grid, name, handle
        self.elementsIndexes = np.asarray(elementsIndexes, dtype=np.int64)

    @property
    def elements(self):
        return _Seq(len(self.elementsIndexes), lambda k: _ElementView(self.grid, self.elementsIndexes[k]))


class _FacetView:
    def __init__(self, boundary, k):
        self.boundary, self.handle = boundary, int(boundary.facet_begin + k)
        g = boundary.grid
        self.verticesIndexes = g.facet_conn[g.facet_ptr[self.handle]:g.facet_ptr[self.handle + 1]]

    @property
    def vertices(self):
        return [_VertexView(self.boundary.grid, v) for v in self.verticesIndexes]


class Boundary:
    """PyEFVLib/Boundary.py:6-25: name, handle, facets and the ordered unique vertex list."""

    def __init__(self, grid, facet_begin, facet_end, name, handle):
        self.grid, self.name, self.handle = grid, name, handle
        self.facet_begin, self.facet_end = int(facet_begin), int(facet_end)
        g = grid
        flat = g.facet_conn[g.facet_ptr[self.facet_begin]:g.facet_ptr[self.facet_end]]
        _, first = np.unique(flat, return_index=True)          # first appearance order (Boundary.py:19-25)
        self.verticesIndexes = flat[np.sort(first)].astype(np.int64)

    @property
    def vertices(self):
        return _Seq(len(self.verticesIndexes), lambda k: _VertexView(self.grid, self.verticesIndexes[k]))

    @property
    def facets(self):
        return _Seq(self.facet_end - self.facet_begin, lambda k: _FacetView(self, k))


# ---------------------------------------------------------------------------------------------------------
# Grid  (PyEFVLib/Grid.py:10-65)
# ---------------------------------------------------------------------------------------------------------
class Grid:
    def __init__(self, gridData):
        if not isinstance(gridData, GridData):
            raise Exception("Grid argument must be of class GridData")                   # Grid.py:12-13
        self.gridData = gridData
        self.dimension = int(gridData.dimension)
        self.coords = np.ascontiguousarray(np.asarray(gridData.verticesCoordinates, dtype=np.float64).reshape(-1, 3))
        self.numberOfVertices = len(self.coords)
        ec = gridData.elementsConnectivities
        if isinstance(ec, np.ndarray) and ec.ndim == 2:
            self.numberOfElements, m = ec.shape
            self.elem_ptr = np.arange(self.numberOfElements + 1, dtype=np.int64) * m
            self.conn = np.ascontiguousarray(ec, dtype=np.int32).ravel()
        else:
            self.numberOfElements = len(ec)
            lens = np.fromiter((len(c) for c in ec), dtype=np.int64, count=len(ec))
            self.elem_ptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
            self.conn = np.fromiter((v for c in ec for v in c), dtype=np.int32, count=int(self.elem_ptr[-1]))
        lens = np.diff(self.elem_ptr)
        for m in np.unique(lens):
            _shapes.shape_code_for(int(m), self.dimension)                               # raises like Element.py:27
        self.region_of_element = np.zeros(self.numberOfElements, dtype=np.int32)
        self.regions = []
        for h, (idx, name) in enumerate(zip(gridData.regionsElementsIndexes, gridData.regionsNames)):
            idx = np.asarray(idx, dtype=np.int64)
            self.region_of_element[idx] = h
            self.regions.append(Region(self, idx, name, h))
        bcs = gridData.boundariesConnectivities
        if isinstance(bcs, np.ndarray) and bcs.ndim == 2:
            nF, mf = bcs.shape
            self.facet_ptr = np.arange(nF + 1, dtype=np.int64) * mf
            self.facet_conn = np.ascontiguousarray(bcs, dtype=np.int32).ravel()
            order = None
        else:
            # facets are re-ordered boundary by boundary (the reference indexes them through boundariesIndexes)
            order = np.fromiter((i for idx in gridData.boundariesIndexes for i in idx), dtype=np.int64)
            flens = np.fromiter((len(bcs[i]) for i in order), dtype=np.int64, count=len(order))
            self.facet_ptr = np.concatenate([[0], np.cumsum(flens)]).astype(np.int64)
            self.facet_conn = np.fromiter((v for i in order for v in bcs[i]), dtype=np.int32, count=int(self.facet_ptr[-1]))
        counts = [len(idx) for idx in gridData.boundariesIndexes]
        self.boundary_ptr = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
        if order is None and len(counts):
            flat = np.concatenate([np.asarray(idx, dtype=np.int64) for idx in gridData.boundariesIndexes])
            if not np.array_equal(flat, np.arange(len(flat))):
                raise Exception("array-form boundariesConnectivities must be ordered boundary by boundary")
        self.boundaries = [Boundary(self, self.boundary_ptr[h], self.boundary_ptr[h + 1], name, h)
                           for h, name in enumerate(gridData.boundariesNames)]
        self.vertices = _Seq(self.numberOfVertices, lambda k: _VertexView(self, k))
        self.elements = _Seq(self.numberOfElements, lambda k: _ElementView(self, k))
        self._device = None

    # -- element groups by shape (grid order inside each group) ---------------------------------------------
    def shape_groups(self):
        lens = np.diff(self.elem_ptr)
        out = {}
        for m in np.unique(lens):
            code = _shapes.shape_code_for(int(m), self.dimension)
            out[code] = np.nonzero(lens == m)[0]
        return out

    @property
    def device(self):
        """The mesh in HBM (created on first use; needs a B200 -- there is no CPU path)."""
        if self._device is None:
            from .device import DeviceMesh
            self._device = DeviceMesh(self)
        return self._device


def read(filePath):
    """PyEFVLib/__init__.py:33-39"""
    extension = filePath.split(".")[-1]
    if extension != "msh":
        raise Exception("File extension not supported yet! Input your extension sugestion at https://github.com/GustavoExel/PyEFVLib")
    return Grid(MSHReader(filePath).getData())


# ---------------------------------------------------------------------------------------------------------
# synthetic structured meshes of SURVEY.md section 8d
# ---------------------------------------------------------------------------------------------------------
_TET_SPLIT = [[(0, 0, 0), (1, 0, 0), (0, 1, 0), (0, 0, 1)], [(1, 0, 0), (1, 0, 1), (0, 1, 1), (0, 0, 1)],
              [(1, 0, 0), (0, 1, 1), (0, 1, 0), (0, 0, 1)], [(1, 0, 0), (1, 1, 0), (0, 1, 0), (0, 1, 1)],
              [(1, 0, 0), (1, 0, 1), (1, 1, 0), (0, 1, 1)], [(1, 1, 0), (0, 1, 1), (1, 0, 1), (1, 1, 1)]]
_HEX_CORNERS = [(0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 1, 0), (0, 0, 1), (1, 0, 1), (1, 1, 1), (0, 1, 1)]


def structured_grid_data(n, kind="hex", perturb=0.0):
    """Unit cube, n^3 vertices with id = i + n j + n^2 k (x fastest, like meshes/msh/3D/Hexas*.msh), hexahedra in
    the bundled local order or the 6-tetrahedra-per-cube split of meshes/msh/3D/Tetras.msh, one region "Body",
    boundaries West East South North Bottom Top.  `perturb` moves interior vertices by a smooth deterministic
    perturb*h*sin(...) field to exercise non-affine hexahedra."""
    ax = np.linspace(0.0, 1.0, n)
    idx = np.arange(n ** 3, dtype=np.int64)
    I, J, K = idx % n, (idx // n) % n, idx // (n * n)
    coords = np.stack([ax[I], ax[J], ax[K]], axis=1)
    if perturb:
        h = 1.0 / (n - 1)
        interior = ((I > 0) & (I < n - 1) & (J > 0) & (J < n - 1) & (K > 0) & (K < n - 1))[:, None]
        x, y, z = coords[:, 0], coords[:, 1], coords[:, 2]
        d = np.stack([np.sin(7 * x + 3 * y + 5 * z), np.sin(4 * x - 6 * y + 2 * z + 1.0), np.sin(5 * x + 2 * y - 7 * z + 2.0)], axis=1)
        coords = coords + perturb * h * d * interior
    ne = n - 1
    e = np.arange(ne ** 3, dtype=np.int64)
    i, j, k = e % ne, (e // ne) % ne, e // (ne * ne)
    vid = lambda a, b, c: a + n * b + n * n * c
    corner = lambda bits: vid(i + bits[0], j + bits[1], k + bits[2])
    if kind == "hex":
        conn = np.stack([corner(b) for b in _HEX_CORNERS], axis=1).astype(np.int32)
    elif kind == "tet":
        conn = np.stack([np.stack([corner(b) for b in t], axis=1) for t in _TET_SPLIT], axis=1).reshape(-1, 4).astype(np.int32)
    else:
        raise Exception("kind must be 'hex' or 'tet'")
    q = np.arange(ne * ne, dtype=np.int64)
    a, b = q // ne, q % ne
    o = np.zeros_like(a)
    L = o + ne
    quads = {
        "West": [vid(o, b, a), vid(o, b, a + 1), vid(o, b + 1, a + 1), vid(o, b + 1, a)],
        "East": [vid(L, b, a), vid(L, b + 1, a), vid(L, b + 1, a + 1), vid(L, b, a + 1)],
        "South": [vid(b, o, a), vid(b + 1, o, a), vid(b + 1, o, a + 1), vid(b, o, a + 1)],
        "North": [vid(b, L, a), vid(b, L, a + 1), vid(b + 1, L, a + 1), vid(b + 1, L, a)],
        "Bottom": [vid(b, a, o), vid(b, a + 1, o), vid(b + 1, a + 1, o), vid(b + 1, a, o)],
        "Top": [vid(b, a, L), vid(b + 1, a, L), vid(b + 1, a + 1, L), vid(b, a + 1, L)],
    }
    names = list(quads.keys())
    if kind == "hex":
        bconn = np.concatenate([np.stack(quads[nm], axis=1) for nm in names]).astype(np.int32)
        per = ne * ne
    else:
        # the split uses the same diagonal on every cube face of one orientation: pick the diagonal that is an
        # edge of the tetrahedra by testing it on cube 0
        tetfaces = set()
        for t in conn[:6].tolist():
            for tri in ((t[0], t[2], t[1]), (t[0], t[3], t[2]), (t[0], t[1], t[3]), (t[1], t[2], t[3])):
                tetfaces.add(tuple(sorted(tri)))
        # faces of cube 0 at x=0,y=0,z=0 come from cube 0 itself; opposite faces use the same split by symmetry
        # of the repeated pattern (verified against the oracle's exhaustive matching in tests)
        parts = []
        all_faces = None
        for nm in names:
            v = np.stack(quads[nm], axis=1)
            parts.append(v)
        # decide per boundary with the first quad of each boundary against the full face set of its cube
        cubefaces = {}
        def faces_of_cube(c):
            if c not in cubefaces:
                s = set()
                for t in conn[6 * c:6 * c + 6].tolist():
                    for tri in ((t[0], t[2], t[1]), (t[0], t[3], t[2]), (t[0], t[1], t[3]), (t[1], t[2], t[3])):
                        s.add(tuple(sorted(tri)))
                cubefaces[c] = s
            return cubefaces[c]
        cube_of_first = {"West": 0, "East": ne - 1, "South": 0, "North": ne * (ne - 1), "Bottom": 0, "Top": ne * ne * (ne - 1)}
        tri_parts = []
        for nm, v in zip(names, parts):
            s = faces_of_cube(cube_of_first[nm])
            v0 = v[0].tolist()
            if tuple(sorted((v0[0], v0[1], v0[2]))) in s:
                t1, t2 = v[:, [0, 1, 2]], v[:, [0, 2, 3]]
            else:
                t1, t2 = v[:, [0, 1, 3]], v[:, [1, 2, 3]]
            tri_parts.append(np.stack([t1, t2], axis=1).reshape(-1, 3))
        bconn = np.concatenate(tri_parts).astype(np.int32)
        per = 2 * ne * ne
    gd = GridData(f"structured_{kind}_{n}")
    gd.setDimension(3)
    gd.setVertices(coords)
    gd.setElementConnectivity(conn)
    gd.setRegions(["Body"], [np.arange(len(conn), dtype=np.int64)])
    gd.setBoundaries(names, [np.arange(h * per, (h + 1) * per, dtype=np.int64) for h in range(6)], bconn)
    gd.setShapes([[]] * 7)
    return gd


def structured_grid(n, kind="hex", perturb=0.0):
    return Grid(structured_grid_data(n, kind, perturb))
