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
        self.lattice = getattr(gridData, "lattice", None)
        self.dimension = int(gridData.dimension)
        self.coords = np.ascontiguousarray(np.asarray(gridData.verticesCoordinates, dtype=np.float64).reshape(-1, 3))
        self.numberOfVertices = len(self.coords)
        ec = gridData.elementsConnectivities
        self._elem_ptr = None
        if isinstance(ec, np.ndarray) and ec.ndim == 2:
            self.numberOfElements, m = ec.shape
            self.uniform_m = int(m)                      # elem_ptr is implicit (built lazily if somebody asks)
            self.conn = np.ascontiguousarray(ec, dtype=np.int32).ravel()
            _shapes.shape_code_for(self.uniform_m, self.dimension)                       # raises like Element.py:27
        else:
            self.numberOfElements = len(ec)
            lens = np.fromiter((len(c) for c in ec), dtype=np.int64, count=len(ec))
            self._elem_ptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
            self.conn = np.fromiter((v for c in ec for v in c), dtype=np.int32, count=int(self._elem_ptr[-1]))
            for m in np.unique(lens):
                _shapes.shape_code_for(int(m), self.dimension)
            self.uniform_m = int(lens[0]) if len(np.unique(lens)) == 1 else 0
        self.region_of_element = np.zeros(self.numberOfElements, dtype=np.int32)
        self.regions = []
        for h, (idx, name) in enumerate(zip(gridData.regionsElementsIndexes, gridData.regionsNames)):
            idx = np.asarray(idx, dtype=np.int64)
            if h > 0:
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

    @property
    def elem_ptr(self):
        if self._elem_ptr is None:
            self._elem_ptr = np.arange(self.numberOfElements + 1, dtype=np.int64) * self.uniform_m
        return self._elem_ptr

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


def _quad_faces(n, a, b, nz=None):
    """Boundary quads of the box in the bundled orientation; a = slow, b = fast facet index (for the four
    lateral faces a is the z layer).  nz = number of vertex planes in z (default n: the unit cube)."""
    ne = n - 1
    vid = lambda i, j, k: i + n * j + n * n * k
    o = np.zeros_like(a)
    L = o + ne
    Lz = o + ((nz or n) - 1)
    return {
        "West": [vid(o, b, a), vid(o, b, a + 1), vid(o, b + 1, a + 1), vid(o, b + 1, a)],
        "East": [vid(L, b, a), vid(L, b + 1, a), vid(L, b + 1, a + 1), vid(L, b, a + 1)],
        "South": [vid(b, o, a), vid(b + 1, o, a), vid(b + 1, o, a + 1), vid(b, o, a + 1)],
        "North": [vid(b, L, a), vid(b, L, a + 1), vid(b + 1, L, a + 1), vid(b + 1, L, a)],
        "Bottom": [vid(b, a, o), vid(b, a + 1, o), vid(b + 1, a + 1, o), vid(b + 1, a, o)],
        "Top": [vid(b, a, Lz), vid(b + 1, a, Lz), vid(b + 1, a + 1, Lz), vid(b, a + 1, Lz)],
    }


BOUNDARY_NAMES_3D = ["West", "East", "South", "North", "Bottom", "Top"]
_tet_diag_cache = {}


def _tet_face_uses_first_diagonal():
    """For each cube face: does the 6-tetrahedra split cut the boundary quad (v0 v1 v2 v3) along v0-v2?  The split is
    the same in every cube, so a 2x2x2-vertex cube decides it."""
    if not _tet_diag_cache:
        n = 2
        vid = lambda i, j, k: i + n * j + n * n * k
        faces = set()
        for t in _TET_SPLIT:
            ids = [vid(*c) for c in t]
            for tri in ((0, 2, 1), (0, 3, 2), (0, 1, 3), (1, 2, 3)):
                faces.add(tuple(sorted(ids[k] for k in tri)))
        z = np.zeros(1, dtype=np.int64)
        for name, q in _quad_faces(2, z, z).items():
            v = [int(x[0]) for x in q]
            _tet_diag_cache[name] = tuple(sorted((v[0], v[1], v[2]))) in faces
    return _tet_diag_cache


def structured_arrays(n, kind="hex", layer_lo=0, layer_hi=None, nz=None, conn_dtype=np.int64):
    """Global-id connectivity of the cube layers [layer_lo, layer_hi) (a layer = all cubes with the same z index) and
    the boundary facets attached to them: dict(conn, facets={name: array}).  Bottom / Top facets are included when
    layer 0 / the last layer is.  nz vertex planes in z (default n).  conn_dtype: np.int32 halves the time when the ids fit."""
    ne = n - 1
    nez = (nz or n) - 1
    layer_hi = nez if layer_hi is None else layer_hi
    # first vertex of every cube, cubes ordered x fastest (element id = i + ne j + ne^2 k): broadcasting instead of
    # div / mod passes over the element ids (3 s -> 0.5 s for the 10 M hexahedra of C2)
    kk = np.arange(layer_lo, layer_hi, dtype=np.int64)
    jj = np.arange(ne, dtype=np.int64)
    v0 = (kk[:, None, None] * (n * n) + jj[None, :, None] * n + jj[None, None, :]).reshape(-1).astype(conn_dtype, copy=False)
    off = lambda bits: bits[0] + n * bits[1] + n * n * bits[2]
    if kind == "hex":
        conn = v0[:, None] + np.array([off(b) for b in _HEX_CORNERS], dtype=conn_dtype)[None, :]
    elif kind == "tet":
        conn = (v0[:, None, None] + np.array([[off(b) for b in t] for t in _TET_SPLIT], dtype=conn_dtype)[None, :, :]).reshape(-1, 4)
    else:
        raise Exception("kind must be 'hex' or 'tet'")
    facets = {}
    ql = np.arange(layer_lo * ne, layer_hi * ne, dtype=np.int64)         # lateral faces: a = z layer
    qf = np.arange(ne * ne, dtype=np.int64)                               # bottom / top: full face
    lat = _quad_faces(n, ql // ne, ql % ne, nz)
    full = _quad_faces(n, qf // ne, qf % ne, nz)
    for name in BOUNDARY_NAMES_3D:
        if name in ("Bottom", "Top"):
            present = (layer_lo == 0) if name == "Bottom" else (layer_hi == nez)
            v = np.stack(full[name], axis=1) if present else np.zeros((0, 4), dtype=np.int64)
        else:
            v = np.stack(lat[name], axis=1)
        if kind == "tet":
            if _tet_face_uses_first_diagonal()[name]:
                t1, t2 = v[:, [0, 1, 2]], v[:, [0, 2, 3]]
            else:
                t1, t2 = v[:, [0, 1, 3]], v[:, [1, 2, 3]]
            v = np.stack([t1, t2], axis=1).reshape(-1, 3)
        facets[name] = v
    return dict(conn=conn, facets=facets)


def structured_coords(n, ids, perturb=0.0, nz=None, planes=None):
    """Coordinates of the global vertex ids `ids` (None: all of them, in order) of the n x n x nz lattice with spacing
    1/(n-1) (unit cube if nz = n).  planes = [(k_lo, k_hi), ...]: `ids` are these whole z planes, in this order (broadcast
    instead of div / mod passes over the ids)."""
    ax = np.linspace(0.0, 1.0, n)
    nz = nz or n
    az = ax if nz == n else np.arange(nz) * (1.0 / (n - 1))
    if ids is None and planes is None:
        planes = [(0, nz)]
    if planes is not None and not perturb:
        parts = []
        for lo, hi in planes:
            c = np.empty((hi - lo, n, n, 3))
            c[..., 0] = ax[None, None, :]
            c[..., 1] = ax[None, :, None]
            c[..., 2] = az[lo:hi, None, None]
            parts.append(c.reshape(-1, 3))
        return parts[0] if len(parts) == 1 else np.concatenate(parts)
    if ids is None:
        ids = np.concatenate([np.arange(lo * n * n, hi * n * n, dtype=np.int64) for lo, hi in planes])
    I, J, K = ids % n, (ids // n) % n, ids // (n * n)
    coords = np.stack([ax[I], ax[J], az[K]], axis=1)
    if perturb:
        h = 1.0 / (n - 1)
        interior = ((I > 0) & (I < n - 1) & (J > 0) & (J < n - 1) & (K > 0) & (K < nz - 1))[:, None]
        x, y, z = coords[:, 0], coords[:, 1], coords[:, 2]
        d = np.stack([np.sin(7 * x + 3 * y + 5 * z), np.sin(4 * x - 6 * y + 2 * z + 1.0), np.sin(5 * x + 2 * y - 7 * z + 2.0)], axis=1)
        coords = coords + perturb * h * d * interior
    return coords


def structured_grid_data(n, kind="hex", perturb=0.0, nz=None):
    """Unit cube (or, with nz, an n x n x nz box of the same spacing), n^3 vertices with id = i + n j + n^2 k (x fastest, like meshes/msh/3D/Hexas*.msh), hexahedra in
    the bundled local order or the 6-tetrahedra-per-cube split of meshes/msh/3D/Tetras.msh, one region "Body",
    boundaries West East South North Bottom Top.  `perturb` moves interior vertices by a smooth deterministic
    perturb*h*sin(...) field to exercise non-affine hexahedra."""
    arr = structured_arrays(n, kind, nz=nz, conn_dtype=np.int32 if n * n * (nz or n) < 2 ** 31 else np.int64)
    coords = structured_coords(n, None, perturb, nz)
    conn = arr["conn"].astype(np.int32, copy=False)
    sizes = [len(arr["facets"][nm]) for nm in BOUNDARY_NAMES_3D]
    starts = np.concatenate([[0], np.cumsum(sizes)])
    bconn = np.concatenate([arr["facets"][nm] for nm in BOUNDARY_NAMES_3D]).astype(np.int32)
    gd = GridData(f"structured_{kind}_{n}")
    gd.setDimension(3)
    gd.setVertices(coords)
    gd.setElementConnectivity(conn)
    gd.setRegions(["Body"], [np.arange(len(conn), dtype=np.int64)])
    gd.setBoundaries(list(BOUNDARY_NAMES_3D), [np.arange(starts[h], starts[h + 1], dtype=np.int64) for h in range(6)], bconn)
    gd.setShapes([[]] * 7)
    gd.lattice = (n, n, nz or n)          # vertex id = i + n j + n^2 k: lets the multigrid preconditioner aggregate arithmetically
    return gd


def structured_grid(n, kind="hex", perturb=0.0, nz=None):
    return Grid(structured_grid_data(n, kind, perturb, nz))
