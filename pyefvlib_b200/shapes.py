"""Reference-element tables of the EbFVM shapes, derived from first principles.

The reference keeps these as literal class attributes (PyEFVLib/Shape.py:7-19 Triangle, :35-47
Quadrilateral, :63-74 Tetrahedron, :98-109 Hexahedron).  Here they are *computed* with exact rational
arithmetic from the definition of the element-based finite-volume sub-control volumes:

* an inner face (sub-control surface) belongs to an element edge (b, f); in 3-D it is the bilinear patch
  spanned by  edge midpoint -> centroid of one adjacent facet -> element centroid -> centroid of the other
  adjacent facet; in 2-D it is the segment edge midpoint -> element centroid.  Its integration point is the
  mean of those corners in parametric space (e.g. hexahedron face 0 sits at xi = (1/2, 1/4, 1/4)).
* a sub-element (the part of the control volume of vertex k inside the element) is the quad / hexahedron
  spanned by the vertex, the adjacent edge midpoints, (3-D) the adjacent facet centroids and the element
  centroid; its integration point is the mean of those corners.
* an outer face is the part of a boundary facet that belongs to one facet vertex; its integration point is
  the sub-element point of the facet seen as a lower-dimensional element.

`tests/test_shapes.py` checks every array against the constants dumped from the reference
(`tests/golden/shape_tables.npz`) bit for bit, and `csrc/shape_tables.inc` (the CUDA copy) is generated
from this module by `python -m pyefvlib_b200.shapes`.
"""
from fractions import Fraction as Fr
from itertools import combinations

import numpy as np

# shape codes shared with the C-ABI (include/efv_b200.h: EFV_SHAPE_*)
TRIANGLE, QUADRILATERAL, TETRAHEDRON, HEXAHEDRON, PRISM, PYRAMID = 0, 1, 2, 3, 4, 5
SHAPE_NAMES = {TRIANGLE: "Triangle", QUADRILATERAL: "Quadrilateral", TETRAHEDRON: "Tetrahedron",
               HEXAHEDRON: "Hexahedron", PRISM: "Prism", PYRAMID: "Pyramid"}


def _N_simplex(xi):
    return [1 - sum(xi)] + list(xi)


def _dN_simplex(xi):
    d = len(xi)
    return [[Fr(-1)] * d] + [[Fr(int(i == j)) for j in range(d)] for i in range(d)]


def _N_tensor(verts):
    def N(xi):
        out = []
        for v in verts:
            p = Fr(1)
            for a, x in zip(v, xi):
                p *= x if a == 1 else (1 - x)
            out.append(p)
        return out
    return N


def _dN_tensor(verts):
    def dN(xi):
        out = []
        for v in verts:
            row = []
            for k in range(len(xi)):
                p = Fr(1 if v[k] == 1 else -1)
                for kk, (a, x) in enumerate(zip(v, xi)):
                    if kk != k:
                        p *= x if a == 1 else (1 - x)
                row.append(p)
            out.append(row)
        return out
    return dN


def _N_prism(xi):
    """Triangle (xi, eta) x line (zeta): vertices 0-2 at zeta = 0, 3-5 at zeta = 1 (Shape.py:141-153)."""
    x, y, z = xi
    tri = [1 - x - y, x, y]
    return [t * (1 - z) for t in tri] + [t * z for t in tri]


def _dN_prism(xi):
    x, y, z = xi
    tri, dtri = [1 - x - y, x, y], [(Fr(-1), Fr(-1)), (Fr(1), Fr(0)), (Fr(0), Fr(1))]
    return [[d[0] * (1 - z), d[1] * (1 - z), -t] for t, d in zip(tri, dtri)] + \
           [[d[0] * z, d[1] * z, t] for t, d in zip(tri, dtri)]


_PYR_SIGNS = [(-1, -1), (1, -1), (1, 1), (-1, 1)]


def _N_pyramid(xi):
    """Rational pyramid functions on the base [0,1]^2 with the apex above its centre: with X = 2 xi - 1, Y = 2 eta - 1,
    N_i = 1/4 [(1 + a X)(1 + b Y) - zeta + a b X Y zeta / (1 - zeta)], N_4 = zeta.  Reproduces Shape.py:187-199."""
    x, y, z = xi
    X, Y = 2 * x - 1, 2 * y - 1
    return [Fr(1, 4) * ((1 + a * X) * (1 + b * Y) - z + a * b * X * Y * z / (1 - z)) for a, b in _PYR_SIGNS] + [z]


def _dN_pyramid(xi):
    x, y, z = xi
    X, Y = 2 * x - 1, 2 * y - 1
    rows = []
    for a, b in _PYR_SIGNS:
        rows.append([Fr(1, 2) * (a * (1 + b * Y) + a * b * Y * z / (1 - z)),
                     Fr(1, 2) * (b * (1 + a * X) + a * b * X * z / (1 - z)),
                     Fr(1, 4) * (-1 + a * b * X * Y / ((1 - z) * (1 - z)))])
    return rows + [[Fr(0), Fr(0), Fr(1)]]


def _mean(points):
    n = len(points)
    return [sum(p[k] for p in points) / n for k in range(len(points[0]))]


class ShapeTable:
    """Exact reference-element data of one shape.  Arrays mirror the reference attribute names."""

    def __init__(self, code, dim, verts, edges, facets, N, dN, sub_volume, centroid=None, facet_a_first=None,
                 triangular_faces=(), sub_point_override=None):
        """centroid: parametric point used as "element centroid" (default: vertex mean; the reference's pyramid uses
        the centre of its base, Shape.py:213-216).  facet_a_first(b, f, facets): which of the two facets adjacent to
        edge (b, f) the reference lists first (default: the one that makes the area vector point from b to f).
        triangular_faces: edges whose inner face is the triangle centroid - facet centroid - edge midpoint
        (pyramid base edges, Shape.py:222-231).  Ragged tables are padded: facet vertices with -1, neighbour
        vertices and outer-face values with 0."""
        self.code = code
        self.name = SHAPE_NAMES[code]
        self.dimension = dim
        self.numberOfVertices = m = len(verts)
        self.numberOfInnerFaces = len(edges)
        self.numberOfFacets = len(facets)
        self.vertices = [[Fr(c) for c in v] for v in verts]
        mf_max = max(len(fc) for fc in facets)
        self.facetSizes = np.array([len(fc) for fc in facets], dtype=np.int32)
        self.facetVerticesIndexes = np.array([list(fc) + [-1] * (mf_max - len(fc)) for fc in facets], dtype=np.int32)
        centroid = [Fr(c) for c in centroid] if centroid is not None else _mean(self.vertices)

        def midpoint(ids):
            return _mean([self.vertices[i] for i in ids])

        # ---- inner faces -----------------------------------------------------------------------
        nbr, ipoints, adjn = [], [], []
        for b, f in edges:
            if dim == 2:
                nbr.append([b, f])
                adjn.append([0, 0])
                ipoints.append(_mean([midpoint([b, f]), centroid]))
                continue
            adj = [list(fc) for fc in facets if b in fc and f in fc]
            assert len(adj) == 2
            if (b, f) in triangular_faces:
                # triangle  centroid -> centroid of the facet that is not the base -> edge midpoint
                fa = [fc for fc in adj if len(fc) == 3][0]
                nbr.append([b, f] + [v for v in fa if v not in (b, f)])
                adjn.append([1, 0])
                ipoints.append(_mean([midpoint([b, f]), midpoint(fa), centroid]))
                continue
            # order the two adjacent facets so that the area vector
            #   s = 1/2 [(p1-p0)x(p3-p0) + (p3-p2)x(p1-p2)]   (Shape.py:84-96, :119-139)
            # points from b to f on the reference element
            p0, p2 = centroid, midpoint([b, f])

            def area(fa, fb):
                p1, p3 = midpoint(fa), midpoint(fb)
                u = [p1[k] - p0[k] for k in range(3)]
                v = [p3[k] - p0[k] for k in range(3)]
                w = [p3[k] - p2[k] for k in range(3)]
                z = [p1[k] - p2[k] for k in range(3)]
                cr = lambda a, c: [a[1] * c[2] - a[2] * c[1], a[2] * c[0] - a[0] * c[2], a[0] * c[1] - a[1] * c[0]]
                c1, c2 = cr(u, v), cr(w, z)
                return [(c1[k] + c2[k]) / 2 for k in range(3)]
            e = [self.vertices[f][k] - self.vertices[b][k] for k in range(3)]
            s = area(adj[0], adj[1])
            if sum(s[k] * e[k] for k in range(3)) < 0:
                adj = [adj[1], adj[0]]
            self_oriented = True
            if facet_a_first is not None and list(facet_a_first(b, f, adj)) != adj[0]:
                adj = [adj[1], adj[0]]
                self_oriented = False
            others = [[v for v in fc if v not in (b, f)] for fc in adj]
            nbr.append([b, f] + others[0] + others[1])
            adjn.append([len(others[0]), len(others[1])])
            ipoints.append(_mean([midpoint([b, f]), midpoint(adj[0]), midpoint(adj[1]), centroid]))
            del self_oriented
        width = max(len(r) for r in nbr)
        nbr = [r + [0] * (width - len(r)) for r in nbr]
        self.innerFaceAdjacentCounts = np.array(adjn, dtype=np.int32)
        self.innerFaceNeighborVertices = np.array(nbr, dtype=np.int32)
        self.innerFacePoints = ipoints
        self.innerFaceShapeFunctionValues = np.array([[float(x) for x in N(p)] for p in ipoints])
        self.innerFaceShapeFunctionDerivatives = np.array([[[float(x) for x in r] for r in dN(p)] for p in ipoints])
        if code == TETRAHEDRON:
            # REFERENCE QUIRK, reproduced for parity: Shape.py:68 lists the shape-function VALUES of
            # tetrahedron faces 4 and 5 in the order (edge 2-3, edge 1-3) while the neighbour table
            # (Shape.py:70) has face 4 = edge 1-3 and face 5 = edge 2-3.  Derivatives are constant, so only
            # users of the values (geomechanics.py:70,107) see it.
            self.innerFaceShapeFunctionValues[[4, 5]] = self.innerFaceShapeFunctionValues[[5, 4]]

        # ---- sub-elements ----------------------------------------------------------------------
        spoints = []
        for k in range(m):
            if sub_point_override and k in sub_point_override:
                spoints.append([Fr(c) for c in sub_point_override[k]])
                continue
            corners = [self.vertices[k], centroid]
            corners += [midpoint(e) for e in edges if k in e]
            if dim == 3:
                corners += [midpoint(fc) for fc in facets if k in fc]
            spoints.append(_mean(corners))
        self.subelementPoints = spoints
        self.subelementShapeFunctionValues = np.array([[float(x) for x in N(p)] for p in spoints])
        self.subelementShapeFunctionDerivatives = np.array([[[float(x) for x in r] for r in dN(p)] for p in spoints])
        sub_volume = list(sub_volume) if isinstance(sub_volume, (list, tuple)) else [sub_volume] * m
        self.subelementTransformedVolumes = np.array([float(v) for v in sub_volume])

        # ---- outer faces (per facet, per facet vertex) -----------------------------------------
        ofv = []
        for fc in facets:
            rows = []
            fc = list(fc)
            for k in fc:
                if len(fc) == 2:       # line facet: 3/4 of the vertex, 1/4 of the other end
                    pt = _mean([self.vertices[k], midpoint(fc)])
                else:                  # sub-element point of the facet as a 2-D element
                    n = len(fc)
                    i = fc.index(k)
                    ring = [midpoint([fc[i], fc[(i + 1) % n]]), midpoint([fc[i], fc[(i - 1) % n]])]
                    pt = _mean([self.vertices[k], midpoint(fc)] + ring)
                rows.append([float(x) for x in N(pt)])
            rows += [[0.0] * m] * (mf_max - len(fc))
            ofv.append(rows)
        self.outerFaceShapeFunctionValues = np.array(ofv)


def _build():
    tri_v = [(0, 0), (1, 0), (0, 1)]
    quad_v = [(0, 0), (1, 0), (1, 1), (0, 1)]
    tet_v = [(0, 0, 0), (1, 0, 0), (0, 1, 0), (0, 0, 1)]
    hex_v = [(0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 1, 0), (0, 0, 1), (1, 0, 1), (1, 1, 1), (0, 1, 1)]
    prism_v = [(0, 0, 0), (1, 0, 0), (0, 1, 0), (0, 0, 1), (1, 0, 1), (0, 1, 1)]
    pyr_v = [(0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 1, 0), (Fr(1, 2), Fr(1, 2), 1)]
    return {
        # edge (b, f) lists and facet vertex lists are topology; they follow Shape.py:13,17 / :41,45 /
        # :69,73 / :104,108 so that per-face arrays line up with the reference's face numbering
        TRIANGLE: ShapeTable(TRIANGLE, 2, tri_v, [(0, 1), (1, 2), (2, 0)], [(1, 0), (2, 1), (0, 2)],
                             _N_simplex, _dN_simplex, Fr(1, 6)),
        QUADRILATERAL: ShapeTable(QUADRILATERAL, 2, quad_v, [(0, 1), (1, 2), (2, 3), (3, 0)],
                                  [(1, 0), (2, 1), (3, 2), (0, 3)], _N_tensor(quad_v), _dN_tensor(quad_v), Fr(1, 4)),
        TETRAHEDRON: ShapeTable(TETRAHEDRON, 3, tet_v, [(0, 1), (1, 2), (2, 0), (0, 3), (1, 3), (2, 3)],
                                [(0, 2, 1), (0, 3, 2), (0, 1, 3), (1, 2, 3)], _N_simplex, _dN_simplex, Fr(1, 24)),
        HEXAHEDRON: ShapeTable(HEXAHEDRON, 3, hex_v,
                               [(0, 1), (1, 2), (2, 3), (3, 0), (4, 5), (5, 6), (6, 7), (7, 4),
                                (4, 0), (5, 1), (6, 2), (7, 3)],
                               [(0, 3, 2, 1), (0, 4, 7, 3), (0, 1, 5, 4), (4, 5, 6, 7), (1, 2, 6, 5), (2, 3, 7, 6)],
                               _N_tensor(hex_v), _dN_tensor(hex_v), Fr(1, 8)),
        # Shape.py:141-153: the quadrilateral facet of an in-triangle edge is listed before the triangle
        PRISM: ShapeTable(PRISM, 3, prism_v, [(0, 1), (1, 2), (2, 0), (4, 3), (5, 4), (3, 5), (3, 0), (4, 1), (5, 2)],
                          [(0, 2, 1), (3, 4, 5), (0, 3, 5, 2), (0, 1, 4, 3), (1, 2, 5, 4)], _N_prism, _dN_prism, Fr(1, 12)),
        # Shape.py:187-199: "centroid" = centre of the base; base edges carry triangular faces; the tabulated apex
        # sub-element point is (1/2, 1/2, 13/24)
        PYRAMID: ShapeTable(PYRAMID, 3, pyr_v, [(0, 1), (1, 2), (2, 3), (3, 0), (0, 4), (1, 4), (2, 4), (3, 4)],
                            [(0, 4, 3), (0, 1, 4), (1, 2, 4), (2, 3, 4), (0, 3, 2, 1)], _N_pyramid, _dN_pyramid,
                            [Fr(1, 18)] * 4 + [Fr(1, 9)], centroid=(Fr(1, 2), Fr(1, 2), 0),
                            triangular_faces=((0, 1), (1, 2), (2, 3), (3, 0)),
                            sub_point_override={4: (Fr(1, 2), Fr(1, 2), Fr(13, 24))}),
    }


SHAPES = _build()
BY_NAME = {t.name: t for t in SHAPES.values()}


def shape_code_for(n_vertices, dimension):
    """Shape of an element from its vertex count (Element.py:22-27 + Shape.py `_is`).  A 4-vertex element is
    a quadrilateral in a 2-D grid and a tetrahedron in a 3-D grid (the reference decides by coplanarity,
    Shape.py:4-5, which on valid meshes is the same thing)."""
    if n_vertices == 3:
        return TRIANGLE
    if n_vertices == 4:
        return QUADRILATERAL if dimension == 2 else TETRAHEDRON
    if n_vertices == 8:
        return HEXAHEDRON
    if n_vertices == 6:
        return PRISM
    if n_vertices == 5:
        return PYRAMID
    raise Exception("This element has not been registered in Grid yet")   # Element.py:27


# ---------------------------------------------------------------------------------------------------------
# CUDA table generator
# ---------------------------------------------------------------------------------------------------------
def _c_array(a, fmt):
    a = np.asarray(a)
    if a.ndim == 1:
        return "{" + ", ".join(fmt(x) for x in a) + "}"
    return "{" + ",\n ".join(_c_array(x, fmt) for x in a) + "}"


def vertex_faces(t):
    """Per local vertex the inner faces that touch it, encoded 2*face + (1 if the vertex is the forward one), padded with
    -1 (the apex of a pyramid touches 4 faces, its base vertices 3).
    The heat flux of face f enters row b with sign -1 and row f with sign +1 (heat_transfer.py:52-54)."""
    rows = []
    for l in range(t.numberOfVertices):
        enc = []
        for f, nb in enumerate(t.innerFaceNeighborVertices):
            if nb[0] == l:
                enc.append(2 * f)
            elif nb[1] == l:
                enc.append(2 * f + 1)
        rows.append(enc)
    width = max(len(r) for r in rows)
    return np.array([r + [-1] * (width - len(r)) for r in rows], dtype=np.int32)


def generate_cuda_tables():
    """Text of csrc/shape_tables.inc: __constant__ tables + one traits struct per shape."""
    fd = lambda x: repr(float(x))
    fi = lambda x: str(int(x))
    out = ["// GENERATED by `python -m pyefvlib_b200.shapes` -- do not edit.",
           "// Reference-element tables (see pyefvlib_b200/shapes.py for the derivation).", "#pragma once", ""]
    for code in (TRIANGLE, QUADRILATERAL, TETRAHEDRON, HEXAHEDRON, PRISM, PYRAMID):
        t = SHAPES[code]
        n = t.name
        m, nf, d = t.numberOfVertices, t.numberOfInnerFaces, t.dimension
        nfc, mf = t.facetVerticesIndexes.shape
        nn = t.innerFaceNeighborVertices.shape[1]
        vf = vertex_faces(t)
        nvf = vf.shape[1]
        out += [
            f"static __constant__ double c_{n}_ifD[{nf}][{m}][{d}] =\n {_c_array(t.innerFaceShapeFunctionDerivatives, fd)};",
            f"static __constant__ double c_{n}_ifN[{nf}][{m}] =\n {_c_array(t.innerFaceShapeFunctionValues, fd)};",
            f"static __constant__ int c_{n}_nbr[{nf}][{nn}] =\n {_c_array(t.innerFaceNeighborVertices, fi)};",
            f"static __constant__ double c_{n}_subD[{m}][{m}][{d}] =\n {_c_array(t.subelementShapeFunctionDerivatives, fd)};",
            f"static __constant__ int c_{n}_facetV[{nfc}][{mf}] =\n {_c_array(t.facetVerticesIndexes, fi)};",
            f"static __constant__ double c_{n}_ofN[{nfc}][{mf}][{m}] =\n {_c_array(t.outerFaceShapeFunctionValues, fd)};",
            f"static __constant__ int c_{n}_vface[{m}][{nvf}] =\n {_c_array(vf, fi)};",
            f"static __constant__ int c_{n}_facetN[{nfc}] = {_c_array(t.facetSizes, fi)};",
            f"struct {n}Tab {{",
            f"  static constexpr int CODE = {code}, DIM = {d}, M = {m}, NF = {nf}, NFACETS = {nfc}, MF = {mf}, NNBR = {nn}, NVF = {nvf};",
            f"  static constexpr double SUBVOL = {fd(t.subelementTransformedVolumes[0])};",
            f"  __host__ __device__ static constexpr double subvol(int l) {{ constexpr double T[{m}] = {_c_array(t.subelementTransformedVolumes, fd)}; return T[l]; }}",
            f"  // extra vertices (beyond b, f) of the two facets adjacent to the edge of inner face f; (1, 0) = triangular face",
            f"  __host__ __device__ static constexpr int adjn(int f, int k) {{ constexpr int T[{nf}][2] = {_c_array(t.innerFaceAdjacentCounts, fi).replace(chr(10), '')}; return T[f][k]; }}",
            f"  __device__ static __forceinline__ int facetN(int f) {{ return c_{n}_facetN[f]; }}",
            f"  __device__ static __forceinline__ double ifD(int f, int j, int a) {{ return c_{n}_ifD[f][j][a]; }}",
            f"  __device__ static __forceinline__ double ifN(int f, int j) {{ return c_{n}_ifN[f][j]; }}",
            f"  // compile-time copy for statically unrolled code (keeps vertex arrays in registers)",
            f"  __host__ __device__ static constexpr int nbr(int f, int k) {{ constexpr int T[{nf}][{nn}] = {_c_array(t.innerFaceNeighborVertices, fi).replace(chr(10), '')}; return T[f][k]; }}",
            f"  __device__ static __forceinline__ int nbr_dyn(int f, int k) {{ return c_{n}_nbr[f][k]; }}",
            f"  __device__ static __forceinline__ double subD(int k, int j, int a) {{ return c_{n}_subD[k][j][a]; }}",
            f"  __device__ static __forceinline__ int facetV(int f, int k) {{ return c_{n}_facetV[f][k]; }}",
            f"  __device__ static __forceinline__ double ofN(int f, int k, int j) {{ return c_{n}_ofN[f][k][j]; }}",
            f"  __device__ static __forceinline__ int vface(int l, int k) {{ return c_{n}_vface[l][k]; }}",
            f"  __host__ __device__ static constexpr int vface_c(int l, int k) {{ constexpr int T[{m}][{nvf}] = {_c_array(vf, fi).replace(chr(10), '')}; return T[l][k]; }}",
            f"  // literal copy of ifD: with static indices the entries become immediates of the DFMAs",
            f"  __host__ __device__ static constexpr double ifD_c(int f, int j, int a) {{ constexpr double T[{nf}][{m}][{d}] = {_c_array(t.innerFaceShapeFunctionDerivatives, fd).replace(chr(10), '')}; return T[f][j][a]; }}",
            "};",
            "",
        ]
    return "\n".join(out) + "\n"


if __name__ == "__main__":
    import os
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", "shape_tables.inc")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "w") as fh:
        fh.write(generate_cuda_tables())
    print("wrote", path)
