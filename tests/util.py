"""Helpers shared by the tests: load a golden fixture into the oracle's and the product's mesh containers."""
import os

import numpy as np
import scipy.sparse as sp

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def lists_from_fixture(z):
    cp, cf = z["conn_ptr"], z["conn_flat"]
    conn = [cf[cp[i]:cp[i + 1]].tolist() for i in range(len(cp) - 1)]
    bp, bf = z["bconn_ptr"], z["bconn_flat"]
    bconn = [bf[bp[i]:bp[i + 1]].tolist() for i in range(len(bp) - 1)]
    rp = z["region_ptr"]
    regions = [z["region_elems"][rp[i]:rp[i + 1]].tolist() for i in range(len(rp) - 1)]
    qp = z["boundary_ptr"]
    bidx = [z["boundary_facets"][qp[i]:qp[i + 1]].tolist() for i in range(len(qp) - 1)]
    return dict(dimension=int(z["dimension"]), coords=z["coords"], conn=conn, region_names=[str(s) for s in z["region_names"]],
                regions=regions, boundary_names=[str(s) for s in z["boundary_names"]], bidx=bidx, bconn=bconn)


def oracle_grid(z):
    from oracle import efv_oracle as O
    d = lists_from_fixture(z)
    gd = O.OracleGridData(d["dimension"], d["coords"].tolist(), d["conn"], d["region_names"], d["regions"],
                          d["boundary_names"], d["bidx"], d["bconn"])
    return O.OracleGrid(gd)


def product_grid(z):
    import pyefvlib_b200 as P
    d = lists_from_fixture(z)
    gd = P.GridData("fixture")
    gd.setDimension(d["dimension"])
    gd.setVertices(d["coords"].tolist())
    gd.setElementConnectivity(d["conn"])
    gd.setRegions(d["region_names"], d["regions"])
    gd.setBoundaries(d["boundary_names"], d["bidx"], d["bconn"])
    gd.setShapes([[]] * 7)
    return P.Grid(gd)


def csr_from(z, prefix):
    n = len(z[prefix + "indptr"]) - 1
    return sp.csr_matrix((z[prefix + "data"], z[prefix + "indices"], z[prefix + "indptr"]), shape=(n, n))


def rel_max(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)) if a.size else 0.0


def rel_l2(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300))


def row_rel_matrix_error(A, B):
    """max over rows of  max_j |A_ij - B_ij| / max_j |B_ij|  (meaningful when rows differ by 20 orders of
    magnitude, as in the Biot system)."""
    A, B = sp.csr_matrix(A), sp.csr_matrix(B)
    D = abs(A - B)
    scale = np.maximum(abs(B).max(axis=1).toarray().ravel(), 1e-300)
    return float((D.max(axis=1).toarray().ravel() / scale).max())


def flatten_by_element(grid_groups, nE, key):
    parts = [None] * nE
    for code, g in grid_groups.items():
        arr = g[key]
        for a, e in enumerate(g["ids"]):
            parts[e] = arr[a].ravel()
    return np.concatenate(parts)
