"""Python handles over the C-ABI objects (efv_mesh, efv_system).  Everything here runs on the GPU."""
import ctypes as C

import numpy as np
import scipy.sparse as sp

from . import _lib
from . import shapes as _shapes

def field_copy_counters():
    """dict(uploads, upload_bytes, downloads, download_bytes): whole-field copies since the process started."""
    c = [C.c_int64(0) for _ in range(4)]
    _lib.check(_lib.load().efv_field_copy_counters(*[C.byref(x) for x in c]))
    return dict(uploads=c[0].value, upload_bytes=c[1].value, downloads=c[2].value, download_bytes=c[3].value)


DIRICHLET_ROWS, DIRICHLET_SYMMETRIC = 0, 1
PRECOND_JACOBI, PRECOND_AMG = 0, 1
VEC_X, VEC_B, VEC_XOLD = 0, 1, 2


class DeviceMesh:
    """efv_mesh: vertices, elements, incidence, boundary facets and (on demand) SoA geometry in HBM."""

    def __init__(self, grid):
        L = _lib.load()
        self.grid = grid
        self._coords, pc = _lib.f64(grid.coords)
        self._conn, pn = _lib.i32(grid.conn)
        nreg = max(1, len(grid.regions))
        pr = None
        if nreg > 1:
            self._reg, pr = _lib.i32(grid.region_of_element)
        if grid.uniform_m:
            self.handle = _lib.check_ptr(L.efv_mesh_create_uniform(grid.dimension, grid.numberOfVertices, pc, grid.numberOfElements,
                                                                   grid.uniform_m, pn, pr, nreg))
        else:
            self._ep, pe = _lib.i64(grid.elem_ptr)
            self.handle = _lib.check_ptr(L.efv_mesh_create(grid.dimension, grid.numberOfVertices, pc, grid.numberOfElements,
                                                           pe, pn, pr, nreg))
        _, pb = _lib.i64(grid.boundary_ptr)
        _, pf = _lib.i64(grid.facet_ptr)
        _, pfc = _lib.i32(grid.facet_conn)
        _lib.check(L.efv_mesh_set_boundaries(self.handle, len(grid.boundaries), pb, len(grid.facet_ptr) - 1, pf, pfc))
        self._geometry = False
        self._vv = None

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                _lib.load().efv_mesh_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    def build_geometry(self):
        if not self._geometry:
            _lib.check(_lib.load().efv_geometry_build(self.handle))
            self._geometry = True

    def release_geometry(self):
        _lib.check(_lib.load().efv_geometry_release(self.handle))
        self._geometry = False
        self._vv = None

    def vertex_volume(self):
        if self._vv is None:
            self.build_geometry()
            out = np.empty(self.grid.numberOfVertices)
            _lib.check(_lib.load().efv_vertex_volume_export(self.handle, out.ctypes.data_as(_lib.c_f64p)))
            self._vv = out
        return self._vv

    def export_geometry(self):
        """Per shape code: dict(ids, G[n,nf,d,m], Jinv[n,nf,d,d], S[n,nf,3], vol[n,m]) -- the arrays the reference
        holds as InnerFace.globalDerivatives / .area and Element.subelementVolumes."""
        self.build_geometry()
        L = _lib.load()
        out = {}
        for code, ids in self.grid.shape_groups().items():
            t = _shapes.SHAPES[code]
            n, nf, d, m = len(ids), t.numberOfInnerFaces, t.dimension, t.numberOfVertices
            G = np.empty((n, nf, d, m)); J = np.empty((n, nf, d, d)); S = np.empty((n, nf, 3)); V = np.empty((n, m))
            _lib.check(L.efv_geometry_export_group(self.handle, code, G.ctypes.data_as(_lib.c_f64p), J.ctypes.data_as(_lib.c_f64p),
                                                   S.ctypes.data_as(_lib.c_f64p), V.ctypes.data_as(_lib.c_f64p)))
            out[code] = dict(ids=ids, G=G, Jinv=J, S=S, vol=V)
        return out

    def export_boundaries(self):
        L = _lib.load()
        nF = L.efv_mesh_num_facets(self.handle)
        nO = L.efv_mesh_num_outer_faces(self.handle)
        fe = np.empty(nF, dtype=np.int64); fl = np.empty(nF, dtype=np.int32); fa = np.empty((nF, 3))
        ov = np.empty(nO, dtype=np.int32); on = np.empty(nO)
        _lib.check(L.efv_boundary_export(self.handle, fe.ctypes.data_as(_lib.c_i64p), fl.ctypes.data_as(_lib.c_i32p),
                                         fa.ctypes.data_as(_lib.c_f64p), ov.ctypes.data_as(_lib.c_i32p), on.ctypes.data_as(_lib.c_f64p)))
        return dict(facet_element=fe, facet_local=fl, facet_area=fa, outer_vertex=ov, outer_area_norm=on)


class DeviceSystem:
    """efv_system: CSR graph + slot map + values + vectors + Krylov workspace for `ndof` unknowns per vertex.
    Row numbering at this API is the reference's field-major one (row = vertex + field * nV)."""

    def __init__(self, grid, ndof=1):
        L = _lib.load()
        self.grid, self.ndof = grid, int(ndof)
        self.mesh = grid.device
        self.handle = _lib.check_ptr(L.efv_system_create(self.mesh.handle, self.ndof))
        self.rows = L.efv_system_rows(self.handle)
        self.nnz = L.efv_system_nnz(self.handle)
        self.graph_nnz = L.efv_system_graph_nnz(self.handle)
        lat = getattr(grid, "lattice", None)
        if lat is not None and int(np.prod(lat)) == grid.numberOfVertices:       # structured numbering: 2x2x2 aggregates on the device
            self.set_lattice(*lat)

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                _lib.load().efv_system_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    # -- structure ---------------------------------------------------------------------------------------------
    def csr_structure(self):
        indptr = np.empty(self.rows + 1, dtype=np.int64)
        indices = np.empty(self.nnz, dtype=np.int32)
        _lib.check(_lib.load().efv_system_export_csr(self.handle, indptr.ctypes.data_as(_lib.c_i64p), indices.ctypes.data_as(_lib.c_i32p)))
        return indptr, indices

    def graph(self):
        gptr = np.empty(self.grid.numberOfVertices + 1, dtype=np.int64)
        gcol = np.empty(self.graph_nnz, dtype=np.int32)
        _lib.check(_lib.load().efv_system_export_graph(self.handle, gptr.ctypes.data_as(_lib.c_i64p), gcol.ctypes.data_as(_lib.c_i32p)))
        return gptr, gcol

    def slot_map(self):
        n = _lib.load().efv_system_slotmap_size(self.handle)
        out = np.empty(n, dtype=np.int64)
        _lib.check(_lib.load().efv_system_export_slotmap(self.handle, out.ctypes.data_as(_lib.c_i64p)))
        return out

    def values(self):
        out = np.empty(self.nnz)
        _lib.check(_lib.load().efv_system_export_values(self.handle, out.ctypes.data_as(_lib.c_f64p)))
        return out

    def to_scipy(self):
        """The assembled matrix as scipy CSR (explicit zeros of Dirichlet rows kept, like matZeroRow,
        LinearSystem.py:163-167)."""
        indptr, indices = self.csr_structure()
        return sp.csr_matrix((self.values(), indices, indptr), shape=(self.rows, self.rows))

    # -- vectors -----------------------------------------------------------------------------------------------
    def upload(self, which, host):
        host = np.ascontiguousarray(host, dtype=np.float64)
        if host.shape != (self.rows,):
            raise _lib.EfvError(f"vector must have {self.rows} entries")
        _lib.check(_lib.load().efv_vec_upload(self.handle, which, host.ctypes.data_as(_lib.c_f64p)))

    def download(self, which):
        out = np.empty(self.rows)
        _lib.check(_lib.load().efv_vec_download(self.handle, which, out.ctypes.data_as(_lib.c_f64p)))
        return out

    # -- assembly ----------------------------------------------------------------------------------------------
    def heat_assemble(self, props, dt, transient):
        """props: [nRegions, 4] = Conductivity, Density, HeatCapacity, HeatGeneration."""
        p, pp = _lib.f64(np.asarray(props, dtype=np.float64).reshape(-1, 4))
        _lib.check(_lib.load().efv_heat_assemble(self.handle, pp, float(dt if dt else 0.0), int(bool(transient))))

    def elasticity_assemble(self, props, gravity=False):
        p, pp = _lib.f64(np.asarray(props, dtype=np.float64).reshape(-1, 4))
        _lib.check(_lib.load().efv_elasticity_assemble(self.handle, pp, int(bool(gravity))))

    def biot_assemble(self, props, dt, gvec=(0.0, 0.0, 0.0)):
        p, pp = _lib.f64(np.asarray(props, dtype=np.float64).reshape(-1, 9))
        g, pg = _lib.f64(np.asarray(gvec, dtype=np.float64))
        _lib.check(_lib.load().efv_biot_assemble(self.handle, pp, float(dt), pg))

    def bc_clear(self):
        _lib.check(_lib.load().efv_bc_clear(self.handle))

    def bc_dirichlet(self, rows, values):
        r, pr = _lib.i64(rows)
        v, pv = _lib.f64(values)
        _lib.check(_lib.load().efv_bc_dirichlet(self.handle, pr, pv, len(r)))

    def bc_neumann(self, boundary, dof, value, sign=1.0):
        _lib.check(_lib.load().efv_bc_neumann(self.handle, int(boundary), int(dof), float(value), float(sign)))

    def bc_neumann_expression(self, bc, dof, time=0.0, sign=1.0):
        """Variable Neumann condition.  The reference evaluates bc.getValue(outerFace.handle) (apps/heat_transfer.py:116-120):
        the expression at the coordinates of the VERTEX whose number is the outer face's global counter (Boundary.py:74) --
        reproduced as is, including its IndexError when the counter exceeds the number of vertices."""
        first = C.c_int64(0)
        n = _lib.load().efv_boundary_outer_faces(self.mesh.handle, int(bc.boundary.handle), C.byref(first))
        if first.value + n > self.grid.numberOfVertices:
            raise IndexError("index %d is out of bounds for axis 0 with size %d" % (first.value + n - 1, self.grid.numberOfVertices))
        vals = np.array([bc.getValue(first.value + k, time) for k in range(n)], dtype=np.float64)
        v, pv = _lib.f64(vals)
        _lib.check(_lib.load().efv_bc_neumann_values(self.handle, int(bc.boundary.handle), int(dof), pv, n, float(sign)))

    def assemble_dirichlet_mode(self, mode):
        """Dirichlet treatment fused into the next assembly kernel (-1 = none)."""
        _lib.check(_lib.load().efv_assemble_dirichlet_mode(self.handle, int(mode)))

    def apply_dirichlet(self, mode):
        _lib.check(_lib.load().efv_apply_dirichlet(self.handle, int(mode)))

    def build_rhs(self, transient=False):
        _lib.check(_lib.load().efv_build_rhs(self.handle, int(bool(transient))))

    # -- preconditioner ------------------------------------------------------------------------------------------
    def set_preconditioner(self, kind):
        """PRECOND_JACOBI (default) or PRECOND_AMG (aggregation multigrid, csrc/amg.cu) for solve_pcg."""
        kind = {"jacobi": PRECOND_JACOBI, "amg": PRECOND_AMG}.get(kind, kind)
        _lib.check(_lib.load().efv_system_set_preconditioner(self.handle, int(kind)))

    def set_lattice(self, nx, ny, nz):
        _lib.check(_lib.load().efv_system_set_lattice(self.handle, int(nx), int(ny), int(nz)))

    def amg_info(self):
        cap = 32
        n = C.c_int(0)
        rows = np.zeros(cap, dtype=np.int64); nnz = np.zeros(cap, dtype=np.int64); lmax = np.zeros(cap); ms = C.c_double(0.0)
        _lib.check(_lib.load().efv_amg_info(self.handle, cap, C.byref(n), rows.ctypes.data_as(_lib.c_i64p), nnz.ctypes.data_as(_lib.c_i64p),
                                            lmax.ctypes.data_as(_lib.c_f64p), C.byref(ms)))
        k = n.value
        return dict(levels=k, rows=rows[:k].tolist(), nnz=nnz[:k].tolist(), lmax=lmax[:k].tolist(), numeric_ms=ms.value)

    def amg_level(self, level, with_aggregates=True):
        """(block CSR as scipy BSR-expanded CSR in vertex-major scalar numbering, aggregates) of one hierarchy level."""
        info = self.amg_info()
        nd = self.ndof
        nrow = info["rows"][level] // nd
        nblk = info["nnz"][level] // (nd * nd)
        gptr = np.empty(nrow + 1, dtype=np.int64); gcol = np.empty(nblk, dtype=np.int32); vals = np.empty(nblk * nd * nd)
        agg = np.empty(nrow, dtype=np.int32) if (with_aggregates and level + 1 < info["levels"]) else None
        _lib.check(_lib.load().efv_amg_level_export(self.handle, level, gptr.ctypes.data_as(_lib.c_i64p), gcol.ctypes.data_as(_lib.c_i32p),
                                                    vals.ctypes.data_as(_lib.c_f64p),
                                                    agg.ctypes.data_as(_lib.c_i32p) if agg is not None else None))
        A = sp.bsr_matrix((vals.reshape(nblk, nd, nd), gcol, gptr), shape=(nrow * nd, nrow * nd)).tocsr()
        return A, agg

    # -- solve ---------------------------------------------------------------------------------------------------
    def solve_pcg(self, rtol=1e-10, maxit=10000, negate=False):
        it, rr = C.c_int(0), C.c_double(0.0)
        _lib.check(_lib.load().efv_solve_pcg(self.handle, float(rtol), int(maxit), int(bool(negate)), C.byref(it), C.byref(rr)))
        return it.value, rr.value

    def solve_bicgstab(self, rtol=1e-10, maxit=10000):
        it, rr = C.c_int(0), C.c_double(0.0)
        _lib.check(_lib.load().efv_solve_bicgstab(self.handle, float(rtol), int(maxit), C.byref(it), C.byref(rr)))
        return it.value, rr.value

    def spmv(self, which_in, which_out):
        _lib.check(_lib.load().efv_spmv(self.handle, which_in, which_out))

    def max_abs_diff(self):
        out = C.c_double(0.0)
        _lib.check(_lib.load().efv_max_abs_diff(self.handle, C.byref(out)))
        return out.value

    def advance(self):
        _lib.check(_lib.load().efv_advance(self.handle))

    def download_xold_begin(self):
        """Start the asynchronous download of XOLD (the field of the step just advanced) on the copy stream."""
        _lib.check(_lib.load().efv_download_xold_begin(self.handle))

    def download_xold_end(self):
        out = np.empty(self.rows)
        _lib.check(_lib.load().efv_download_xold_end(self.handle, out.ctypes.data_as(_lib.c_f64p)))
        return out

    def true_residual(self, negate=False):
        """||B - A X|| / ||B|| recomputed with one extra SpMV (not the solver's recurrence residual)."""
        out = C.c_double(0.0)
        _lib.check(_lib.load().efv_true_residual(self.handle, int(bool(negate)), C.byref(out)))
        return out.value

    def vec_sums(self, which):
        """(sum, sum of squares) of the owned entries, all-reduced over the ranks of a partitioned system."""
        a, b = C.c_double(0.0), C.c_double(0.0)
        _lib.check(_lib.load().efv_vec_sums(self.handle, int(which), C.byref(a), C.byref(b)))
        return a.value, b.value

    def symmetry_defect(self):
        out = C.c_double(0.0)
        _lib.check(_lib.load().efv_symmetry_defect(self.handle, C.byref(out)))
        return out.value

    def timings(self):
        a, s = C.c_double(0.0), C.c_double(0.0)
        _lib.check(_lib.load().efv_last_timings(self.handle, C.byref(a), C.byref(s)))
        return dict(assemble_ms=a.value, solve_ms=s.value)
