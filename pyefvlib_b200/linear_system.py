"""LinearSystem / LinearSystemCSR with the reference's interface (PyEFVLib/LinearSystem.py:26-183) and a `gpu` backend.

The reference switches backends with `LinearSystem(stencil, nDOF, PETSc_backend=False)` ('numpy/scipy' | 'petsc',
LinearSystem.py:27-35).  Here the only backend is 'gpu': add/set/get and matZeroRow edit a host staging copy of the
CSR values exactly like the reference's numpy/scipy branch (structure: field-major blocks, columns in stencil order,
LinearSystem.py:114-128), and `solve()` ships pattern, values and RHS through the C-ABI
(efv_system_create_from_csr / efv_system_set_values) and runs PCG or BiCGStab on the device -- the call shape the stale
apps/csr/heat_transfer_csr.py:57-133 intended (restartRHS / add* / matZeroRow / spsolve).
Unlike the reference, touching an entry outside the pattern raises (LinearSystem.py:139-140 only prints and then
corrupts every entry)."""
import ctypes as C

import numpy as np
from scipy.sparse import csr_matrix

from . import _lib


class LinearSystem(object):
    """Dense variant (LinearSystem.py:26-111): kept for interface completeness; it converts to CSR before solving."""

    def __init__(self, stencil, nDOF, backend="gpu"):
        if backend != "gpu":
            raise Exception("pyefvlib_b200 only ships the 'gpu' backend")
        self.stencil, self.nDOF = stencil, nDOF
        self.nVertices = len(stencil)
        self.size = self.nDOF * self.nVertices
        self.backend = backend

    def initialize(self):
        self.rhs = np.zeros(self.size)
        self.matrix = np.zeros((self.size, self.size))

    def addValueToMatrix(self, row, col, value):
        self.matrix[row][col] += value

    def setValueToMatrix(self, row, col, value):
        self.matrix[row][col] = value

    def getValueFromMatrix(self, row, col):
        return self.matrix[row][col]

    def addValueToRHS(self, row, value):
        self.rhs[row] += value

    def setValueToRHS(self, row, value):
        self.rhs[row] = value

    def getValueFromRHS(self, row):
        return self.rhs[row]

    def restartRHS(self):
        self.rhs = np.zeros(self.size)

    def matZeroRow(self, row, diagonal_value):
        self.matrix[row, :] = 0
        self.matrix[row, row] = diagonal_value

    def getDense(self):
        return self.matrix

    def getRHS(self):
        return self.rhs

    def assembly(self):
        return

    def _csr(self):
        return csr_matrix(self.matrix)

    def solve(self, method="bicgstab", rtol=1e-10, maxit=20000, x0=None):
        A = self._csr()
        return _solve_csr(A.indptr, A.indices, A.data, self.rhs, method, rtol, maxit, x0)


class LinearSystemCSR(LinearSystem):
    def initialize(self):
        self.rhs = np.zeros(self.size)
        indPtr, indices = [0], []
        for i in range(self.nDOF):                                          # LinearSystem.py:117-124
            for vertex_stencil in self.stencil:
                indPtr.append(indPtr[-1] + self.nDOF * len(vertex_stencil))
                for j in range(self.nDOF):
                    for vertex in vertex_stencil:
                        indices.append(vertex + j * self.nVertices)
        indPtr, indices = np.array(indPtr), np.array(indices, dtype=np.int64)
        self.matrix = csr_matrix((np.zeros(indices.size), indices, indPtr), shape=(self.size, self.size))

    def _index(self, row, col):
        pos1, pos2 = self.matrix.indptr[row], self.matrix.indptr[row + 1]
        hit = np.where(self.matrix.indices[pos1:pos2] == col)[0]
        if hit.size == 0:
            raise Exception("New entries not allowed.")                     # the reference only prints this
        return pos1 + hit[0]

    def addValueToMatrix(self, row, col, value):
        self.matrix.data[self._index(row, col)] += value

    def setValueToMatrix(self, row, col, value):
        self.matrix.data[self._index(row, col)] = value

    def getValueFromMatrix(self, row, col):
        return self.matrix.data[self._index(row, col)]

    def matZeroRow(self, row, diagonal_value):                              # LinearSystem.py:163-167
        self.matrix.data[self.matrix.indptr[row]:self.matrix.indptr[row + 1]] = 0.0
        self.setValueToMatrix(row, row, diagonal_value)

    def getDense(self):
        return self.matrix.todense()

    def getSparse(self):
        return (self.matrix.indptr, self.matrix.indices, self.matrix.data)

    def _csr(self):
        return self.matrix


def _solve_csr(indptr, indices, data, rhs, method, rtol, maxit, x0):
    """Solve on the device; returns (x, iterations, relative residual)."""
    L = _lib.load()
    n = len(rhs)
    ip, pip = _lib.i64(indptr)
    ix, pix = _lib.i32(indices)
    handle = _lib.check_ptr(L.efv_system_create_from_csr(n, pip, pix))
    try:
        d, pd = _lib.f64(data)
        _lib.check(L.efv_system_set_values(handle, pd))
        b, pb = _lib.f64(rhs)
        _lib.check(L.efv_vec_upload(handle, 1, pb))
        x, px = _lib.f64(np.zeros(n) if x0 is None else x0)
        _lib.check(L.efv_vec_upload(handle, 0, px))
        it, rr = C.c_int(0), C.c_double(0.0)
        if method == "pcg":
            _lib.check(L.efv_solve_pcg(handle, float(rtol), int(maxit), 0, C.byref(it), C.byref(rr)))
        elif method == "bicgstab":
            _lib.check(L.efv_solve_bicgstab(handle, float(rtol), int(maxit), C.byref(it), C.byref(rr)))
        else:
            raise Exception("method must be 'pcg' or 'bicgstab'")
        out = np.empty(n)
        _lib.check(L.efv_vec_download(handle, 0, out.ctypes.data_as(_lib.c_f64p)))
        return out, it.value, rr.value
    finally:
        L.efv_system_destroy(handle)
