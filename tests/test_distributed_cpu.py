"""Host-side logic of the vertex-block partition, run with world_size 2 over gloo on the CPU (no GPU):
the two partitioners agree, every owned row sees its complete stencil, and a halo exchange with the registered
send/receive lists reproduces the global matrix-vector product."""
import os
import socket

import numpy as np
import pytest
import scipy.sparse as sp
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import pyefvlib_b200 as P
from pyefvlib_b200 import distributed as efd
from oracle import efv_oracle as O
from tests.cases import HEAT_PROPS


def oracle_of(gd):
    return O.OracleGrid(O.OracleGridData(3, np.asarray(gd.verticesCoordinates).tolist(), np.asarray(gd.elementsConnectivities).tolist(),
                                         gd.regionsNames, [list(r) for r in gd.regionsElementsIndexes], gd.boundariesNames,
                                         [list(r) for r in gd.boundariesIndexes], np.asarray(gd.boundariesConnectivities).tolist()))


@pytest.mark.parametrize("kind", ["hex", "tet"])
@pytest.mark.parametrize("world", [2, 3])
def test_structured_part_equals_generic_partition(kind, world):
    n = 7
    g = P.structured_grid(n, kind)
    ranges = efd.block_ranges(n ** 3, world, align=n * n)
    for rank in range(world):
        a = efd.partition_grid(g, rank, world, ranges)
        b = efd.structured_part(n, kind, rank, world)
        assert a.n_owned == b.n_owned and np.array_equal(a.local_to_global, b.local_to_global)
        assert np.array_equal(a.grid.coords, b.grid.coords) and np.array_equal(a.grid.conn, b.grid.conn)
        assert a.nbr == b.nbr and np.array_equal(a.recv_ptr, b.recv_ptr)
        for x, y in zip(a.send_lists, b.send_lists):
            assert np.array_equal(x, y)
        assert np.array_equal(a.grid.facet_conn, b.grid.facet_conn) and np.array_equal(a.grid.boundary_ptr, b.grid.boundary_ptr)
        for x, y in zip(a.boundary_vertices, b.boundary_vertices):
            assert set(x.tolist()) == set(y.tolist())


def test_owned_rows_are_complete():
    """Rows of owned vertices assembled on the local mesh equal the corresponding rows of the global matrix."""
    n, world = 6, 3
    gd = P.structured_grid_data(n, "hex", 0.15)
    g = P.Grid(gd)
    og = oracle_of(gd)
    R, C, V = O.heat_interior_coo(og, HEAT_PROPS, True, 1e-2)
    A = sp.csr_matrix((V, (R, C)), shape=(n ** 3, n ** 3))
    for rank in range(world):
        part = efd.partition_grid(g, rank, world, efd.block_ranges(n ** 3, world, align=n * n))
        lgd = part.grid.gridData
        ol = oracle_of(lgd)
        R, C, V = O.heat_interior_coo(ol, HEAT_PROPS, True, 1e-2)
        Al = sp.csr_matrix((V, (R, C)), shape=(ol.numberOfVertices,) * 2)
        l2g = part.local_to_global
        sub = A[l2g[:part.n_owned]][:, l2g]
        assert abs(Al[:part.n_owned] - sub).max() <= 1e-13 * abs(A).max()


@pytest.mark.parametrize("fixture", ["heat_Prisms_3d", "heat_Pyrams_3d", "heat_eight_sphere"])
def test_owned_rows_are_complete_on_bundled_meshes(fixture):
    """Same property on the reference's prism, pyramid and unstructured tetrahedron meshes (ragged boundary facets,
    5/6-vertex elements): the generic partitioner keeps every element incident to an owned vertex."""
    from tests import util
    z = util.load(fixture)
    g = util.product_grid(z)
    og = util.oracle_grid(z)
    nV = g.numberOfVertices
    R, C, V = O.heat_interior_coo(og, HEAT_PROPS, True, 1e-2)
    A = sp.csr_matrix((V, (R, C)), shape=(nV, nV))
    world = 2
    seen = np.zeros(nV, dtype=int)
    for rank in range(world):
        part = efd.partition_grid(g, rank, world, efd.block_ranges(nV, world))
        lgd = part.grid.gridData
        ol = O.OracleGrid(O.OracleGridData(3, np.asarray(lgd.verticesCoordinates, dtype=float).reshape(-1, 3).tolist(),
                                           [list(map(int, c)) for c in lgd.elementsConnectivities], lgd.regionsNames,
                                           [list(map(int, r)) for r in lgd.regionsElementsIndexes], lgd.boundariesNames,
                                           [list(map(int, r)) for r in lgd.boundariesIndexes],
                                           [list(map(int, c)) for c in lgd.boundariesConnectivities]))
        R, C, V = O.heat_interior_coo(ol, HEAT_PROPS, True, 1e-2)
        Al = sp.csr_matrix((V, (R, C)), shape=(ol.numberOfVertices,) * 2)
        l2g = part.local_to_global
        seen[l2g[:part.n_owned]] += 1
        sub = A[l2g[:part.n_owned]][:, l2g]
        assert abs(Al[:part.n_owned] - sub).max() <= 1e-13 * abs(A).max()
    assert (seen == 1).all()                       # every vertex is owned exactly once


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, n, kind, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        part = efd.structured_part(n, kind, rank, world)
        nV = n ** 3
        x_global = np.sin(np.arange(nV) * 0.37) + 2.0
        x = np.zeros(part.grid.numberOfVertices)
        x[:part.n_owned] = x_global[part.local_to_global[:part.n_owned]]
        # halo exchange exactly as the CUDA library does it: pack send lists, receive into the ghost block
        reqs = []
        recv_bufs = []
        for i, q in enumerate(part.nbr):
            sb = torch.from_numpy(x[part.send_lists[i]].copy())
            rb = torch.zeros(int(part.recv_ptr[i + 1] - part.recv_ptr[i]), dtype=torch.float64)
            recv_bufs.append(rb)
            reqs.append(dist.isend(sb, q)); reqs.append(dist.irecv(rb, q))
        for r in reqs:
            r.wait()
        for i, rb in enumerate(recv_bufs):
            x[part.n_owned + part.recv_ptr[i]:part.n_owned + part.recv_ptr[i + 1]] = rb.numpy()
        assert np.array_equal(x, x_global[part.local_to_global])
        # local SpMV on owned rows == rows of the global product; global dot product by all-reduce
        ol = oracle_of(part.grid.gridData)
        R, C, V = O.heat_interior_coo(ol, HEAT_PROPS, True, 1e-2)
        Al = sp.csr_matrix((V, (R, C)), shape=(ol.numberOfVertices,) * 2)
        y = (Al @ x)[:part.n_owned]
        dot = torch.tensor([float(y @ x[:part.n_owned])], dtype=torch.float64)
        dist.all_reduce(dot)
        if rank == 0:
            og = O.OracleGrid(O.structured_grid_data(n, kind))
            R, C, V = O.heat_interior_coo(og, HEAT_PROPS, True, 1e-2)
            A = sp.csr_matrix((V, (R, C)), shape=(nV, nV))
            yg = A @ x_global
            assert np.allclose(y, yg[:part.n_owned], rtol=1e-13, atol=1e-13 * abs(yg).max())
            assert abs(dot.item() - float(yg @ x_global)) <= 1e-11 * abs(float(yg @ x_global))
        out.put((rank, "ok"))
    except Exception as ex:  # pragma: no cover
        out.put((rank, repr(ex)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("kind", ["hex", "tet"])
def test_halo_exchange_and_allreduce_gloo(kind):
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, 6, kind, out)) for r in range(world)]
    for p in procs:
        p.start()
    results = [out.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] == "ok" for r in results), results
