"""Multi-GPU: vertex-block domain decomposition.  The reference has no distributed code.

One process per GPU (torch.distributed for the plumbing).  The global vertex range is cut into contiguous blocks;
rank r owns the vertices (= matrix rows) of block r and assembles every element incident to one of them, so the
layer of elements straddling an interface is assembled redundantly on both sides and assembly needs NO
communication.  Local vertices are numbered [owned (global order) | ghosts (global order, hence grouped by owner)].
The Krylov solver exchanges the halo of its search direction before every SpMV (NCCL send/recv inside the CUDA
library, on its stream) and all-reduces its dot products.

`partition_grid` cuts an arbitrary global Grid (host numpy; used by the tests and for unstructured meshes),
`structured_part` builds the local slab of the synthetic benchmark meshes directly, without the global mesh.
Both give the same LocalPart (tests/test_distributed_cpu.py).
"""
import os

import numpy as np

from . import grid as _grid


def block_ranges(n_vertices, world, align=1):
    """Contiguous vertex blocks, as even as possible, boundaries multiples of `align` (n^2 = one mesh plane for the
    structured meshes)."""
    units = n_vertices // align
    assert units * align == n_vertices and units >= world
    cuts = [(units * r) // world * align for r in range(world + 1)]
    return np.array(cuts, dtype=np.int64)


class LocalPart:
    """What one rank holds: its local Grid, the owned/ghost split and the exchange lists."""

    def __init__(self, rank, world, ranges, grid, n_owned, local_to_global, nbr, send_lists, recv_ptr, boundary_vertices):
        self.rank, self.world, self.ranges = rank, world, ranges
        self.grid, self.n_owned, self.local_to_global = grid, int(n_owned), local_to_global
        self.nbr = list(nbr)                      # neighbour ranks, ascending
        self.send_lists = send_lists              # per neighbour: local ids of owned vertices it needs (global order)
        self.recv_ptr = np.asarray(recv_ptr, dtype=np.int64)   # per neighbour slice of the ghost block
        self.boundary_vertices = boundary_vertices   # per boundary: local ids (owned and ghost) lying on it
        self.lattice = None                          # (nx, ny, nz) when the OWNED vertices are numbered like a lattice

    @property
    def n_ghost(self):
        return self.grid.numberOfVertices - self.n_owned


def _make_local_grid(dimension, coords, conn_local, region_names, region_of_elem, boundary_names, facets_local):
    gd = _grid.GridData("partition")
    gd.setDimension(dimension)
    gd.setVertices(coords)
    gd.setElementConnectivity(conn_local)
    gd.setRegions(list(region_names), [np.nonzero(region_of_elem == h)[0] for h in range(len(region_names))])
    sizes = [len(f) for f in facets_local]
    starts = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    if isinstance(conn_local, np.ndarray) and all(isinstance(f, np.ndarray) and f.ndim == 2 for f in facets_local) \
            and len({f.shape[1] for f in facets_local if len(f)}) <= 1:
        width = next((f.shape[1] for f in facets_local if len(f)), 2)
        bconn = np.concatenate([f.reshape(-1, width) for f in facets_local]) if sum(sizes) else np.zeros((0, width), dtype=np.int32)
        gd.setBoundaries(list(boundary_names), [np.arange(starts[h], starts[h + 1], dtype=np.int64) for h in range(len(sizes))],
                         bconn.astype(np.int32))
    else:
        flat = [list(map(int, f)) for fl in facets_local for f in fl]
        gd.setBoundaries(list(boundary_names), [list(range(starts[h], starts[h + 1])) for h in range(len(sizes))], flat)
    gd.setShapes([[]] * 7)
    return _grid.Grid(gd)


def partition_grid(g, rank, world, ranges=None):
    """Cut the global Grid `g` for `rank` (uniform element type per grid or mixed; host numpy)."""
    nV = g.numberOfVertices
    ranges = block_ranges(nV, world) if ranges is None else np.asarray(ranges, dtype=np.int64)
    a, b = int(ranges[rank]), int(ranges[rank + 1])
    ep, conn = g.elem_ptr, g.conn.astype(np.int64)
    owner_of_vertex = np.searchsorted(ranges, np.arange(nV), side="right") - 1
    elem_of_entry = np.repeat(np.arange(g.numberOfElements), np.diff(ep))
    mine_entry = (conn >= a) & (conn < b)
    elem_mask = np.zeros(g.numberOfElements, dtype=bool)
    elem_mask[elem_of_entry[mine_entry]] = True
    local_elems = np.nonzero(elem_mask)[0]
    entry_mask = elem_mask[elem_of_entry]
    used = np.unique(conn[entry_mask])
    ghosts = used[(used < a) | (used >= b)]
    n_owned = b - a
    l2g = np.concatenate([np.arange(a, b, dtype=np.int64), ghosts])
    g2l = np.full(nV, -1, dtype=np.int64)
    g2l[l2g] = np.arange(len(l2g))
    lens = np.diff(ep)[local_elems]
    if len(np.unique(np.diff(ep))) == 1:
        m = int(lens[0])
        conn_local = g2l[conn.reshape(-1, m)[local_elems]].astype(np.int32)
    else:
        conn_local = [g2l[conn[ep[e]:ep[e + 1]]].tolist() for e in local_elems]
    # boundary facets with at least one owned vertex (their owning element is then local); global order preserved
    facets_local, bverts = [], []
    for bd in g.boundaries:
        fl = []
        for k in range(bd.facet_begin, bd.facet_end):
            fv = g.facet_conn[g.facet_ptr[k]:g.facet_ptr[k + 1]].astype(np.int64)
            if ((fv >= a) & (fv < b)).any():
                fl.append(g2l[fv])
        widths = {len(f) for f in fl}
        facets_local.append(np.array(fl, dtype=np.int32).reshape(len(fl), widths.pop()) if len(widths) == 1 else fl)
        lv = g2l[bd.verticesIndexes]
        bverts.append(lv[lv >= 0])
    # neighbours and exchange lists (computed identically on both sides: sorted by global id)
    ghost_owner = owner_of_vertex[ghosts]
    nbr = sorted(set(ghost_owner.tolist()))
    recv_ptr = np.concatenate([[0], np.cumsum([(ghost_owner == q).sum() for q in nbr])])
    send_lists = []
    vert_owner_entry = owner_of_vertex[conn]
    for q in nbr:
        # my owned vertices that appear in an element which also has a vertex owned by q
        has_q = np.zeros(g.numberOfElements, dtype=bool)
        has_q[elem_of_entry[vert_owner_entry == q]] = True
        sel = has_q[elem_of_entry] & mine_entry
        send_lists.append((np.unique(conn[sel]) - a).astype(np.int32))
    grid = _make_local_grid(g.dimension, g.coords[l2g], conn_local, [r.name for r in g.regions], g.region_of_element[local_elems],
                            [bd.name for bd in g.boundaries], facets_local)
    return LocalPart(rank, world, ranges, grid, n_owned, l2g, nbr, send_lists, recv_ptr, bverts)


def structured_part(n, kind, rank, world, perturb=0.0, nz=None):
    """Local slab of the n x n x nz (default n^3) structured mesh for `rank`: planes are dealt out in contiguous blocks."""
    n2 = n * n
    nz = nz or n
    ranges = block_ranges(n2 * nz, world, align=n2)
    k0, k1 = int(ranges[rank] // n2), int(ranges[rank + 1] // n2)           # owned planes [k0, k1)
    lo, hi = max(k0 - 1, 0), min(k1 + 1, nz)                                # local planes [lo, hi)
    arr = _grid.structured_arrays(n, kind, layer_lo=lo, layer_hi=min(hi - 1, nz - 1), nz=nz,
                                  conn_dtype=np.int32 if n2 * nz < 2 ** 31 else np.int64)
    n_owned = (k1 - k0) * n2
    below = np.arange(lo * n2, k0 * n2, dtype=np.int64)
    above = np.arange(k1 * n2, hi * n2, dtype=np.int64)
    l2g = np.concatenate([np.arange(k0 * n2, k1 * n2, dtype=np.int64), below, above])

    def to_local(gid):
        gid = np.asarray(gid, dtype=np.int64)
        out = gid - k0 * n2
        out = np.where(gid < k0 * n2, n_owned + (gid - lo * n2), out)
        out = np.where(gid >= k1 * n2, n_owned + len(below) + (gid - k1 * n2), out)
        return out
    # connectivity: one affine pass (owned ids), then the general map on the two element layers that touch a ghost plane
    # (2.2 s -> 0.3 s per 10 M-vertex slab)
    conn_g = arr["conn"]
    conn_local = (conn_g - conn_g.dtype.type(k0 * n2)).astype(np.int32, copy=False)
    per_layer = (n - 1) * (n - 1) * (6 if kind == "tet" else 1)
    if len(below):
        conn_local[:per_layer] = to_local(conn_g[:per_layer])
    if len(above):
        conn_local[-per_layer:] = to_local(conn_g[-per_layer:])
    facets_local = []
    for name in _grid.BOUNDARY_NAMES_3D:
        f = arr["facets"][name]
        if len(f):
            keep = ((f >= k0 * n2) & (f < k1 * n2)).any(axis=1)
            f = f[keep]
        facets_local.append(to_local(f).astype(np.int32).reshape(len(f), 4 if kind == "hex" else 3))
    plane_ranges = [(k0, k1)] + ([(lo, k0)] if len(below) else []) + ([(k1, hi)] if len(above) else [])     # = l2g, plane by plane
    coords = _grid.structured_coords(n, l2g, perturb, nz, planes=plane_ranges)
    npl = len(l2g) // n2
    I = np.broadcast_to(np.arange(n, dtype=np.int64)[None, None, :], (npl, n, n)).reshape(-1)
    J = np.broadcast_to(np.arange(n, dtype=np.int64)[None, :, None], (npl, n, n)).reshape(-1)
    K = np.repeat(np.concatenate([np.arange(a, b, dtype=np.int64) for a, b in plane_ranges]), n2)
    on = {"West": I == 0, "East": I == n - 1, "South": J == 0, "North": J == n - 1, "Bottom": K == 0, "Top": K == nz - 1}
    bverts = [np.nonzero(on[name])[0].astype(np.int64) for name in _grid.BOUNDARY_NAMES_3D]
    nbr, send_lists, recv = [], [], [0]
    if len(below):
        nbr.append(rank - 1); send_lists.append(np.arange(0, n2, dtype=np.int32)); recv.append(recv[-1] + len(below))
    if len(above):
        nbr.append(rank + 1); send_lists.append(np.arange(n_owned - n2, n_owned, dtype=np.int32)); recv.append(recv[-1] + len(above))
    grid = _make_local_grid(3, coords, conn_local, ["Body"], np.zeros(len(conn_local), dtype=np.int32), _grid.BOUNDARY_NAMES_3D, facets_local)
    part = LocalPart(rank, world, ranges, grid, n_owned, l2g, nbr, send_lists, recv, bverts)
    part.lattice = (n, n, k1 - k0)            # the owned vertices are numbered like an n x n x (owned planes) lattice
    return part


# ---------------------------------------------------------------------------------------------------------------------
# device side
# ---------------------------------------------------------------------------------------------------------------------
def nccl_library_path():
    """The NCCL shared object torch itself uses, so that the process holds one NCCL instance."""
    try:
        import torch
        root = os.path.dirname(os.path.dirname(torch.__file__))
        cand = os.path.join(root, "nvidia", "nccl", "lib", "libnccl.so.2")
        if os.path.isfile(cand):
            return cand
    except Exception:
        pass
    return "libnccl.so.2"


_comm_ready = False
_peer_ok = False


def init_comm(use_nccl=True):
    """Create the library's communicator.  With NCCL (default) the unique id travels through torch.distributed; with
    `use_nccl=False` the communicator is peer-memory only (CUDA IPC), which also works for several ranks on ONE GPU.
    The peer-memory path (EFV_COMM=peer, the default) is taken only if EVERY rank is on the same host and managed to
    map every other rank's region; otherwise all ranks fall back to NCCL together."""
    global _comm_ready, _peer_ok
    import ctypes as C
    import socket
    import torch.distributed as dist
    from . import _lib
    if _comm_ready:
        return
    L = _lib.load()
    rank, world = dist.get_rank(), dist.get_world_size()
    path = nccl_library_path().encode()
    if use_nccl:
        buf = C.create_string_buffer(128)
        if rank == 0:
            _lib.check(L.efv_comm_unique_id(path, buf))
        box = [bytes(buf.raw)]
        dist.broadcast_object_list(box, src=0)
        _lib.check(L.efv_comm_init(path, rank, world, box[0]))
    else:
        _lib.check(L.efv_comm_init(None, rank, world, None))
    want_peer = os.environ.get("EFV_COMM", "peer") == "peer" and world <= 16
    # every rank takes part in the same collectives below, whatever it decides locally
    mine = C.create_string_buffer(64)
    ok = want_peer
    if want_peer:
        try:
            _lib.check(L.efv_comm_peer_export(mine))
        except _lib.EfvError:
            ok = False
    table = [None] * world
    dist.all_gather_object(table, (socket.gethostname(), ok, bytes(mine.raw)))
    ok = all(t[1] for t in table) and len({t[0] for t in table}) == 1
    if ok:
        try:
            _lib.check(L.efv_comm_peer_import(b"".join(t[2] for t in table)))
        except _lib.EfvError:
            ok = False
    flags = [None] * world
    dist.all_gather_object(flags, ok)
    _peer_ok = all(flags)
    if not _peer_ok:
        _lib.check(L.efv_comm_peer_release())
        if not use_nccl:
            raise _lib.EfvError("peer-memory communicator requested without NCCL, but the ranks cannot map each other's memory")
    _comm_ready = True


def make_system(part, ndof=1):
    """DeviceSystem on the rank's local mesh, restricted to its owned rows, with the halo registered."""
    import ctypes as C
    from . import _lib
    from .device import DeviceSystem
    s = DeviceSystem(part.grid, ndof)
    L = _lib.load()
    _lib.check(L.efv_system_set_owned(s.handle, part.n_owned))
    if part.lattice is not None:
        s.set_lattice(*part.lattice)
    nb = np.asarray(part.nbr, dtype=np.int32)
    sp = np.concatenate([[0], np.cumsum([len(x) for x in part.send_lists])]).astype(np.int64)
    si = np.concatenate(part.send_lists).astype(np.int32) if part.send_lists else np.zeros(0, dtype=np.int32)
    rp = np.asarray(part.recv_ptr, dtype=np.int64)
    _lib.check(L.efv_system_set_halo(s.handle, len(nb), nb.ctypes.data_as(_lib.c_intp), sp.ctypes.data_as(_lib.c_i64p),
                                     si.ctypes.data_as(_lib.c_i32p), rp.ctypes.data_as(_lib.c_i64p)))
    if _peer_ok:                      # a global decision (init_comm): every rank runs the collective below or none does
        import torch.distributed as dist
        # staging buffers: everybody publishes (handle, ghost count, {neighbour: (ghost offset, flag slot) reserved for it})
        mine = C.create_string_buffer(64)
        _lib.check(L.efv_halo_stage_export(s.handle, mine))
        info = {int(q): (int(part.recv_ptr[i]), i) for i, q in enumerate(part.nbr)}
        table = [None] * part.world
        dist.all_gather_object(table, (bytes(mine.raw), int(part.n_ghost), info))
        handles = b"".join(table[int(q)][0] for q in part.nbr)
        offs = np.array([table[int(q)][2][part.rank][0] for q in part.nbr], dtype=np.int64)
        slots = np.array([table[int(q)][2][part.rank][1] for q in part.nbr], dtype=np.int32)
        ghosts = np.array([table[int(q)][1] for q in part.nbr], dtype=np.int64)
        _lib.check(L.efv_halo_stage_import(s.handle, handles, offs.ctypes.data_as(_lib.c_i64p), slots.ctypes.data_as(_lib.c_intp),
                                           ghosts.ctypes.data_as(_lib.c_i64p)))
    return s


def send_heat_bcs(system, part, bcs):
    """Boundary conditions of the heat problem on a partition: Neumann through the local facets, Dirichlet on every
    local vertex (owned or ghost) of the boundary, boundaries in grid order (heat_transfer.py:115-126)."""
    system.bc_clear()
    for b in part.grid.boundaries:
        kind, val = bcs[b.name]
        if kind == "neumann":
            system.bc_neumann(b.handle, 0, val, 1.0)
    for h, b in enumerate(part.grid.boundaries):
        kind, val = bcs[b.name]
        if kind == "dirichlet":
            idx = part.boundary_vertices[h]
            system.bc_dirichlet(idx, np.full(len(idx), val))


def _peak_gbs():
    import json
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(path))["hbm_gbs"])
    except Exception:
        return 6650.0


def _allreduce_max(value, world):
    if world == 1:
        return float(value)
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(value)], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def run_partitioned_case(n, nz, rank, world, BC3D, PROPS, DT, RTOL, transient, steps, warmup, precond, stream, sampler=None):
    """The hot path on the n x n x nz structured hexahedral mesh cut into `world` vertex-block slabs (world = 1: the whole
    mesh on one GPU): slab generation on the host, upload + geometry + CSR (setup), then `warmup` + `steps` passes of
    assembly + BCs + RHS + PCG solve (+ advance).  Transient runs restart from T0 = 0 after the warm-up; steady runs
    solve from x0 = 0 every time.  Times are CUDA events on the library stream, max over ranks."""
    import time
    import torch
    import torch.distributed as dist
    from . import _lib
    from .device import DIRICHLET_SYMMETRIC, VEC_X, VEC_XOLD
    L = _lib.load()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()
    t0 = time.perf_counter()
    part = structured_part(n, "hex", rank, world, nz=nz)
    host_s = time.perf_counter() - t0
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    system = make_system(part, 1)
    system.set_preconditioner(precond)
    e1.record(stream)
    barrier()
    setup_ms = _allreduce_max(e0.elapsed_time(e1), world)
    props = np.array([[PROPS["Body"]["Conductivity"], PROPS["Body"]["Density"], PROPS["Body"]["HeatCapacity"], PROPS["Body"]["HeatGeneration"]]])
    zeros = np.zeros(system.rows)
    system.upload(VEC_XOLD, zeros); system.upload(VEC_X, zeros)

    def step():
        send_heat_bcs(system, part, BC3D)
        system.assemble_dirichlet_mode(DIRICHLET_SYMMETRIC)
        system.heat_assemble(props, DT, transient)
        system.build_rhs(transient)
        if not transient:
            system.upload(VEC_X, zeros)          # steady: every step is the full solve from zero
        it, rr = system.solve_pcg(RTOL, 20000)
        system.advance()
        return it, rr
    if sampler:
        sampler.start()                          # clocks under load: warm-up + timed steps (not the host-side slab generation)
    warm = [step()[0] for _ in range(warmup)]
    if transient:
        system.upload(VEC_XOLD, zeros); system.upload(VEC_X, zeros)      # timed steps = time steps 1..K from T0
    launches1 = L.efv_kernel_launch_count()
    barrier()
    e0.record(stream)
    iters, solve_ms, asm_ms = [], [], []
    for _ in range(steps):
        it, rr = step()
        assert rr <= RTOL, f"PCG did not converge on rank {rank}: {rr}"
        iters.append(it)
        tm = system.timings(); solve_ms.append(tm["solve_ms"]); asm_ms.append(tm["assemble_ms"])
    e1.record(stream)
    barrier()
    launches2 = L.efv_kernel_launch_count()
    total_ms = _allreduce_max(e0.elapsed_time(e1), world)
    clocks = sampler.stop() if sampler else None
    true_relres = system.true_residual()                     # b - A x with one extra SpMV, all-reduced
    assert true_relres <= 2e-10, f"true residual {true_relres:.3e} exceeds 2e-10"
    sums = system.vec_sums(VEC_X)
    out = dict(n=n, nz=nz, vertices=n * n * nz, ms_per_step=total_ms / steps, iterations=iters, warmup_iterations=warm,
               solve_ms=float(np.mean(solve_ms)), assemble_ms=float(np.mean(asm_ms)), setup_ms=setup_ms, host_slab_generation_s=host_s,
               true_relres=true_relres, field_sum=sums[0], field_sum_of_squares=sums[1], launches=int(launches2 - launches1),
               iterations_total=int(np.sum(iters)), ms_per_iteration=float(np.sum(solve_ms)) / max(1, int(np.sum(iters))),
               n_owned=part.n_owned, clocks=clocks)
    if precond == "amg":
        out["amg"] = system.amg_info()
    return out, system, part


def c5_steady(args, rank, world, BC3D, PROPS, DT, RTOL, stream):
    """BASELINE.json configs[4] / the north_star target: STEADY heat conduction on the 368^3-vertex mesh (49.8 M vertices),
    assembled and solved to 1e-10 from x0 = 0, strong scaling (the same mesh on `world` GPUs)."""
    n = args.c5_n
    res, system, part = run_partitioned_case(n, n, rank, world, BC3D, PROPS, DT, RTOL, False, 2, 1, args.precond, stream)
    del system
    part.grid._device = None
    return {"workload": f"C5: 3D STEADY heat conduction, structured hex mesh {n}^3 vertices, 3-D BC set, assembled + solved to 1e-10 from "
                        f"x0 = 0 with PCG({args.precond}); strong scaling over {world} GPU(s)",
            "n_gpus": world, "vertices": res["vertices"], "time_to_solution_s": res["ms_per_step"] * 1e-3,
            "dof_per_s": res["vertices"] / (res["ms_per_step"] * 1e-3), "iterations": res["iterations"],
            "assemble_ms": res["assemble_ms"], "solve_ms": res["solve_ms"], "setup_ms_once": res["setup_ms"],
            "host_slab_generation_s": res["host_slab_generation_s"], "true_relres": res["true_relres"],
            "field_sum": res["field_sum"], "field_sum_of_squares": res["field_sum_of_squares"], "amg": res.get("amg")}


def cross_check_small(rank, world, BC3D, PROPS, DT, RTOL, precond, stream):
    """N-GPU against 1-GPU on the same global mesh (24^3): the owned-row checksums of the partitioned solve must equal
    the sums of a plain single-GPU solve (every rank runs its own copy of it next to the partitioned system)."""
    from .device import DeviceSystem, DIRICHLET_SYMMETRIC, VEC_X, VEC_XOLD
    n = 24
    res, system, part = run_partitioned_case(n, n, rank, world, BC3D, PROPS, DT, RTOL, True, 2, 0, precond, stream)
    g = _grid.Grid(_grid.structured_grid_data(n, "hex"))
    single = DeviceSystem(g, 1)
    single.set_preconditioner(precond)
    props = np.array([[PROPS["Body"]["Conductivity"], PROPS["Body"]["Density"], PROPS["Body"]["HeatCapacity"], PROPS["Body"]["HeatGeneration"]]])
    single.upload(VEC_XOLD, np.zeros(single.rows)); single.upload(VEC_X, np.zeros(single.rows))
    for _ in range(2):
        single.bc_clear()
        for b in g.boundaries:
            kind, val = BC3D[b.name]
            if kind == "neumann":
                single.bc_neumann(b.handle, 0, val, 1.0)
        for b in g.boundaries:
            kind, val = BC3D[b.name]
            if kind == "dirichlet":
                single.bc_dirichlet(b.verticesIndexes, np.full(len(b.verticesIndexes), val))
        single.assemble_dirichlet_mode(DIRICHLET_SYMMETRIC)
        single.heat_assemble(props, DT, True)
        single.build_rhs(True)
        single.solve_pcg(RTOL, 20000)
        single.advance()
    s1 = single.vec_sums(VEC_X)
    rel = max(abs(res["field_sum"] - s1[0]) / abs(s1[0]), abs(res["field_sum_of_squares"] - s1[1]) / abs(s1[1]))
    assert rel <= 1e-9, f"partitioned and single-GPU solutions differ: {rel:.3e}"
    return {"mesh": f"{n}^3", "n_gpus": world, "field_sum_partitioned": res["field_sum"], "field_sum_single_gpu": s1[0],
            "field_sum_of_squares_partitioned": res["field_sum_of_squares"], "field_sum_of_squares_single_gpu": s1[1], "max_rel_diff": rel}


def bench_multi_gpu(args, BC3D, PROPS, DT, RTOL):
    """N > 1 leg of bench.py.  Headline line: WEAK scaling of C2 (every rank owns an n x n x n block of an n x n x (n N) box,
    same per-GPU work as the N = 1 run), so that the driver's per-N series is one workload.  Nested: the north_star's
    C5 steady STRONG-scaling case and the N-GPU vs 1-GPU cross-check."""
    import json
    import torch
    import torch.distributed as dist
    from . import _lib
    rank, world, local_rank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    L = _lib.load()
    _lib.check(L.efv_set_device(local_rank))
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    init_comm()
    stream = torch.cuda.ExternalStream(L.efv_get_stream())
    transient = not getattr(args, "steady", False)
    strong = args.workload == "c5"
    # weak scaling: the C2 cube refined isotropically so that every rank keeps the per-GPU size of the N = 1 run
    # (216^3 -> 273^3 -> 343^3 -> 433^3 vertices on 1, 2, 4, 8 GPUs; z-slabs of n^2 planes), or, with --weak-family stack,
    # one n^3 block per rank stacked into an n x n x (n N) box (the round-1 family, whose aspect ratio grows with N).
    # The refined sizes are made ODD: the SpMV gathers x from 9 lines that are n and n^2 entries apart, and line strides
    # that are multiples of 128 B alias in the caches (same rows, one GPU: 65 ns per 1 000 rows at n = 432, 53 at n = 433;
    # tools/spmv_shape_probe.py, profiles/r02_spmv_vs_plane_size.json)
    stack = getattr(args, "weak_family", "cube") == "stack"
    n = args.c5_n if strong else (args.n if stack else int(round(args.n * world ** (1.0 / 3.0))) | 1)
    nz = n if (strong or not stack) else n * world
    sampler = args.clock_sampler(local_rank) if (rank == 0 and getattr(args, "clock_sampler", None)) else None
    res, system, part = run_partitioned_case(n, nz, rank, world, BC3D, PROPS, DT, RTOL, transient, args.steps, args.warmup, args.precond, stream,
                                             sampler=sampler)
    clocks = res.get("clocks")
    nV = res["vertices"]
    ms_per_step = res["ms_per_step"]
    # e2e: slab upload + setup + one step + owned field download through the same Python API, host arrays in / host field out
    import time
    from .device import DIRICHLET_SYMMETRIC, VEC_X, VEC_XOLD
    del system
    part.grid._device = None
    torch.cuda.synchronize(); dist.barrier()
    props = np.array([[PROPS["Body"]["Conductivity"], PROPS["Body"]["Density"], PROPS["Body"]["HeatCapacity"], PROPS["Body"]["HeatGeneration"]]])
    t0 = time.perf_counter()
    sys2 = make_system(part, 1)
    sys2.set_preconditioner(args.precond)
    sys2.upload(VEC_XOLD, np.zeros(sys2.rows)); sys2.upload(VEC_X, np.zeros(sys2.rows))
    send_heat_bcs(sys2, part, BC3D)
    sys2.assemble_dirichlet_mode(DIRICHLET_SYMMETRIC)
    sys2.heat_assemble(props, DT, transient)
    sys2.build_rhs(transient)
    it0, rr0 = sys2.solve_pcg(RTOL, 20000)
    T = sys2.download(VEC_X)[:part.n_owned]
    _lib.check(L.efv_synchronize())
    e2e_s = _allreduce_max(time.perf_counter() - t0, world)
    g = part.grid
    h2d = g.coords.nbytes + g.conn.nbytes + g.elem_ptr.nbytes + g.region_of_element.nbytes + g.facet_conn.nbytes + g.facet_ptr.nbytes + 2 * g.numberOfVertices * 8
    hb = torch.tensor([float(h2d), float(part.n_owned * 8)], device="cuda", dtype=torch.float64)
    dist.all_reduce(hb)
    del sys2
    part.grid._device = None
    nested = {}
    if not args.no_c5 and not strong:
        nested["c5_steady"] = c5_steady(args, rank, world, BC3D, PROPS, DT, RTOL, stream)
    nested["cross_check"] = cross_check_small(rank, world, BC3D, PROPS, DT, RTOL, args.precond, stream)
    if rank == 0:
        # aggregate algorithmic bytes of one solver iteration over the ranks: Jacobi-PCG 12 nnz + 112 rows; multigrid-PCG
        # 5 passes over the matrix on the finest level + the coarser levels
        peak = _peak_gbs()
        line = {
            "metric": "heat_transfer_3d_dof_per_s", "value": nV / (ms_per_step / 1e3), "unit": "DOF/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if strong else "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": (f"C5 strong scaling: 3D heat conduction, structured hex mesh {n}^3 vertices cut into {world} "
                                    "vertex-block slabs" if strong else
                                    (f"C2 weak scaling: each of the {world} ranks owns a {n}x{n}x{n}-vertex block of an {n}x{n}x{nz} hex box "
                                     "(same spacing, same per-GPU work as the N=1 C2 run)" if stack else
                                     f"C2 weak scaling: the C2 cube refined to {n}^3 vertices ({nV / world / 1e6:.2f} M per GPU, N=1: 216^3 = 10.08 M), "
                                     f"cut into {world} z-slabs")) +
                                   (f", dt=1e-2, 3-D BC set, PCG({args.precond}) rtol 1e-10, 1 time step per step from T0 = 0" if transient else
                                    f", STEADY, 3-D BC set, PCG({args.precond}) rtol 1e-10 from x0 = 0, one full assemble + solve per step"),
                       "vertices": nV, "pcg_iterations_per_step": res["iterations"], "warmup_iterations": res["warmup_iterations"],
                       "iterations_total": res["iterations_total"], "ms_per_iteration": res["ms_per_iteration"],
                       "assemble_ms_per_step": res["assemble_ms"], "solve_ms_per_step": res["solve_ms"],
                       "solve_dof_per_s": nV / (res["solve_ms"] * 1e-3), "time_to_solution_s_per_step": ms_per_step * 1e-3,
                       "setup_ms": res["setup_ms"], "true_relres": res["true_relres"], "field_sum": res["field_sum"],
                       "field_sum_of_squares": res["field_sum_of_squares"], "amg": res.get("amg"),
                       "host_slab_generation_s": res["host_slab_generation_s"],
                       "host_mesh_outside_e2e": "every rank generates its slab arrays on the host before the timed regions",
                       "halo": ("peer-memory stores over NVLink (CUDA IPC) + peer-memory all-reduce" if _peer_ok else "NCCL send/recv + ncclAllReduce") + ", one n^2 plane per neighbour per SpMV (and per multigrid level)",
                       "l2_policy": "inputs larger than L2"},
            "roofline": {"bound": "hbm", "kernel": "k_spmv_sell_tma on the finest level, all ranks (every solver iteration passes over the matrix "
                                                   + ("5 times: 1 CG product + 4 fused Chebyshev / residual passes of the V-cycle)" if args.precond == "amg" else "once)"),
                         "achieved": None, "peak": peak * world, "unit": "GB/s", "frac": None, "traffic": None,
                         "note": "the per-kernel roofline (bytes moved / launch time / measured peak) is measured in the N = 1 line; "
                                 "here: see ms_per_iteration against the N = 1 value"},
            "e2e": {"value": nV / e2e_s, "unit": "DOF/s", "h2d_bytes_per_step": int(hb[0].item()),
                    "d2h_bytes_per_step": int(hb[1].item()), "ms_per_step": 1e3 * e2e_s,
                    "what": "per-rank slab upload + setup + assembly + one solve + owned field download"},
            "gpu_launches": res["launches"],
            "clocks": clocks,
        }
        line.update(nested)
        print(json.dumps(line))
    dist.barrier()
    L.efv_comm_destroy()
    dist.destroy_process_group()
