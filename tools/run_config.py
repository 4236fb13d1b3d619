"""Run BASELINE.json configs[2] (C3: 3-D elasticity, 5M-vertex tet mesh) or configs[3] (C4: coupled Biot, 2M-vertex hex
mesh) on one GPU and print one JSON line with sizes, iteration counts and timings (SURVEY.md section 8d set-ups)."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pyefvlib_b200 as P
from pyefvlib_b200 import _lib
from pyefvlib_b200.device import DeviceSystem, DIRICHLET_SYMMETRIC, DIRICHLET_ROWS, VEC_X, VEC_XOLD

which = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else (171 if which == "c3" else 126)
L = _lib.load()
sync = lambda: _lib.check(L.efv_synchronize())
N0 = ("neumann", 0.0)
t0 = time.perf_counter()
grid = P.structured_grid(n, "tet" if which == "c3" else "hex")
host_s = time.perf_counter() - t0
nV = grid.numberOfVertices
t0 = time.perf_counter()
if which == "c3":
    # apps/stress_equilibrium.py:191-218 extended to 3-D: u=0 on West/East, v=0 on South, w=0 on Bottom, v-traction -1e4 on North
    bcs = {"u": {"West": ("dirichlet", 0.0), "East": ("dirichlet", 0.0)}, "v": {"South": ("dirichlet", 0.0), "North": ("neumann", -1e4)},
           "w": {"Bottom": ("dirichlet", 0.0)}}
    s = DeviceSystem(grid, 3)
    sync(); setup_s = time.perf_counter() - t0
    s.bc_clear()
    for b in grid.boundaries:
        for i, var in enumerate("uvw"):
            kind, val = bcs[var].get(b.name, N0)
            if kind == "neumann":
                s.bc_neumann(b.handle, i, val, -1.0)
        for i, var in enumerate("uvw"):
            kind, val = bcs[var].get(b.name, N0)
            if kind == "dirichlet":
                s.bc_dirichlet(b.verticesIndexes + i * nV, np.full(len(b.verticesIndexes), val))
    s.set_preconditioner(os.environ.get("PRECOND", "amg"))
    s.assemble_dirichlet_mode(DIRICHLET_SYMMETRIC)
    s.elasticity_assemble(np.array([[6.0e6, 0.4, 1800.0, 0.0]]), False)
    first_assemble_ms = s.timings()["assemble_ms"]        # includes the one-time face records of the block tile kernel
    s.elasticity_assemble(np.array([[6.0e6, 0.4, 1800.0, 0.0]]), False)
    s.build_rhs(False)
    s.upload(VEC_X, np.zeros(s.rows))
    it, rr = s.solve_pcg(1e-10, 100000, negate=True)
    tm = s.timings()
    x = s.download(VEC_X)
    out = dict(config="C3 elasticity tets", n=n, vertices=nV, elements=grid.numberOfElements, rows=s.rows, nnz=s.nnz,
               pcg_iterations=it, relres=rr, assemble_ms=tm["assemble_ms"], first_assemble_ms=first_assemble_ms, solve_ms=tm["solve_ms"], setup_s=setup_s,
               host_mesh_s=host_s, v_min=float(x[nV:2 * nV].min()), true_relres=s.true_residual(negate=True),
               preconditioner=os.environ.get("PRECOND", "amg"), amg=(s.amg_info() if os.environ.get("PRECOND", "amg") == "amg" else None), dof_per_s=s.rows / (1e-3 * (tm["assemble_ms"] + tm["solve_ms"])),
               iteration_GBs=(s.graph_nnz * (72 + 4) + s.rows * 112) * it / (tm["solve_ms"] * 1e-3) / 1e9)
else:
    props = np.array([[0.2, 6.0e9, 0.0, 0.19, 1.9e-15, 3.0303e-10, 1000.0, 2700.0, 1000.0]])     # apps/geomechanics.py:292-304
    bcs = {"u_x": {"West": ("dirichlet", 0.0), "East": ("dirichlet", 0.0)}, "u_y": {"South": ("dirichlet", 0.0), "North": ("neumann", 1e5)},
           "u_z": {"Bottom": ("dirichlet", 0.0)}, "p": {"North": ("dirichlet", 0.0)}}
    names = ["u_x", "u_y", "u_z", "p"]
    s = DeviceSystem(grid, 4)
    sync(); setup_s = time.perf_counter() - t0
    s.bc_clear()
    for i, var in enumerate(names):
        for b in grid.boundaries:
            kind, val = bcs[var].get(b.name, N0)
            if kind == "neumann":
                s.bc_neumann(b.handle, i, val, 1.0)
    for i, var in enumerate(names):
        for b in grid.boundaries:
            kind, val = bcs[var].get(b.name, N0)
            if kind == "dirichlet":
                s.bc_dirichlet(b.verticesIndexes + i * nV, np.full(len(b.verticesIndexes), val))
    s.set_preconditioner(os.environ.get("PRECOND", "amg"))
    s.assemble_dirichlet_mode(DIRICHLET_ROWS)
    s.biot_assemble(props, 20.0)
    first_asm = s.timings()["assemble_ms"]                  # includes the one-time pair schedule and face records
    s.biot_assemble(props, 20.0)
    asm = s.timings()["assemble_ms"]
    s.upload(VEC_XOLD, np.zeros(s.rows)); s.upload(VEC_X, np.zeros(s.rows))
    its, ms = [], []
    for step in range(5):
        s.build_rhs(True)
        it, rr = s.solve_bicgstab(1e-10, 20000)
        its.append(it); ms.append(s.timings()["solve_ms"])
        assert rr <= 1e-10, rr
        s.advance()
    x = s.download(VEC_X)
    out = dict(config="C4 Biot hex", n=n, vertices=nV, elements=grid.numberOfElements, rows=s.rows, nnz=s.nnz, bicgstab_iterations=its,
               assemble_ms=asm, first_assemble_ms=first_asm, solve_ms_per_step=ms, setup_s=setup_s, host_mesh_s=host_s, p_max=float(x[3 * nV:].max()),
               uy_min=float(x[nV:2 * nV].min()), dof_per_s=s.rows * 5 / (1e-3 * (asm + sum(ms))), true_relres=s.true_residual(),
               preconditioner=os.environ.get("PRECOND", "amg"))
print(json.dumps(out))
