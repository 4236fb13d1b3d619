"""Worker of tests/test_gpu_distributed.py (launched with torch.distributed.run, one rank per GPU): partitioned heat
solve, checked against the CPU oracle on the global mesh."""
import os
import sys

import numpy as np
import scipy.sparse as sp
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import pyefvlib_b200 as P  # noqa: E402
from pyefvlib_b200 import _lib, distributed as efd  # noqa: E402
from pyefvlib_b200.device import DIRICHLET_ROWS, DIRICHLET_SYMMETRIC, VEC_X, VEC_XOLD, VEC_B  # noqa: E402
from oracle import efv_oracle as O  # noqa: E402
from tests.cases import BC3D, HEAT_PROPS  # noqa: E402
from tests import util  # noqa: E402


def main():
    kind = sys.argv[1] if len(sys.argv) > 1 else "hex"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    rank, world, local_rank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    # EFV_TEST_SHARED_GPU=1: every rank runs on cuda:0 (a 1-GPU box).  NCCL refuses duplicate devices, so the library's
    # communicator is created in peer-memory-only mode (CUDA IPC between the processes) and torch uses gloo.
    shared = os.environ.get("EFV_TEST_SHARED_GPU", "0") == "1"
    if shared:
        local_rank = 0
    torch.cuda.set_device(local_rank)
    _lib.check(_lib.load().efv_set_device(local_rank))
    if shared:
        dist.init_process_group("gloo")
        efd.init_comm(use_nccl=False)
    else:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        efd.init_comm()
    tdev = "cpu" if shared else "cuda"
    part = efd.structured_part(n, kind, rank, world, perturb=0.0)
    s = efd.make_system(part, 1)
    props = np.array([[1.0, 1.0, 1.0, 100.0]])
    dt = 1e-2
    nV = n ** 3
    Told_global = 300.0 + 40.0 * np.sin(0.01 * np.arange(nV))
    # oracle on the global mesh (every rank computes it; small)
    og = O.OracleGrid(O.structured_grid_data(n, kind))
    Aref = sp.csr_matrix(O.heat_matrix(og, HEAT_PROPS, BC3D, transient=True, dt=dt))
    bref = O.heat_rhs(og, HEAT_PROPS, BC3D, transient=True, dt=dt, Told=Told_global)
    xref = O.solve_direct(Aref, bref)
    l2g = part.local_to_global
    own = l2g[:part.n_owned]
    # 1. owned rows of the reference-form matrix and RHS
    efd.send_heat_bcs(s, part, BC3D)
    s.assemble_dirichlet_mode(DIRICHLET_ROWS)
    s.heat_assemble(props, dt, True)
    s.upload(VEC_XOLD, Told_global[l2g])
    s.build_rhs(True)
    A = s.to_scipy()[:part.n_owned]
    sub = Aref[own][:, l2g]
    assert util.row_rel_matrix_error(A, sub) <= 1e-12, "owned rows differ from the global matrix"
    assert util.rel_max(s.download(VEC_B)[:part.n_owned], bref[own]) <= 1e-12
    # 2. distributed PCG on the symmetric form
    s.assemble_dirichlet_mode(DIRICHLET_SYMMETRIC)
    s.heat_assemble(props, dt, True)
    s.build_rhs(True)
    s.upload(VEC_X, Told_global[l2g])
    iters, relres = s.solve_pcg(1e-10, 5000)
    assert relres <= 1e-10, (iters, relres)
    x = s.download(VEC_X)[:part.n_owned]
    err = np.array([np.sum((x - xref[own]) ** 2), np.sum(xref[own] ** 2)])
    t = torch.tensor(err, device=tdev, dtype=torch.float64)
    dist.all_reduce(t)
    rel = float(torch.sqrt(t[0] / t[1]).item())
    assert rel <= 1e-8, rel
    # 2a. the same solve with the aggregation multigrid preconditioner (ghost aggregates, per-level halos, gathered coarsest level)
    s.set_preconditioner("amg")
    s.upload(VEC_X, Told_global[l2g])
    it_a, rr_a = s.solve_pcg(1e-10, 500)
    assert rr_a <= 1e-10 and s.true_residual() <= 2e-10, (it_a, rr_a)
    xa = s.download(VEC_X)[:part.n_owned]
    err = np.array([np.sum((xa - xref[own]) ** 2), np.sum(xref[own] ** 2)])
    t = torch.tensor(err, device=tdev, dtype=torch.float64)
    dist.all_reduce(t)
    rel_a = float(torch.sqrt(t[0] / t[1]).item())
    assert rel_a <= 1e-8 and it_a < iters, (rel_a, it_a, iters)
    s.set_preconditioner("jacobi")
    # 2b. distributed BiCGStab on the reference (row-replaced, non-symmetric) form: halo before both SpMVs of an
    #     iteration, per-block partials all-reduced in place
    s.assemble_dirichlet_mode(DIRICHLET_ROWS)
    s.heat_assemble(props, dt, True)
    s.build_rhs(True)
    s.upload(VEC_X, Told_global[l2g])
    it_b, rr_b = s.solve_bicgstab(1e-10, 5000)
    assert rr_b <= 1e-10, (it_b, rr_b)
    xb = s.download(VEC_X)[:part.n_owned]
    err = np.array([np.sum((xb - xref[own]) ** 2), np.sum(xref[own] ** 2)])
    t = torch.tensor(err, device=tdev, dtype=torch.float64)
    dist.all_reduce(t)
    rel_b = float(torch.sqrt(t[0] / t[1]).item())
    assert rel_b <= 1e-8, rel_b
    # 3. max|x - xold| is a global reduction
    d = s.max_abs_diff()
    assert abs(d - np.abs(xref - Told_global).max()) <= 1e-6 * np.abs(xref - Told_global).max()
    # 4. partitioned linear elasticity (3 DOF per vertex: block halo, block SELL SpMV, negated PCG)
    N0 = ("neumann", 0.0)
    ebc = {"u": {"West": ("dirichlet", 0.0), "East": ("dirichlet", 0.0), "South": N0, "North": N0, "Bottom": N0, "Top": N0},
           "v": {"West": N0, "East": N0, "South": ("dirichlet", 0.0), "North": ("neumann", -1e4), "Bottom": N0, "Top": N0},
           "w": {"West": N0, "East": N0, "South": N0, "North": N0, "Bottom": ("dirichlet", 0.0), "Top": N0}}
    eprops = {"Body": {"Density": 1800.0, "PoissonsRatio": 0.4, "ShearModulus": 6.0e6, "Gravity": 10.0}}
    Ae, be = O.elasticity_system(og, eprops, ebc, gravity=True)
    xe = O.solve_equilibrated(Ae, be)
    se = efd.make_system(part, 3)
    nl = part.grid.numberOfVertices
    se.bc_clear()
    for h, b in enumerate(part.grid.boundaries):
        for i, var in enumerate("uvw"):
            bkind, val = ebc[var][b.name]
            if bkind == "neumann":
                se.bc_neumann(b.handle, i, val, -1.0)
        for i, var in enumerate("uvw"):
            bkind, val = ebc[var][b.name]
            if bkind == "dirichlet":
                idx = part.boundary_vertices[h]
                se.bc_dirichlet(idx + i * nl, np.full(len(idx), val))
    se.assemble_dirichlet_mode(DIRICHLET_SYMMETRIC)
    se.elasticity_assemble(np.array([[6.0e6, 0.4, 1800.0, 10.0]]), True)
    se.build_rhs(False)
    se.upload(VEC_X, np.zeros(se.rows))
    it_e, rr_e = se.solve_pcg(1e-10, 50000, negate=True)
    assert rr_e <= 1e-10, (it_e, rr_e)
    se.set_preconditioner("amg")
    se.upload(VEC_X, np.zeros(se.rows))
    it_ea, rr_ea = se.solve_pcg(1e-10, 5000, negate=True)
    assert rr_ea <= 1e-10 and it_ea < it_e, (it_ea, it_e, rr_ea)
    xl = se.download(VEC_X).reshape(3, nl)[:, :part.n_owned]
    xg = xe.reshape(3, nV)[:, own]
    t = torch.tensor([np.sum((xl - xg) ** 2), np.sum(xg ** 2)], device=tdev, dtype=torch.float64)
    dist.all_reduce(t)
    rel_e = float(torch.sqrt(t[0] / t[1]).item())
    assert rel_e <= 1e-8, rel_e
    if rank == 0:
        print(f"DIST_OK world={world} shared_gpu={int(shared)} kind={kind} n={n} iters={iters} relres={relres:.2e} rel_l2={rel:.2e} "
              f"amg iters={it_a} rel_l2={rel_a:.2e} elasticity iters={it_e} (amg {it_ea}) rel_l2={rel_e:.2e}")
    dist.barrier()
    _lib.load().efv_comm_destroy()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
