#!/usr/bin/env python
"""bench.py -- BASELINE.json metric: 3-D heat-transfer DOF/s (assemble + PCG solve) on B200.

Workload (BASELINE.json configs[1], SURVEY.md section 8d "C2"): transient heat conduction on a synthetic
structured hexahedral mesh of n^3 vertices (default n = 216 -> 10,077,696 vertices, 9,938,375 elements,
269,586,136 non-zeros), k = rho = cp = 1, q = 100, dt = 1e-2, T0 = 0, 3-D boundary set West=300 / East=400 /
Bottom=500 Dirichlet, Top Neumann 1000, South/North Neumann 0, PCG to ||r||/||b|| <= 1e-10, fp64.

A "step" is one pass of the hot path = one transient time step: fused assembly of the matrix (north_star kernel
3), Dirichlet/Neumann kernels (4), RHS, Jacobi-PCG solve (5) from the previous field, advance.  The one-time
geometry pass (kernel 1) and CSR + slot-map build (kernel 2) are reported as `setup_ms` and are inside `e2e`.

  value : DOF/s with the mesh resident in HBM (vertices solved per second, whole job)
  e2e   : same metric through the public API (HeatTransferSolver on host arrays): H2D of the mesh, geometry,
          CSR build, assembly, one time-step solve, D2H of the field -- everything inside the timed region
  roofline     : the SpMV kernel (dominant): algorithmic bytes 12*nnz + 24*rows per launch / CUDA-event time
  cpu_baseline : the oracle port (numpy assembly + SuperLU solve = what the reference computes) on a bounded sample

`--impl reference` times the CPU oracle port alone (the reference is pure Python and cannot travel to the GPU
box), same metric and unit.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BC3D = {"West": ("dirichlet", 300.0), "East": ("dirichlet", 400.0), "Bottom": ("dirichlet", 500.0),
        "Top": ("neumann", 1000.0), "South": ("neumann", 0.0), "North": ("neumann", 0.0)}
PROPS = {"Body": {"HeatCapacity": 1.0, "Conductivity": 1.0, "Density": 1.0, "HeatGeneration": 100.0}}
DT = 1e-2
RTOL = 1e-10


# ---------------------------------------------------------------------------------------------------------------------
# CPU leg: oracle port (numpy restatement of the reference assembly + SuperLU, i.e. the reference's direct solve)
# ---------------------------------------------------------------------------------------------------------------------
def cpu_step(n):
    """One full pass of the reference's path on an n^3-vertex sample: Grid geometry, assembleMatrix
    (apps/heat_transfer.py:40-87, with the sparse inverse replaced by its LU factorisation),
    addToIndependentVector + solveLinearSystem for one time step.  Returns (seconds, dof)."""
    from oracle import efv_oracle as O
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    t0 = time.perf_counter()
    gd = O.structured_grid_data(n, "hex")
    og = O.OracleGrid(gd)
    A = O.heat_matrix(og, PROPS, BC3D, transient=True, dt=DT)
    lu = spla.splu(sp.csc_matrix(A))
    b = O.heat_rhs(og, PROPS, BC3D, transient=True, dt=DT, Told=np.zeros(og.numberOfVertices))
    T = lu.solve(b)
    assert np.isfinite(T).all()
    return time.perf_counter() - t0, og.numberOfVertices


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # bounded sample: SuperLU on a 3-D 27-point operator costs ~ n^6; keep steps * t(n) within ~150 s
    n = int(min(args.cpu_n, max(12, 24 * (150.0 / (3.4 * max(1, args.steps))) ** (1.0 / 6.0))))
    for _ in range(args.warmup):
        cpu_step(max(8, n // 2))
    times = []
    dof = 0
    for _ in range(args.steps):
        t, dof = cpu_step(n)
        times.append(t)
    ms = 1e3 * float(np.mean(times))
    value = dof / (ms / 1e3)
    sample = f"structured hex mesh n={n} ({dof} vertices), 1 transient step: oracle geometry+assembly (numpy) + SuperLU factor+solve"
    line = {
        "impl": "reference", "metric": "heat_transfer_3d_dof_per_s", "value": value, "unit": "DOF/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "3D transient heat conduction, structured hex mesh, 1 time step per step (bounded CPU sample)",
                   "n_vertices_per_side": n, "vertices": dof},
        "cpu_baseline": {"value": value, "unit": "DOF/s", "cores": 1, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "DOF/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1])); mx.append(float(p[2]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if p[5 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------------------------------
# GPU legs
# ---------------------------------------------------------------------------------------------------------------------
def send_bcs(system, grid):
    system.bc_clear()
    for b in grid.boundaries:
        kind, val = BC3D[b.name]
        if kind == "neumann":
            system.bc_neumann(b.handle, 0, val, 1.0)
    for b in grid.boundaries:
        kind, val = BC3D[b.name]
        if kind == "dirichlet":
            system.bc_dirichlet(b.verticesIndexes, np.full(len(b.verticesIndexes), val))


def hot_path_step(system, grid, props):
    """assemble + BCs + RHS + PCG + advance for one time step (everything on the device)."""
    from pyefvlib_b200.device import DIRICHLET_SYMMETRIC
    send_bcs(system, grid)
    system.assemble_dirichlet_mode(DIRICHLET_SYMMETRIC)     # Dirichlet treatment fused into the assembly kernel
    system.heat_assemble(props, DT, True)
    system.build_rhs(True)
    iters, relres = system.solve_pcg(RTOL, 20000)
    system.advance()
    return iters, relres


def pinned_like(a):
    import torch
    t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    return t.numpy(), t


def run_gpu(args):
    import torch
    import torch.distributed as dist
    import pyefvlib_b200 as P
    from pyefvlib_b200 import _lib
    from pyefvlib_b200.device import DeviceSystem, VEC_X, VEC_XOLD, VEC_B
    from pyefvlib_b200.apps.heat_transfer import HeatTransferSolver

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        from pyefvlib_b200 import distributed as efd
        return efd.bench_multi_gpu(args, BC3D, PROPS, DT, RTOL)
    torch.cuda.set_device(local_rank)
    L = _lib.load()
    _lib.check(L.efv_set_device(local_rank))
    stream = torch.cuda.ExternalStream(L.efv_get_stream())
    n = 368 if (args.workload == "c5" and args.n == 216) else args.n
    t_host = time.perf_counter()
    gd = P.structured_grid_data(n, "hex")
    keep = []
    for attr in ("verticesCoordinates", "elementsConnectivities", "boundariesConnectivities"):
        arr, t = pinned_like(getattr(gd, attr))
        setattr(gd, attr, arr)
        keep.append(t)
    grid = P.Grid(gd)
    host_mesh_s = time.perf_counter() - t_host
    nV, nE = grid.numberOfVertices, grid.numberOfElements
    props = np.array([[1.0, 1.0, 1.0, 100.0]])

    def ev():
        return torch.cuda.Event(enable_timing=True)

    # ---- setup (one-time kernels 1 and 2) --------------------------------------------------------------------------
    launches0 = L.efv_kernel_launch_count()
    e0, e1, e2 = ev(), ev(), ev()
    torch.cuda.synchronize()
    e0.record(stream)
    mesh = grid.device                       # H2D + incidence + boundary matching
    e1.record(stream)
    mesh.build_geometry()                    # kernel 1
    system = DeviceSystem(grid, 1)           # kernel 2
    e2.record(stream)
    torch.cuda.synchronize()
    upload_ms, setup_ms = e0.elapsed_time(e1), e1.elapsed_time(e2)
    nnz, rows = system.nnz, system.rows
    system.upload(VEC_XOLD, np.zeros(rows))
    system.upload(VEC_X, np.zeros(rows))

    # ---- value: K time steps with resident inputs ------------------------------------------------------------------
    iters_log = []
    for _ in range(args.warmup):
        iters_log.append(hot_path_step(system, grid, props)[0])
    sampler = ClockSampler(local_rank)
    sampler.start()
    torch.cuda.synchronize()
    launches1 = L.efv_kernel_launch_count()
    s0, s1 = ev(), ev()
    s0.record(stream)
    timed_iters, solve_ms, asm_ms = [], [], []
    for _ in range(args.steps):
        it, rr = hot_path_step(system, grid, props)
        assert rr <= RTOL, f"PCG did not converge: {rr}"
        timed_iters.append(it)
        tm = system.timings()
        solve_ms.append(tm["solve_ms"]); asm_ms.append(tm["assemble_ms"])
    s1.record(stream)
    torch.cuda.synchronize()
    launches2 = L.efv_kernel_launch_count()
    total_ms = s0.elapsed_time(s1)
    ms_per_step = total_ms / args.steps
    value = nV / (ms_per_step / 1e3)

    # ---- roofline of the dominant kernel (SpMV), timed live with CUDA events on the library stream -----------------
    reps = 30
    for _ in range(3):
        system.spmv(VEC_X, VEC_B)
    torch.cuda.synchronize()
    r0, r1 = ev(), ev()
    r0.record(stream)
    for _ in range(reps):
        system.spmv(VEC_X, VEC_B)
    r1.record(stream)
    torch.cuda.synchronize()
    clocks = sampler.stop()
    spmv_ms = r0.elapsed_time(r1) / reps
    spmv_bytes = 12.0 * nnz + 24.0 * rows
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(peaks_path):
        peak = float(json.load(open(peaks_path))["hbm_gbs"]); peak_src = "measured (MEASURED_PEAKS.json)"
    else:
        peak = 6650.0; peak_src = "fallback (B200_PROFILING.md)"
    achieved = spmv_bytes / (spmv_ms * 1e-3) / 1e9
    it_bytes = 12.0 * nnz + 112.0 * rows
    pcg_gbs = it_bytes * float(np.sum(timed_iters)) / (float(np.sum(solve_ms)) * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "spmv_traffic_bytes.json")
    if os.path.isfile(tpath):
        try:
            tj = json.load(open(tpath))
            if int(tj.get("n", -1)) == n:
                traffic = tj["dram_bytes_per_launch"]
        except Exception:
            traffic = None
    roofline = {"bound": "hbm",
                "kernel": "k_spmv_sell_tma (SpMV on the solver's sliced-ELL copy of the CSR matrix, values staged by "
                          "TMA bulk copies, + fused p.Ap partial)",
                "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": spmv_bytes, "launch_ms": spmv_ms,
                "pcg_iteration_gbs": pcg_gbs, "pcg_iteration_frac": pcg_gbs / peak,
                "pcg_algorithmic_bytes_per_iteration": it_bytes}
    roofline["spec_peak_gbs"] = 8000.0          # report against both the 8 TB/s spec and the measured copy peak
    roofline["frac_of_spec"] = achieved / 8000.0
    if traffic:
        # the SELL copy replaces the 32 column indices of a slice column by one offset wherever they are row + const
        # (lossless), so the kernel moves fewer bytes than the CSR-based algorithmic count; this is the fraction of
        # the measured peak that the bytes actually moved (ncu dram read+write) amount to
        roofline["moved_gbs"] = traffic / (spmv_ms * 1e-3) / 1e9
        roofline["moved_frac"] = roofline["moved_gbs"] / peak
        roofline["moved_frac_of_spec"] = roofline["moved_gbs"] / 8000.0
        roofline["note"] = ("frac > 1 is index compression, not skipped work: algorithmic bytes count 4 B/nnz of "
                            "column indices that the compressed SELL copy does not read; moved_frac uses ncu DRAM bytes")

    # ---- e2e: public API, host buffers in, host field out, everything timed ----------------------------------------
    del system
    grid._device = None
    del mesh
    e2e_steps = max(1, min(args.steps, 3))
    Tfinal = None
    h2d = grid.coords.nbytes + grid.conn.nbytes + grid.elem_ptr.nbytes + grid.region_of_element.nbytes + \
        grid.facet_conn.nbytes + grid.facet_ptr.nbytes + grid.boundary_ptr.nbytes + 2 * nV * 8
    d2h = nV * 8
    e2e_ms = []
    for k in range(e2e_steps + 1):
        pd = P.ProblemData(meshFilePath=grid, outputFilePath=None,
                           numericalSettings=P.NumericalSettings(timeStep=DT, finalTime=0.5 * DT, tolerance=None, maxNumberOfIterations=None),
                           propertyData=P.PropertyData(PROPS),
                           boundaryConditions=P.BoundaryConditions({"temperature": dict(
                               {"InitialValue": 0.0}, **{name: {"condition": P.Dirichlet if kind == "dirichlet" else P.Neumann,
                                                                "type": P.Constant, "value": val} for name, (kind, val) in BC3D.items()})}))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        solver = HeatTransferSolver(pd, saverType="none", transient=True, verbosity=False, rtol=RTOL)
        solver.solve()
        Tfinal = solver.temperatureField
        _lib.check(L.efv_synchronize())
        dt_ms = 1e3 * (time.perf_counter() - t0)
        if k > 0:                                  # first pass is the warm-up
            e2e_ms.append(dt_ms)
        del solver
        grid._device = None
    e2e_value = nV / (float(np.mean(e2e_ms)) / 1e3)
    assert np.isfinite(Tfinal).all() and Tfinal.min() > -1e-6 and Tfinal.max() < 1e4

    # ---- CPU baseline on a bounded sample (rank 0, N = 1) ----------------------------------------------------------
    cpu = None
    if not args.no_cpu:
        cpu_step(12)
        t, dof = cpu_step(args.cpu_n)
        cpu = {"value": dof / t, "unit": "DOF/s", "cores": 1, "kind": "port",
               "sample": f"structured hex mesh n={args.cpu_n} ({dof} vertices), 1 transient step: oracle geometry+assembly (numpy) "
                         f"+ SuperLU factor+solve, {t:.1f} s; host has {os.cpu_count()} cores, path is single-threaded like the reference"}

    # assembly kernel against the HBM roofline with the three byte counts SURVEY 8(d) asks for side by side
    asm_s = float(np.mean(asm_ms)) * 1e-3
    asm_bytes = {"compulsory (coords + connectivity + region in, values + rhs out)": 24.0 * nV + 36.0 * nE + 8.0 * nnz + 8.0 * rows,
                 "as_specified (geometry SoA with J^-1 + area vectors + sub-volumes, slot map, values)": 1816.0 * nE,
                 "kernel_streams (3 reduced face vectors k J^-T s, sub-volumes, byte slot map, incidence, values)":
                     (288.0 + 64.0 + 64.0 + 32.0) * nE + 8.0 * nnz + 16.0 * rows}
    assembly_roofline = {"kernel": "k_heat_tiles<3>", "ms": asm_s * 1e3, "peak": peak, "unit": "GB/s",
                         "achieved": {k: v / asm_s / 1e9 for k, v in asm_bytes.items()},
                         "frac": {k: v / asm_s / 1e9 / peak for k, v in asm_bytes.items()}}

    line = {
        "metric": "heat_transfer_3d_dof_per_s", "value": value, "unit": "DOF/s", "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{'C5' if n == 368 else 'C2'}: 3D transient heat conduction, structured hex mesh {n}^3 vertices, dt=1e-2, 3-D BC set, "
                               "PCG(Jacobi) rtol 1e-10, 1 time step per step (assembly + BCs + RHS + solve)",
                   "vertices": nV, "elements": nE, "nnz": nnz, "pcg_iterations_per_step": timed_iters,
                   "warmup_iterations": iters_log, "assemble_ms_per_step": float(np.mean(asm_ms)),
                   "solve_ms_per_step": float(np.mean(solve_ms)), "solve_dof_per_s": nV / (float(np.mean(solve_ms)) * 1e-3),
                   "time_to_solution_s_per_step": ms_per_step * 1e-3, "setup_ms": setup_ms, "mesh_upload_ms": upload_ms,
                   "host_mesh_generation_s": host_mesh_s, "l2_policy": "inputs (matrix of 12 B/nnz) larger than the 126 MB L2",
                   "assembled_elements_per_s": nE / (float(np.mean(asm_ms)) * 1e-3)},
        "roofline": roofline,
        "assembly_roofline": assembly_roofline,
        "cpu_baseline": cpu,
        "e2e": {"value": e2e_value, "unit": "DOF/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "ms_per_step": float(np.mean(e2e_ms)), "what": "HeatTransferSolver(problemData).solve(), one time step, host arrays in / host field out"},
        "gpu_launches": int(launches2 - launches1),
        "clocks": clocks,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=216, help="vertices per side of the structured mesh (216 = C2)")
    ap.add_argument("--cpu-n", type=int, default=30, help="vertices per side of the bounded CPU sample")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--steady", action="store_true", help="(N>1 leg) steady problem instead of one transient step")
    ap.add_argument("--workload", default="c2", choices=["c2", "c5"],
                    help="c2: configs[1] (N>1: weak scaling, one C2 block per GPU); c5: configs[4] 368^3 strong scaling")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    run_gpu(args)


if __name__ == "__main__":
    main()
