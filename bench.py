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
  roofline     : the SpMV kernel (dominant): bytes moved per launch / CUDA-event time (frac), and the algorithmic
                 12*nnz + 24*rows figure beside it (frac_algorithmic)
  cpu_baseline : the unmodified reference (kind "reference") or, without its sources, the oracle port, on a bounded sample

`--impl reference` times the reference's own CPU path: the UNMODIFIED reference (baseline/_ref, a git-ignored copy of
/root/reference/{PyEFVLib,apps} made by __graft_entry__.build(), behind the three import shims of oracle/ref_harness.py)
on a bounded sample, with the numpy oracle port beside it; same metric and unit.
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
# CPU legs (the only places where bench.py executes oracle/): the UNMODIFIED reference when its sources are on the box
# (baseline/_ref, copied by __graft_entry__.build(); /root/reference in the build container) and the numpy oracle port
# ---------------------------------------------------------------------------------------------------------------------
def reference_available():
    from oracle import ref_harness as R
    return R.available()


def cpu_port(n):
    from oracle import ref_bench as RB
    RB.port_step(8, PROPS, BC3D, DT)
    r = RB.port_step(n, PROPS, BC3D, DT)
    return {"value": r["vertices"] / r["total_s"], "unit": "DOF/s", "cores": 1, "kind": "port", "n": n, "vertices": r["vertices"],
            "seconds_per_step": r["total_s"],
            "sample": f"structured hex mesh n={n} ({r['vertices']} vertices), 1 transient step: oracle assembly (numpy) + SuperLU "
                      f"factor+solve, {r['total_s']:.2f} s (+ {r['grid_s']:.2f} s geometry); single-threaded like the reference; "
                      f"host has {os.cpu_count()} cores"}


def reference_n_for(seconds_per_step):
    """Largest n whose predicted reference step fits the budget.  Measured law of the reference's assembleMatrix (the
    Dirichlet filter of heat_transfer.py:70-76 rebuilds the whole triplet list once per Dirichlet vertex): t(n) ~ n^5.3,
    1.2 s at n = 6, 11.7 s at n = 9, 38 s at n = 11 (this container, 1 core)."""
    n = 6
    while n < 21 and 1.2 * ((n + 1) / 6.0) ** 5.3 <= seconds_per_step:
        n += 1
    return n


def cpu_reference(n, steps=1):
    """steps passes of the reference's own assembleMatrix + addToIndependentVector + solveLinearSystem on an n^3 mesh."""
    from oracle import ref_bench as RB
    case = RB.ReferenceHeatCase(n, PROPS, BC3D, DT)
    times = [case.step() for _ in range(steps)]
    t = float(np.mean([x["total_s"] for x in times]))
    return {"value": case.vertices / t, "unit": "DOF/s", "cores": 1, "kind": "reference", "n": n, "vertices": case.vertices,
            "seconds_per_step": t, "grid_build_s": case.grid_s, "assemble_s": float(np.mean([x["assemble_s"] for x in times])),
            "rhs_solve_s": float(np.mean([x["rhs_solve_s"] for x in times])),
            "sample": f"UNMODIFIED reference (PyEFVLib + apps/heat_transfer.py behind the 3 import shims of oracle/ref_harness.py) on a "
                      f"structured hex mesh n={n} ({case.vertices} vertices): HeatTransferSolver.assembleMatrix (element loops, Dirichlet "
                      f"filter, sparse inverse) + addToIndependentVector + solveLinearSystem, {t:.1f} s per step; Grid() took "
                      f"{case.grid_s:.1f} s once (not in the step, like the GPU arm's setup); pure Python, 1 core of {os.cpu_count()}"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    total = max(1, args.steps + min(args.warmup, 1))
    budget = 170.0 / total                                       # whole run within a few minutes
    if reference_available():
        n = min(args.cpu_n, reference_n_for(budget)) if args.cpu_n else reference_n_for(budget)
        if args.warmup:
            cpu_reference(5, 1)
        cb = cpu_reference(n, args.steps)
        cb["port_beside_it"] = {k: v for k, v in cpu_port(n).items() if k in ("value", "seconds_per_step", "kind")}
    else:
        n = int(min(args.cpu_n or 30, max(12, 24 * (budget / 3.4) ** (1.0 / 6.0))))
        cb = cpu_port(n)
        cb["sample"] += " (reference sources not on this box: baseline/_ref missing)"
    ms = 1e3 * cb["seconds_per_step"]
    line = {
        "impl": "reference", "metric": "heat_transfer_3d_dof_per_s", "value": cb["value"], "unit": "DOF/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "3D transient heat conduction, structured hex mesh, dt=1e-2, 3-D BC set, 1 time step per step "
                               "(assembly + BCs + RHS + solve): bounded CPU sample of the GPU arm's workload",
                   "n_vertices_per_side": cb["n"], "vertices": cb["vertices"]},
        "cpu_baseline": cb,
        "e2e": {"value": cb["value"], "unit": "DOF/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
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
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
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


def same_config_run(n, cpu, precond="jacobi"):
    """The GPU arm at exactly the size of the CPU sample (one `same_config: true` comparison; at this size the GPU step is
    launch-latency bound, which is why the headline config is not this one)."""
    import pyefvlib_b200 as P
    from pyefvlib_b200.device import DeviceSystem, VEC_X, VEC_XOLD
    grid = P.Grid(P.structured_grid_data(n, "hex"))
    system = DeviceSystem(grid, 1)
    system.set_preconditioner(precond)
    props = np.array([[1.0, 1.0, 1.0, 100.0]])
    system.upload(VEC_XOLD, np.zeros(system.rows)); system.upload(VEC_X, np.zeros(system.rows))
    for _ in range(3):
        hot_path_step(system, grid, props)
    system.upload(VEC_XOLD, np.zeros(system.rows)); system.upload(VEC_X, np.zeros(system.rows))
    from pyefvlib_b200 import _lib
    L = _lib.load()
    _lib.check(L.efv_synchronize())
    t0 = time.perf_counter()
    reps = 20
    for _ in range(reps):
        it, rr = hot_path_step(system, grid, props)
    _lib.check(L.efv_synchronize())
    t = (time.perf_counter() - t0) / reps
    return {"same_config": True, "n": n, "vertices": grid.numberOfVertices, "gpu_ms_per_step": 1e3 * t, "gpu_dof_per_s": grid.numberOfVertices / t,
            "cpu_kind": cpu["kind"], "cpu_dof_per_s": cpu["value"], "gpu_over_cpu": grid.numberOfVertices / t / cpu["value"],
            "true_relres": system.true_residual()}


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
        args.clock_sampler = ClockSampler          # rank 0 samples its GPU's clocks / throttle reasons during the run
        return efd.bench_multi_gpu(args, BC3D, PROPS, DT, RTOL)
    torch.cuda.set_device(local_rank)
    L = _lib.load()
    _lib.check(L.efv_set_device(local_rank))
    stream = torch.cuda.ExternalStream(L.efv_get_stream())
    n = args.c5_n if (args.workload == "c5" and args.n == 216) else args.n
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
    system.set_preconditioner(args.precond)
    e2.record(stream)
    torch.cuda.synchronize()
    upload_ms, setup_ms = e0.elapsed_time(e1), e1.elapsed_time(e2)
    nnz, rows = system.nnz, system.rows
    system.upload(VEC_XOLD, np.zeros(rows))
    system.upload(VEC_X, np.zeros(rows))

    # ---- value: K time steps with resident inputs ------------------------------------------------------------------
    iters_log = []
    # clocks / throttle reasons under this load: nvidia-smi needs ~0.1 s to deliver its first sample and the timed region is
    # ~0.3 s long, so the sampler (every 50 ms) already runs through the warm-up steps, which are the same work
    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(args.warmup):
        iters_log.append(hot_path_step(system, grid, props)[0])
    # the timed steps are time steps 1..K of the configuration (from T0 = 0), not a later sub-window
    system.upload(VEC_XOLD, np.zeros(rows))
    system.upload(VEC_X, np.zeros(rows))
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
    # the solver's own residual is a recurrence; this one is recomputed from b - A x with one extra SpMV
    true_relres = system.true_residual()
    assert true_relres <= 2e-10, f"true residual {true_relres:.3e} of the last timed step exceeds 2e-10"
    field_sums = system.vec_sums(VEC_X)
    amg_info = system.amg_info() if args.precond == "amg" else None

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
    roofline = {"bound": "hbm", "frac_algorithmic": achieved / peak,
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
        # headline fraction = bytes actually moved (ncu DRAM read + write of this kernel at this size) / time / measured peak;
        # frac_algorithmic keeps the SURVEY 8(d) formula, which counts 4 B/nnz of column indices the compressed copy never reads
        roofline["achieved"] = roofline["moved_gbs"]
        roofline["frac"] = roofline["moved_frac"]
        roofline["note"] = ("frac = ncu DRAM bytes per launch / live CUDA-event launch time / measured copy peak; frac_algorithmic "
                            "(12 nnz + 24 rows per launch) exceeds it because the SELL copy replaces 85 % of the column indices "
                            "by one offset per 32 entries (lossless)")

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
        solver = HeatTransferSolver(pd, saverType="none", transient=True, verbosity=False, rtol=RTOL, preconditioner=args.precond)
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

    # ---- the north_star's target case on this GPU count: C5 steady (368^3), strong-scaling base -------------------------
    c5 = None
    if not args.no_c5 and n != args.c5_n:
        from pyefvlib_b200 import distributed as efd
        c5 = efd.c5_steady(args, 0, 1, BC3D, PROPS, DT, RTOL, stream)

    # ---- CPU baseline on a bounded sample (rank 0, N = 1), and the GPU arm once more at exactly that size -------------
    cpu = None
    same = None
    if not args.no_cpu:
        if reference_available():
            n_cpu = min(args.cpu_n, 9) if args.cpu_n else 9           # ~12 s of the reference's pure-Python path
            cpu = cpu_reference(n_cpu, 1)
            cpu["port_beside_it"] = {k: v for k, v in cpu_port(n_cpu).items() if k in ("value", "seconds_per_step", "kind")}
            cpu["port_at_n30"] = {k: v for k, v in cpu_port(30).items() if k in ("value", "seconds_per_step", "kind", "vertices")}
        else:
            n_cpu = args.cpu_n or 30
            cpu = cpu_port(n_cpu)
        same = same_config_run(n_cpu, cpu, args.precond)
    # assembly kernel against the HBM roofline with the three byte counts SURVEY 8(d) asks for side by side
    asm_s = float(np.mean(asm_ms)) * 1e-3
    asm_bytes = {"compulsory (coords + connectivity + region in, values + rhs out)": 24.0 * nV + 36.0 * nE + 8.0 * nnz + 8.0 * rows,
                 "kernel_streams (3 reduced face vectors k J^-T s, sub-volumes, byte slot map, incidence, values)":
                     (288.0 + 64.0 + 64.0 + 32.0) * nE + 8.0 * nnz + 16.0 * rows}
    # bytes the kernel MOVES: ncu dram__bytes_read + dram__bytes_write of one k_heat_lanes<3> launch at n = 216
    # (profiles/r02_assembly_lanes.json); scaled with the element count at other sizes
    asm_moved = 7.74e9 * nE / 9938375.0
    assembly_roofline = {"kernel": "k_heat_lanes<3> (lane-per-row assembly on the pair schedule, operands one step ahead in registers, fused Dirichlet + lumped terms)",
                         "ms": asm_s * 1e3, "peak": peak, "unit": "GB/s", "bound": "hbm",
                         "moved_bytes_per_launch_ncu": asm_moved, "moved_gbs": asm_moved / asm_s / 1e9,
                         "moved_frac": asm_moved / asm_s / 1e9 / peak,
                         "achieved": {k: v / asm_s / 1e9 for k, v in asm_bytes.items()},
                         "frac": {k: v / asm_s / 1e9 / peak for k, v in asm_bytes.items()}}

    line = {
        "metric": "heat_transfer_3d_dof_per_s", "value": value, "unit": "DOF/s", "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{'C5' if n == 368 else 'C2'}: 3D transient heat conduction, structured hex mesh {n}^3 vertices, dt=1e-2, 3-D BC set, "
                               f"PCG({args.precond}) rtol 1e-10, 1 time step per step (assembly + BCs + RHS + solve)",
                   "vertices": nV, "elements": nE, "nnz": nnz, "pcg_iterations_per_step": timed_iters,
                   "iterations_total": int(np.sum(timed_iters)), "ms_per_iteration": float(np.sum(solve_ms)) / max(1, int(np.sum(timed_iters))),
                   "amg": amg_info, "true_relres": true_relres, "field_sum": field_sums[0], "field_sum_of_squares": field_sums[1],
                   "timed_steps": f"time steps 1..{args.steps} from T0 = 0 (fields reset after the {args.warmup} warm-up steps)",
                   "host_mesh_outside_e2e": "the structured mesh arrays are generated on the host before the timed regions "
                                            "(host_mesh_generation_s); e2e starts from those host arrays in pinned memory",
                   "warmup_iterations": iters_log, "assemble_ms_per_step": float(np.mean(asm_ms)),
                   "solve_ms_per_step": float(np.mean(solve_ms)), "solve_dof_per_s": nV / (float(np.mean(solve_ms)) * 1e-3),
                   "time_to_solution_s_per_step": ms_per_step * 1e-3, "setup_ms": setup_ms, "mesh_upload_ms": upload_ms,
                   "host_mesh_generation_s": host_mesh_s, "l2_policy": "inputs (matrix of 12 B/nnz) larger than the 126 MB L2",
                   "assembled_elements_per_s": nE / (float(np.mean(asm_ms)) * 1e-3)},
        "roofline": roofline,
        "assembly_roofline": assembly_roofline,
        "cpu_baseline": cpu,
        "same_config_as_cpu_baseline": same,
        "c5_steady": c5,
        "e2e": {"value": e2e_value, "unit": "DOF/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "ms_per_step": float(np.mean(e2e_ms)), "ms_runs": [float(x) for x in e2e_ms], "what": "HeatTransferSolver(problemData).solve(), one time step, host arrays in / host field out"},
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
    ap.add_argument("--n", "--mesh-n", dest="n", type=int, default=216, help="vertices per side of the structured mesh (216 = C2)")
    ap.add_argument("--cpu-n", type=int, default=0, help="vertices per side of the bounded CPU sample (0: chosen from the time budget)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--precond", default="amg", choices=["amg", "jacobi"], help="preconditioner of the CG solve")
    ap.add_argument("--no-c5", action="store_true", help="skip the nested C5 steady case (368^3, the north_star target)")
    ap.add_argument("--c5-n", type=int, default=368, help="vertices per side of the nested C5 case")
    ap.add_argument("--steady", action="store_true", help="(N>1 leg) steady problem instead of one transient step")
    ap.add_argument("--workload", default="c2", choices=["c2", "c5"],
                    help="c2: configs[1] (N>1: weak scaling at the per-GPU size of C2); c5: configs[4] 368^3 strong scaling")
    ap.add_argument("--weak-family", default="cube", choices=["cube", "stack"],
                    help="N>1 weak scaling: cube = the C2 cube refined to n N^(1/3) vertices per side; stack = one n^3 block per GPU in a row")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    run_gpu(args)


if __name__ == "__main__":
    main()
