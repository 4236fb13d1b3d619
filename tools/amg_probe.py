"""Jacobi-PCG vs multigrid-PCG on the bench problem: iterations, solve time, numeric set-up time, hierarchy."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import pyefvlib_b200 as P  # noqa: E402
from pyefvlib_b200.device import DeviceSystem, VEC_X, VEC_XOLD, DIRICHLET_SYMMETRIC  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 216
transient = (sys.argv[2] if len(sys.argv) > 2 else "t") == "t"
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
grid = P.Grid(P.structured_grid_data(n, "hex"))
props = np.array([[1.0, 1.0, 1.0, 100.0]])
out = {"n": n, "transient": transient}
for pc in (("amg",) if os.environ.get("PROBE_ONLY_AMG") else ("jacobi", "amg")):
    s = DeviceSystem(grid, 1)
    s.set_preconditioner(pc)
    s.upload(VEC_XOLD, np.zeros(s.rows)); s.upload(VEC_X, np.zeros(s.rows))
    res = []
    for k in range(steps):
        bench.send_bcs(s, grid)
        s.assemble_dirichlet_mode(DIRICHLET_SYMMETRIC)
        s.heat_assemble(props, bench.DT, transient)
        s.build_rhs(transient)
        if not transient:
            s.upload(VEC_X, np.zeros(s.rows))
        t0 = time.perf_counter()
        it, rr = s.solve_pcg(1e-10, 20000)
        wall = time.perf_counter() - t0
        tm = s.timings()
        res.append(dict(iters=it, relres=rr, solve_ms=tm["solve_ms"], wall_ms=1e3 * wall, true_relres=s.true_residual()))
        s.advance()
    out[pc] = res
    if pc == "amg":
        out["amg_info"] = s.amg_info()
    del s
print(json.dumps(out))
