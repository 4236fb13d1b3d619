"""Time the heat assembly kernels (EFV_ASSEMBLY = lanes (default) | rows) on a structured mesh and check that they agree."""
import os, sys, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pyefvlib_b200 as P
from pyefvlib_b200.device import DeviceSystem, DIRICHLET_SYMMETRIC
import bench

n = int(sys.argv[1]) if len(sys.argv) > 1 else 216
kind = sys.argv[2] if len(sys.argv) > 2 else "hex"
grid = P.structured_grid(n, kind)
s = DeviceSystem(grid, 1)
props = np.array([[1.0, 1.0, 1.0, 100.0]])
bench.send_bcs(s, grid)
s.assemble_dirichlet_mode(DIRICHLET_SYMMETRIC)
ref = None
out = {}
modes = tuple(sys.argv[3].split(",")) if len(sys.argv) > 3 else ("rows", "lanes")
for mode in modes:   # "lanes" = anything else = the default lane-per-row kernel
    os.environ["EFV_ASSEMBLY"] = mode
    ts = []
    for _ in range(4):
        s.heat_assemble(props, 1e-2, True)
        ts.append(s.timings()["assemble_ms"])
    vals = s.values()
    if ref is None:
        ref = vals
    err = float(np.abs(vals - ref).max() / np.abs(ref).max())
    assert err < 1e-15, (mode, err)
    out[mode] = dict(ms=min(ts[1:]), max_rel_diff_vs_first=err)
    print(mode, ts, err, flush=True)
nE = grid.numberOfElements
print(json.dumps(dict(n=n, kind=kind, elements=nE, **out)))
