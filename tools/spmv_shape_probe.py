"""Plain SpMV time per row on structured hex boxes of the same size and different plane sizes (one GPU): the x gathers of a
row come from three planes that are n^2 entries apart."""
import os, sys, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pyefvlib_b200 as P
from pyefvlib_b200 import _lib
from pyefvlib_b200.device import DeviceSystem, VEC_X, VEC_B

L = _lib.load()
stream = torch.cuda.ExternalStream(L.efv_get_stream())
out = []
import gc
for n, nz in [(int(a), int(b)) for a, b in (p.split("x") for p in (sys.argv[1].split(",") if len(sys.argv) > 1 else "216x216,272x136,343x86,432x54,608x27".split(",")))]:
    grid = P.structured_grid(n, "hex", nz=nz)
    s = DeviceSystem(grid, 1)
    s.heat_assemble(np.array([[1.0, 1.0, 1.0, 100.0]]), 1e-2, True)
    s.upload(VEC_X, np.random.default_rng(0).random(s.rows))
    for _ in range(3):
        s.spmv(VEC_X, VEC_B)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(20):
        s.spmv(VEC_X, VEC_B)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    out.append(dict(n=n, nz=nz, rows=s.rows, spmv_ms=ms, ns_per_1000_rows=1e6 * ms / (s.rows / 1000.0)))
    print(out[-1], flush=True)
    del s
    grid._device = None
    del grid
    gc.collect()
print(json.dumps(out))
