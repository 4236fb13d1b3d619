"""Sweep the (chunk, stages, warps) configurations of the TMA-staged SELL SpMV on the C2 matrix."""
import os, sys, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pyefvlib_b200 as P
from pyefvlib_b200 import _lib
from pyefvlib_b200.device import DeviceSystem, VEC_X, VEC_B

n = int(sys.argv[1]) if len(sys.argv) > 1 else 216
kind = sys.argv[2] if len(sys.argv) > 2 else "hex"
L = _lib.load()
grid = P.structured_grid(n, kind)
s = DeviceSystem(grid, 1)
s.heat_assemble(np.array([[1.0, 1.0, 1.0, 100.0]]), 1e-2, True)
s.upload(VEC_X, np.random.default_rng(0).random(s.rows))
stream = torch.cuda.ExternalStream(L.efv_get_stream())
bytes_ = 12.0 * s.nnz + 24.0 * s.rows
os.environ["EFV_SPMV_FORMAT"] = "csr"
s.spmv(VEC_X, VEC_B)
ycsr = s.download(VEC_B)
os.environ.pop("EFV_SPMV_FORMAT")
names = {0: "ch16 st2 w24", 6: "ch16 st3 w16", 1: "ch32 st2 w12", 4: "ch8 st4 w24", 5: "ch8 st3 w32"}
out = {}
for delta in (1, 0):
    os.environ["EFV_SELL_DELTA"] = str(delta)
    os.environ.pop("EFV_SELL_KERNEL", None)
    s.heat_assemble(np.array([[1.0, 1.0, 1.0, 100.0]]), 1e-2, True)
    for cfg in list(names) + ["reg"]:
        if cfg == "reg":
            os.environ["EFV_SELL_KERNEL"] = "reg"
        else:
            os.environ["EFV_SELL_TMA"] = str(cfg)
        for _ in range(3):
            s.spmv(VEC_X, VEC_B)
        torch.cuda.synchronize()
        err = float(np.abs(s.download(VEC_B) - ycsr).max() / np.abs(ycsr).max())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(20):
            s.spmv(VEC_X, VEC_B)
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        out[f"delta{delta} " + names.get(cfg, "reg U8")] = round(ms, 4)
        print(f"delta={delta} {names.get(cfg, 'reg U8'):14s} {ms:.4f} ms  {bytes_/ms/1e6:8.1f} GB/s (algorithmic bytes)  err vs CSR {err:.1e}", flush=True)
print(json.dumps(out))
