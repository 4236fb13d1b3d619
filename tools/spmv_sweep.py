"""Sweep SpMV launch variants (lanes per row x unroll) on the C2 matrix; prints GB/s of algorithmic bytes."""
import os, sys, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pyefvlib_b200 as P
from pyefvlib_b200 import _lib
from pyefvlib_b200.device import DeviceSystem, VEC_X, VEC_B, DIRICHLET_SYMMETRIC

n = int(sys.argv[1]) if len(sys.argv) > 1 else 216
kind = sys.argv[2] if len(sys.argv) > 2 else "hex"
L = _lib.load()
grid = P.structured_grid(n, kind)
s = DeviceSystem(grid, 1)
s.heat_assemble(np.array([[1.0, 1.0, 1.0, 100.0]]), 1e-2, True)
s.upload(VEC_X, np.random.default_rng(0).random(s.rows))
stream = torch.cuda.ExternalStream(L.efv_get_stream())
bytes_ = 12.0 * s.nnz + 24.0 * s.rows
out = {}
os.environ["EFV_SPMV_FORMAT"] = "sell"
for delta in (1, 0):
    os.environ["EFV_SELL_DELTA"] = str(delta)
    s.heat_assemble(np.array([[1.0, 1.0, 1.0, 100.0]]), 1e-2, True)      # new values version -> SELL copy rebuilt
    for U in ("tma", 4, 8, 16):
        os.environ["EFV_SELL_KERNEL"] = "tma" if U == "tma" else "reg"
        os.environ["EFV_SELL_UNROLL"] = str(U)
        for _ in range(3):
            s.spmv(VEC_X, VEC_B)
        torch.cuda.synchronize()
        ysell = s.download(VEC_B)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(20):
            s.spmv(VEC_X, VEC_B)
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        out[f"SELL32_delta{delta}_U{U}"] = round(bytes_ / ms / 1e6, 1)
        print(f"SELL-32 delta={delta} unroll={U!s:>3}    {ms:.4f} ms  {bytes_/ms/1e6:8.1f} GB/s (algorithmic bytes)", flush=True)
os.environ["EFV_SELL_DELTA"] = "1"
os.environ.pop("EFV_SELL_UNROLL")
os.environ.pop("EFV_SELL_KERNEL")
os.environ["EFV_SPMV_FORMAT"] = "csr"
s.spmv(VEC_X, VEC_B)
ycsr = s.download(VEC_B)
print("max |y_sell - y_csr| / max|y| =", float(np.abs(ysell - ycsr).max() / np.abs(ycsr).max()))
for lanes, unroll in [(8, 4), (4, 8), (16, 2)]:
    os.environ["EFV_SPMV_LANES"] = str(lanes); os.environ["EFV_SPMV_UNROLL"] = str(unroll)
    for _ in range(3):
        s.spmv(VEC_X, VEC_B)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(20):
        s.spmv(VEC_X, VEC_B)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    out[f"L{lanes}_U{unroll}"] = round(bytes_ / ms / 1e6, 1)
    print(f"lanes={lanes:2d} unroll={unroll}  {ms:.4f} ms  {bytes_/ms/1e6:8.1f} GB/s", flush=True)
print(json.dumps(out))
