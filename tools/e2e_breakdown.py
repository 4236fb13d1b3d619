"""Wall-clock breakdown of the e2e path (host arrays -> field) on the C2 mesh."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pyefvlib_b200 as P
from pyefvlib_b200 import _lib
from pyefvlib_b200.device import DeviceSystem, DeviceMesh, VEC_X, VEC_XOLD, DIRICHLET_SYMMETRIC
import bench

n = int(sys.argv[1]) if len(sys.argv) > 1 else 216
L = _lib.load()
gd = P.structured_grid_data(n, "hex")
keep = []
for attr in ("verticesCoordinates", "elementsConnectivities", "boundariesConnectivities"):
    arr, t = bench.pinned_like(getattr(gd, attr)); setattr(gd, attr, arr); keep.append(t)
grid = P.Grid(gd)
props = np.array([[1.0, 1.0, 1.0, 100.0]])
sync = lambda: _lib.check(L.efv_synchronize())
for rep in range(2):
    grid._device = None
    sync(); t = [time.perf_counter()]
    def lap(name):
        sync(); t.append(time.perf_counter()); print(f"  {name:28s} {1e3*(t[-1]-t[-2]):8.2f} ms", flush=True)
    print(f"rep {rep}")
    mesh = grid.device; lap("DeviceMesh (H2D+incidence+bnd)")
    mesh.build_geometry(); lap("geometry")
    s = DeviceSystem(grid, 1); lap("DeviceSystem (graph+slotmap)")
    z = np.zeros(s.rows); lap("np.zeros")
    s.upload(VEC_XOLD, z); s.upload(VEC_X, z); lap("upload x, xold")
    bench.send_bcs(s, grid); lap("send bcs")
    s.assemble_dirichlet_mode(DIRICHLET_SYMMETRIC); s.heat_assemble(props, 1e-2, True); lap("assemble")
    s.build_rhs(True); lap("rhs")
    it, rr = s.solve_pcg(1e-10, 20000); lap(f"pcg ({it} its)")
    T = s.download(VEC_X); lap("download")
    d = s.max_abs_diff(); lap("max_abs_diff")
    print(f"  total {1e3*(t[-1]-t[0]):.1f} ms")
    del s, mesh
