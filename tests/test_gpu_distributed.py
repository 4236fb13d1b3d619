"""Partitioned (multi-GPU) heat path on real GPUs: needs >= 2 devices, otherwise skipped (the host-side logic is
covered on the CPU by tests/test_distributed_cpu.py)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _run_worker(world, kind, n, shared, port, extra_env=None):
    env = dict(os.environ, EFV_TEST_SHARED_GPU="1" if shared else "0")
    env.update(extra_env or {})
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "dist_gpu_worker.py"), kind, str(n)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
    assert r.returncode == 0 and "DIST_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
    line = [ln for ln in r.stdout.splitlines() if "DIST_OK" in ln][-1]
    out = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out):
        with open(os.path.join(out, "dist_ok.log"), "a") as f:
            f.write(line + "\n")
    return line


@pytest.mark.parametrize("kind,n", [("hex", 20), ("tet", 14)])
def test_partitioned_heat_matches_oracle(kind, n):
    world = min(_n_gpus(), 4)
    if world < 2:
        pytest.skip("needs at least 2 GPUs (the shared-GPU variant below covers the 1-GPU box)")
    _run_worker(world, kind, n, False, 29517)


@pytest.mark.parametrize("kind,n,world,extra", [("hex", 14, 2, {}), ("tet", 10, 3, {}),
                                                # SELL / TMA kernels forced on the small mesh: interior + boundary launches of the split product
                                                ("hex", 16, 2, {"EFV_SPMV_FORMAT": "sell", "EFV_HALO_SPLIT": "1"}),
                                                ("hex", 14, 3, {"EFV_SPMV_FORMAT": "sell", "EFV_HALO_SPLIT": "0", "EFV_HALO_FUSED": "1"}),
                                                # every multigrid level partitioned down to the gathered dense solve, ghosts copied by a kernel
                                                ("hex", 14, 2, {"EFV_AMG_GATHER_ROWS": "0", "EFV_HALO_FUSED": "0"})])
def test_partitioned_heat_matches_oracle_shared_gpu(kind, n, world, extra):
    """The same N>1 parity run with all ranks time-sliced on ONE GPU: partition, ghost assembly, peer-memory halo and
    all-reduce (CUDA IPC), distributed PCG / BiCGStab / elasticity against the oracle.  This is the N>1 correctness
    evidence on a single-GPU box."""
    if _n_gpus() < 1:
        pytest.skip("needs a GPU")
    _run_worker(world, kind, n, True, 29519, extra)
