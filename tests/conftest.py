import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def pytest_sessionstart(session):
    """Test-infrastructure convenience: build the CUDA library if the in-tree .so is missing and nvcc is available
    (the package itself never builds or falls back; see pyefvlib_b200/_lib.py)."""
    import shutil
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    so = os.path.join(root, "pyefvlib_b200", "libefv_b200.so")
    nvcc = shutil.which("nvcc") or ("/usr/local/cuda/bin/nvcc" if os.path.isfile("/usr/local/cuda/bin/nvcc") else None)
    if not os.path.isfile(so) and nvcc:
        subprocess.run(["bash", os.path.join(root, "pyefvlib_b200", "csrc", "build.sh")], check=False)
