"""TEST INFRASTRUCTURE ONLY -- golden output files of the reference's savers (build container only).

Runs the UNMODIFIED reference `VtuSaver` / `CsvSaver` (PyEFVLib/VtuSaver.py, PyEFVLib/CsvSaver.py) on bundled meshes
with a closed-form field and writes their output text to tests/golden/saver_<mesh>.{vtu,csv}; tests/test_savers.py
regenerates the same field and requires pyefvlib_b200's array-based writers to produce the same bytes.
"""
import os
import shutil
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import ref_harness  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(HERE), "tests", "golden")
MESHES = ["2D/Square.msh", "2D/10x10.msh", "3D/Hexas3x3x3.msh", "3D/Tetras.msh", "3D/Prisms.msh", "3D/Pyrams.msh"]


def field(coords, step):
    """Closed-form nodal field (shared with tests/test_savers.py)."""
    x, y, z = coords[:, 0], coords[:, 1], coords[:, 2]
    return 300.0 + 100.0 * x + 7.0 * y - 3.0 * z + (step + 1) * np.sin(3.0 * x + 2.0 * y + z)


def main():
    ref = ref_harness.load()
    P = ref["PyEFVLib"]
    for rel in MESHES:
        tag = os.path.splitext(os.path.basename(rel))[0]
        grid = P.read(ref_harness.mesh_path(rel))
        coords = np.array([v.getCoordinates() for v in grid.vertices], dtype=np.float64)
        tmp = tempfile.mkdtemp()
        for cls, ext in ((P.VtuSaver, "vtu"), (P.CsvSaver, "csv")):
            saver = cls(grid, tmp, "")
            for step in range(2):
                saver.save("temperature", field(coords, step), 0.01 * (step + 1))
            saver.finalize()
            shutil.copy(saver.outputPath, os.path.join(GOLDEN, f"saver_{tag}.{ext}"))
            print("wrote", f"saver_{tag}.{ext}", os.path.getsize(saver.outputPath), "bytes")
        shutil.rmtree(tmp)


if __name__ == "__main__":
    main()
