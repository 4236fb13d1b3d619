"""TEST INFRASTRUCTURE ONLY -- harness that imports the *unmodified* reference from /root/reference.

Imports the reference from /root/reference (build container) or from the git-ignored copy baseline/_ref that
__graft_entry__.build() makes (travels to the GPU box).  Used by oracle/make_golden.py to pin the numpy
restatement (oracle/efv_oracle.py) and generate the committed fixtures in tests/golden/, and by
oracle/ref_bench.py (the CPU arm of bench.py).  Nothing in the product package imports this file.

The three import shims are the ones listed in SURVEY.md Appendix C:
  * np.object alias                   (PyEFVLib/Shape.py:151,152,194,197,198)
  * stub meshio / xmltodict modules   (PyEFVLib/XDMFReader.py:2-4, PyEFVLib/MeshioSaver.py:3)
  * tuple-ify the zip handed to csc_matrix (apps/heat_transfer.py:80, apps/stress_equilibrium.py:171)
"""
import os
import sys
import types

import numpy as np
import scipy.sparse
import scipy.sparse.linalg

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# /root/reference exists only in the build container; baseline/_ref is the git-ignored copy of its PyEFVLib/ and
# apps/ directories that __graft_entry__.build() makes so that the reference arm of bench.py can run on the GPU box
_CANDIDATES = [os.environ.get("EFV_REFERENCE_ROOT", ""), "/root/reference", os.path.join(_REPO, "baseline", "_ref")]
REFERENCE_ROOT = next((c for c in _CANDIDATES if c and os.path.isdir(os.path.join(c, "PyEFVLib"))), "/root/reference")


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "PyEFVLib"))


_loaded = {}


def load():
    """Import the reference package + apps behind the shims; returns a namespace dict."""
    if _loaded:
        return _loaded
    if not available():
        raise RuntimeError("reference not present at %s (only exists in the build container)" % REFERENCE_ROOT)
    os.chdir("/")  # ProblemData.py:152-153 drops the leading "/" of absolute paths
    sys.path[:0] = [REFERENCE_ROOT, os.path.join(REFERENCE_ROOT, "apps")]
    if not hasattr(np, "object"):
        np.object = object
    for m in ("meshio", "xmltodict"):
        if m not in sys.modules:
            sys.modules[m] = types.ModuleType(m)
    import PyEFVLib  # noqa
    import heat_transfer  # noqa
    import stress_equilibrium  # noqa
    import geomechanics  # noqa
    import fixed_stress  # noqa

    class _Sparse:
        linalg = scipy.sparse.linalg

        @staticmethod
        def csc_matrix(arg, **kw):
            if isinstance(arg, tuple) and not isinstance(arg[1], (tuple, list, np.ndarray)):
                arg = (arg[0], tuple(arg[1]))
            return scipy.sparse.csc_matrix(arg, **kw)

    heat_transfer.sparse = stress_equilibrium.sparse = _Sparse
    _loaded.update(PyEFVLib=PyEFVLib, heat_transfer=heat_transfer,
                   stress_equilibrium=stress_equilibrium, geomechanics=geomechanics, fixed_stress=fixed_stress)
    return _loaded


def mesh_path(rel):
    return os.path.join(REFERENCE_ROOT, "meshes", "msh", rel)
