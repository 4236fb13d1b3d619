"""Result writers that work from arrays.

`VtuSaver` writes the same ASCII .vtu text as the reference (`PyEFVLib/VtuSaver.py:10-47`: last saved step of every
field, points with 8 decimals, point data with 15) but formats whole arrays at once instead of looping over
Vertex/Element objects.  `VtuSeriesSaver` is the streaming variant the reference lacks: one .vtu per saved step plus
a ParaView .pvd index, keeping only the latest field in memory instead of `Saver.fields` growing by one row per step
(`PyEFVLib/Saver.py:16-22`), so a 10^7-vertex transient run does not hold every step on the host.
"""
import os

import numpy as np

from .solver import Saver

# VTK cell types (`VtuSaver.py:11`): triangle 5, quadrilateral 9, tetrahedron 10, hexahedron 12, prism 13, pyramid 14
_VTK_2D = {3: 5, 4: 9}
_VTK_3D = {4: 10, 8: 12, 6: 13, 5: 14}


def _cell_arrays(grid):
    ptr = np.asarray(grid.elem_ptr, dtype=np.int64)
    lens = np.diff(ptr)
    table = _VTK_2D if grid.dimension == 2 else _VTK_3D
    types = np.zeros(len(lens), dtype=np.int64)
    for m, code in table.items():
        types[lens == m] = code
    if (types == 0).any():
        raise Exception("element with an unsupported number of vertices for the .vtu writer")
    return np.asarray(grid.conn, dtype=np.int64), ptr[1:], types


def _join_ints(a):
    return "".join(f"{v} " for v in a.tolist())


def write_vtu(path, grid, fields):
    """fields: {name: ndarray[nV]}; text layout of `VtuSaver.finalize` (`VtuSaver.py:24-45`)."""
    conn, offsets, types = _cell_arrays(grid)
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    t4, t5 = "\t" * 4, "\t" * 5
    with open(path, "w") as f:
        f.write("<?xml version=\"1.0\"?>\n<VTKFile type=\"UnstructuredGrid\" version=\"0.1\" byte_order=\"LittleEndian\">\n"
                "\t<UnstructuredGrid>\n")
        f.write(f"\t\t<Piece NumberOfPoints=\"{grid.numberOfVertices}\" NumberOfCells=\"{grid.numberOfElements}\">\n"
                "\t\t\t<Points>\n")
        f.write(f"{t4}<DataArray type=\"Float64\" Name=\"Points\" NumberOfComponents=\"3\" format=\"ascii\">\n")
        f.write("".join(f"{t5}{c:.8f}\n" for c in np.asarray(grid.coords, dtype=np.float64).ravel().tolist()))
        f.write(f"{t4}</DataArray>\n\t\t\t</Points>\n\t\t\t<Cells>\n")
        for name, arr in (("connectivity", conn), ("offsets", offsets), ("types", types)):
            f.write(f"{t4}<DataArray type=\"Int32\" Name=\"{name}\" format=\"ascii\">\n")
            f.write(t5 + _join_ints(arr) + "\n")
            f.write(f"{t4}</DataArray>\n")
        f.write("\t\t\t</Cells>\n\t\t\t<PointData>\n")
        for name, values in fields.items():
            f.write(f"{t4}<DataArray type=\"Float64\" Name=\"{name}\" format=\"ascii\">\n")
            f.write(t5 + "".join(f"{d:.15f} " for d in np.asarray(values, dtype=np.float64).tolist()) + "\n")
            f.write(f"{t4}</DataArray>\n")
        f.write("\t\t\t</PointData>\n\t\t</Piece>\n\t</UnstructuredGrid>\n</VTKFile>\n")


class VtuSaver(Saver):
    """Reference behaviour: buffer every step, write the last one at finalize (`VtuSaver.py:10-47`)."""

    def __init__(self, grid, outputPath, basePath, fileName="Results", **kwargs):
        Saver.__init__(self, grid, outputPath, basePath, "vtu", fileName)

    def finalize(self):
        if self.outputPath and self.fields:
            if os.path.isfile(self.outputPath):
                os.remove(self.outputPath)
            write_vtu(self.outputPath, self.grid, {k: v[-1] for k, v in self.fields.items()})
        self.finalized = True


class VtuSeriesSaver(Saver):
    """One `<fileName>_<step>.vtu` per saved time plus `<fileName>.pvd`; memory = the latest step only."""

    def __init__(self, grid, outputPath, basePath, fileName="Results", **kwargs):
        Saver.__init__(self, grid, outputPath, basePath, "pvd", fileName)
        self._pending, self._pending_time, self._files = {}, None, []

    def _flush(self):
        if not self._pending or not self.outputPath:
            return
        name = f"{self.fileName}_{len(self._files):05d}.vtu"
        write_vtu(os.path.join(os.path.dirname(self.outputPath), name), self.grid, self._pending)
        self._files.append((self._pending_time, name))
        self._pending = {}

    def save(self, fieldName, fieldValues, currentTime):
        if self._pending_time is not None and currentTime != self._pending_time:
            self._flush()
        if not self.timeSteps.size or self.timeSteps[-1] != currentTime:
            self.timeSteps = np.append(self.timeSteps, currentTime)
        self._pending_time = currentTime
        self._pending[fieldName] = np.array(fieldValues, dtype=np.float64)
        self.fields[fieldName] = self._pending[fieldName][None, :]        # latest step only

    def finalize(self):
        self._flush()
        if self.outputPath and self._files:
            with open(self.outputPath, "w") as f:
                f.write("<?xml version=\"1.0\"?>\n<VTKFile type=\"Collection\" version=\"0.1\" byte_order=\"LittleEndian\">\n"
                        "\t<Collection>\n")
                for t, name in self._files:
                    f.write(f"\t\t<DataSet timestep=\"{float(t)!r}\" group=\"\" part=\"0\" file=\"{name}\"/>\n")
                f.write("\t</Collection>\n</VTKFile>\n")
        self.finalized = True
