"""Solver scaffold: the reference's template method (PyEFVLib/Solver.py:10-86) plus a `backend` switch.

`backend="gpu"` is the only implementation shipped here -- the CUDA path through include/efv_b200.h.  The hook
names (init / mainloop / assembleMatrix / addToIndependentVector / solveLinearSystem / saveIterationResults /
checkConvergence), printed lines and result attributes are the reference's."""
import os
import time

import numpy as np


class Saver:
    """Buffers every saved field like PyEFVLib/Saver.py:16-22; CSV output like CsvSaver.py:9-41.
    (Output formatting is outside the hot path; only the in-memory `fields` contract matters to the apps.)"""
    wants_fields = True          # the time loop downloads a field from the device only for savers that ask for it

    def __init__(self, grid, outputPath, basePath, extension="csv", fileName="Results", **kwargs):
        self.grid, self.fileName, self.extension, self.basePath = grid, fileName, extension, basePath
        self.outputPath = os.path.join(outputPath, f"{fileName}.{extension}") if outputPath else None
        self.timeSteps = np.array([])
        self.fields = dict()
        self.finalized = False

    def save(self, fieldName, fieldValues, currentTime):
        if not self.timeSteps.size or self.timeSteps[-1] != currentTime:
            self.timeSteps = np.append(self.timeSteps, currentTime)
        if fieldName not in self.fields:
            self.fields[fieldName] = np.zeros((0, self.grid.numberOfVertices))
        self.fields[fieldName] = np.vstack([self.fields[fieldName], fieldValues])

    def finalize(self):
        if self.outputPath and self.extension == "csv" and self.fields:
            os.makedirs(os.path.dirname(self.outputPath), exist_ok=True)
            labels = ["X", "Y", "Z"] + [f"TimeStep{t + 1} - {name}" for t in range(self.timeSteps.size) for name in self.fields]
            cols = [self.grid.coords[:, 0], self.grid.coords[:, 1], self.grid.coords[:, 2]]
            cols += [self.fields[name][t] for t in range(self.timeSteps.size) for name in self.fields if t < len(self.fields[name])]
            with open(self.outputPath, "w") as f:
                f.write("\"{}\"\n".format("\",\"".join(labels)))
                for line in np.array(cols).T:
                    f.write("{}\n".format(",".join(str(n) for n in line)))
        self.finalized = True


CsvSaver = Saver


class NullSaver(Saver):
    """Keeps only the latest field (large meshes: the reference's keep-everything buffer would not fit).  The time loop
    does not download anything for it; the solver hands it the final field once, when the run ends."""
    wants_fields = False

    def save(self, fieldName, fieldValues, currentTime):
        self.fields[fieldName] = np.asarray(fieldValues)[None, :]

    def finalize(self):
        self.finalized = True


class Solver:
    def __init__(self, problemData, extension="csv", saverType="default", transient=True, verbosity=False, backend="gpu"):
        if backend != "gpu":
            raise Exception("pyefvlib_b200 only ships the CUDA backend (backend='gpu'); the CPU path is the reference itself")
        self.problemData = problemData
        self.extension = extension
        self.transient = transient
        self.verbosity = verbosity
        self.saverType = saverType
        self.backend = backend
        self.grid = self.problemData.grid

    def solve(self):
        self.settings()
        self.printHeader()
        self.init()
        self.mainloop()
        self.finalizeSaver()
        self.printFooter()

    def settings(self):
        self.initialTime = time.time()
        self.propertyData = self.problemData.propertyData
        self.outputPath = self.problemData.outputFilePath
        if self.saverType == "none":
            self.saver = NullSaver(self.grid, self.outputPath, self.problemData.libraryPath)
        elif self.saverType in ("stream", "series"):
            from .savers import VtuSeriesSaver
            self.saver = VtuSeriesSaver(self.grid, self.outputPath, self.problemData.libraryPath)
        elif self.extension == "vtu":
            from .savers import VtuSaver
            self.saver = VtuSaver(self.grid, self.outputPath, self.problemData.libraryPath)
        elif self.extension == "csv":
            self.saver = Saver(self.grid, self.outputPath, self.problemData.libraryPath, extension=self.extension)
        else:
            raise Exception(f"Output extension '{self.extension}' is not supported by the GPU backend (csv, vtu, or saverType "
                            "'stream' / 'none'); the reference's xdmf / vtm writers need meshio")
        self.numberOfVertices = self.grid.numberOfVertices
        self.dimension = self.grid.dimension
        self.currentTime = 0.0
        self.timeStep = self.problemData.timeStep
        self.tolerance = self.problemData.tolerance
        self.iteration = 0
        self.converged = False

    def printHeader(self):
        if self.verbosity:
            for key, path in zip(["output", "grids"], [self.problemData.outputFilePath, self.problemData.meshFilePath]):
                print("\t{}\n\t\t{}\n".format(key, path))
            print("\tsolid")
            for region in self.grid.regions:
                print("\t\t{}".format(region.name))
                colSize = len(max(self.problemData.propertyData.properties, key=lambda w: len(w)))
                for propertyName in self.problemData.propertyData.properties:
                    print(f"\t\t{propertyName:>{colSize}} : {self.problemData.propertyData.get(region.handle, propertyName)}")
                print("")
            print("\n{:>9}\t{:>14}\t{:>14}\t{:>14}".format("Iteration", "CurrentTime", "TimeStep", "Difference"))

    def printFooter(self):
        if self.verbosity:
            print("Ended Simultaion, elapsed {:.2f}s".format(time.time() - self.initialTime))
            if self.saver.outputPath:
                print("\n\tresult: ", end="")
                print(os.path.realpath(self.saver.outputPath), "\n")

    def printIterationData(self):
        if self.verbosity:
            print("{:>9}\t{:>14.2e}\t{:>14.2e}\t{:>14.2e}".format(self.iteration, self.currentTime, self.timeStep, self.difference))

    def finalizeSaver(self):
        self.saver.finalize()
