"""Minimal ``meshio`` stand-in: ``read(filename)`` for MSH 2.2 ASCII files,
returning an object with ``points`` (Nv x 3 float64) and ``cells_dict``
({'triangle': (Nt,3) int64, 'line': (Nl,2) int64}) -- the part of meshio the
reference uses (crbe.py:59,63,676).  Registered as ``meshio`` by the top-level
``crbe`` shim only when the real package is not installed."""
from __future__ import annotations

import numpy as np

_NODES_PER_TYPE = {1: ("line", 2), 2: ("triangle", 3), 3: ("quad", 4), 15: ("vertex", 1)}


class Mesh:
    def __init__(self, points, cells_dict):
        self.points = points
        self.cells_dict = cells_dict

    @property
    def cells(self):
        return [(k, v) for k, v in self.cells_dict.items()]

    def __repr__(self):
        return "<meshio stand-in mesh: %d points, %s>" % (
            len(self.points), ", ".join("%d %s" % (len(v), k) for k, v in self.cells_dict.items()))


def read(filename, file_format=None):
    with open(filename) as f:
        tokens = f.read().split("\n")
    it = iter(tokens)
    points = None
    cells = {}
    version = None
    for line in it:
        line = line.strip()
        if line == "$MeshFormat":
            version = next(it).split()[0]
            if not version.startswith("2"):
                raise ValueError(f"meshio stand-in reads MSH 2.x ASCII only, got version {version}")
        elif line == "$Nodes":
            n = int(next(it))
            ids = np.empty(n, dtype=np.int64)
            points = np.empty((n, 3), dtype=np.float64)
            for k in range(n):
                parts = next(it).split()
                ids[k] = int(parts[0])
                points[k] = [float(parts[1]), float(parts[2]), float(parts[3])]
            remap = None
            if not np.array_equal(ids, np.arange(1, n + 1)):
                remap = {int(i): k for k, i in enumerate(ids)}
        elif line == "$Elements":
            m = int(next(it))
            for _ in range(m):
                parts = next(it).split()
                etype, ntags = int(parts[1]), int(parts[2])
                if etype not in _NODES_PER_TYPE:
                    continue
                name, nn = _NODES_PER_TYPE[etype]
                nodes = [int(x) for x in parts[3 + ntags:3 + ntags + nn]]
                cells.setdefault(name, []).append(nodes)
    if points is None:
        raise ValueError(f"{filename}: no $Nodes section")
    out = {}
    for name, rows in cells.items():
        arr = np.asarray(rows, dtype=np.int64)
        if remap is not None:
            arr = np.vectorize(remap.__getitem__)(arr)
        else:
            arr = arr - 1
        out[name] = arr
    return Mesh(points, out)
