"""Synthetic mesh inputs for the CRBE path.

The reference obtains its meshes from gmsh through a ``.msh`` file read back
with meshio (reference crbe.py:14-44, :675-676); ``MeshData`` only ever
touches ``mesh.points`` (Nv x 3) and ``mesh.cells_dict['triangle']`` (Nt x 3)
(crbe.py:59,63).  This module produces objects with exactly those two
attributes: deterministic structured triangulations (the benchmark inputs) and
seeded unstructured ones (the parity-test inputs).  Host/numpy: this is input
construction, not part of the solve.
"""
from __future__ import annotations

import numpy as np


class TriMesh:
    """Duck-typed stand-in for ``meshio.Mesh`` as far as ``MeshData`` reads it."""

    def __init__(self, points, triangles):
        points = np.asarray(points, dtype=np.float64)
        if points.shape[1] == 2:
            points = np.hstack([points, np.zeros((points.shape[0], 1))])
        self.points = points
        self.cells_dict = {"triangle": np.asarray(triangles)}

    @property
    def triangles(self):
        return self.cells_dict["triangle"]


def structured_mesh(nx, ny=None, lo=(-0.5, -0.5), hi=(0.5, 0.5), index_dtype=np.int64,
                    row_range=None, strip=None):
    """Row-major structured triangulation of ``[lo,hi]`` with ``nx*ny`` cells.

    Vertices ``vid = j*(nx+1)+i``; cell (i,j) with ``a=vid(i,j), b=a+1,
    c=a+nx+2, d=a+nx+1`` is split into the counter-clockwise triangles
    ``(a,b,c)`` then ``(a,c,d)``; cells are emitted row by row.  ``row_range``
    = (j0, j1) emits only the triangles of cell rows j0..j1-1 (vertex ids stay
    global) -- used by the strip partitioner.
    """
    ny = nx if ny is None else ny
    xs = np.linspace(lo[0], hi[0], nx + 1)
    ys = np.linspace(lo[1], hi[1], ny + 1)
    if strip is not None:
        # stand-alone mesh of the cell rows [s0, s1) with LOCAL vertex ids; the coordinates are slices of
        # the global linspace so that every vertex has bit-identical coordinates in every strip
        s0, s1 = strip
        ys = ys[s0:s1 + 1]
        ny = s1 - s0
    X, Y = np.meshgrid(xs, ys)
    points = np.stack([X.reshape(-1), Y.reshape(-1), np.zeros(X.size)], axis=1)
    j0, j1 = (0, ny) if row_range is None else row_range
    i = np.arange(nx, dtype=index_dtype)[None, :]
    j = np.arange(j0, j1, dtype=index_dtype)[:, None]
    a = j * (nx + 1) + i
    b = a + 1
    c = a + nx + 2
    d = a + nx + 1
    tri = np.empty((j1 - j0, nx, 2, 3), dtype=index_dtype)
    tri[:, :, 0, 0] = a
    tri[:, :, 0, 1] = b
    tri[:, :, 0, 2] = c
    tri[:, :, 1, 0] = a
    tri[:, :, 1, 1] = c
    tri[:, :, 1, 2] = d
    return TriMesh(points, tri.reshape(-1, 3))


def structured_counts(nx, ny=None):
    """(Nv, Nt, N_dofs, N_boundary, nnz_struct) of :func:`structured_mesh`."""
    ny = nx if ny is None else ny
    nv = (nx + 1) * (ny + 1)
    nt = 2 * nx * ny
    n = 3 * nx * ny + nx + ny
    nb = 2 * (nx + ny)
    nnz = 5 * (n - nb) + 3 * nb
    return nv, nt, n, nb, nnz


def delaunay_mesh(n_points, seed=0, lo=(-1.0, -1.0), hi=(1.0, 1.0), shuffle=True,
                  flip_fraction=0.0):
    """Seeded unstructured triangulation: random interior points plus a
    boundary frame, Delaunay-triangulated (scipy/Qhull).  ``shuffle`` permutes
    the triangle order (the numbering depends on it); ``flip_fraction`` makes
    that share of triangles clockwise (the reference's advection matrix is
    orientation-sensitive, crbe.py:296)."""
    from scipy.spatial import Delaunay

    rng = np.random.default_rng(seed)
    k = max(2, int(np.sqrt(n_points)))
    edge = np.linspace(0.0, 1.0, k + 1)
    frame = np.concatenate([
        np.stack([edge, np.zeros_like(edge)], 1), np.stack([edge, np.ones_like(edge)], 1),
        np.stack([np.zeros_like(edge[1:-1]), edge[1:-1]], 1),
        np.stack([np.ones_like(edge[1:-1]), edge[1:-1]], 1)])
    inner = rng.uniform(0.02, 0.98, size=(n_points, 2))
    pts = np.concatenate([frame, inner])
    pts = np.asarray(lo) + pts * (np.asarray(hi) - np.asarray(lo))
    tri = Delaunay(pts).simplices.astype(np.int64)
    # drop degenerate slivers on the frame
    p = pts
    area = 0.5 * np.abs((p[tri[:, 1], 0] - p[tri[:, 0], 0]) * (p[tri[:, 2], 1] - p[tri[:, 0], 1])
                        - (p[tri[:, 2], 0] - p[tri[:, 0], 0]) * (p[tri[:, 1], 1] - p[tri[:, 0], 1]))
    tri = tri[area > 1e-12]
    if shuffle:
        tri = tri[rng.permutation(len(tri))]
    if flip_fraction > 0:
        flip = rng.random(len(tri)) < flip_fraction
        tri[flip] = tri[flip][:, [0, 2, 1]]
    return TriMesh(pts, tri)
