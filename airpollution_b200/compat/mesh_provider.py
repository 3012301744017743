"""``create_mesh`` back-ends (reference crbe.py:14-44)."""
from __future__ import annotations

import numpy as np


def _have_gmsh():
    try:
        import gmsh  # noqa: F401
        return hasattr(gmsh, "initialize")
    except Exception:
        return False


def create_mesh(n_points_per_axis=20, domain_size=2.0, filename="square_mesh.msh"):
    if _have_gmsh():
        return _create_mesh_gmsh(n_points_per_axis, domain_size, filename)
    return _create_mesh_structured(n_points_per_axis, domain_size, filename)


def _create_mesh_gmsh(n_points_per_axis, domain_size, filename):
    """The reference's own gmsh recipe: OCC rectangle, uniform target size."""
    import gmsh
    gmsh.initialize()
    try:
        gmsh.model.add("rectangle")
        side = 2 * domain_size
        gmsh.model.occ.addRectangle(-domain_size, -domain_size, 0, side, side)
        gmsh.model.occ.synchronize()
        size = side / (n_points_per_axis - 1)
        gmsh.option.setNumber("Mesh.CharacteristicLengthMin", size)
        gmsh.option.setNumber("Mesh.CharacteristicLengthMax", size)
        gmsh.model.mesh.generate(2)
        gmsh.write(filename)
    finally:
        gmsh.finalize()
    return filename


def _create_mesh_structured(n_points_per_axis, domain_size, filename):
    """gmsh is absent: a structured triangulation with the same target edge
    length ``2*domain_size/(n_points_per_axis-1)``, written as MSH 2.2 ASCII
    (nodes, boundary lines, triangles) so any MSH reader can load it."""
    from ..meshgen import structured_mesh
    n = max(int(n_points_per_axis) - 1, 1)
    mesh = structured_mesh(n, lo=(-domain_size, -domain_size), hi=(domain_size, domain_size))
    write_msh22(filename, mesh.points, mesh.triangles, nx=n, ny=n)
    return filename


def write_msh22(filename, points, triangles, nx=None, ny=None):
    pts = np.asarray(points, dtype=np.float64)
    tri = np.asarray(triangles, dtype=np.int64)
    lines = []
    if nx is not None:  # boundary edges of the structured grid, like gmsh's physical-less line elements
        row = nx + 1
        bottom = [(i, i + 1) for i in range(nx)]
        top = [(ny * row + i, ny * row + i + 1) for i in range(nx)]
        left = [(j * row, (j + 1) * row) for j in range(ny)]
        right = [(j * row + nx, (j + 1) * row + nx) for j in range(ny)]
        lines = bottom + right + top + left
    with open(filename, "w") as f:
        f.write("$MeshFormat\n2.2 0 8\n$EndMeshFormat\n$Nodes\n%d\n" % len(pts))
        for i, p in enumerate(pts):
            f.write("%d %.17g %.17g %.17g\n" % (i + 1, p[0], p[1], p[2] if len(p) > 2 else 0.0))
        f.write("$EndNodes\n$Elements\n%d\n" % (len(lines) + len(tri)))
        eid = 1
        for a, b in lines:
            f.write("%d 1 2 0 1 %d %d\n" % (eid, a + 1, b + 1))
            eid += 1
        for t in tri:
            f.write("%d 2 2 0 1 %d %d %d\n" % (eid, t[0] + 1, t[1] + 1, t[2] + 1))
            eid += 1
        f.write("$EndElements\n")
