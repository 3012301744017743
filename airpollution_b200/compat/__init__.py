"""Stand-ins for the mesh tooling the reference depends on (gmsh, meshio) when
it is not installed: a structured mesher writing the MSH 2.2 container and a
reader returning the ``points`` / ``cells_dict`` view ``MeshData`` consumes
(reference crbe.py:14-44, :59, :63, :675-676).  SURVEY.md section 8f-1."""
