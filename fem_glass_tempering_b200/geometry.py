"""Offline equivalent of the reference's gmsh script (/root/reference/geometry.py:3-29): writes the graded
1-D through-thickness mesh as a gmsh .msh file (MSH 4.1 ASCII, gmsh.write's default) that ThermoViscoProblem(mesh_path=...)
reads back."""
from .mesh import graded_line_mesh
from .meshio import write_msh


def create_mesh(path: str):
    m = graded_line_mesh()
    write_msh(path, m, physical_name="cells")
    return m
