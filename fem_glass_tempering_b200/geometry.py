"""Offline equivalent of the reference's gmsh script (/root/reference/geometry.py:3-29): writes the graded
1-D through-thickness mesh as a gmsh 2.2 ASCII .msh file that ThermoViscoProblem(mesh_path=...) reads back."""
from .mesh import graded_line_mesh
from .meshio import write_msh


def create_mesh(path: str):
    m = graded_line_mesh()
    write_msh(path, m, physical_name="cells")
    return m
