"""gmsh .msh (format 2.2 ASCII) reader/writer for simplicial meshes — what the reference gets from
dolfinx.io.gmshio.read_from_msh (ThermoViscoProblem.py:27-28) and gmsh.write (geometry.py:29)."""
from __future__ import annotations

import numpy as np

from .mesh import Mesh

_GMSH_SIMPLEX = {1: (1, 2), 2: (2, 3), 3: (4, 4)}   # dim -> (element type id, nodes per element)


def write_msh(path: str, mesh: Mesh, physical_name: str = "cells") -> None:
    etype, npe = _GMSH_SIMPLEX[mesh.dim]
    with open(path, "w") as fh:
        fh.write("$MeshFormat\n2.2 0 8\n$EndMeshFormat\n")
        fh.write(f"$PhysicalNames\n1\n{mesh.dim} 0 \"{physical_name}\"\n$EndPhysicalNames\n")
        fh.write(f"$Nodes\n{mesh.n_vertices}\n")
        for i, p in enumerate(mesh.x):
            xyz = [float(v) for v in p] + [0.0] * (3 - mesh.dim)
            fh.write(f"{i + 1} {xyz[0]!r} {xyz[1]!r} {xyz[2]!r}\n")
        fh.write(f"$EndNodes\n$Elements\n{mesh.n_cells}\n")
        for i, c in enumerate(mesh.cells):
            fh.write(f"{i + 1} {etype} 2 0 0 " + " ".join(str(int(v) + 1) for v in c) + "\n")
        fh.write("$EndElements\n")


def read_msh(path: str) -> Mesh:
    """Reads the highest-dimensional simplices of a gmsh 2.2 ASCII file (lower-dimensional elements are
    physical-group markers and are skipped, like gmshio does for the cell mesh)."""
    with open(path) as fh:
        lines = [ln.strip() for ln in fh]
    def section(name):
        a = lines.index(f"${name}") + 1
        b = lines.index(f"$End{name}")
        return lines[a:b]
    fmt = section("MeshFormat")[0].split()
    if not fmt[0].startswith("2"):
        raise NotImplementedError(f"only gmsh format 2.x ASCII is supported (file is {fmt[0]})")
    nodes = section("Nodes")
    ids, xyz = [], []
    for ln in nodes[1:]:
        t = ln.split()
        ids.append(int(t[0]))
        xyz.append([float(t[1]), float(t[2]), float(t[3])])
    ids, xyz = np.array(ids), np.array(xyz)
    remap = np.full(ids.max() + 1, -1, dtype=np.int64)
    remap[ids] = np.arange(ids.size)
    by_dim = {1: [], 2: [], 3: []}
    type_dim = {1: 1, 2: 2, 4: 3}
    for ln in section("Elements")[1:]:
        t = [int(v) for v in ln.split()]
        if t[1] in type_dim:
            by_dim[type_dim[t[1]]].append(t[3 + t[2]:])
    dim = max(d for d, v in by_dim.items() if v)
    cells = remap[np.array(by_dim[dim], dtype=np.int64)]
    used = np.unique(cells)
    renum = np.full(ids.size, -1, dtype=np.int64)
    renum[used] = np.arange(used.size)
    x = xyz[used][:, :dim]
    if dim == 1:                       # gmshio keeps file order; sort 1-D cells left to right for a lattice numbering
        order = np.argsort(x[renum[cells]].mean(axis=1)[:, 0])
        cells = cells[order]
    return Mesh(x, renum[cells])
