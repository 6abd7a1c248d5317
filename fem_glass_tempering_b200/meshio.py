"""gmsh .msh reader/writer for simplicial meshes (ASCII, formats 4.1 and 2.2) — what the reference gets from
dolfinx.io.gmshio.read_from_msh (ThermoViscoProblem.py:27-28) and gmsh.write (geometry.py:29).

geometry.py calls gmsh.write(path) with default options, which writes MSH 4.1 ASCII: $Entities plus block-structured
$Nodes / $Elements sections, and — because a physical group is defined (geometry.py:23-24) — only the elements of that
group.  Both directions are supported here in 4.1 (default of write_msh, like gmsh) and in the legacy 2.2 layout.
Binary files and the short-lived 4.0 layout are rejected with an explicit error."""
from __future__ import annotations

import numpy as np

from .mesh import Mesh

_GMSH_SIMPLEX = {1: (1, 2), 2: (2, 3), 3: (4, 4)}   # dim -> (element type id, nodes per element)
_TYPE_DIM = {1: 1, 2: 2, 4: 3}                      # first-order line / triangle / tetrahedron
_TYPE_NODES = {1: 2, 2: 3, 4: 4, 15: 1}


def write_msh(path: str, mesh: Mesh, physical_name: str = "cells", version: str = "4.1") -> None:
    if version.startswith("2"):
        return _write_msh2(path, mesh, physical_name)
    if version != "4.1":
        raise ValueError("write_msh: version must be '4.1' or '2.2'")
    d = mesh.dim
    etype, _ = _GMSH_SIMPLEX[d]
    x3 = np.zeros((mesh.n_vertices, 3))
    x3[:, :d] = mesh.x
    lo, hi = x3.min(axis=0), x3.max(axis=0)
    with open(path, "w") as fh:
        fh.write("$MeshFormat\n4.1 0 8\n$EndMeshFormat\n")
        fh.write(f"$PhysicalNames\n1\n{d} 1 \"{physical_name}\"\n$EndPhysicalNames\n")
        # one entity of the mesh dimension carrying the physical group; no bounding entities
        counts = [0, 0, 0, 0]
        counts[d] = 1
        fh.write("$Entities\n" + " ".join(str(c) for c in counts) + "\n")
        box = " ".join(repr(float(v)) for v in (*lo, *hi))
        fh.write(f"1 {box} 1 1 0\n$EndEntities\n")
        fh.write(f"$Nodes\n1 {mesh.n_vertices} 1 {mesh.n_vertices}\n{d} 1 0 {mesh.n_vertices}\n")
        fh.write("\n".join(str(i + 1) for i in range(mesh.n_vertices)) + "\n")
        fh.write("\n".join(f"{p[0]!r} {p[1]!r} {p[2]!r}" for p in x3.tolist()) + "\n")
        fh.write(f"$EndNodes\n$Elements\n1 {mesh.n_cells} 1 {mesh.n_cells}\n{d} 1 {etype} {mesh.n_cells}\n")
        fh.write("\n".join(f"{i + 1} " + " ".join(str(int(v) + 1) for v in c) for i, c in enumerate(mesh.cells.tolist())) + "\n")
        fh.write("$EndElements\n")


def _write_msh2(path: str, mesh: Mesh, physical_name: str) -> None:
    etype, _ = _GMSH_SIMPLEX[mesh.dim]
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


def _sections(path: str) -> dict:
    with open(path, "rb") as fh:
        head = fh.read(64)
    if b"$MeshFormat" not in head:
        raise ValueError(f"{path}: not a gmsh .msh file")
    with open(path, errors="replace") as fh:
        lines = [ln.strip() for ln in fh]
    out, name, start = {}, None, 0
    for i, ln in enumerate(lines):
        if ln.startswith("$End"):
            if name is not None and ln == f"$End{name}":
                out[name] = lines[start:i]
                name = None
        elif ln.startswith("$") and name is None:
            name, start = ln[1:], i + 1
    return out


def _read_nodes_elements_2(sec):
    ids, xyz = [], []
    for ln in sec["Nodes"][1:]:
        t = ln.split()
        ids.append(int(t[0]))
        xyz.append([float(t[1]), float(t[2]), float(t[3])])
    by_dim = {1: [], 2: [], 3: []}
    for ln in sec["Elements"][1:]:
        t = [int(v) for v in ln.split()]
        if t[1] in _TYPE_DIM:
            by_dim[_TYPE_DIM[t[1]]].append(t[3 + t[2]:])
    return np.array(ids), np.array(xyz), by_dim


def _read_nodes_elements_41(sec):
    """MSH 4.1: '$Nodes' = header 'numBlocks numNodes minTag maxTag', then per block 'entityDim entityTag parametric n',
    n node tags, n coordinate lines; '$Elements' = header, then per block 'entityDim entityTag elementType n' and n lines
    'elementTag node...'."""
    L = sec["Nodes"]
    n_blocks, n_nodes = (int(v) for v in L[0].split()[:2])
    ids, xyz, p = [], [], 1
    for _ in range(n_blocks):
        _, _, parametric, n = (int(v) for v in L[p].split())
        p += 1
        ids.extend(int(L[p + i]) for i in range(n))
        p += n
        for i in range(n):
            t = L[p + i].split()
            xyz.append([float(t[0]), float(t[1]), float(t[2])])   # parametric coordinates (if any) follow and are ignored
        p += n
    if len(ids) != n_nodes:
        raise ValueError(f"$Nodes: header announces {n_nodes} nodes, blocks hold {len(ids)}")
    L = sec["Elements"]
    n_blocks = int(L[0].split()[0])
    by_dim, p = {1: [], 2: [], 3: []}, 1
    for _ in range(n_blocks):
        _, _, etype, n = (int(v) for v in L[p].split())
        p += 1
        if etype in _TYPE_DIM:
            for i in range(n):
                by_dim[_TYPE_DIM[etype]].append([int(v) for v in L[p + i].split()[1:]])
        elif etype not in _TYPE_NODES:
            raise NotImplementedError(f"gmsh element type {etype}: only first-order points, lines, triangles and tetrahedra are "
                                      "supported (the reference meshes with gmsh's default order 1)")
        p += n
    return np.array(ids), np.array(xyz), by_dim


def read_msh(path: str) -> Mesh:
    """Reads the highest-dimensional simplices of a gmsh ASCII file, format 4.1 (what geometry.py:29 writes) or 2.2;
    lower-dimensional elements are physical-group markers and are skipped, like gmshio does for the cell mesh."""
    sec = _sections(path)
    fmt = sec["MeshFormat"][0].split()
    if int(fmt[1]) != 0:
        raise NotImplementedError(f"{path}: binary .msh files are not supported (gmsh.write writes ASCII by default)")
    if fmt[0].startswith("2"):
        ids, xyz, by_dim = _read_nodes_elements_2(sec)
    elif fmt[0] == "4.1":
        ids, xyz, by_dim = _read_nodes_elements_41(sec)
    else:
        raise NotImplementedError(f"{path}: gmsh format {fmt[0]} is not supported (4.1 and 2.2 ASCII are)")
    if not any(by_dim.values()):
        raise ValueError(f"{path}: no line, triangle or tetrahedron elements")
    remap = np.full(ids.max() + 1, -1, dtype=np.int64)
    remap[ids] = np.arange(ids.size)
    dim = max(d for d, v in by_dim.items() if v)
    cells = remap[np.array(by_dim[dim], dtype=np.int64)]
    used = np.unique(cells)
    renum = np.full(ids.size, -1, dtype=np.int64)
    renum[used] = np.arange(used.size)
    x = xyz[used][:, :dim]
    cells = renum[cells]
    if dim == 1:                       # gmshio keeps file order; sort 1-D vertices and cells left to right (lattice numbering)
        vorder = np.argsort(x[:, 0], kind="stable")
        inv = np.empty_like(vorder)
        inv[vorder] = np.arange(vorder.size)
        x, cells = x[vorder], np.sort(inv[cells], axis=1)
        cells = cells[np.argsort(cells[:, 0], kind="stable")]
    return Mesh(x, cells)
