"""CPU tests of the .msh reader/writer: the caller-side input format of the hot path (geometry.py:29 -> TVP:27-28)."""
import numpy as np
import pytest

from fem_glass_tempering_b200 import mesh as msh
from fem_glass_tempering_b200.meshio import read_msh, write_msh


def gmsh41_like_line_file(path, m):
    """The file gmsh.write() produces for geometry.py: MSH 4.1 ASCII, 5 point + 4 curve entities, nodes in one block per
    entity (points first, then the curves' interior nodes), only the line elements of physical group "cells"."""
    x = m.x[:, 0]
    corner = [0.0, 5.0, 25.0, 45.0, 50.0]
    cid = [int(np.argmin(np.abs(x - c))) for c in corner]
    tag = {}                                     # vertex -> gmsh node tag: corner points 1..5, interior nodes after
    for k, v in enumerate(cid):
        tag[v] = k + 1
    nxt = 6
    curves = []
    for a, b in zip(cid[:-1], cid[1:]):
        inner = list(range(a + 1, b))
        for v in inner:
            tag[v] = nxt
            nxt += 1
        curves.append((a, b, inner))
    with open(path, "w") as fh:
        fh.write("$MeshFormat\n4.1 0 8\n$EndMeshFormat\n$PhysicalNames\n1\n1 0 \"cells\"\n$EndPhysicalNames\n")
        fh.write("$Entities\n5 4 0 0\n")
        for k, c in enumerate(corner):
            fh.write(f"{k} {c} 0 0 0 \n")
        for k, (a, b, _) in enumerate(curves):
            fh.write(f"{k} {x[a]} 0 0 {x[b]} 0 0 1 0 2 {k} -{k + 1} \n")
        fh.write("$EndEntities\n")
        fh.write(f"$Nodes\n9 {x.size} 1 {x.size}\n")
        for k, v in enumerate(cid):
            fh.write(f"0 {k} 0 1\n{tag[v]}\n{float(x[v])!r} 0 0\n")
        for k, (_, _, inner) in enumerate(curves):
            fh.write(f"1 {k} 0 {len(inner)}\n" + "".join(f"{tag[v]}\n" for v in inner) + "".join(f"{float(x[v])!r} 0 0\n" for v in inner))
        fh.write(f"$EndNodes\n$Elements\n4 {m.n_cells} 1 {m.n_cells}\n")
        e = 1
        for k, (a, b, _) in enumerate(curves):
            fh.write(f"1 {k} 1 {b - a}\n")
            for v in range(a, b):
                fh.write(f"{e} {tag[v]} {tag[v + 1]}\n")
                e += 1
        fh.write("$EndElements\n")


def test_reads_the_msh41_file_the_reference_workflow_writes(tmp_path):
    m = msh.graded_line_mesh()
    p = str(tmp_path / "mesh1d.msh")
    gmsh41_like_line_file(p, m)
    r = read_msh(p)
    assert r.dim == 1 and r.n_cells == 48
    assert np.array_equal(r.cells, m.cells)
    assert np.array_equal(r.x, m.x)            # repr() round-trips float64 exactly


@pytest.mark.parametrize("version", ["4.1", "2.2"])
@pytest.mark.parametrize("dim", [1, 2, 3])
def test_write_read_round_trip(tmp_path, version, dim):
    m = {1: msh.graded_line_mesh(), 2: msh.rectangle_mesh(4, 3, 2.0, 1.5), 3: msh.box_mesh(3, 2, 2, 1.5, 1.0, 0.8)}[dim]
    p = str(tmp_path / f"m{dim}.msh")
    write_msh(p, m, version=version)
    r = read_msh(p)
    assert np.array_equal(r.x, m.x) and np.array_equal(r.cells, m.cells)


def test_create_mesh_writes_gmsh_default_format(tmp_path):
    from fem_glass_tempering_b200 import create_mesh
    p = str(tmp_path / "mesh1d.msh")
    create_mesh(p)
    with open(p) as fh:
        assert fh.read().splitlines()[1].split()[0] == "4.1"      # gmsh.write's default (geometry.py:29)


def test_rejects_what_it_cannot_read(tmp_path):
    p = str(tmp_path / "bin.msh")
    with open(p, "w") as fh:
        fh.write("$MeshFormat\n4.1 1 8\n$EndMeshFormat\n")
    with pytest.raises(NotImplementedError):
        read_msh(p)
    with open(p, "w") as fh:
        fh.write("$MeshFormat\n4 0 8\n$EndMeshFormat\n")
    with pytest.raises(NotImplementedError):
        read_msh(p)
    with open(p, "w") as fh:
        fh.write("hello\n")
    with pytest.raises(ValueError):
        read_msh(p)


def test_missing_mesh_file_raises_instead_of_substituting_a_mesh():
    from fem_glass_tempering_b200.problem import ThermoViscoProblem
    with pytest.raises(FileNotFoundError):
        ThermoViscoProblem._load_mesh("no_such_mesh.msh", None)
    assert ThermoViscoProblem._load_mesh("", None).n_cells == 48      # explicit sentinel: the built-in graded line
