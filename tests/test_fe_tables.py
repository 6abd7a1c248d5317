"""CPU: host-side FE tables / mesh topology of the product against the independent thermal oracle."""
import itertools
import math

import numpy as np
import pytest

import kernel_mirror
from fem_glass_tempering_b200 import fe
from fem_glass_tempering_b200 import mesh as msh
from oracle import thermal_oracle as to
from oracle.visco_oracle import MAIN_PARAMS


def _mesh(dim, n=None):
    if dim == 1:
        return msh.graded_line_mesh() if n is None else msh.interval_mesh(n)
    if dim == 2:
        return msh.rectangle_mesh(*(n or (5, 3)), 5.0, 2.5)
    return msh.box_mesh(*(n or (3, 2, 2)), 1.5, 1.0, 0.8)


def test_graded_line_matches_gmsh_counts():
    m = msh.graded_line_mesh()
    assert m.n_cells == 48                       # 13 + 11 + 11 + 13
    x = m.x[:, 0]
    assert x[0] == 0.0 and x[-1] == 50.0 and np.all(np.diff(x) > 0)
    assert abs(x[1] - x[0] - 0.1) < 0.02 and abs(np.diff(x).max() - 3.0) < 0.5


@pytest.mark.parametrize("dim", [1, 2, 3])
@pytest.mark.parametrize("degree", [0, 1, 2, 3, 4, 5, 10])
def test_simplex_quadrature_exact(dim, degree):
    P, W = fe.simplex_quadrature(dim, degree)
    assert abs(W.sum() - 1 / math.factorial(dim)) < 1e-14
    for e in itertools.product(range(degree + 1), repeat=dim):
        if sum(e) > degree:
            continue
        exact = math.prod(math.factorial(k) for k in e) / math.factorial(dim + sum(e))
        assert abs(np.sum(W * np.prod(P ** np.array(e), axis=1)) - exact) < 1e-14


@pytest.mark.parametrize("fdim,degree", [(1, 2), (1, 4), (2, 2), (2, 4)])
def test_symmetric_facet_rule(fdim, degree):
    B, W = fe.symmetric_facet_rule(fdim, degree)
    assert abs(W.sum() - 1) < 1e-15 and np.allclose(B.sum(axis=1), 1)
    for e in itertools.product(range(degree + 1), repeat=fdim):   # monomials in the first fdim barycentrics
        if sum(e) > degree:
            continue
        exact = math.factorial(fdim) * math.prod(math.factorial(k) for k in e) / math.factorial(fdim + sum(e))
        assert abs(np.sum(W * np.prod(B[:, 1:] ** np.array(e), axis=1)) - exact) < 1e-14
    perms, tab = fe.facet_permutation_table(fdim, B)
    assert len(perms) == math.factorial(fdim + 1)
    for p in range(len(perms)):
        assert sorted(tab[p]) == list(range(B.shape[0]))


@pytest.mark.parametrize("dim", [1, 2, 3])
@pytest.mark.parametrize("degree", [1, 2])
def test_basis_matches_vandermonde_oracle(dim, degree):
    el = fe.LagrangeElement(dim, degree)
    v, g = el.tabulate(el.nodes)
    assert np.allclose(v, np.eye(el.n_ld), atol=1e-14)
    rng = np.random.default_rng(0)
    pts = rng.dirichlet(np.ones(dim + 1), 7)[:, 1:]
    v, g = el.tabulate(pts)
    nb = to.NodalBasis(dim, degree, el.nodes)
    assert np.allclose(v, nb.values(pts), atol=1e-13)
    assert np.allclose(g, nb.grads(pts), atol=1e-12)
    assert np.allclose(v.sum(axis=1), 1) and np.allclose(g.sum(axis=2), 0, atol=1e-13)


@pytest.mark.parametrize("dim", [1, 2, 3])
def test_facet_topology(dim):
    m = _mesh(dim)
    t = fe.facet_topology(m)
    nc = m.n_cells
    for c in range(nc):
        for f in range(dim + 1):
            nb = t.neighbor[c, f]
            if nb >= 0:
                assert t.neighbor[nb, t.nb_facet[c, f]] == c and t.nb_facet[nb, t.nb_facet[c, f]] == f
    geo = fe.cell_geometry(m)
    area = fe.facet_measures(m, geo, t.bnd_cell, t.bnd_facet).sum()
    if dim == 1:
        assert len(t.bnd_cell) == 2
    elif dim == 2:
        assert abs(area - 2 * (5.0 + 2.5)) < 1e-12
    else:
        assert abs(area - 2 * (1.5 * 1.0 + 1.5 * 0.8 + 1.0 * 0.8)) < 1e-12
    assert abs(geo.detJ.sum() / math.factorial(dim) - {1: 50.0, 2: 12.5, 3: 1.2}[dim]) < 1e-10


@pytest.mark.parametrize("dim", [2, 3])
def test_p2_lattice_numbering_equals_generic(dim):
    m = _mesh(dim)
    a = fe.ScalarSpace(m, "CG", 2)
    lat, m.lattice = m.lattice, None
    b = fe.ScalarSpace(m, "CG", 2)
    m.lattice = lat
    assert a.n_nodes == b.n_nodes
    xa, xb = a.tabulate_dof_coordinates(), b.tabulate_dof_coordinates()
    # same node set; the dofmaps agree up to the renumbering
    ren = np.full(a.n_nodes, -1)
    ren[a.dofmap.ravel()] = b.dofmap.ravel()
    assert (ren >= 0).all() and np.allclose(xa, xb[ren])


def _setup(dim, family, degree, n=None):
    m = _mesh(dim, n)
    space = fe.ScalarSpace(m, family, degree)
    tabs = fe.operator_tables(dim, degree)
    geo = fe.cell_geometry(m)
    topo = fe.facet_topology(m)
    orc = to.ThermalOracle(m.x, m.cells, space.dofmap, space.element.nodes, family, degree, MAIN_PARAMS, 0.1)
    return m, space, tabs, geo, topo, orc


CASES = [(1, "CG", 1), (1, "DG", 1), (1, "CG", 2), (1, "DG", 2), (2, "CG", 1), (2, "CG", 2), (2, "DG", 1),
         (2, "DG", 2), (3, "CG", 1), (3, "CG", 2), (3, "DG", 1), (3, "DG", 2)]


@pytest.mark.parametrize("dim,family,degree", CASES)
def test_table_driven_operator_equals_assembled_oracle(dim, family, degree):
    m, space, tabs, geo, topo, orc = _setup(dim, family, degree)
    rng = np.random.default_rng(dim * 10 + degree)
    T = 700 + 100 * rng.random(space.n_nodes)
    Tp = T + rng.random(space.n_nodes)
    x = rng.standard_normal(space.n_nodes)
    J = orc.jacobian(T)
    assert abs(J - J.T).max() < 1e-12 * abs(J).max()
    y = kernel_mirror.jac_apply(space, tabs, geo, topo, MAIN_PARAMS, 0.1, x, T_lin=T)
    yo = J @ x
    assert np.max(np.abs(y - yo)) < 1e-11 * np.max(np.abs(yo))
    r = kernel_mirror.jac_apply(space, tabs, geo, topo, MAIN_PARAMS, 0.1, T, residual=True, T_prev=Tp)
    ro = orc.residual(T, Tp)
    assert np.max(np.abs(r - ro)) < 1e-11 * np.max(np.abs(ro))


@pytest.mark.parametrize("dim,family,degree", [(1, "DG", 1), (2, "CG", 2), (3, "DG", 1)])
def test_oracle_jacobian_is_derivative_of_residual(dim, family, degree):
    m, space, tabs, geo, topo, orc = _setup(dim, family, degree)
    rng = np.random.default_rng(1)
    T = 700 + 100 * rng.random(space.n_nodes)
    Tp = T + 1.0
    v = rng.standard_normal(space.n_nodes)
    eps = 1e-4
    fd = (orc.residual(T + eps * v, Tp) - orc.residual(T - eps * v, Tp)) / (2 * eps)
    assert np.max(np.abs(fd - orc.jacobian(T) @ v)) < 1e-7 * np.max(np.abs(fd))


def test_oracle_steady_state_and_newton():
    """T == T_prev == T_ambient is a fixed point; from T_0 = 800 the Newton iteration converges in a few steps."""
    m, space, tabs, geo, topo, orc = _setup(1, "DG", 1)
    Ta = np.full(space.n_nodes, MAIN_PARAMS["T_ambient"])
    assert np.max(np.abs(orc.residual(Ta, Ta))) < 1e-12
    T0 = np.full(space.n_nodes, 800.0)
    T, its, ok = orc.newton(T0, T0)
    assert ok and 2 <= its <= 6
    assert np.max(np.abs(orc.residual(T, T0))) < 1e-9
    assert T.min() > 600 and T.max() < 800.01 and T[0] < 799.9   # surfaces cool first


@pytest.mark.parametrize("dim,n", [(1, None), (2, (8, 4)), (3, (4, 4, 2))])
def test_dg_operator_is_spd_on_the_benchmark_plates(dim, n):
    """CG needs an SPD Jacobian (TVP:342); check the SIP-DG operator with the reference's penalty 5/h."""
    if dim == 3:
        m = msh.box_mesh(*n, 4 * 1.0, 4 * 1.0, 2 * 1.0)
    elif dim == 2:
        m = msh.rectangle_mesh(*n, 8 * 50 / 408, 4 * 25 / 204)
    else:
        m = msh.graded_line_mesh()
    space = fe.ScalarSpace(m, "DG", 1)
    orc = to.ThermalOracle(m.x, m.cells, space.dofmap, space.element.nodes, "DG", 1, MAIN_PARAMS, 0.1)
    J = orc.jacobian(np.full(space.n_nodes, 800.0)).toarray()
    ev = np.linalg.eigvalsh(0.5 * (J + J.T))
    assert ev.min() > 0, f"indefinite: min eigenvalue {ev.min()}"


@pytest.mark.parametrize("dim", [1, 2, 3])
def test_stability_of_the_dg_time_stepping(dim):
    """Implicit Euler amplifies a mode by 1/lambda(M^-1 J): the scheme is stable iff lambda_min(M^-1 J) >= 1, i.e. iff the
    SIP form is coercive.  The reference's penalty 5.0/CellDiameter (TVP:313-320) is coercive on lines and on the
    right triangles of the 2-D plate but NOT on Kuhn tetrahedra (lambda_min = 0.63 at 1 mm, dt = 0.1: a 3-D DG run with
    the reference's own formulation grows by 1.6x per step and diverges after ~15 steps, whatever the mesh size —
    coercivity is scale invariant).  model_params["sip_penalty"] = 6.0 restores coercivity; the 3-D DG benchmark plate
    uses it (bench.py), every parity test keeps the reference's 5.0."""
    import scipy.linalg as sla
    m = {1: msh.graded_line_mesh(), 2: msh.rectangle_mesh(8, 4, 8.0, 4.0), 3: msh.box_mesh(4, 4, 2, 4.0, 4.0, 2.0)}[dim]
    space = fe.ScalarSpace(m, "DG", 1)

    def lam_min(params):
        orc = to.ThermalOracle(m.x, m.cells, space.dofmap, space.element.nodes, "DG", 1, params, 0.1)
        J = orc.jacobian(np.full(space.n_nodes, 800.0)).toarray()
        return sla.eigh(0.5 * (J + J.T), orc.M.toarray(), eigvals_only=True)[0]

    ref = lam_min(MAIN_PARAMS)
    if dim < 3:
        assert ref >= 1.0 - 1e-9, ref
    else:
        assert ref < 0.9, f"expected the reference penalty to be non-coercive on tetrahedra, lambda_min = {ref}"
        assert lam_min(dict(MAIN_PARAMS, sip_penalty=6.0)) >= 1.0 - 1e-9



@pytest.mark.parametrize("dims", [(3, 5, 9), (2, 8, 8), (4, 3, 2)])
def test_class_uniform_cell_tiles(dims):
    """mesh.py numbers the simplices of every x-column in tiles of 32 elements, Kuhn type by Kuhn type: the same set of
    cells as the plain element-major order, x-columns stay contiguous (slab partition), and 32 consecutive cells of a
    full tile share their shape (one warp of the operator kernels reads one class-table entry)."""
    tiled, plain = msh.box_mesh(*dims), msh.box_mesh(*dims, tile=1)
    assert {tuple(c) for c in tiled.cells.tolist()} == {tuple(c) for c in plain.cells.tolist()}
    nx, ny, nz = dims
    col = ny * nz * 6
    for i in range(nx):
        xs = tiled.x[tiled.cells[i * col:(i + 1) * col]][:, :, 0]
        assert xs.min() >= tiled.x[:, 0].max() * i / nx - 1e-12 and xs.max() <= tiled.x[:, 0].max() * (i + 1) / nx + 1e-12
    geo = fe.cell_geometry(tiled)
    per_col = ny * nz
    for i in range(nx):
        for t0 in range(0, per_col, 32):
            ts = min(32, per_col - t0)
            for k in range(6):
                a = i * col + t0 * 6 + k * ts
                J = geo.Jinv[a:a + ts]
                assert np.allclose(J, J[0], atol=1e-12), "cells of one tile row differ in shape"
    r = msh.rectangle_mesh(5, 40)
    assert r.n_cells == 400 and len({tuple(c) for c in r.cells.tolist()}) == 400


def test_oracle_and_mesher_against_the_hand_evaluated_1d_vectors():
    """The assembled oracle (and the graded-line mesher) against tests/golden/thermal_kat.json: residual and Jacobian-vector
    product of TVP:293-325 on the reference's own mesh, hand-evaluated with closed-form P1 element matrices in plain
    Python (tests/golden/make_thermal_kat.py).  1e-12 relative: the summation orders differ."""
    from helpers import load_thermal_kat
    kat = load_thermal_kat()
    m = msh.graded_line_mesh()
    assert m.n_cells == 48 and np.max(np.abs(m.x[:, 0] - kat["points"])) <= 1e-13
    for c in kat["cases"]:
        space = fe.ScalarSpace(m, c["family"], c["degree"])
        orc = to.ThermalOracle(m.x, m.cells, space.dofmap, space.element.nodes, c["family"], c["degree"], MAIN_PARAMS, kat["dt"])
        F = orc.residual(c["T"], c["T_prev"])
        Jx = orc.jacobian(c["T"]) @ c["x"]
        assert np.max(np.abs(F - c["residual"])) <= 1e-12 * np.max(np.abs(c["residual"])), c["family"]
        assert np.max(np.abs(Jx - c["jac_x"])) <= 1e-12 * np.max(np.abs(c["jac_x"])), c["family"]


def test_oracle_against_the_hand_evaluated_2d_3d_vectors():
    """The assembled oracle against tests/golden/thermal_kat_2d3d.json: the weak form of TVP:293-325 (mass, stiffness,
    Robin + radiation, SIP with the '+' = lower-cell-index convention and h = CellDiameter('+')) hand-evaluated in plain
    Python on small PERTURBED triangle / tetrahedron meshes, DG1 and CG1.  1e-12 relative (summation orders differ)."""
    from helpers import load_thermal_kat_simplex
    from fem_glass_tempering_b200.mesh import Mesh
    kat = load_thermal_kat_simplex()
    assert {(c["dim"], c["family"]) for c in kat["cases"]} == {(2, "DG"), (2, "CG"), (3, "DG"), (3, "CG")}
    for c in kat["cases"]:
        m = Mesh(c["x"], c["cells"])
        space = fe.ScalarSpace(m, c["family"], c["degree"])
        orc = to.ThermalOracle(m.x, m.cells, space.dofmap, space.element.nodes, c["family"], c["degree"], MAIN_PARAMS, kat["dt"])
        F = orc.residual(c["T"], c["T_prev"])
        Jx = orc.jacobian(c["T"]) @ c["v"]
        assert np.max(np.abs(F - c["residual"])) <= 1e-12 * np.max(np.abs(c["residual"])), (c["dim"], c["family"])
        assert np.max(np.abs(Jx - c["jac_x"])) <= 1e-12 * np.max(np.abs(c["jac_x"])), (c["dim"], c["family"])


def test_oracle_and_p2_numbering_against_the_hand_evaluated_cg2_vectors():
    """CG2 (the element of BASELINE configs 2 and 4): the oracle and the product's P2 dof numbering against
    tests/golden/thermal_kat_p2.json (plain-Python P2 basis, Duffy-Gauss integration, TVP:293-306).  1e-12 relative."""
    from helpers import load_thermal_kat_p2
    from fem_glass_tempering_b200.mesh import Mesh
    kat = load_thermal_kat_p2()
    assert [c["dim"] for c in kat["cases"]] == [2, 3]
    for c in kat["cases"]:
        m = Mesh(c["x"], c["cells"])
        space = fe.ScalarSpace(m, "CG", 2)
        assert np.array_equal(space.dofmap, c["dofmap"])        # vertices, then edges sorted by (min, max) vertex
        orc = to.ThermalOracle(m.x, m.cells, space.dofmap, space.element.nodes, "CG", 2, MAIN_PARAMS, kat["dt"])
        F = orc.residual(c["T"], c["T_prev"])
        Jx = orc.jacobian(c["T"]) @ c["v"]
        assert np.max(np.abs(F - c["residual"])) <= 1e-12 * np.max(np.abs(c["residual"])), c["dim"]
        assert np.max(np.abs(Jx - c["jac_x"])) <= 1e-12 * np.max(np.abs(c["jac_x"])), c["dim"]


@pytest.mark.parametrize("dim,n,degree", [(2, (10, 6), 1), (2, (10, 6), 2), (3, (4, 7, 3), 1), (3, (4, 7, 3), 2)])
def test_row_stencil_form_of_the_cg_apply(dim, n, degree):
    """The gather form csrc/stencil.cu builds (numpy emulation, tests/kernel_mirror.py): on a lattice-numbered plate the
    rows fall into a handful of classes and ONE (offset, coefficient) list per class reproduces the scattered cell
    matrices to rounding; with a scrambled numbering the classes do not repeat (the library then keeps the cell kernel)."""
    import kernel_mirror
    from fem_glass_tempering_b200 import mesh as msh
    m = msh.plate_mesh(dim, n, tuple(float(k) for k in n))
    space = fe.ScalarSpace(m, "CG", degree)
    tabs, geo = fe.operator_tables(dim, degree), fe.cell_geometry(m)
    A = kernel_mirror.cell_matrices(space, tabs, geo, MAIN_PARAMS, 0.1)
    gkey = np.round(np.concatenate([geo.Jinv.reshape(m.n_cells, -1), geo.detJ[:, None]], axis=1), 9)
    _, first, ccls = np.unique(gkey, axis=0, return_index=True, return_inverse=True)
    rcls, R = kernel_mirror.row_stencil_classes(space.dofmap, ccls, space.n_nodes)
    assert R <= (4 if degree == 2 else 3) ** dim                      # (lo, odd, even, hi) per axis for P2, (lo, inner, hi) for P1
    x = np.random.default_rng(0).standard_normal(space.n_nodes)
    y, table = kernel_mirror.row_stencil_apply(space.dofmap, ccls, A[first], rcls, R, x)
    ys = np.zeros(space.n_nodes)
    np.add.at(ys, space.dofmap.ravel(), np.einsum("cij,cj->ci", A, x[space.dofmap]).ravel())
    assert np.max(np.abs(y - ys)) <= 1e-13 * np.max(np.abs(ys))
    assert max(len(t) for t in table) == {(2, 1): 7, (2, 2): 19, (3, 1): 15, (3, 2): 65}[(dim, degree)]   # Kuhn vertex stars
    perm = np.random.default_rng(1).permutation(space.n_nodes)
    _, R_scrambled = kernel_mirror.row_stencil_classes(perm[space.dofmap], ccls, space.n_nodes)
    assert R_scrambled > space.n_nodes // 2


@pytest.mark.parametrize("dim,n,degree", [(2, (10, 6), 2), (3, (4, 7, 3), 1), (3, (4, 7, 3), 2)])
def test_row_stencil_form_of_the_cg_residual(dim, n, degree):
    """Large CG meshes evaluate the cell part of the residual as S_J T - S_M T_prev: the row classes of the Jacobian tables
    applied to T minus the row classes of the class MASS matrices |detJ| Mhat applied to T_prev (csrc/thermal.cu
    build_stencil_t, numpy emulation).  Must equal the cell-by-cell residual (f = 0, exterior facets apart)."""
    import kernel_mirror
    from fem_glass_tempering_b200 import mesh as msh
    m = msh.plate_mesh(dim, n, tuple(float(k) for k in n))
    space = fe.ScalarSpace(m, "CG", degree)
    tabs, geo = fe.operator_tables(dim, degree), fe.cell_geometry(m)
    A = kernel_mirror.cell_matrices(space, tabs, geo, MAIN_PARAMS, 0.1)
    Mc = geo.detJ[:, None, None] * tabs.mass[None]
    gkey = np.round(np.concatenate([geo.Jinv.reshape(m.n_cells, -1), geo.detJ[:, None]], axis=1), 9)
    _, first, ccls = np.unique(gkey, axis=0, return_index=True, return_inverse=True)
    rcls, R = kernel_mirror.row_stencil_classes(space.dofmap, ccls, space.n_nodes)
    rng = np.random.default_rng(2)
    T, Tp = 700 + 100 * rng.random(space.n_nodes), 700 + 100 * rng.random(space.n_nodes)
    yJ, _ = kernel_mirror.row_stencil_apply(space.dofmap, ccls, A[first], rcls, R, T)
    yM, _ = kernel_mirror.row_stencil_apply(space.dofmap, ccls, Mc[first], rcls, R, Tp)
    F = np.zeros(space.n_nodes)
    np.add.at(F, space.dofmap.ravel(), (np.einsum("cij,cj->ci", A, T[space.dofmap]) - np.einsum("cij,cj->ci", Mc, Tp[space.dofmap])).ravel())
    assert np.max(np.abs((yJ - yM) - F)) <= 1e-12 * np.max(np.abs(F))
    # ... which is the cell part of the oracle's residual: M (T - T_prev) + dt alpha K T
    orc = to.ThermalOracle(m.x, m.cells, space.dofmap, space.element.nodes, "CG", degree, MAIN_PARAMS, 0.1)
    Fo = orc.M @ (T - Tp) + 0.1 * float(MAIN_PARAMS["alpha"]) * (orc.K @ T)
    assert np.max(np.abs(F - Fo)) <= 1e-12 * np.max(np.abs(Fo))


def test_p2_lattice_numbering_keeps_planes_and_runs():
    """The P2 numbering of the plate meshes: x half-planes are contiguous id ranges (distributed.slab_partition relies on
    it) and inside a plane consecutive ids run along the longest in-plane axis with even half-steps before odd ones, so
    that neighbouring ids are nodes of the same type (one row class per warp in csrc/stencil.cu)."""
    from fem_glass_tempering_b200 import mesh as msh
    m = msh.box_mesh(3, 8, 2, 3.0, 8.0, 2.0)
    space = fe.ScalarSpace(m, "CG", 2)
    X = space.tabulate_dof_coordinates()
    h2 = np.rint(2 * X).astype(int)                                  # half-step lattice coordinates (unit cells)
    plane = 17 * 5
    assert space.n_nodes == 7 * plane
    ids = np.arange(space.n_nodes)
    assert np.array_equal(h2[:, 0], ids // plane)                    # x slowest
    inp = ids % plane
    assert np.array_equal(h2[:, 2], inp // 17)                       # then z (the short axis)
    run = inp % 17
    assert np.array_equal(h2[:, 1], np.where(run < 9, 2 * run, 2 * (run - 9) + 1))   # y: 9 even half-steps, then 8 odd


@pytest.mark.parametrize("dim", [1, 2, 3])
def test_device_side_setup_equals_the_numpy_statements(dim):
    """fe.facet_topology / fe.cell_geometry run as torch array code (on the GPU when there is one); the numpy/LAPACK
    statements they replaced stay as the checkers."""
    from fem_glass_tempering_b200 import mesh as msh
    rng = np.random.default_rng(dim)
    m = {1: msh.graded_line_mesh(), 2: msh.rectangle_mesh(7, 5, 3.5, 2.0), 3: msh.box_mesh(4, 3, 5, 2.0, 1.5, 2.5)}[dim]
    if dim > 1:                                     # perturb the interior so that every cell has its own geometry
        inner = np.all((m.x > 1e-9) & (m.x < m.x.max(axis=0) - 1e-9), axis=1)
        m.x[inner] += 0.05 * rng.uniform(-1, 1, (int(inner.sum()), dim))
    a, b = fe.facet_topology(m, device="cpu"), fe._facet_topology_np(m)
    for k in ("neighbor", "nb_facet", "nb_perm", "bnd_cell", "bnd_facet"):
        assert np.array_equal(getattr(a, k), getattr(b, k)), k
    mask = lambda mid: mid[:, 0] > 0.5 * m.x[:, 0].max()
    a, b = fe.facet_topology(m, mask, device="cpu"), fe._facet_topology_np(m, mask)
    assert np.array_equal(a.bnd_cell, b.bnd_cell) and np.array_equal(a.bnd_facet, b.bnd_facet)
    g, h = fe.cell_geometry(m, device="cpu"), fe._cell_geometry_np(m)
    assert np.allclose(g.detJ, h.detJ, rtol=1e-14, atol=0) and np.allclose(g.h, h.h, rtol=1e-15, atol=0)
    assert np.max(np.abs(g.Jinv - h.Jinv)) <= 1e-13 * np.max(np.abs(h.Jinv))
