"""GPU parity of hot path (B): matrix-free CUDA operator / PCG / Newton (through the C ABI) against the
assembled scipy oracle.  Tolerances are written next to each assertion; the north_star bar is 1e-10."""
import numpy as np
import pytest
import torch

from fem_glass_tempering_b200 import fe
from fem_glass_tempering_b200 import mesh as msh
from fem_glass_tempering_b200.thermal_op import ThermalOperator
from oracle import thermal_oracle as to
from oracle.visco_oracle import MAIN_PARAMS

pytestmark = pytest.mark.gpu

CASES = [(1, "CG", 1), (1, "DG", 1), (1, "CG", 2), (1, "DG", 2), (2, "CG", 1), (2, "CG", 2), (2, "DG", 1),
         (2, "DG", 2), (3, "CG", 1), (3, "CG", 2), (3, "DG", 1), (3, "DG", 2)]


def make_mesh(dim):
    if dim == 1:
        return msh.graded_line_mesh()
    if dim == 2:
        return msh.rectangle_mesh(9, 5, 9.0, 5.0)
    return msh.box_mesh(5, 4, 3, 5.0, 4.0, 3.0)


def setup(sg_ctx, dim, family, degree, dt=0.1, params=MAIN_PARAMS, use_classes=True, cheb_degree=3):
    m = make_mesh(dim)
    space = fe.ScalarSpace(m, family, degree)
    op = ThermalOperator(sg_ctx, space, params, dt, use_classes=use_classes, cheb_degree=cheb_degree)
    orc = to.ThermalOracle(m.x, m.cells, space.dofmap, space.element.nodes, family, degree, params, dt)
    return m, space, op, orc


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to("cuda:0")


@pytest.mark.parametrize("use_classes", [True, False], ids=["class_tables", "per_cell_geometry"])
@pytest.mark.parametrize("dim,family,degree", CASES)
def test_operator_matches_assembled_oracle(sg_ctx, dim, family, degree, use_classes):
    m, space, op, orc = setup(sg_ctx, dim, family, degree, use_classes=use_classes)
    info = op.class_info()
    assert info["active"] == use_classes          # the plates have a handful of cell shapes: the fast path must engage
    if use_classes and dim > 1:
        # Kuhn triangulation of a uniform box: d! cell shapes (rounding noise of the coordinates may split a few)
        assert 1 <= info["geometry"] <= 4 * (2 if dim == 2 else 6), info
    rng = np.random.default_rng(dim * 10 + degree)
    n = space.n_nodes
    T = 700 + 100 * rng.random(n)
    Tp = T + rng.random(n)
    x = rng.standard_normal(n)
    J = orc.jacobian(T)
    out = torch.empty(n, dtype=torch.float64, device="cuda:0")
    y = op.jac_apply(dev(T), dev(x), out).cpu().numpy()
    yo = J @ x
    assert np.max(np.abs(y - yo)) <= 1e-12 * np.max(np.abs(yo))        # Jacobian apply: 1e-12 relative
    r = op.residual(dev(T), dev(Tp), out).cpu().numpy()
    ro = orc.residual(T, Tp)
    assert np.max(np.abs(r - ro)) <= 1e-12 * np.max(np.abs(ro))        # residual: 1e-12 relative
    dg = op.jac_diag(dev(T), out).cpu().numpy()
    do = J.diagonal()
    assert np.max(np.abs(dg - do)) <= 1e-12 * np.max(np.abs(do))       # Jacobi diagonal: 1e-12 relative


@pytest.mark.parametrize("use_classes", [True, False], ids=["class_tables", "per_cell_geometry"])
def test_hand_evaluated_vectors_on_the_reference_mesh(sg_ctx, use_classes):
    """CUDA residual / Jacobian apply against tests/golden/thermal_kat.json (hand-evaluated on the graded 1-D line of
    geometry.py with closed-form element matrices, independent of the oracle): 1e-12 relative."""
    from helpers import load_thermal_kat
    kat = load_thermal_kat()
    m = msh.graded_line_mesh()
    for c in kat["cases"]:
        space = fe.ScalarSpace(m, c["family"], c["degree"])
        op = ThermalOperator(sg_ctx, space, MAIN_PARAMS, kat["dt"], use_classes=use_classes)
        out = torch.empty(space.n_nodes, dtype=torch.float64, device="cuda:0")
        F = op.residual(dev(c["T"]), dev(c["T_prev"]), out).cpu().numpy()
        assert np.max(np.abs(F - c["residual"])) <= 1e-12 * np.max(np.abs(c["residual"])), c["family"]
        Jx = op.jac_apply(dev(c["T"]), dev(c["x"]), out).cpu().numpy()
        assert np.max(np.abs(Jx - c["jac_x"])) <= 1e-12 * np.max(np.abs(c["jac_x"])), c["family"]


@pytest.mark.parametrize("use_classes", [True, False], ids=["class_tables", "per_cell_geometry"])
def test_hand_evaluated_vectors_on_perturbed_2d_3d_meshes(sg_ctx, use_classes):
    """CUDA residual / Jacobian apply against tests/golden/thermal_kat_2d3d.json (hand-evaluated in plain Python on small
    perturbed triangle / tetrahedron meshes stored in the file; DG1 and CG1): 1e-12 relative."""
    from helpers import load_thermal_kat_simplex
    from fem_glass_tempering_b200.mesh import Mesh
    kat = load_thermal_kat_simplex()
    for c in kat["cases"]:
        space = fe.ScalarSpace(Mesh(c["x"], c["cells"]), c["family"], c["degree"])
        op = ThermalOperator(sg_ctx, space, MAIN_PARAMS, kat["dt"], use_classes=use_classes)
        out = torch.empty(space.n_nodes, dtype=torch.float64, device="cuda:0")
        F = op.residual(dev(c["T"]), dev(c["T_prev"]), out).cpu().numpy()
        assert np.max(np.abs(F - c["residual"])) <= 1e-12 * np.max(np.abs(c["residual"])), (c["dim"], c["family"])
        Jx = op.jac_apply(dev(c["T"]), dev(c["v"]), out).cpu().numpy()
        assert np.max(np.abs(Jx - c["jac_x"])) <= 1e-12 * np.max(np.abs(c["jac_x"])), (c["dim"], c["family"])


@pytest.mark.parametrize("use_classes", [True, False], ids=["class_tables", "per_cell_geometry"])
def test_hand_evaluated_cg2_vectors(sg_ctx, use_classes):
    """CUDA residual / Jacobian apply for CG2 (BASELINE configs 2 and 4) against tests/golden/thermal_kat_p2.json
    (plain-Python P2 basis and Duffy-Gauss integration on perturbed meshes): 1e-12 relative."""
    from helpers import load_thermal_kat_p2
    from fem_glass_tempering_b200.mesh import Mesh
    kat = load_thermal_kat_p2()
    for c in kat["cases"]:
        space = fe.ScalarSpace(Mesh(c["x"], c["cells"]), "CG", 2)
        op = ThermalOperator(sg_ctx, space, MAIN_PARAMS, kat["dt"], use_classes=use_classes)
        out = torch.empty(space.n_nodes, dtype=torch.float64, device="cuda:0")
        F = op.residual(dev(c["T"]), dev(c["T_prev"]), out).cpu().numpy()
        assert np.max(np.abs(F - c["residual"])) <= 1e-12 * np.max(np.abs(c["residual"])), c["dim"]
        Jx = op.jac_apply(dev(c["T"]), dev(c["v"]), out).cpu().numpy()
        assert np.max(np.abs(Jx - c["jac_x"])) <= 1e-12 * np.max(np.abs(c["jac_x"])), c["dim"]


@pytest.mark.parametrize("dim,family,degree", [(1, "DG", 1), (2, "CG", 2), (3, "DG", 1), (3, "CG", 2), (2, "DG", 2)])
def test_pcg_solves_the_linear_system(sg_ctx, dim, family, degree):
    import scipy.sparse.linalg as spla
    m, space, op, orc = setup(sg_ctx, dim, family, degree)
    n = space.n_nodes
    rng = np.random.default_rng(3)
    T = np.full(n, 800.0) - rng.random(n)
    b = rng.standard_normal(n)
    Td, bd, xd = dev(T), dev(b), torch.zeros(n, dtype=torch.float64, device="cuda:0")
    op.prepare_preconditioner(Td)
    its, res = op.pcg(Td, bd, xd, rtol=1e-13)
    xo = spla.spsolve(orc.jacobian(T).tocsc(), b)
    assert res <= 1e-13 and 1 <= its < 2000
    assert np.max(np.abs(xd.cpu().numpy() - xo)) <= 1e-10 * np.max(np.abs(xo))   # solution: 1e-10 relative


@pytest.mark.parametrize("dim,family,degree", [(1, "DG", 1), (1, "CG", 1), (2, "CG", 2), (3, "DG", 1), (3, "CG", 1)])
def test_time_steps_match_oracle_newton(sg_ctx, dim, family, degree):
    """Several implicit-Euler steps from T_0 = 800 K: temperature within 1e-10 relative of the oracle
    (the oracle solves each Newton system directly)."""
    m, space, op, orc = setup(sg_ctx, dim, family, degree)
    n = space.n_nodes
    T_o = np.full(n, 800.0)
    T_d, Tp_d = dev(T_o), dev(T_o)
    for step in range(4):
        Tp_o = T_o.copy()
        T_o, its_o, ok = orc.newton(T_o, Tp_o)
        assert ok
        st = op.timestep(T_d, Tp_d)
        assert st.converged == 1 and st.newton_its <= its_o + 2
        err = np.max(np.abs(T_d.cpu().numpy() - T_o)) / np.max(np.abs(T_o))
        assert err <= 1e-10, f"step {step}: rel err {err}"
        dT_o = T_o - Tp_o                                  # the increment drives the stress (H3): 1e-8 of its size
        dT_d = T_d.cpu().numpy() - Tp_d.cpu().numpy()
        assert np.max(np.abs(dT_d - dT_o)) <= 1e-8 * np.max(np.abs(dT_o))
        Tp_d.copy_(T_d)


@pytest.mark.parametrize("dim,family,degree", [(3, "DG", 1), (3, "CG", 2), (2, "DG", 1), (2, "CG", 1), (1, "DG", 2)])
def test_fused_dot_product_of_the_solver(sg_ctx, dim, family, degree):
    """The PCG iteration count and solution must not depend on which apply kernel runs (class tables with the
    fused x.Ax reduction, or per-cell geometry + separate dot kernel)."""
    res = {}
    for uc in (True, False):
        m, space, op, orc = setup(sg_ctx, dim, family, degree, use_classes=uc, cheb_degree=0)
        n = space.n_nodes
        rng = np.random.default_rng(11)
        T = np.full(n, 790.0) - rng.random(n)
        b = rng.standard_normal(n)
        Td, bd, xd = dev(T), dev(b), torch.zeros(n, dtype=torch.float64, device="cuda:0")
        op.prepare_preconditioner(Td)
        its, rr = op.pcg(Td, bd, xd, rtol=1e-12)
        res[uc] = (its, xd.cpu().numpy())
    assert abs(res[True][0] - res[False][0]) <= 2 + res[False][0] // 16      # rounding may shift a long solve by a few iterations
    assert np.max(np.abs(res[True][1] - res[False][1])) <= 1e-10 * np.max(np.abs(res[False][1]))


@pytest.mark.parametrize("dim,degree", [(3, 1), (2, 1), (1, 1), (1, 2), (3, 2)])
def test_chebyshev_preconditioned_pcg(sg_ctx, dim, degree):
    """DG: CG preconditioned with the Chebyshev polynomial in M^-1 J reaches the same solution as the element-mass
    preconditioner in fewer (outer) iterations; the spectrum interval comes from the library's Lanczos estimate."""
    import scipy.sparse.linalg as spla
    m = make_mesh(dim) if degree == 1 else (msh.graded_line_mesh() if dim == 1 else msh.box_mesh(4, 3, 2, 12.0, 9.0, 6.0))
    space = fe.ScalarSpace(m, "DG", degree)
    orc = to.ThermalOracle(m.x, m.cells, space.dofmap, space.element.nodes, "DG", degree, MAIN_PARAMS, 0.1)
    n = space.n_nodes
    rng = np.random.default_rng(5)
    T = np.full(n, 790.0) - rng.random(n)
    b = rng.standard_normal(n)
    xo = spla.spsolve(orc.jacobian(T).tocsc(), b)
    its = {}
    for k in (0, 1, 2, 3, 4):
        op = ThermalOperator(sg_ctx, space, MAIN_PARAMS, 0.1, cheb_degree=k)
        assert op.chebyshev_degree == k
        Td, bd, xd = dev(T), dev(b), torch.zeros(n, dtype=torch.float64, device="cuda:0")
        its[k], res = op.pcg(Td, bd, xd, rtol=1e-12)
        assert res <= 1e-12
        assert np.max(np.abs(xd.cpu().numpy() - xo)) <= 1e-9 * np.max(np.abs(xo)), (k, its)
        info = op.chebyshev_info()
        assert info["degree"] == k, f"degree {k}: fell back ({info})"
        if k:
            assert 0 < info["lo"] < info["hi"]
    assert its[1] < its[0] and its[3] <= its[1], its
    if dim > 1 and degree == 1:      # on the plates the operator applications stay within ~1.6x of the plain iteration's (the 96-dof graded
        assert its[4] * 5 <= its[0] * 1.6 + 5, its      # line is solved by plain CG in far fewer steps than its conditioning suggests)


def test_chebyshev_with_too_small_bound_falls_back(sg_ctx):
    m, space, op, orc = setup(sg_ctx, 3, "DG", 1)
    assert op.set_chebyshev(3, lo=0.5, hi=2.0)            # the true largest eigenvalue of M^-1 J is ~15 here
    n = space.n_nodes
    rng = np.random.default_rng(2)
    T = np.full(n, 800.0)
    b = rng.standard_normal(n)
    Td, bd, xd = dev(T), dev(b), torch.zeros(n, dtype=torch.float64, device="cuda:0")
    its, res = op.pcg(Td, bd, xd, rtol=1e-12)
    assert res <= 1e-12 and op.chebyshev_info()["degree"] == 0
    y = torch.empty_like(xd)
    op.jac_apply(Td, xd, y)
    assert float((y - bd).abs().max()) <= 1e-10 * float(bd.abs().max())


def test_chebyshev_is_not_offered_for_cg_spaces(sg_ctx):
    m, space, op, orc = setup(sg_ctx, 2, "CG", 2)
    assert op.chebyshev_degree == 0 and not op.set_chebyshev(3)


@pytest.mark.parametrize("dim,family,degree", [(2, "CG", 2), (3, "CG", 2), (3, "DG", 1), (1, "DG", 1)])
def test_cuda_graph_batches_match_plain_launches(sg_ctx, dim, family, degree, monkeypatch):
    """The plain PCG replays batches of 8 iterations as a CUDA graph on the solver's own stream; SG_NO_GRAPHS=1 launches
    the same kernels one by one.  Same iteration count, same solution; repeated solves reuse the graph."""
    res = {}
    for no_graphs in ("0", "1"):
        monkeypatch.setenv("SG_NO_GRAPHS", no_graphs)
        m, space, op, orc = setup(sg_ctx, dim, family, degree, cheb_degree=0)
        n = space.n_nodes
        rng = np.random.default_rng(21)
        T = np.full(n, 790.0) - rng.random(n)
        Td = dev(T)
        op.prepare_preconditioner(Td)
        sols = []
        for rep in range(3):
            b = rng.standard_normal(n)
            xd = torch.zeros(n, dtype=torch.float64, device="cuda:0")
            its, rr = op.pcg(Td, dev(b), xd, rtol=1e-12)
            sols.append((its, xd.cpu().numpy()))
        res[no_graphs] = sols
    for (i0, x0), (i1, x1) in zip(res["0"], res["1"]):
        assert abs(i0 - i1) <= 1     # consecutive kernels alternate their sweep direction (summation order): +-1 iteration
        assert np.max(np.abs(x0 - x1)) <= 1e-11 * np.max(np.abs(x1))
    assert res["0"][0][0] > 8          # more than one batch: the graph path was exercised


@pytest.mark.parametrize("dim,degree,dims", [(2, 1, (9, 5)), (2, 2, (9, 5)), (3, 1, (5, 4, 3)), (3, 2, (5, 4, 3)),
                                             (3, 2, (37, 30, 6)), (2, 2, (130, 70)), (1, 2, None)])
def test_cg_row_stencil_form_matches_the_cell_kernel(sg_ctx, dim, degree, dims):
    """CG spaces: the gather form of the apply (row-stencil classes, csrc/stencil.cu) against the cell-centric class kernel
    (RED.ADD scatter), and on the small meshes against the assembled oracle: 1e-12 relative.  The gather form is
    deterministic (bit-identical repeats); PCG must not care which form runs."""
    m = msh.graded_line_mesh() if dim == 1 else msh.plate_mesh(dim, dims, tuple(float(k) for k in dims))
    space = fe.ScalarSpace(m, "CG", degree)
    ops = {st: ThermalOperator(sg_ctx, space, MAIN_PARAMS, 0.1, use_stencil=st, cheb_degree=0) for st in (True, False)}
    info = ops[True].stencil_info()
    assert info["active"] and not ops[False].stencil_info()["active"], info
    if dim > 1:
        # lattice-numbered plate: (boundary-lo, odd, even, boundary-hi) per axis for P2, (lo, interior, hi) for P1
        assert info["classes"] <= (4 if degree == 2 else 3) ** dim * 2, info
    n = space.n_nodes
    rng = np.random.default_rng(3)
    T = 700 + 100 * rng.random(n)
    x = rng.standard_normal(n)
    Td, xd = dev(T), dev(x)
    ys = ops[True].jac_apply(Td, xd, torch.empty(n, dtype=torch.float64, device="cuda:0")).cpu().numpy()
    ys2 = ops[True].jac_apply(Td, xd, torch.full((n,), 7.0, dtype=torch.float64, device="cuda:0")).cpu().numpy()
    yc = ops[False].jac_apply(Td, xd, torch.empty(n, dtype=torch.float64, device="cuda:0")).cpu().numpy()
    # y need not be zeroed; the cell part has no atomics, only the exterior-facet part adds in arbitrary order
    assert np.max(np.abs(ys - ys2)) <= 1e-15 * np.max(np.abs(ys))
    assert np.max(np.abs(ys - yc)) <= 1e-12 * np.max(np.abs(yc))
    if m.n_cells < 5000:
        orc = to.ThermalOracle(m.x, m.cells, space.dofmap, space.element.nodes, "CG", degree, MAIN_PARAMS, 0.1)
        yo = orc.jacobian(T) @ x
        assert np.max(np.abs(ys - yo)) <= 1e-12 * np.max(np.abs(yo))
    res = {}
    b = rng.standard_normal(n)
    for st, op in ops.items():
        xs = torch.zeros(n, dtype=torch.float64, device="cuda:0")
        op.prepare_preconditioner(Td)
        its, _ = op.pcg(Td, dev(b), xs, rtol=1e-12)
        res[st] = (its, xs.cpu().numpy())
    assert abs(res[True][0] - res[False][0]) <= 2 + res[False][0] // 16
    assert np.max(np.abs(res[True][1] - res[False][1])) <= 1e-10 * np.max(np.abs(res[False][1]))


def test_cg_row_stencil_on_a_rank_slab(sg_ctx):
    """One rank's slab of a partitioned plate (ghost columns on both sides, cell range inside the local mesh): the gather
    form must reproduce the cell kernel on every local row, including the incomplete ghost rows."""
    from fem_glass_tempering_b200 import distributed
    m, part, _ = distributed.slab_partition(3, (12, 5, 3), (12.0, 5.0, 3.0), "CG", 2, 1, 3)
    part = dict(part, halo=[])
    space = fe.ScalarSpace(m, "CG", 2)
    ops = {st: ThermalOperator(sg_ctx, space, MAIN_PARAMS, 0.1, partition=part, use_stencil=st, cheb_degree=0) for st in (True, False)}
    assert ops[True].stencil_info()["active"]
    n = space.n_nodes
    rng = np.random.default_rng(4)
    Td, xd = dev(700 + 100 * rng.random(n)), dev(rng.standard_normal(n))
    ys = ops[True].jac_apply(Td, xd, torch.empty(n, dtype=torch.float64, device="cuda:0")).cpu().numpy()
    yc = ops[False].jac_apply(Td, xd, torch.empty(n, dtype=torch.float64, device="cuda:0")).cpu().numpy()
    assert np.max(np.abs(ys - yc)) <= 1e-12 * np.max(np.abs(yc))


def test_many_shapes_fall_back_to_per_cell_geometry(sg_ctx):
    """A mesh whose cells all differ (randomly perturbed vertices) has too many classes for the shared-memory tables:
    the library must keep the general kernel and still match the oracle."""
    m = msh.box_mesh(7, 6, 5, 7.0, 6.0, 5.0)
    rng = np.random.default_rng(0)
    m.x += 0.05 * rng.standard_normal(m.x.shape)
    space = fe.ScalarSpace(m, "DG", 1)
    op = ThermalOperator(sg_ctx, space, MAIN_PARAMS, 0.1)
    orc = to.ThermalOracle(m.x, m.cells, space.dofmap, space.element.nodes, "DG", 1, MAIN_PARAMS, 0.1)
    info = op.class_info()
    assert not info["active"] and info["geometry"] == m.n_cells
    n = space.n_nodes
    T = 700 + 100 * rng.random(n)
    x = rng.standard_normal(n)
    out = torch.empty(n, dtype=torch.float64, device="cuda:0")
    y = op.jac_apply(dev(T), dev(x), out).cpu().numpy()
    yo = orc.jacobian(T) @ x
    assert np.max(np.abs(y - yo)) <= 1e-12 * np.max(np.abs(yo))


def test_nonconvergence_is_reported(sg_ctx):
    from fem_glass_tempering_b200 import _lib
    m, space, op, orc = setup(sg_ctx, 2, "CG", 1)
    n = space.n_nodes
    T = dev(np.full(n, 800.0))
    op.opts.newton_max_it = 1
    with pytest.raises(_lib.SgError) as e:
        op.timestep(T, T.clone())
    assert e.value.code == _lib.SG_E_NOCONV          # TVP:390 assert(converged)


def test_cg_residual_in_gather_form(sg_ctx):
    """Large CG meshes evaluate the residual's cell part as S_J T - S_M T_prev with two row-stencil launches (no RED
    scatter in the time loop).  Forced on small meshes here (subprocess: the threshold is read once per process) and
    compared with the cell-centric scatter kernel and, through the existing tests, the oracle."""
    import os
    import subprocess
    import sys
    code = r"""
import numpy as np, torch, sys
sys.path.insert(0, %r)
from fem_glass_tempering_b200 import _lib, fe
from fem_glass_tempering_b200 import mesh as msh
from fem_glass_tempering_b200.thermal_op import ThermalOperator
from oracle import thermal_oracle as to
from oracle.visco_oracle import MAIN_PARAMS
ctx = _lib.Context(0)
for dim, degree, dims in ((2, 2, (9, 5)), (3, 1, (5, 4, 3)), (3, 2, (5, 4, 3)), (3, 2, (21, 18, 4))):
    m = msh.plate_mesh(dim, dims, tuple(float(k) for k in dims))
    space = fe.ScalarSpace(m, "CG", degree)
    op = ThermalOperator(ctx, space, MAIN_PARAMS, 0.1, cheb_degree=0)
    assert op.stencil_info()["active"]
    n = space.n_nodes
    rng = np.random.default_rng(5)
    T, Tp = 700 + 100 * rng.random(n), 700 + 100 * rng.random(n)
    F = op.residual(torch.from_numpy(T).cuda(), torch.from_numpy(Tp).cuda(), torch.full((n,), 3.0, dtype=torch.float64, device="cuda")).cpu().numpy()
    if m.n_cells < 5000:
        orc = to.ThermalOracle(m.x, m.cells, space.dofmap, space.element.nodes, "CG", degree, MAIN_PARAMS, 0.1)
        Fo = orc.residual(T, Tp)
        assert np.max(np.abs(F - Fo)) <= 1e-12 * np.max(np.abs(Fo)), (dim, degree, np.max(np.abs(F - Fo)))
    print("ok", dim, degree, dims)
print("GATHER_RESID_OK")
""" % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    outs = {}
    for rows in ("0", "1000000000"):
        env = dict(os.environ, SG_GATHER_RESID_MIN_ROWS=rows)
        res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=600)
        assert res.returncode == 0 and "GATHER_RESID_OK" in res.stdout, res.stdout[-1500:] + res.stderr[-1500:]
        outs[rows] = res.stdout
    assert outs["0"].count("ok") == outs["1000000000"].count("ok") == 4
