"""GPU: the mechanical-equilibrium step (csrc/mech.cu through the C ABI; SURVEY §8(f) row 4 — an extension, the
reference sets total_strain = -thermal_strain, VM:135-139) against oracle/mechanics_oracle.py, the hand-evaluated answers
of tests/golden/mech_kat.json and mesh-independent properties.

Tolerances: operator application, right-hand side and correction 1e-12 relative (same arithmetic, different summation
order — atomics); solved displacement and equilibrated stress 1e-8 relative with the PCG run to rtol 1e-12 (the tangent's
condition number amplifies the residual tolerance; stated per test)."""
import json
import os

import numpy as np
import pytest
import torch

from fem_glass_tempering_b200 import ThermoViscoProblem, _lib, fe
from fem_glass_tempering_b200 import mesh as M
from fem_glass_tempering_b200.mechanics import MechanicalEquilibrium
from helpers import GOLDEN, rel_err, unhex
from oracle import mechanics_oracle as mo
from oracle import visco_oracle as vo
from oracle.reference_problem import OracleProblem

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")

MESHES = {1: lambda: M.interval_mesh(9, 3.0), 2: lambda: M.perturb_interior(M.rectangle_mesh(5, 4, 5.0, 4.0), 0.15, seed=3),
          3: lambda: M.perturb_interior(M.box_mesh(3, 3, 2, 3.0, 3.0, 2.0), 0.12, seed=5)}


def gpu(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(DEV)


def make_plan(ctx, dim, mode="reference", dt=0.1):
    p = vo.ViscoParams(dim=dim, dt=dt)
    plan = _lib.ViscoPlan(ctx, dim=dim, dt=dt, H=p.H, Rg=p.Rg, Tb=p.Tb, alpha_solid=p.alpha_solid, alpha_liquid=p.alpha_liquid,
                          m=p.m, lambda_m=p.lambda_m, g=p.g, lambda_g=p.lambda_g, k=p.k, lambda_k=p.lambda_k,
                          mode=_lib.VISCO_CORRECTED if mode == "corrected" else _lib.VISCO_REFERENCE)
    return p, plan


def make_pair(ctx, mesh, family, degree, mode="reference", **opts):
    S = fe.ScalarSpace(mesh, family, degree)
    p, plan = make_plan(ctx, mesh.dim, mode)
    me = MechanicalEquilibrium(ctx, mesh, S, plan, DEV, dict(rtol=1e-12, **opts))
    O = mo.MechanicsOracle(mesh.x, mesh.cells, S.dofmap, mo.sigma_weights(S.element.nodes, degree, mesh.dim),
                           mo.symmetry_planes(mesh.x))
    return S, p, plan, me, O


def load_mech_kat():
    with open(os.path.join(GOLDEN, "mech_kat.json")) as fh:
        return {c["name"]: c for c in json.load(fh)["cases"]}


@pytest.mark.parametrize("dim", [1, 2, 3])
@pytest.mark.parametrize("space", [("DG", 1), ("CG", 1), ("CG", 2)])
def test_apply_and_rhs_equal_the_assembled_oracle(sg_ctx, dim, space):
    mesh = MESHES[dim]()
    S, p, plan, me, O = make_pair(sg_ctx, mesh, *space)
    rng = np.random.default_rng(7 * dim)
    G, K = rng.uniform(20, 30, S.n_nodes), rng.uniform(30, 45, S.n_nodes)
    me.set_moduli(gpu(G), gpu(K))
    x = rng.normal(size=mesh.n_vertices * dim)
    y = me.apply(gpu(x), torch.empty(x.size, dtype=torch.float64, device=DEV)).cpu().numpy()
    assert rel_err(y, O.apply(G, K, x)) < 1e-12
    s0 = rng.normal(size=S.n_nodes * dim * dim)
    b = me.rhs(gpu(s0), torch.empty(x.size, dtype=torch.float64, device=DEV)).cpu().numpy()
    _, b_ref = O.constrained(O.stiffness(G, K), O.rhs(s0))
    assert rel_err(b, b_ref) < 1e-12
    assert me.apply_bytes() > 0


@pytest.mark.parametrize("mode", ["reference", "corrected"])
def test_coefficients_follow_the_chain(sg_ctx, mode):
    mesh = MESHES[2]()
    S, p, plan, me, O = make_pair(sg_ctx, mesh, "DG", 1, mode=mode)
    rng = np.random.default_rng(1)
    xi = rng.uniform(-5e-2, 5e-2, S.n_nodes) if mode == "reference" else rng.uniform(1e-4, 5e-1, S.n_nodes)
    xi[3] = 0.0                                       # T_cur == T_prev bitwise: the factor's limit, not 0/0
    G, K = me.coefficients(gpu(xi))
    Gr, Kr = mo.tangent_moduli(xi, p.g, p.lambda_g, p.k, p.lambda_k, mode)
    assert rel_err(G.cpu().numpy(), Gr) < 1e-13 and rel_err(K.cpu().numpy(), Kr) < 1e-13
    assert abs(G[3].item() - sum(p.g)) < 1e-13 * sum(p.g)


def test_bar_closed_form(sg_ctx):
    c = load_mech_kat()["bar"]
    mesh = M.line_mesh(np.array(unhex(c["x"])))
    S, p, plan, me, O = make_pair(sg_ctx, mesh, "DG", 1)
    K, G, s0 = (np.array(unhex(c[k])) for k in ("K", "G", "sigma0"))
    me.set_moduli(gpu(G), gpu(K))
    du = torch.zeros(mesh.n_vertices, dtype=torch.float64, device=DEV)
    sig = gpu(s0)
    me.solve(sig, du)
    assert rel_err(du.cpu().numpy(), np.array(unhex(c["u"]))) < 1e-11
    eps = torch.empty_like(sig)
    me.correct(du, gpu(np.full(S.n_nodes, 1e-3)), dict(sigma=sig, mech_strain=eps), gpu(G), gpu(K))
    assert np.abs(sig.cpu().numpy() - np.array(unhex(c["sigma"]))).max() < 1e-11 * np.abs(s0).max()
    assert rel_err(eps.cpu().numpy(), np.repeat(unhex(c["eps"]), 2)) < 1e-10


@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("space", [("DG", 1), ("CG", 2)])
def test_patch_test(sg_ctx, dim, space):
    """Constant diagonal eigenstrain, constant moduli: u_i = E_ii (x_i - min x_i) exactly, zero stress afterwards."""
    c = load_mech_kat()[f"patch{dim}"]
    mesh = MESHES[dim]()
    S, p, plan, me, O = make_pair(sg_ctx, mesh, *space)
    E = np.array(unhex(c["E"]))
    G = gpu(np.full(S.n_nodes, float.fromhex(c["G"])))
    K = gpu(np.full(S.n_nodes, float.fromhex(c["K"])))
    s0 = np.tile(np.array(unhex(c["sigma0"])).ravel(), S.n_nodes)
    me.set_moduli(G, K)
    du = torch.zeros(mesh.n_vertices * dim, dtype=torch.float64, device=DEV)
    sig = gpu(s0)
    me.solve(sig, du)
    exact = (mesh.x - mesh.x.min(axis=0)) * E
    assert np.abs(du.cpu().numpy().reshape(-1, dim) - exact).max() < 1e-10 * np.abs(exact).max()
    eps = torch.empty_like(sig)
    me.correct(du, gpu(np.full(S.n_nodes, 1e-3)), dict(sigma=sig, mech_strain=eps), G, K)
    assert np.abs(sig.cpu().numpy()).max() < 1e-9 * np.abs(s0).max()


@pytest.mark.parametrize("dim", [1, 2, 3])
@pytest.mark.parametrize("space", [("DG", 1), ("CG", 1), ("CG", 2)])
def test_solve_and_correct_against_the_oracle(sg_ctx, dim, space):
    mesh = MESHES[dim]()
    S, p, plan, me, O = make_pair(sg_ctx, mesh, *space)
    rng = np.random.default_rng(100 + dim)
    xi = rng.uniform(-2e-2, -1e-3, S.n_nodes)                           # cooling: phi decreases, xi < 0 (VM:170)
    G, K = mo.tangent_moduli(xi, p.g, p.lambda_g, p.k, p.lambda_k)
    eth = rng.uniform(1e-5, 1e-4, S.n_nodes)
    s0 = ((K * dim * eth))[:, None, None] * np.eye(dim) + rng.normal(0, 1e-4, (S.n_nodes, dim, dim))
    me.coefficients(gpu(xi))
    me.set_moduli()
    du = torch.zeros(mesh.n_vertices * dim, dtype=torch.float64, device=DEV)
    sig = gpu(s0.ravel())
    me.solve(sig, du)
    du_ref = O.solve(G, K, s0.ravel())
    assert rel_err(du.cpu().numpy(), du_ref) < 1e-8, (me.last_iters, me.last_rel_res)
    assert 0 < me.last_iters < 5000 and me.last_rel_res <= 1e-12
    # the correction of the full-materialisation arrays, given the ORACLE's displacement (isolates the kernel)
    N, dd = p.N, dim * dim
    names = ("total_strain", "deviatoric_strain", "ds_partial", "dsigma_partial", "s_partial", "sigma_partial")
    before = {k: rng.normal(0, 1e-3, S.n_nodes * (dd if "strain" in k else N * dd)) for k in names}
    t = {k: gpu(v) for k, v in before.items()}
    t["sigma"], t["mech_strain"] = gpu(s0.ravel()), torch.empty(S.n_nodes * dd, dtype=torch.float64, device=DEV)
    me.correct(gpu(du_ref), gpu(xi), t)
    sig_ref, eps_ref = O.correct(du_ref, G, K, s0.ravel())
    assert rel_err(t["sigma"].cpu().numpy(), sig_ref) < 1e-12
    assert rel_err(t["mech_strain"].cpu().numpy(), eps_ref) < 1e-12
    eps = eps_ref.reshape(-1, dim, dim)
    tr = np.trace(eps, axis1=1, axis2=2)
    dev = eps - (tr / dim)[:, None, None] * np.eye(dim)
    Ag, Ak = mo.term_factors(xi, p.lambda_g, "reference"), mo.term_factors(xi, p.lambda_k, "reference")
    ds = 2.0 * np.asarray(p.g)[None, :, None, None] * Ag[:, :, None, None] * dev[:, None]
    dk = np.asarray(p.k)[None, :, None, None] * Ak[:, :, None, None] * (tr[:, None, None] * np.eye(dim))[:, None]
    expect = dict(total_strain=eps.ravel(), deviatoric_strain=dev.ravel(), ds_partial=ds.ravel(), dsigma_partial=dk.ravel(),
                  s_partial=ds.ravel(), sigma_partial=dk.ravel())
    for k in names:
        got = t[k].cpu().numpy() - before[k]
        assert np.abs(got - expect[k]).max() < 1e-9 * max(np.abs(expect[k]).max(), 1e-300), k
    # the per-term increments add up to the nodal correction (what ties G_eff/K_eff to the chain)
    np.testing.assert_allclose((ds + dk).sum(axis=1).ravel(), sig_ref - s0.ravel(), rtol=0, atol=1e-9 * np.abs(sig_ref - s0.ravel()).max())


def test_corrected_scheme_feeds_the_history(sg_ctx):
    """physics = "corrected": the history IS the partial stress, so eps(du) must enter s_tilde / sigma_tilde."""
    mesh = MESHES[3]()
    S, p, plan, me, O = make_pair(sg_ctx, mesh, "DG", 1, mode="corrected")
    rng = np.random.default_rng(5)
    dim, N = 3, p.N
    xi = rng.uniform(1e-3, 1e-1, S.n_nodes)
    G, K = mo.tangent_moduli(xi, p.g, p.lambda_g, p.k, p.lambda_k, "corrected")
    du = rng.normal(0, 1e-4, mesh.n_vertices * dim)
    s_t, k_t = rng.normal(0, 1e-3, S.n_nodes * N * 9), rng.normal(0, 1e-3, S.n_nodes * N * 9)
    t = dict(sigma=gpu(np.zeros(S.n_nodes * 9)), mech_strain=gpu(np.zeros(S.n_nodes * 9)), s_tilde=gpu(s_t), sigma_tilde=gpu(k_t))
    me.coefficients(gpu(xi))
    me.correct(gpu(du), gpu(xi), t)
    sig_ref, eps_ref = O.correct(du, G, K, np.zeros(S.n_nodes * 9))
    assert rel_err(t["sigma"].cpu().numpy(), sig_ref) < 1e-12
    inc = (t["s_tilde"].cpu().numpy() - s_t) + (t["sigma_tilde"].cpu().numpy() - k_t)
    assert rel_err(inc.reshape(S.n_nodes, N, 9).sum(axis=1).ravel(), sig_ref) < 1e-11
    with pytest.raises(_lib.SgError):
        me.correct(gpu(du), gpu(xi), dict(sigma=t["sigma"], mech_strain=t["mech_strain"]))


def test_not_converged_is_an_error(sg_ctx):
    mesh = M.box_mesh(6, 6, 2, 6.0, 6.0, 2.0)
    S, p, plan, me, O = make_pair(sg_ctx, mesh, "DG", 1, max_it=3)
    rng = np.random.default_rng(0)
    me.set_moduli(gpu(rng.uniform(20, 30, S.n_nodes)), gpu(rng.uniform(30, 45, S.n_nodes)))
    with pytest.raises(_lib.SgError) as e:
        me.solve(gpu(rng.normal(size=S.n_nodes * 9)), torch.zeros(mesh.n_vertices * 3, dtype=torch.float64, device=DEV))
    assert e.value.code == _lib.SG_E_NOCONV and me.last_iters == 3


# ------------------------------------------------------------------------------------------ through the problem API
def make_problem(sg_ctx, mesh, config, params, dt=0.1, materialize="all"):
    prob = ThermoViscoProblem(mesh_path="", time=(0.0, 50.0), dt=dt, config=config, model_parameters=params, mesh=mesh,
                              ctx=sg_ctx, verbose=False, materialize=materialize)
    prob.setup(dirichlet_bc=False)
    return prob


@pytest.mark.parametrize("case", [(1, "DG", 1, "CG", 1), (2, "CG", 2, "CG", 2), (3, "DG", 1, "DG", 1)])
def test_time_steps_with_mechanics_against_the_oracle_pipeline(sg_ctx, case):
    """solve_timestep with model_parameters["mechanics"]: the reference's phases (checked elsewhere) followed by the
    equilibrium step = OracleProblem.step + MechanicsOracle on its stress, step by step."""
    dim, fT, dT, fS, dS = case
    mesh = {1: lambda: M.graded_line_mesh(), 2: lambda: M.rectangle_mesh(6, 4, 6.0, 4.0), 3: lambda: M.box_mesh(4, 4, 2, 4.0, 4.0, 2.0)}[dim]()
    config = {"T": {"element": fT, "degree": dT}, "sigma": {"element": fS, "degree": dS}}
    params = dict(vo.MAIN_PARAMS, mechanics={"rtol": 1e-12})
    if fT == "DG" and dim > 1:
        params["sip_penalty"] = 8.0
    prob = make_problem(sg_ctx, mesh, config, params)
    sp = lambda s: dict(dofmap=s.dofmap, ref_nodes=s.element.nodes, family=s.family, degree=s.degree)
    T, S = prob.functionSpaces["T"].scalar, prob.functionSpaces["sigma"].scalar
    ref_params = {k: v for k, v in params.items() if k != "mechanics"}
    orc = OracleProblem(mesh.x, mesh.cells, sp(T), sp(S), ref_params, 0.1)
    O = mo.MechanicsOracle(mesh.x, mesh.cells, S.dofmap, mo.sigma_weights(S.element.nodes, dS, dim), mo.symmetry_planes(mesh.x))
    u_ref = np.zeros(mesh.n_vertices * dim)
    p = orc.vp
    for step in range(3):
        prob.t += prob.dt
        prob.solve_timestep(prob.t)
        orc.step()
        xi_s = np.zeros(S.n_nodes)
        xi_s[S.dofmap.ravel()] = orc._T_at_sigma_points(orc.f["xi"])            # last cell wins
        G, K = mo.tangent_moduli(xi_s, p.g, p.lambda_g, p.k, p.lambda_k)
        s0 = orc.f["sigma_next"]
        du_ref = O.solve(G, K, s0)
        sig_ref, eps_ref = O.correct(du_ref, G, K, s0)
        u_ref += du_ref
        du = prob.functions["displacement_increment"].x.array.cpu().numpy()
        scale = np.abs(du_ref).max()
        assert np.abs(du - du_ref).max() < 1e-7 * scale, (step, prob.mechanics.last_iters)
        assert np.abs(prob.functions["displacement"].x.array.cpu().numpy() - u_ref).max() < 1e-7 * np.abs(u_ref).max()
        sig = prob.functions_next["sigma"].x.array.cpu().numpy()
        assert np.abs(sig - sig_ref).max() < 1e-7 * np.abs(np.nan_to_num(s0)).max(), step
        assert rel_err(prob.functions["mechanical_strain"].x.array.cpu().numpy(), eps_ref) < 1e-7
        # the plate relaxes: the equilibrated stress is far below the restrained one
        assert np.abs(sig).max() < np.abs(np.nan_to_num(s0)).max()
        orc.end_step()


def test_dg_stress_is_in_discrete_equilibrium_after_every_step(sg_ctx):
    """Size-independent property on a plate the CPU oracle would not solve in seconds: B^T sigma_h = 0 on the free
    components (exact for DG sigma spaces), with the corrected physics, several steps, warm-started PCG."""
    mesh = M.box_mesh(24, 24, 4, 24.0, 24.0, 4.0)
    config = {"T": {"element": "DG", "degree": 1}, "sigma": {"element": "DG", "degree": 1}}
    params = dict(vo.MAIN_PARAMS, mechanics=True, physics="corrected", sip_penalty=6.0)
    prob = make_problem(sg_ctx, mesh, config, params, materialize="minimal")
    me = prob.mechanics
    its = []
    for step in range(4):
        prob.t += prob.dt
        prob.solve_timestep(prob.t)
        sig = prob.functions_next["sigma"].x.array
        b = me.rhs(sig, torch.empty(mesh.n_vertices * 3, dtype=torch.float64, device=DEV))
        # scale: the out-of-balance force of the same stress field with every cell taken alone
        scale = (sig.abs().max() * mesh.x.max()).item()
        assert b.abs().max().item() < 1e-8 * scale, (step, b.abs().max().item(), scale)
        assert torch.isfinite(sig).all()
        its.append(me.last_iters)
    assert its[-1] <= its[0]          # warm start from the previous increment never costs more than the cold start
    assert prob.functions["displacement"].x.array.abs().max().item() > 0.0


def test_field_writer_records_the_displacement(sg_ctx, tmp_path):
    """With mechanics on, the per-step output (TVP:357-364 schedule, output.FieldWriter) carries the displacement too."""
    from fem_glass_tempering_b200.output import read_series
    mesh = M.box_mesh(4, 4, 2, 4.0, 4.0, 2.0)
    config = {"T": {"element": "DG", "degree": 1}, "sigma": {"element": "DG", "degree": 1}}
    prob = ThermoViscoProblem(mesh_path="", time=(0.0, 50.0), dt=0.1, config=config, mesh=mesh, ctx=sg_ctx, verbose=False,
                              model_parameters=dict(vo.MAIN_PARAMS, mechanics=True, sip_penalty=8.0))
    prob.output_dir = str(tmp_path)
    prob.setup(dirichlet_bc=False)
    for _ in range(3):
        prob.t += prob.dt
        prob.solve_timestep(prob.t)
    prob._finalize()
    times, u = read_series(str(tmp_path), "displacement")
    assert len(times) == 4 and u.shape == (4, mesh.n_vertices * 3)      # initial state + three steps
    assert np.abs(u[0]).max() == 0.0
    np.testing.assert_array_equal(u[-1], prob.functions["displacement"].x.array.cpu().numpy())
    _, sig = read_series(str(tmp_path), "sigma")
    np.testing.assert_array_equal(np.nan_to_num(sig[-1]), np.nan_to_num(prob.functions_next["sigma"].x.array.cpu().numpy()))


def test_mechanics_off_is_the_reference_behaviour(sg_ctx):
    mesh = M.graded_line_mesh()
    config = {"T": {"element": "DG", "degree": 1}, "sigma": {"element": "CG", "degree": 1}}
    a = make_problem(sg_ctx, mesh, config, dict(vo.MAIN_PARAMS))
    b = make_problem(sg_ctx, mesh, config, dict(vo.MAIN_PARAMS, mechanics=False))
    assert a.mechanics is None and b.mechanics is None and "displacement" not in a.functions
    for prob in (a, b):
        prob.t += prob.dt
        prob.solve_timestep(prob.t)
    assert torch.equal(torch.nan_to_num(a.functions_next["sigma"].x.array), torch.nan_to_num(b.functions_next["sigma"].x.array))
