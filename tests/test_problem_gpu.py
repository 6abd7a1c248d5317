"""GPU end-to-end parity: ThermoViscoProblem (reference API, CUDA hot path) against the CPU oracle of the whole
time step (oracle/reference_problem.py).

Tolerances (north_star: 1e-10 relative for temperature, fictive temperature and stress histories), all with the DEFAULT
solver settings — the ones bench.py runs with (tools/parity_probe.py, profiles/r2a_parity_probe.jsonl is the data):
  * T, Tf, Tf_partial: 1e-12 relative asserted (measured 1e-14..1e-13); phi 1e-10;
  * xi: 1e-10 (it is proportional to the temperature INCREMENT of the step; measured <= 1.3e-11);
  * stress, per node:  |sigma_gpu - sigma_oracle| <= 1e-10 * max|sigma| + 2 * floor_i, where floor_i is the rounding noise
    of the reference's own formula lambda*(1 - taylor)/xi at that node (helpers.stress_rounding_floor): the
    sum (1.0 + a) is rounded to ulp(1) whatever the inputs' accuracy, so two evaluations whose xi differ in the last
    bits differ by up to that much, for ANY solver accuracy (tightening the solves 10x does not move the measured
    differences).  The norm-wise figure max|d sigma|/max|sigma| is returned and held to 1e-9 on top (measured 4e-12..4e-10);
  * nodes whose temperature did not change this step (|dT| below 1e-6 K: 0/0 = NaN or round-off noise in the
    reference too) are excluded from the stress comparison and must be NaN or negligible.
"""
import numpy as np
import pytest
import torch

from fem_glass_tempering_b200 import ThermoViscoProblem, fe
from fem_glass_tempering_b200 import mesh as msh
from helpers import assert_same, rel_err, stress_rounding_floor
from oracle.reference_problem import OracleProblem
from oracle.visco_oracle import MAIN_PARAMS

pytestmark = pytest.mark.gpu

MAIN_CONFIG = {"T": {"element": "DG", "degree": 1}, "sigma": {"element": "CG", "degree": 1}}   # main.py:24-27


def make_pair(sg_ctx, mesh, config, dt=0.1, params=MAIN_PARAMS, materialize="all"):
    prob = ThermoViscoProblem(mesh_path="", time=(0.0, 50.0), dt=dt, config=config, model_parameters=params,
                              jit_options={"cffi_extra_compile_args": ["-O3", "-march=native"]}, mesh=mesh,
                              ctx=sg_ctx, verbose=False, materialize=materialize)
    prob.setup(dirichlet_bc=False)
    sp = lambda s: dict(dofmap=s.dofmap, ref_nodes=s.element.nodes, family=s.family, degree=s.degree)
    T, S = prob.functionSpaces["T"].scalar, prob.functionSpaces["sigma"].scalar
    orc = OracleProblem(mesh.x, mesh.cells, sp(T), sp(S), params, dt)
    return prob, orc


def cpu(F):
    return F.x.array.cpu().numpy()


def compare_step(prob, orc, tol_T=1e-12, tol_inc=1e-10, tol_sigma=1e-10):
    f = orc.f
    assert rel_err(cpu(prob.functions_current["T"]), f["T_cur"]) <= tol_T
    assert rel_err(cpu(prob.functions_current["Tf"]), f["Tf_cur"]) <= tol_T
    assert rel_err(cpu(prob.functions_current["Tf_partial"]), f["Tf_partial_cur"]) <= tol_T
    assert rel_err(cpu(prob.functions["phi"]), f["phi"]) <= 100 * tol_T      # d(ln phi)/dT ~ 0.12/K amplifies T error
    assert rel_err(cpu(prob.functions["xi"]), f["xi"]) <= tol_inc
    # stress: exclude nodes whose temperature is stationary this step
    d = orc.d
    S = prob.functionSpaces["sigma"].scalar
    dT_at_S = np.abs(orc._T_at_sigma_points(f["T_cur"]) - orc._T_at_sigma_points(f["T_prev"]))
    xi_at_S = np.abs(orc._T_at_sigma_points(f["xi"]))
    node_dT, node_xi = np.zeros(S.n_nodes), np.zeros(S.n_nodes)
    node_dT[S.dofmap.ravel()] = dT_at_S
    node_xi[S.dofmap.ravel()] = xi_at_S
    good = node_dT > 1e-6
    sig_g = cpu(prob.functions_next["sigma"]).reshape(-1, d * d)
    sig_o = f["sigma_next"].reshape(-1, d * d)
    scale = np.max(np.abs(sig_o[good]))
    assert np.isfinite(sig_g[good]).all() and np.isfinite(sig_o[good]).all()
    err = np.max(np.abs(sig_g[good] - sig_o[good]), axis=1)
    floor = stress_rounding_floor(orc.vp, node_dT[good], node_xi[good])
    assert np.all(err <= tol_sigma * scale + 2.0 * floor), float(np.max(err - 2.0 * floor) / scale)
    assert np.max(err) <= 1e-9 * scale
    rest = sig_g[~good]
    assert np.all(np.isnan(rest) | (np.abs(rest) <= 1e-9))   # |sigma| ~ sum(k) * alpha_s * |dT| < 3.2e-10 there
    return float(np.max(err) / scale)


def test_main_py_default_run(sg_ctx):
    """BASELINE config 1: main.py — graded 1-D line, T = DG1, sigma = CG1, 60 of the 500 steps."""
    prob, orc = make_pair(sg_ctx, msh.graded_line_mesh(), MAIN_CONFIG)
    assert prob.n_steps == 500 and prob.dim == 1 and prob.mesh.n_cells == 48
    worst = 0.0
    for step in range(60):
        prob.t += prob.dt
        prob.solve_timestep(t=prob.t)
        orc.step()
        worst = max(worst, compare_step(prob, orc))
        orc.end_step()
        assert_same(cpu(prob.functions_previous["T"]), cpu(prob.functions_current["T"]), "T_prev <- T_cur (TVP:378)")
    print("worst stress error (default solver settings):", worst)


@pytest.mark.parametrize("dim,config,n", [
    (1, {"T": {"element": "CG", "degree": 1}, "sigma": {"element": "CG", "degree": 1}}, None),    # config 1 as worded
    (2, {"T": {"element": "CG", "degree": 2}, "sigma": {"element": "CG", "degree": 2}}, (12, 6)),   # config 2 shape
    (3, {"T": {"element": "DG", "degree": 1}, "sigma": {"element": "DG", "degree": 1}}, (6, 6, 3)),  # config 3 shape
    (3, {"T": {"element": "CG", "degree": 2}, "sigma": {"element": "CG", "degree": 2}}, (4, 4, 2)),  # config 4 shape
    (2, {"T": {"element": "CG", "degree": 2}, "sigma": {"element": "CG", "degree": 1}}, (8, 4)),    # P2 -> P1 gather
    (2, {"T": {"element": "DG", "degree": 1}, "sigma": {"element": "CG", "degree": 2}}, (8, 4)),    # P1 -> P2 gather
])
def test_other_configs(sg_ctx, dim, config, n):
    if dim == 1:
        m = msh.graded_line_mesh()
    elif dim == 2:
        m = msh.rectangle_mesh(*n, n[0] * 0.5, n[1] * 0.5)
    else:
        m = msh.box_mesh(*n, *(float(k) for k in n))
    prob, orc = make_pair(sg_ctx, m, config)
    for step in range(8):
        prob.solve_timestep(t=0.0)
        orc.step()
        compare_step(prob, orc)
        orc.end_step()


def test_phase_methods_equal_fused_step_and_full_materialisation(sg_ctx):
    """_solve_Tf/_solve_strains/_solve_shifted_time/_solve_stress (TVP:370-373) one by one == the fused launch,
    and every one of the reference's 24 Functions holds the oracle's value."""
    m = msh.rectangle_mesh(6, 4, 3.0, 2.0)
    cfg = {"T": {"element": "CG", "degree": 1}, "sigma": {"element": "CG", "degree": 1}}
    a, orc = make_pair(sg_ctx, m, cfg)
    b, _ = make_pair(sg_ctx, m, cfg)
    for step in range(3):
        a.solve_timestep(t=0.0)
        b._solve_T()
        b._solve_Tf()
        b._solve_strains()
        b._solve_shifted_time()
        b._solve_stress()
        b._update_values(current=b.functions_current["T"], previous=b.functions_previous["T"])
        orc.step()
        for grp in ("functions", "functions_current", "functions_previous", "functions_next"):
            for k in getattr(a, grp):
                assert_same(cpu(getattr(a, grp)[k]), cpu(getattr(b, grp)[k]), f"{grp}[{k}]")
        names = {"thermal_strain": "thermal_strain", "total_strain": "total_strain", "deviatoric_strain": "deviatoric_strain"}
        for k, o in names.items():
            assert rel_err(cpu(a.functions[k]), orc.f[o]) <= 1e-10
        assert rel_err(cpu(a.functions_next["T"]), orc.f["T_next"]) <= 1e-10
        orc.end_step()


def test_expression_dict_is_usable_like_dolfinx(sg_ctx):
    """Function.interpolate(material_model.expressions[k]) — the reference's call pattern (TVP:456-591) — gives
    the same arrays as the fused kernel (bit-exact except through exp)."""
    m = msh.rectangle_mesh(5, 3, 2.5, 1.5)
    cfg = {"T": {"element": "CG", "degree": 2}, "sigma": {"element": "CG", "degree": 2}}
    a, _ = make_pair(sg_ctx, m, cfg)
    b, _ = make_pair(sg_ctx, m, cfg)
    E = b.material_model.expressions
    assert set(E) == {"Tf_partial", "Tf", "thermal_strain", "total_strain", "deviatoric_strain", "T_next", "phi",
                      "phi_next", "xi", "ds_partial", "dsigma_partial", "s_tilde_partial_next",
                      "sigma_tilde_partial_next", "s_partial_next", "sigma_partial_next", "sigma_next"}
    for step in range(2):
        a.solve_timestep(t=0.0)
        b._solve_T()
        fc, fp, fn, f = b.functions_current, b.functions_previous, b.functions_next, b.functions
        f["phi"].interpolate(E["phi"])
        tmp = E["Tf_partial"]._fn().reshape(-1)            # cur and prev share a buffer here: evaluate, then store
        fc["Tf_partial"].x.array.copy_(tmp)
        fc["Tf"].interpolate(E["Tf"])
        for k in ("thermal_strain", "total_strain", "deviatoric_strain"):
            f[k].interpolate(E[k])
        fn["T"].interpolate(E["T_next"])
        f["phi"].interpolate(E["phi"])
        fn["phi"].interpolate(E["phi_next"])
        f["xi"].interpolate(E["xi"])
        f["ds_partial"].interpolate(E["ds_partial"])
        fn["s_tilde_partial"].interpolate(E["s_tilde_partial_next"])
        fn["s_partial"].interpolate(E["s_partial_next"])
        f["dsigma_partial"].interpolate(E["dsigma_partial"])
        fn["sigma_tilde_partial"].interpolate(E["sigma_tilde_partial_next"])
        fn["sigma_partial"].interpolate(E["sigma_partial_next"])
        fn["sigma"].interpolate(E["sigma_next"])
        b._update_values(current=fc["T"], previous=fp["T"])
        assert rel_err(cpu(b.functions_current["Tf"]), cpu(a.functions_current["Tf"])) <= 1e-14
        assert rel_err(cpu(b.functions["xi"]), cpu(a.functions["xi"])) <= 1e-10
        assert rel_err(cpu(b.functions_next["sigma"]), cpu(a.functions_next["sigma"])) <= 1e-8


def test_api_surface(sg_ctx, tmp_path):
    """Names and attributes the reference exposes."""
    from fem_glass_tempering_b200 import create_mesh
    path = str(tmp_path / "mesh1d.msh")
    create_mesh(path)                                                 # geometry.py:3
    prob = ThermoViscoProblem(mesh_path=path, config=MAIN_CONFIG, time=(0.0, 50.0), dt=0.1,
                              model_parameters=MAIN_PARAMS, jit_options=None, ctx=sg_ctx, verbose=False)
    for attr in ("mesh", "dim", "dt", "time", "t", "n_steps", "material_model", "physical_model", "finiteElements",
                 "functionSpaces", "functions", "functions_previous", "functions_current", "functions_next"):
        assert hasattr(prob, attr)
    prob.setup(dirichlet_bc=False)
    assert hasattr(prob, "F") and hasattr(prob, "solver") and prob.solver.convergence_criterion == "incremental"
    assert prob.solver.rtol == 1e-12
    assert prob.finiteElements["T"].family() == "Discontinuous Lagrange"
    mm, pm = prob.material_model, prob.physical_model
    assert mm.chi == 0.5 and mm.tableau_size == 6 and mm.dim == 1 and float(mm.Tb) == 869.0
    assert [float(getattr(pm, k)) for k in ("f", "epsilon", "sigma", "alpha", "htc", "rho", "cp", "k", "T_ambient")] == \
           [0.0, 0.93, 5.670e-8, 1.0, 280.1, 2500.0, 1433.0, 1.0, 600.0]
    assert set(prob.functions_current) == {"T", "Tf_partial", "Tf", "s_tilde_partial", "sigma_tilde_partial",
                                           "s_partial", "sigma_partial"}
    assert set(prob.functions_next) == {"T", "phi", "s_tilde_partial", "sigma_tilde_partial", "s_partial",
                                        "sigma_partial", "sigma"}
    assert float(prob.functions_current["Tf_partial"].x.array[7]) == 800.0           # TVP:224
    its, converged = prob.solver.solve(prob.functions_current["T"])
    assert converged and its >= 2
    with pytest.raises(AttributeError):
        prob.setup(dirichlet_bc=True)
    with pytest.raises(AssertionError):
        ThermoViscoProblem(mesh_path=path, config={"T": {"element": "RT", "degree": 1}, "sigma": MAIN_CONFIG["sigma"]},
                           time=(0, 1), dt=0.1, model_parameters=MAIN_PARAMS, ctx=sg_ctx)   # TVP:70-71


def test_minimal_materialisation_matches_full(sg_ctx):
    m = msh.box_mesh(4, 4, 2, 4.0, 4.0, 2.0)
    cfg = {"T": {"element": "DG", "degree": 1}, "sigma": {"element": "DG", "degree": 1}}
    a, _ = make_pair(sg_ctx, m, cfg, materialize="all")
    b, _ = make_pair(sg_ctx, m, cfg, materialize="minimal")
    for step in range(3):
        a.solve_timestep(t=0.0)
        b.solve_timestep(t=0.0)
    # the two runs differ only by the atomic summation order of the exterior-facet kernel in the thermal solve
    for k in ("sigma", "s_tilde_partial", "sigma_tilde_partial"):
        assert rel_err(cpu(a.functions_next[k]), cpu(b.functions_next[k])) <= 1e-10, k
    with pytest.raises(RuntimeError):
        b.functions["ds_partial"].x.array


def test_field_writer_records_every_step(sg_ctx, tmp_path):
    """output_dir switches on the per-step output of TVP:357-364: T, phi, Tf, xi, sigma captured after the
    viscoelastic update and before T_prev <- T_cur, through pinned double buffers on a side stream."""
    from fem_glass_tempering_b200.output import read_series
    prob = ThermoViscoProblem(mesh_path="", time=(0.0, 0.5), dt=0.1, config=MAIN_CONFIG, model_parameters=MAIN_PARAMS,
                              mesh=msh.graded_line_mesh(), ctx=sg_ctx, verbose=False)
    prob.output_dir = str(tmp_path / "output")
    prob.setup(dirichlet_bc=False)
    snaps = {k: [cpu(f).copy()] for k, f in (("T", prob.functions_current["T"]), ("sigma", prob.functions_next["sigma"]))}
    for _ in range(prob.n_steps):
        prob.t += prob.dt
        prob.solve_timestep(t=prob.t)
        snaps["T"].append(cpu(prob.functions_current["T"]).copy())
        snaps["sigma"].append(cpu(prob.functions_next["sigma"]).copy())
    prob._finalize()
    for key in ("T", "sigma"):
        t, v = read_series(prob.output_dir, key)
        assert len(t) == prob.n_steps + 1 and abs(t[-1] - 0.5) < 1e-12
        for a, b in zip(v, snaps[key]):
            assert_same(a, b, f"written {key}")
    t, xi = read_series(prob.output_dir, "xi")
    assert xi.shape == (prob.n_steps + 1, prob.functionSpaces["T"].n_nodes)


def test_chebyshev_solver_gives_the_same_histories(sg_ctx):
    """The polynomial-preconditioned DG solve (default on large meshes) lands on the same discrete solution."""
    cfg = {"T": {"element": "DG", "degree": 1}, "sigma": {"element": "DG", "degree": 1}}
    prob, orc = make_pair(sg_ctx, msh.box_mesh(6, 6, 3, 6.0, 6.0, 3.0), cfg)
    assert prob._thermal_op.set_chebyshev(3)
    for step in range(6):
        prob.solve_timestep(t=0.0)
        orc.step()
        compare_step(prob, orc)              # T, Tf, Tf_partial, phi, xi and the stress
        orc.end_step()
    assert prob._thermal_op.chebyshev_info()["degree"] == 3


def test_bench_path_combination_on_a_coercive_plate(sg_ctx):
    """What bench.py's headline runs — 32-cell class tiles of plate_mesh, sip_penalty 6.0, degree-4 Chebyshev with the
    Lanczos interval, minimal materialisation, default tolerances — against the oracle, stress included."""
    cfg = {"T": {"element": "DG", "degree": 1}, "sigma": {"element": "DG", "degree": 1}}
    params = dict(MAIN_PARAMS, sip_penalty=6.0)
    prob, orc = make_pair(sg_ctx, msh.plate_mesh(3, (12, 12, 4), (12.0, 12.0, 4.0)), cfg, params=params, materialize="minimal")
    assert prob._thermal_op.set_chebyshev(4)
    for step in range(8):
        prob.solve_timestep(t=0.0)
        orc.step()
        compare_step(prob, orc)
        orc.end_step()
    info = prob._thermal_op.chebyshev_info()
    assert info["degree"] == 4 and info["hi"] > info["lo"] > 0.0


@pytest.mark.parametrize("cheb", [4, 0])
def test_layer_grouped_plate_tiles_against_the_oracle(sg_ctx, cheb):
    """A plate whose columns hold a multiple of 32 cubes (8 x 4): mesh.py groups the bottom / interior / top layers so that
    most 32-cell tiles share their whole local-matrix class; whole time steps against the oracle with the Chebyshev solver
    (fused step kernel) and the plain one (apply kernel)."""
    cfg = {"T": {"element": "DG", "degree": 1}, "sigma": {"element": "DG", "degree": 1}}
    params = dict(MAIN_PARAMS, sip_penalty=6.0)
    prob, orc = make_pair(sg_ctx, msh.plate_mesh(3, (5, 8, 4), (5.0, 8.0, 4.0)), cfg, params=params, materialize="minimal")
    prob._thermal_op.set_chebyshev(cheb)
    for step in range(5):
        prob.solve_timestep(t=0.0)
        orc.step()
        compare_step(prob, orc)
        orc.end_step()


def test_corrected_physics_through_the_problem_api(sg_ctx):
    """model_params["physics"] = "corrected" (an extension, see ViscoelasticModel): the problem-level run equals the CPU
    statement of that scheme fed with the GPU's own temperature history; stresses carry memory and stay finite."""
    from oracle import visco_oracle as vo
    cfg = {"T": {"element": "CG", "degree": 1}, "sigma": {"element": "CG", "degree": 1}}
    params = dict(MAIN_PARAMS, physics="corrected")
    prob = ThermoViscoProblem(mesh_path="", time=(0.0, 50.0), dt=0.1, config=cfg, model_parameters=params,
                              mesh=msh.graded_line_mesh(), ctx=sg_ctx, verbose=False, materialize="minimal")
    prob.setup(dirichlet_bc=False)
    n = prob.functionSpaces["T"].n_nodes
    p = vo.ViscoParams(dim=1, dt=0.1)
    o = dict(Tfp=cpu(prob.functions_current["Tf_partial"]).copy(), Tf=cpu(prob.functions_current["Tf"]).copy(),
             s=np.zeros(n * 6), k=np.zeros(n * 6), phi=np.zeros(n), xi=np.zeros(n), sig=np.zeros(n))
    for step in range(12):
        T_prev = cpu(prob.functions_previous["T"]).copy()
        prob.solve_timestep(t=0.0)
        T_cur = cpu(prob.functions_current["T"]).copy()
        vo.step_corrected(p, 0.5, T_cur, T_prev, o["Tfp"], o["Tf"], o["phi"], o["xi"], o["s"], o["k"], o["sig"])
        assert rel_err(cpu(prob.functions_current["Tf"]), o["Tf"]) <= 1e-12
        assert rel_err(cpu(prob.functions["xi"]), o["xi"]) <= 1e-12
        sig = cpu(prob.functions_next["sigma"])
        assert np.isfinite(sig).all()
        assert rel_err(sig, o["sig"]) <= 1e-11
    with pytest.raises(NotImplementedError):
        prob._solve_Tf()


def test_default_run_against_the_hand_evaluated_history(sg_ctx):
    """ThermoViscoProblem with main.py's configuration against tests/golden/main_py_history.json (first five steps of the
    reference's default run, hand-evaluated in plain Python: closed-form P1 matrices, dense Newton solves, the 16
    expressions, last-cell-wins DG1 -> CG1).  T, Tf: 1e-12 relative; stress where the temperature moved: 1e-9 norm-wise
    (the fixture's own stress agrees with the oracle to 1e-11; the per-node bound with the rounding floor is asserted
    against the oracle in test_main_py_default_run)."""
    from helpers import load_main_py_history
    g = load_main_py_history()
    prob = ThermoViscoProblem(mesh_path="", time=(0.0, 50.0), dt=g["dt"], config=MAIN_CONFIG, model_parameters=MAIN_PARAMS,
                              mesh=msh.graded_line_mesh(), ctx=sg_ctx, verbose=False)
    prob.setup(dirichlet_bc=False)
    for st in g["steps"]:
        prob.solve_timestep(t=0.0)
        assert rel_err(cpu(prob.functions_current["T"]), st["T"]) <= 1e-12
        assert rel_err(cpu(prob.functions_current["Tf"]), st["Tf"]) <= 1e-12
        moved = np.abs(st["T"] - st["T_prev"])[g["winner_dof"]] > 1e-6
        sig = cpu(prob.functions_next["sigma"])
        assert np.max(np.abs(sig[moved] - st["sigma"][moved])) <= 1e-9 * np.max(np.abs(st["sigma"][moved]))


def test_extrapolated_newton_start_reaches_the_same_solution(sg_ctx):
    """model_parameters["newton_initial_guess"] = "extrapolate" starts Newton from 2 T_n - T_{n-1} instead of T_n (TVP:389):
    a different path to the same discrete solution — histories must agree to the solver tolerance."""
    mesh = msh.box_mesh(6, 6, 3, 6.0, 6.0, 3.0)
    cfg = {"T": {"element": "DG", "degree": 1}, "sigma": {"element": "DG", "degree": 1}}
    runs = {}
    for guess in ("previous", "extrapolate"):
        prob = ThermoViscoProblem(mesh_path="", time=(0.0, 50.0), dt=0.1, config=cfg, mesh=mesh, ctx=sg_ctx, verbose=False,
                                  model_parameters=dict(MAIN_PARAMS, newton_initial_guess=guess))
        prob.setup(dirichlet_bc=False)
        its = 0
        for _ in range(6):
            prob.t += prob.dt
            prob.solve_timestep(prob.t)
            its += prob.solver.last_stats.lin_its
        runs[guess] = (cpu(prob.functions_current["T"]), cpu(prob.functions_current["Tf"]), its)
    assert rel_err(runs["extrapolate"][0], runs["previous"][0]) <= 1e-12
    assert rel_err(runs["extrapolate"][1], runs["previous"][1]) <= 1e-12
    assert runs["extrapolate"][2] <= runs["previous"][2]
    with pytest.raises(ValueError):
        ThermoViscoProblem(mesh_path="", time=(0.0, 1.0), dt=0.1, config=cfg, mesh=mesh, ctx=sg_ctx, verbose=False,
                           model_parameters=dict(MAIN_PARAMS, newton_initial_guess="secant"))
