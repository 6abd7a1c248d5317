"""CPU: the C oracle for hot path (A) against the hand-evaluated golden vectors."""
import numpy as np
import pytest

from helpers import assert_same, load_visco_kat, random_visco_state
from oracle import visco_oracle as vo

CASES = load_visco_kat()


def _run_case(c):
    d, inp = c["d"], c["inputs"]
    p = vo.ViscoParams(dim=d, dt=inp["dt"])
    st = vo.new_state(p, 1)
    st["T_cur"][:] = inp["T_cur"]
    st["T_prev"][:] = inp["T_prev"]
    st["Tf_partial_prev"][:] = inp["Tfp_prev"]
    st["Tf_partial_cur"][:] = inp["Tfp_prev"]
    st["s_tilde_cur"][:] = np.ravel(inp["s_tilde"])
    st["sigma_tilde_cur"][:] = np.ravel(inp["sigma_tilde"])
    vo.step_passes(p, st)
    return st


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_passes_match_golden(case):
    st, exp = _run_case(case), case["expected"]
    d = case["d"]
    assert_same(st["phi"], [exp["phi"]], "phi")
    assert_same(st["Tf_partial_cur"], exp["Tf_partial"], "Tf_partial")
    assert_same(st["Tf_partial_prev"], exp["Tf_partial"], "Tf_partial_prev (TVP:469)")
    assert_same(st["Tf_cur"], [exp["Tf"]], "Tf")
    assert_same(st["Tf_prev"], [exp["Tf"]], "Tf_prev (TVP:481)")
    assert_same(st["thermal_strain"][:: d + 1], [exp["thermal_strain"]] * d, "thermal strain diagonal")
    assert_same(st["T_next"], [exp["T_next"]], "T_next")
    assert_same(st["phi_next"], [exp["phi_next"]], "phi_next")
    assert_same(st["xi"], [exp["xi"]], "xi")
    assert_same(st["s_tilde_cur"], np.ravel(exp["s_tilde"]), "s_tilde")
    assert_same(st["sigma_tilde_cur"], np.ravel(exp["sigma_tilde"]), "sigma_tilde")
    assert_same(st["s_partial_next"], np.ravel(exp["s_partial"]), "s_partial")
    assert_same(st["sigma_partial_next"], np.ravel(exp["sigma_partial"]), "sigma_partial")
    assert_same(st["sigma_next"], exp["sigma"], "sigma")


def test_survey_kat_decimal_values():
    """The decimal known answers quoted in SURVEY.md §4(1)."""
    p = vo.ViscoParams(dim=3, dt=0.1)
    assert vo.phi(p, np.array([790.0, 800.0, 869.0])).tolist() == [1.6835520133935624e-04, 5.560583791350996e-04, 1.0]
    c = next(c for c in CASES if c["name"] == "survey_kat_d3")
    st = _run_case(c)
    assert st["Tf_partial_cur"].tolist() == [799.725508802584, 799.9843925309675, 799.9987640646569,
                                             799.9988814859098, 799.999975047461, 799.9999943180862]
    assert st["Tf_cur"][0] == 799.8871449764878
    assert st["thermal_strain"][0] == -9.099999999999999e-05
    assert st["T_next"][0] == 780.0
    assert st["xi"][0] == -5.946048890740225e-06
    assert st["sigma_next"][0] == 9.343600766769931e-03
    for d, v in ((1, 3.11453358892331e-03), (2, 6.22906717784662e-03)):
        assert _run_case(next(c for c in CASES if c["name"] == f"survey_kat_d{d}"))["sigma_next"][0] == v


def test_table_sums():
    """SURVEY §8(c) analytic checks on the tableaux."""
    assert abs(sum(vo.PRONY_M) - 0.99988) < 1e-12
    assert abs(sum(vo.PRONY_G) - 28.686) < 1e-12
    assert abs(sum(vo.PRONY_K) - 34.1754) < 1e-12


@pytest.mark.parametrize("d", [1, 2, 3])
@pytest.mark.parametrize("N", [3, 6, 12])
def test_fused_equals_passes(d, N):
    """The one-sweep restatement is bit-identical to the 17-pass replay over several steps."""
    n = 257
    p = vo.ViscoParams(dim=d, dt=0.1, **vo.prony_tables(N))
    T_cur, T_prev, Tfp, s, k = random_visco_state(n, d, N)
    T_cur[:5] = T_prev[:5]  # NaN rows (Q5)
    st = vo.new_state(p, n)
    st["T_cur"][:], st["T_prev"][:] = T_cur, T_prev
    st["Tf_partial_prev"][:] = Tfp
    st["Tf_partial_cur"][:] = Tfp
    st["s_tilde_cur"][:], st["sigma_tilde_cur"][:] = s, k
    Tf, ph, x, sig = np.zeros(n), np.zeros(n), np.zeros(n), np.zeros(n * d * d)
    fT_cur, fT_prev = T_cur.copy(), T_prev.copy()
    for _ in range(3):
        vo.step_passes(p, st)
        vo.step_fused(p, fT_cur, fT_prev, Tfp, Tf, ph, x, s, k, sig)
        assert_same(Tfp, st["Tf_partial_cur"], "Tf_partial")
        assert_same(Tf, st["Tf_cur"], "Tf")
        assert_same(x, st["xi"], "xi")
        assert_same(s, st["s_tilde_cur"], "s_tilde")
        assert_same(k, st["sigma_tilde_cur"], "sigma_tilde")
        assert_same(sig, st["sigma_next"], "sigma")
        # advance T like TVP:378 after a (fake) thermal solve
        new_T = st["T_cur"] - 0.3
        st["T_prev"][:] = st["T_cur"]
        st["T_cur"][:] = new_T
        fT_prev[:] = fT_cur
        fT_cur[:] = new_T


def test_memoryless_stress_from_zero_history():
    """with zero s_tilde/sigma_tilde the history stays zero and sigma = sum(ds + dsigma)."""
    p = vo.ViscoParams(dim=3, dt=0.1)
    st = vo.new_state(p, 16)
    st["T_cur"][:] = np.linspace(700, 799, 16)
    vo.step_passes(p, st)
    assert not st["s_tilde_cur"].any() and not st["sigma_tilde_cur"].any()
    assert_same(st["s_partial_next"], st["ds_partial"])
    assert_same(st["sigma_partial_next"], st["dsigma_partial"])


def test_nan_where_T_unchanged():
    """T_cur == T_prev gives xi = 0 and 0/0 = NaN in every stress component."""
    p = vo.ViscoParams(dim=2, dt=0.1)
    st = vo.new_state(p, 4)
    st["T_cur"][1:] = 790.0
    vo.step_passes(p, st)
    assert st["xi"][0] == 0.0
    assert np.isnan(st["sigma_next"][:4]).all()
    assert np.isfinite(st["sigma_next"][4:]).all()


def test_whole_step_oracle_against_the_hand_evaluated_default_run():
    """oracle/reference_problem.py (assembled heat solve + 17-pass replay + last-cell-wins DG1 -> CG1 interpolation) against
    tests/golden/main_py_history.json: the first five steps of main.py's default run evaluated by hand in plain Python.
    T and Tf 1e-12; stress 1e-9 where the node's temperature moved (its formula cancels)."""
    from fem_glass_tempering_b200 import fe
    from fem_glass_tempering_b200 import mesh as msh
    from helpers import load_main_py_history
    from oracle.reference_problem import OracleProblem
    g = load_main_py_history()
    m = msh.graded_line_mesh()
    T, S = fe.ScalarSpace(m, "DG", 1), fe.ScalarSpace(m, "CG", 1)
    sp = lambda s: dict(dofmap=s.dofmap, ref_nodes=s.element.nodes, family=s.family, degree=s.degree)
    orc = OracleProblem(m.x, m.cells, sp(T), sp(S), vo.MAIN_PARAMS, g["dt"])
    for st in g["steps"]:
        orc.step()
        assert np.max(np.abs(orc.f["T_cur"] - st["T"])) <= 1e-12 * 800
        assert np.max(np.abs(orc.f["Tf_cur"] - st["Tf"])) <= 1e-12 * 800
        moved = np.abs(st["T"] - st["T_prev"])[g["winner_dof"]] > 1e-6
        so, sg = orc.f["sigma_next"].reshape(-1), st["sigma"]
        assert moved.sum() >= 20 and np.max(np.abs(so[moved] - sg[moved])) <= 1e-9 * np.max(np.abs(sg[moved]))
        orc.end_step()
