"""Shared helpers for the parity tests (test infrastructure)."""
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def unhex(v):
    if isinstance(v, list):
        return [unhex(x) for x in v]
    return float.fromhex(v)


def load_visco_kat():
    with open(os.path.join(GOLDEN, "visco_kat.json")) as fh:
        data = json.load(fh)
    cases = []
    for c in data["cases"]:
        cases.append(dict(name=c["name"], d=c["d"], inputs={k: unhex(v) for k, v in c["inputs"].items()},
                          expected={k: unhex(v) for k, v in c["expected"].items()}))
    return cases


def assert_same(a, b, what=""):
    """Exact float64 equality (0.0 == -0.0), NaNs must sit at the same positions."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    nan_a, nan_b = np.isnan(a), np.isnan(b)
    assert np.array_equal(nan_a, nan_b), f"{what}: NaN positions differ"
    ok = (a == b) | nan_a
    if not ok.all():
        i = int(np.argmin(ok.ravel()))
        raise AssertionError(f"{what}: {int((~ok).sum())} of {a.size} values differ; first at flat index {i}: "
                             f"{a.ravel()[i]!r} vs {b.ravel()[i]!r}")


def rel_err(a, b):
    """max |a-b| / max |b| over finite entries (NaN positions must match)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert np.array_equal(np.isnan(a), np.isnan(b)), "NaN positions differ"
    m = ~np.isnan(b)
    if not m.any():
        return 0.0
    scale = np.max(np.abs(b[m]))
    return float(np.max(np.abs(a[m] - b[m])) / (scale if scale > 0 else 1.0))


def random_visco_state(n, d, N=6, seed=1234):
    """SURVEY §8(d) micro-benchmark state."""
    rng = np.random.default_rng(seed)
    T_cur = rng.uniform(650.0, 850.0, n)
    T_prev = T_cur + rng.uniform(0.05, 1.0, n)
    Tfp = np.repeat(T_prev, N) + rng.uniform(0.0, 5.0, n * N)
    s = rng.normal(0.0, 1e-3, n * N * d * d)
    k = rng.normal(0.0, 1e-3, n * N * d * d)
    return T_cur, T_prev, Tfp, s, k


def load_thermal_kat():
    """tests/golden/thermal_kat.json: hand-evaluated residual / Jacobian-vector product of the heat equation on the
    reference's graded 1-D line (generator: tests/golden/make_thermal_kat.py)."""
    with open(os.path.join(GOLDEN, "thermal_kat.json")) as fh:
        data = json.load(fh)
    arr = lambda v: np.array(unhex(v))
    cases = [dict(family=c["family"], degree=c["degree"], T=arr(c["T"]), T_prev=arr(c["T_prev"]), x=arr(c["x"]),
                  residual=arr(c["residual"]), jac_x=arr(c["jac_x"])) for c in data["cases"]]
    return dict(dt=data["dt"], points=arr(data["points"]), cases=cases)


def load_thermal_kat_simplex():
    """tests/golden/thermal_kat_2d3d.json: hand-evaluated residual / Jacobian-vector product on small perturbed 2-D and 3-D
    meshes, DG1 and CG1 (generator: tests/golden/make_thermal_kat_simplex.py).  The meshes come from the file."""
    with open(os.path.join(GOLDEN, "thermal_kat_2d3d.json")) as fh:
        data = json.load(fh)
    arr = lambda v: np.array(unhex(v))
    cases = [dict(dim=c["dim"], family=c["family"], degree=c["degree"], x=np.array([unhex(p) for p in c["x"]]),
                  cells=np.array(c["cells"]), T=arr(c["T"]), T_prev=arr(c["T_prev"]), v=arr(c["v"]),
                  residual=arr(c["residual"]), jac_x=arr(c["jac_x"])) for c in data["cases"]]
    return dict(dt=data["dt"], cases=cases)


def load_main_py_history():
    """tests/golden/main_py_history.json: hand-evaluated first time steps of the reference's default run
    (generator: tests/golden/make_main_py_history.py)."""
    with open(os.path.join(GOLDEN, "main_py_history.json")) as fh:
        data = json.load(fh)
    arr = lambda v: np.array([float("nan") if q == "nan" else float.fromhex(q) for q in v])
    steps = [{k: (arr(v) if isinstance(v, list) else v) for k, v in s.items()} for s in data["steps"]]
    return dict(dt=data["dt"], T_0=data["T_0"], points=arr(data["points"]), winner_dof=np.array(data["winner_dof"]), steps=steps)


def load_thermal_kat_p2():
    """tests/golden/thermal_kat_p2.json: hand-evaluated CG2 residual / Jacobian-vector product on small perturbed 2-D and
    3-D meshes (generator: tests/golden/make_thermal_kat_p2.py); includes the P2 dofmap it assumed."""
    with open(os.path.join(GOLDEN, "thermal_kat_p2.json")) as fh:
        data = json.load(fh)
    arr = lambda v: np.array(unhex(v))
    cases = [dict(dim=c["dim"], family=c["family"], degree=c["degree"], x=np.array([unhex(p) for p in c["x"]]),
                  cells=np.array(c["cells"]), dofmap=np.array(c["dofmap"]), T=arr(c["T"]), T_prev=arr(c["T_prev"]),
                  v=arr(c["v"]), residual=arr(c["residual"]), jac_x=arr(c["jac_x"])) for c in data["cases"]]
    return dict(dt=data["dt"], cases=cases)


def stress_rounding_floor(vp, dT_nodes, xi_nodes):
    """Rounding noise of the reference's own stress formula, per sigma node (absolute, same unit as sigma).

    VM:185-191 evaluates  k_n tr(eps) / xi * lambda_n * (1.0 - taylor_n)  with  taylor_n = (1.0 + a) + 0.5 a^2,
    a = -xi/lambda_n (VM:233-242): the sum (1.0 + a) is rounded to ulp(1) = 2.2e-16 ABSOLUTE, so the factor
    (1 - taylor_n) ~ |a| carries a relative error of 2.2e-16/|a| whatever the inputs' accuracy.  Two
    evaluations whose xi differ in the last bits draw that error independently; their stresses can therefore differ by
    up to twice   sum_n |k_n tr(eps)| * min(1, 2.2e-16 * lambda_n / |xi|)   (the deviator of the isotropic strain is
    round-off itself and adds nothing visible).  tr(eps) = -dim * alpha_solid * dT (VM:128-139).
    Parity tests hold the stress to  max(1e-10 * max|sigma|, 2 * this floor)  — the first is north_star's bar, the second
    the part of the difference no solver accuracy can remove.
    """
    import numpy as np
    dT = np.abs(np.asarray(dT_nodes, dtype=np.float64))
    xi = np.abs(np.asarray(xi_nodes, dtype=np.float64))
    tr = vp.dim * vp.alpha_solid * dT
    floor = np.zeros_like(dT)
    with np.errstate(divide="ignore", invalid="ignore"):
        for k_n, lam_n in zip(vp.k, vp.lambda_k):
            # once |xi/lambda_n| < ulp the factor quantises to 0 or ulp: the term is wrong by at most itself
            floor += abs(k_n) * tr * np.minimum(1.0, 2.220446049250313e-16 * lam_n / xi)
    return floor
