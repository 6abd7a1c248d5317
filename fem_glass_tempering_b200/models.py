"""ThermalModel and ViscoelasticModel with the reference's constructor signatures and attributes
(/root/reference/ThermalModel.py:6-29, /root/reference/ViscoelasticModel.py:9-242).

The models hold constants; the arithmetic of the viscoelastic chain lives in csrc/visco.cu and is reached
through `ViscoelasticModel.plan` (sg_visco_plan).  `expressions[...]` keeps the reference's 16 keys; each
entry can be passed to Function.interpolate like a dolfinx Expression (a slow, reference-shaped path made of
elementwise torch float64 operations in the source's association order — the time loop never uses it).
"""
from __future__ import annotations

from math import factorial

import numpy as np

from .function import Constant

# ViscoelasticModel.py:19-68
_PRONY = dict(
    m_n=(5.523e-2, 8.205e-2, 1.215e-1, 2.286e-1, 2.860e-1, 2.265e-1),
    lambda_m_n=(5.965e-4, 1.077e-2, 1.362e-1, 1.505e-1, 6.747e+0, 2.963e+1),
    g_n=(1.585, 2.354, 3.486, 6.558, 8.205, 6.498),
    lambda_g_n=(6.658e-5, 1.197e-3, 1.514e-2, 1.672e-1, 7.497e-1, 3.292e+0),
    k_n=(7.588e-1, 7.650e-1, 9.806e-1, 7.301e+0, 1.347e+1, 1.090e+1),
    lambda_k_n=(5.009e-5, 9.945e-4, 2.022e-3, 1.925e-2, 1.199e-1, 2.033e+0),
)


def prony_tables(n_terms: int) -> dict:
    """Tables for the Prony-term sweep (BASELINE config 5): the first N reference entries for N <= 6,
    log-spaced relaxation times in [1e-5, 1e2] with normalised weights for N > 6."""
    if n_terms <= 6:
        return {k: v[:n_terms] for k, v in _PRONY.items()}
    lam = tuple(float(v) for v in np.logspace(-5, 2, n_terms))
    w = np.linspace(1.0, 2.0, n_terms)
    w = w / w.sum()
    return dict(m_n=tuple(float(v) for v in w), lambda_m_n=lam,
                g_n=tuple(float(v) for v in w * sum(_PRONY["g_n"])), lambda_g_n=lam,
                k_n=tuple(float(v) for v in w * sum(_PRONY["k_n"])), lambda_k_n=lam)


class ThermalModel:
    """ThermalModel.py:6-29 — the nine constants of the heat equation."""

    def __init__(self, mesh, model_parameters: dict) -> None:
        p = model_parameters
        self.f = Constant(mesh, p["f"])
        self.epsilon = Constant(mesh, p["epsilon"])
        self.sigma = Constant(mesh, p["sigma"])
        self.alpha = Constant(mesh, p["alpha"])
        self.htc = Constant(mesh, p["htc"])
        self.rho = Constant(mesh, p["rho"])      # stored, unused by the weak form
        self.cp = Constant(mesh, p["cp"])
        self.k = Constant(mesh, p["k"])
        self.T_ambient = Constant(mesh, p["T_ambient"])


class PointwiseExpression:
    """One entry of ViscoelasticModel.expressions (the reference builds dolfinx Expressions, VM:100-228)."""

    def __init__(self, name: str, space_key: str, fn):
        self.name, self.space_key, self._fn = name, space_key, fn

    def evaluate_into(self, target) -> None:
        target.x.array.copy_(self._fn().reshape(-1))


class ViscoelasticModel:
    """ViscoelasticModel.py:9-84; optional model_parameters["prony"] = dict of six tables or an int N, and
    model_parameters["physics"] = "reference" (default: the reference's expressions exactly as it executes them, quirks
    Q1-Q5 of SURVEY 3.5 included) | "corrected" (the scheme its comments cite: Eq. 25 shift function with chi and the
    old fictive temperature, structural strain term, trapezoidal shifted time, exact exponentials, history = partial
    stress; csrc/visco.cu).  The corrected scheme is an extension: it needs equal T and sigma spaces and runs only as
    the fused update of solve_timestep."""

    def __init__(self, mesh, model_parameters: dict) -> None:
        self.chi = float(model_parameters.get("chi", 0.5))   # VM:15 (unused by the reference at run time)
        self.physics = model_parameters.get("physics", "reference")
        if self.physics not in ("reference", "corrected"):
            raise ValueError('model_parameters["physics"] must be "reference" or "corrected"')
        self.dim = mesh.topology.dim
        tabs = model_parameters.get("prony", 6)
        if isinstance(tabs, int):
            tabs = prony_tables(tabs)
        self.tableau_size = len(tabs["m_n"])             # VM:16
        self.m_n_tableau = Constant(mesh, tabs["m_n"])
        self.lambda_m_n_tableau = Constant(mesh, tabs["lambda_m_n"])
        self.g_n_tableau = Constant(mesh, tabs["g_n"])
        self.lambda_g_n_tableau = Constant(mesh, tabs["lambda_g_n"])
        self.k_n_tableau = Constant(mesh, tabs["k_n"])
        self.lambda_k_n_tableau = Constant(mesh, tabs["lambda_k_n"])
        self.I = np.eye(self.dim)
        self.T_init = Constant(mesh, model_parameters["T_0"])
        self.H = Constant(mesh, model_parameters["H"])
        self.Rg = Constant(mesh, model_parameters["Rg"])
        self.Tb = Constant(mesh, model_parameters["Tb"])
        self.alpha_solid = Constant(mesh, model_parameters["alpha_solid"])
        self.alpha_liquid = Constant(mesh, model_parameters["alpha_liquid"])
        self.plan = None
        self.expressions = {}

    def make_plan(self, ctx, dt: float):
        """sg_visco_plan for these constants (created by ThermoViscoProblem once a GPU context exists)."""
        from . import _lib
        self.plan = _lib.ViscoPlan(
            ctx, dim=self.dim, dt=dt, H=float(self.H), Rg=float(self.Rg), Tb=float(self.Tb),
            alpha_solid=float(self.alpha_solid), alpha_liquid=float(self.alpha_liquid),
            m=list(self.m_n_tableau), lambda_m=list(self.lambda_m_n_tableau), g=list(self.g_n_tableau),
            lambda_g=list(self.lambda_g_n_tableau), k=list(self.k_n_tableau), lambda_k=list(self.lambda_k_n_tableau),
            mode=_lib.VISCO_CORRECTED if self.physics == "corrected" else _lib.VISCO_REFERENCE, chi=self.chi)
        return self.plan

    # -- reference-shaped expressions -------------------------------------------------------------
    def _init_expressions(self, functions: dict, functions_next: dict, functions_current: dict,
                          functions_previous: dict, functionSpaces: dict, dt: float, to_sigma=None) -> None:
        """VM:86-230.  `to_sigma(array)` evaluates a T-space array at the sigma nodes (identity when the two
        spaces coincide; the last-cell-wins gather of SURVEY Q13 otherwise)."""
        import torch
        N, d = self.tableau_size, self.dim
        f, fn, fc, fp = functions, functions_next, functions_current, functions_previous
        ts = to_sigma or (lambda a: a)
        arr = lambda F: F.x.array
        H_Rg = float(self.H) / float(self.Rg)
        inv_Tb = 1.0 / float(self.Tb)
        lm = torch.tensor(list(self.lambda_m_n_tableau), dtype=torch.float64)
        eye = torch.eye(d, dtype=torch.float64)

        def on(t, like):
            return t.to(like.device)

        def phi_of(Tarr):
            return torch.exp(H_Rg * (inv_Tb - 1.0 / Tarr))

        def tf_partial():
            T, ph = arr(fc["T"]), arr(f["phi"])
            prev = arr(fp["Tf_partial"]).view(-1, N)
            l = on(lm, T)
            return (l * prev + ((T * dt) * ph)[:, None]) / (l + (dt * ph)[:, None])

        def tf():
            cur = arr(fc["Tf_partial"]).view(-1, N)
            m = list(self.m_n_tableau)
            acc = m[0] * cur[:, 0]
            for i in range(1, N):
                acc = acc + m[i] * cur[:, i]
            return acc

        def thermal_strain():
            a_s, a_l = float(self.alpha_solid), float(self.alpha_liquid)
            s = a_s * (ts(arr(fc["T"])) - ts(arr(fp["T"]))) + (a_l - a_s) * (ts(arr(fc["Tf"])) - ts(arr(fp["Tf"])))
            return s[:, None, None] * on(eye, s)

        def deviatoric():
            tot = arr(f["total_strain"]).view(-1, d, d)
            tr = tot[:, 0, 0]
            for i in range(1, d):
                tr = tr + tot[:, i, i]
            return tot - (1 / d * tr)[:, None, None] * on(eye, tot)

        def taylor(xi, lam):
            a = (-1.0 * xi) / lam
            return (1.0 + a) + 0.5 * (a * a)

        def partial(kind):
            xi = ts(arr(f["xi"]))
            out = []
            if kind == "ds":
                dev = arr(f["deviatoric_strain"]).view(-1, d, d)
                for lam, g in zip(self.lambda_g_n_tableau, self.g_n_tableau):
                    out.append((2.0 * g * dev) / xi[:, None, None] * lam * (1.0 - taylor(xi, lam))[:, None, None])
            else:
                tot = arr(f["total_strain"]).view(-1, d, d)
                tr = tot[:, 0, 0]
                for i in range(1, d):
                    tr = tr + tot[:, i, i]
                for lam, k in zip(self.lambda_k_n_tableau, self.k_n_tableau):
                    v = (k * tr) / xi * lam * (1.0 - taylor(xi, lam))
                    out.append(v[:, None, None] * on(eye, v))
            return torch.stack(out, dim=1)

        def tilde(kind):
            xi = ts(arr(f["xi"]))
            src = arr(fc["s_tilde_partial" if kind == "g" else "sigma_tilde_partial"]).view(-1, N, d, d)
            lams = self.lambda_g_n_tableau if kind == "g" else self.lambda_k_n_tableau
            return torch.stack([src[:, n] * taylor(xi, lam)[:, None, None] for n, lam in enumerate(lams)], dim=1)

        def sigma_next():
            s = arr(fn["s_partial"]).view(-1, N, d, d)
            k = arr(fn["sigma_partial"]).view(-1, N, d, d)
            acc = s[:, 0] + k[:, 0]
            for n in range(1, N):
                acc = acc + (s[:, n] + k[:, n])
            return acc

        E = PointwiseExpression
        self.expressions = {
            "Tf_partial": E("Tf_partial", "Tf_partial", tf_partial),                                     # VM:111
            "Tf": E("Tf", "T", tf),                                                                      # VM:122
            "thermal_strain": E("thermal_strain", "sigma", thermal_strain),                              # VM:128
            "total_strain": E("total_strain", "sigma", lambda: -1.0 * arr(f["thermal_strain"])),         # VM:136
            "deviatoric_strain": E("deviatoric_strain", "sigma", deviatoric),                            # VM:142
            "T_next": E("T_next", "T", lambda: arr(fc["T"]) + (arr(fc["T"]) - arr(fp["T"]))),            # VM:150
            "phi": E("phi", "T", lambda: phi_of(arr(fc["T"]))),                                          # VM:156 (live def.)
            "phi_next": E("phi_next", "T", lambda: phi_of(arr(fn["T"]))),                                # VM:162
            "xi": E("xi", "T", lambda: dt / 2 * (arr(fn["phi"]) - arr(f["phi"]))),                       # VM:170
            "ds_partial": E("ds_partial", "sigma_partial", lambda: partial("ds")),                       # VM:176
            "dsigma_partial": E("dsigma_partial", "sigma_partial", lambda: partial("dsigma")),           # VM:185
            "s_tilde_partial_next": E("s_tilde_partial_next", "sigma_partial", lambda: tilde("g")),      # VM:195
            "sigma_tilde_partial_next": E("sigma_tilde_partial_next", "sigma_partial", lambda: tilde("k")),  # VM:203
            "s_partial_next": E("s_partial_next", "sigma_partial",
                                lambda: arr(f["ds_partial"]) + arr(fn["s_tilde_partial"])),              # VM:212
            "sigma_partial_next": E("sigma_partial_next", "sigma_partial",
                                    lambda: arr(f["dsigma_partial"]) + arr(fn["sigma_tilde_partial"])),  # VM:218
            "sigma_next": E("sigma_next", "sigma", sigma_next),                                          # VM:224
        }

    def _taylor_exponential(self, functions: dict, lambda_value):
        """VM:233-242: sum_{k<3} 1/k! (-xi/lambda)^k on the xi array."""
        xi = functions["xi"].x.array
        a = (-1.0 * xi) / float(lambda_value)
        return (1.0 / factorial(0) + a) + 1.0 / factorial(2) * (a * a)
