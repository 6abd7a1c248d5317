"""TEST INFRASTRUCTURE — CPU restatement of the whole per-timestep path of the reference:
ThermoViscoProblem.solve_timestep (/root/reference/ThermoViscoProblem.py:367-381) on top of
oracle/thermal_oracle.py (hot path B) and oracle/visco_oracle.c (hot path A).

dolfinx's Function.interpolate(Expression) is restated literally [dolfinx-recall, SURVEY §3.3]: every
expression is evaluated CELL BY CELL at the target element's interpolation points — coefficients living in
another space are evaluated through that space's basis inside the cell — and the results are scattered
through the target dofmap in cell order, later cells overwriting earlier ones.

PARITY UNPINNED (see the headers of the two oracle modules).  Only tests/, smoke() and bench.py's CPU legs
may import this module.
"""
from __future__ import annotations

import numpy as np

from . import thermal_oracle as to
from . import visco_oracle as vo


class OracleProblem:
    def __init__(self, x, cells, T_space, S_space, params: dict, dt: float, prony: dict | None = None):
        """T_space / S_space: dicts(dofmap, ref_nodes, family, degree) of the T and sigma node sets."""
        self.x, self.cells = np.asarray(x, dtype=np.float64), np.asarray(cells)
        self.d = self.cells.shape[1] - 1
        self.dt, self.params = dt, params
        self.T, self.S = T_space, S_space
        tabs = prony or {}
        self.vp = vo.ViscoParams(dim=self.d, dt=dt, H=params["H"], Rg=params["Rg"], Tb=params["Tb"],
                                 alpha_solid=params["alpha_solid"], alpha_liquid=params["alpha_liquid"], **tabs)
        self.thermal = to.ThermalOracle(self.x, self.cells, T_space["dofmap"], T_space["ref_nodes"], T_space["family"],
                                        T_space["degree"], params, dt)
        self.nT = int(np.max(T_space["dofmap"])) + 1
        self.nS = int(np.max(S_space["dofmap"])) + 1
        N, dd = self.vp.N, self.d * self.d
        z = np.zeros
        T0 = float(params["T_0"])
        self.f = {  # the Functions of TVP:106-173 with the ICs of TVP:187-233
            "T_cur": np.full(self.nT, T0), "T_prev": np.full(self.nT, T0), "T_next": z(self.nT),
            "Tf_cur": np.full(self.nT, T0), "Tf_prev": np.full(self.nT, T0),
            "Tf_partial_cur": np.full(self.nT * N, T0), "Tf_partial_prev": np.full(self.nT * N, T0),
            "phi": z(self.nT), "phi_next": z(self.nT), "xi": z(self.nT),
            "thermal_strain": z(self.nS * dd), "total_strain": z(self.nS * dd), "deviatoric_strain": z(self.nS * dd),
            "ds_partial": z(self.nS * N * dd), "dsigma_partial": z(self.nS * N * dd),
            "s_tilde_cur": z(self.nS * N * dd), "s_tilde_next": z(self.nS * N * dd),
            "sigma_tilde_cur": z(self.nS * N * dd), "sigma_tilde_next": z(self.nS * N * dd),
            "s_partial_cur": z(self.nS * N * dd), "s_partial_next": z(self.nS * N * dd),
            "sigma_partial_cur": z(self.nS * N * dd), "sigma_partial_next": z(self.nS * N * dd),
            "sigma_next": z(self.nS * dd),
        }
        # T basis tabulated at the sigma element's interpolation points (FFCx clamps table entries to 0/1)
        W = to.NodalBasis(self.d, T_space["degree"], T_space["ref_nodes"]).values(np.asarray(S_space["ref_nodes"]))
        W[np.abs(W) < 1e-12] = 0.0
        W[np.abs(W - 1.0) < 1e-12] = 1.0
        self.W = W
        self.newton_its = []

    # -- cell-wise evaluation helpers -------------------------------------------------------------------
    def _T_at_sigma_points(self, arr):
        """[n_cells * n_pts] values of a scalar T-space function at the sigma points of every cell."""
        loc = arr[self.T["dofmap"]]                                   # [nc, n_ld_T]
        out = np.zeros((loc.shape[0], self.W.shape[0]))
        for j in range(self.W.shape[1]):                              # accumulate in j order, zero entries dropped
            nz = self.W[:, j] != 0.0
            out[:, nz] = out[:, nz] + self.W[nz, j][None, :] * loc[:, j][:, None]
        return np.ascontiguousarray(out.ravel())

    def _S_cellwise(self, arr, bs):
        return np.ascontiguousarray(arr.reshape(self.nS, bs)[self.S["dofmap"].ravel()].ravel())

    def _scatter_S(self, name, vals, bs):
        tgt = self.f[name].reshape(self.nS, bs)
        tgt[self.S["dofmap"].ravel()] = vals.reshape(-1, bs)         # numpy: the last occurrence wins

    # -- one time step: TVP:367-381 ------------------------------------------------------------------------
    def step(self):
        f, p = self.f, self.vp
        N, dd = p.N, self.d * self.d
        # _solve_T (TVP:384-391)
        T, its, ok = self.thermal.newton(f["T_cur"], f["T_prev"])
        assert ok
        self.newton_its.append(its)
        f["T_cur"] = T
        # _solve_Tf (TVP:393-407) — T space
        f["phi"] = vo.phi(p, f["T_cur"])
        f["Tf_partial_cur"] = vo.Tf_partial(p, f["Tf_partial_prev"], f["T_cur"], f["phi"])
        f["Tf_partial_prev"] = f["Tf_partial_cur"].copy()
        f["Tf_cur"] = vo.Tf(p, f["Tf_partial_cur"])
        f["Tf_prev"] = f["Tf_cur"].copy()
        # _solve_strains (TVP:409-423) — sigma space, T-space coefficients evaluated cell by cell
        Tc, Tp = self._T_at_sigma_points(f["T_cur"]), self._T_at_sigma_points(f["T_prev"])
        Tfc, Tfp = self._T_at_sigma_points(f["Tf_cur"]), self._T_at_sigma_points(f["Tf_prev"])
        self._scatter_S("thermal_strain", vo.thermal_strain(p, Tc, Tp, Tfc, Tfp), dd)
        self._scatter_S("total_strain", vo.total_strain(p, self._S_cellwise(f["thermal_strain"], dd)), dd)
        self._scatter_S("deviatoric_strain", vo.deviatoric_strain(p, self._S_cellwise(f["total_strain"], dd)), dd)
        # _solve_shifted_time (TVP:426-435) — T space
        f["T_next"] = vo.T_next(p, f["T_cur"], f["T_prev"])
        f["phi"] = vo.phi(p, f["T_cur"])
        f["phi_next"] = vo.phi(p, f["T_next"])
        f["xi"] = vo.xi(p, f["phi_next"], f["phi"])
        # _solve_stress (TVP:438-452) — sigma space
        xi_c = self._T_at_sigma_points(f["xi"])
        self._scatter_S("ds_partial", vo.ds_partial(p, self._S_cellwise(f["deviatoric_strain"], dd), xi_c), N * dd)
        self._scatter_S("s_tilde_next", vo.tilde_next(p, "g", self._S_cellwise(f["s_tilde_cur"], N * dd), xi_c), N * dd)
        self._scatter_S("s_partial_next", vo.add(self._S_cellwise(f["ds_partial"], N * dd),
                                                 self._S_cellwise(f["s_tilde_next"], N * dd)), N * dd)
        f["s_tilde_cur"] = f["s_tilde_next"].copy()
        f["s_partial_cur"] = f["s_partial_next"].copy()
        self._scatter_S("dsigma_partial", vo.dsigma_partial(p, self._S_cellwise(f["total_strain"], dd), xi_c), N * dd)
        self._scatter_S("sigma_tilde_next", vo.tilde_next(p, "k", self._S_cellwise(f["sigma_tilde_cur"], N * dd), xi_c), N * dd)
        self._scatter_S("sigma_partial_next", vo.add(self._S_cellwise(f["dsigma_partial"], N * dd),
                                                     self._S_cellwise(f["sigma_tilde_next"], N * dd)), N * dd)
        f["sigma_tilde_cur"] = f["sigma_tilde_next"].copy()
        f["sigma_partial_cur"] = f["sigma_partial_next"].copy()
        self._scatter_S("sigma_next", vo.sigma_next(p, self._S_cellwise(f["s_partial_next"], N * dd),
                                                    self._S_cellwise(f["sigma_partial_next"], N * dd)), dd)

    def end_step(self):
        """TVP:378-379: T_prev <- T_cur (after the output is written)."""
        self.f["T_prev"] = self.f["T_cur"].copy()
