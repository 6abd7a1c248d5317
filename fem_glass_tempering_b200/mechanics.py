"""Mechanical equilibrium of the plate — an EXTENSION behind model_parameters["mechanics"].

The reference never solves for a displacement: /root/reference/ViscoelasticModel.py:135-139 sets
total_strain = -thermal_strain, i.e. it treats every point as fully restrained.  With

    model_parameters["mechanics"] = True | {"fixed": "symmetry" | bool array [n_vertices, dim], "rtol": 1e-10,
                                            "atol": 0.0, "max_it": 50000, "warm_start": True}

ThermoViscoProblem.solve_timestep runs `_solve_mechanics` after the reference's phases: the displacement increment du
(vector P1 on the mesh vertices) for which the stresses of the reference's own Prony chain, evaluated with
total_strain = eps(du) - thermal_strain, are in weak equilibrium (csrc/mech.cu; CPU statement oracle/mechanics_oracle.py).
Default (key absent / False): the reference's behaviour, nothing here runs.

New Functions: functions["displacement"] (accumulated u), functions["displacement_increment"] (du of the last step),
functions["mechanical_strain"] (eps(du) at the sigma nodes); functions_next["sigma"] then holds the equilibrated stress
and, when materialised, total/deviatoric strain and the partial stresses include eps(du).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib, fe
from ._lib_mech import MechDescC, MechFieldsC


def symmetry_planes(x: np.ndarray, tol: float = 1e-12) -> np.ndarray:
    """Default constraint: component c is held on the plane x_c = min x_c (a quarter / eighth model of a symmetric
    plate): removes every rigid-body motion without restraining the thermal expansion.  [n_vertices, dim] bool."""
    x = np.asarray(x, dtype=np.float64)
    lo = x.min(axis=0)
    span = np.maximum(x.max(axis=0) - lo, 1e-300)
    return np.abs(x - lo) <= tol * span


def sigma_cell_weights(dim: int, degree: int) -> np.ndarray:
    """w_l = int phi_l / |K| of the sigma element: how a nodal field enters a cell integral against a constant."""
    load = fe.operator_tables(dim, degree).load
    return np.ascontiguousarray(load / load.sum())


def winner_cells(sigma: fe.ScalarSpace) -> np.ndarray:
    """The cell whose (cell-wise constant) strain a sigma node takes: the LAST cell touching it, as dolfinx's
    interpolate leaves it (for DG spaces simply the node's own cell)."""
    nc = sigma.mesh.n_cells
    w = np.zeros(sigma.n_nodes, dtype=np.int32)
    w[sigma.dofmap.ravel()] = np.repeat(np.arange(nc, dtype=np.int32), sigma.n_ld)   # last occurrence wins
    return w


class MechanicalEquilibrium:
    def __init__(self, ctx: _lib.Context, mesh, sigma_space: fe.ScalarSpace, visco_plan, device, options=None):
        import torch
        opts = dict(options) if isinstance(options, dict) else {}
        self.rtol = float(opts.get("rtol", 1e-10))
        self.atol = float(opts.get("atol", 0.0))
        self.max_it = int(opts.get("max_it", 50000))
        self.warm_start = bool(opts.get("warm_start", True))
        if mesh.cell_owned is not None or ctx.nranks != 1:
            raise NotImplementedError('model_parameters["mechanics"]: the equilibrium solve runs on one GPU')
        fixed = opts.get("fixed", "symmetry")
        if isinstance(fixed, str):
            if fixed != "symmetry":
                raise ValueError('mechanics["fixed"] must be "symmetry" or a bool array [n_vertices, dim]')
            fixed = symmetry_planes(mesh.x)
        fixed = np.ascontiguousarray(np.asarray(fixed, dtype=bool).reshape(mesh.n_vertices, mesh.dim))
        self.ctx, self.plan, self.device = ctx, visco_plan, device
        self.dim, self.n_vertices, self.n_sigma = mesh.dim, mesh.n_vertices, sigma_space.n_nodes
        up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(device)
        self._coords = up(mesh.x)
        self._cells = up(mesh.cells.astype(np.int32))
        self._fixed = up(fixed.astype(np.uint8).ravel())
        self._sdm = up(sigma_space.dofmap.astype(np.int32))
        self._winner = up(winner_cells(sigma_space))
        self._w = sigma_cell_weights(mesh.dim, sigma_space.degree)
        d = MechDescC()
        d.dim, d.n_vertices, d.n_cells = mesh.dim, mesh.n_vertices, mesh.n_cells
        d.coords, d.cells, d.fixed = _lib.ptr(self._coords), _lib.ptr(self._cells), _lib.ptr(self._fixed)
        d.n_ld_sigma, d.n_sigma_nodes = sigma_space.n_ld, sigma_space.n_nodes
        d.sigma_dofmap, d.winner_cell = _lib.ptr(self._sdm), _lib.ptr(self._winner)
        d.sigma_weights = self._w.ctypes.data
        self._h = C.c_void_p()
        with torch.cuda.device(device):
            _lib.check(_lib.lib().sg_mech_op_create(ctx.handle, C.byref(d), C.byref(self._h)))
        z = lambda n: torch.zeros(n, dtype=torch.float64, device=device)
        self.G_eff, self.K_eff = z(self.n_sigma), z(self.n_sigma)
        self.last_iters, self.last_rel_res = 0, 0.0
        self._du_prev = None

    # -- single operations (tests, benchmarks) --------------------------------------------------------------
    def coefficients(self, xi_sigma):
        _lib.check(_lib.lib().sg_mech_coefficients(self.plan.handle, self.n_sigma, _lib.ptr(xi_sigma), _lib.ptr(self.G_eff),
                                                   _lib.ptr(self.K_eff), _lib.current_stream_ptr()))
        return self.G_eff, self.K_eff

    def set_moduli(self, G_eff=None, K_eff=None):
        G = self.G_eff if G_eff is None else G_eff
        K = self.K_eff if K_eff is None else K_eff
        _lib.check(_lib.lib().sg_mech_set_moduli(self._h, _lib.ptr(G), _lib.ptr(K), _lib.current_stream_ptr()))

    def apply(self, x, y):
        _lib.check(_lib.lib().sg_mech_apply(self._h, _lib.ptr(x), _lib.ptr(y), _lib.current_stream_ptr()))
        return y

    def rhs(self, sigma0, b):
        _lib.check(_lib.lib().sg_mech_rhs(self._h, _lib.ptr(sigma0), _lib.ptr(b), _lib.current_stream_ptr()))
        return b

    def solve(self, sigma0, du):
        it, res = C.c_int32(0), C.c_double(0.0)
        rc = _lib.lib().sg_mech_solve(self._h, _lib.ptr(sigma0), _lib.ptr(du), self.rtol, self.atol, self.max_it,
                                      C.byref(it), C.byref(res), _lib.current_stream_ptr())
        self.last_iters, self.last_rel_res = it.value, res.value
        _lib.check(rc)
        return du

    def correct(self, du, xi_sigma, tensors: dict, G_eff=None, K_eff=None):
        f = MechFieldsC()
        for k, _ in MechFieldsC._fields_:
            setattr(f, k, _lib.ptr(tensors.get(k)))
        G = self.G_eff if G_eff is None else G_eff
        K = self.K_eff if K_eff is None else K_eff
        _lib.check(_lib.lib().sg_mech_correct(self._h, self.plan.handle, _lib.ptr(du), _lib.ptr(xi_sigma), _lib.ptr(G), _lib.ptr(K),
                                              C.byref(f), _lib.current_stream_ptr()))

    def apply_bytes(self) -> int:
        return int(_lib.lib().sg_mech_apply_bytes(self._h))

    # -- one time step ----------------------------------------------------------------------------------------
    def step(self, xi_sigma, tensors: dict, du, u):
        """sigma (= tensors["sigma"], the reference's stress on entry) -> equilibrated stress; du, u updated."""
        self.coefficients(xi_sigma)
        self.set_moduli()
        if not self.warm_start:
            du.zero_()
        elif self._du_prev is not None:
            # the increments change slowly from step to step: start from 2 du_n - du_{n-1}
            guess = 2.0 * du - self._du_prev
            self._du_prev.copy_(du)
            du.copy_(guess)
        else:
            self._du_prev = du.clone()
        self.solve(tensors["sigma"], du)
        self.correct(du, xi_sigma, tensors)
        u.add_(du)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            _lib.lib().sg_mech_op_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
