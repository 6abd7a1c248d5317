"""TEST INFRASTRUCTURE — CPU statement of the mechanical equilibrium step.

The reference has NO equilibrium solve: it sets total_strain = -thermal_strain (/root/reference/ViscoelasticModel.py:
135-139, "Eq. 28"), i.e. it assumes a fully restrained body.  This module states the extension the framework adds behind
model_parameters["mechanics"]: a displacement increment du in the vector-P1 space on the mesh vertices such that the
stresses of the reference's own Prony chain (VM:176-228), evaluated with

    total_strain = eps(du) - thermal_strain        instead of        - thermal_strain,

are in weak equilibrium, sum_K int_K sigma_h : eps(v) = 0 for all test displacements v.  The chain is linear in
total_strain, so with sigma0 = the reference's stress (what VM:224-228 yields with du = 0) and the nodal tangent moduli
    G_eff = sum_n g_n A(lambda_g_n),  K_eff = sum_n k_n A(lambda_k_n),   A(l) = l (1 - taylor(xi, l)) / xi = 1 - xi/(2 l)   (VM:176-191)
the stress is  sigma = sigma0 + 2 G_eff dev(eps(du)) + K_eff tr(eps(du)) I  at every sigma node.  eps(du) is constant per
cell; nodal coefficient fields enter the cell integrals through w_l = int phi_l / |K| of the sigma element; a sigma node
takes the strain of the LAST cell that touches it (the rule dolfinx's interpolate applies to cell-wise discontinuous
expressions).

Nothing here is checked against a reference run (there is nothing to run and nothing to compare with): PARITY UNPINNED by
construction; the pins are analytic (free uniform expansion is stress free, 1-D bar closed form; tests/test_mechanics_cpu.py).
Only tests/, smoke() and bench.py's CPU legs may import this module.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla


def taylor3(xi, lam):
    a = (-1.0 * xi) / lam                                            # VM:233-242
    return (1.0 + a) + 0.5 * (a * a)


def term_factors(xi, lams, mode: str = "reference"):
    """A(lambda_n) per node and term, [n_nodes, N]: what multiplies 2 g_n dev / k_n tr in VM:176-191 ("reference") or
    (1 - exp(-x))/x, x = xi/lambda_n, of the corrected scheme (csrc/visco.cu decay_fac)."""
    xi = np.asarray(xi, dtype=np.float64)[:, None]
    lam = np.asarray(lams, dtype=np.float64)[None, :]
    if mode == "reference":
        # lambda (1 - taylor)/xi == 1 + a/2, a = -xi/lambda, identically; the closed form has neither the cancellation noise
        # nor the 0/0 at xi = 0 of the reference's sequence - see csrc/mech.cu term_factor
        return 1.0 + 0.5 * ((-1.0 * xi) / lam)
    x = xi / lam
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.where(x != 0.0, -np.expm1(-x) / x, 1.0)


def tangent_moduli(xi, g, lam_g, k, lam_k, mode: str = "reference"):
    """Nodal G_eff, K_eff."""
    Ag, Ak = term_factors(xi, lam_g, mode), term_factors(xi, lam_k, mode)
    return Ag @ np.asarray(g, dtype=np.float64), Ak @ np.asarray(k, dtype=np.float64)


def p1_gradients(x, cells):
    """Physical gradients of the P1 hat functions, [nc, d+1, d], and cell volumes [nc]."""
    x, cells = np.asarray(x, dtype=np.float64), np.asarray(cells)
    d = cells.shape[1] - 1
    xv = x[cells]                                                    # [nc, d+1, d]
    J = np.transpose(xv[:, 1:] - xv[:, :1], (0, 2, 1))               # J[:, i, a] = x_a[i] - x_0[i]
    Jinv = np.linalg.inv(J)                                          # rows = gradients of lambda_1..lambda_d
    g = np.concatenate([-Jinv.sum(axis=1, keepdims=True), Jinv], axis=1)
    fact = (1, 1, 2, 6)[d]
    return g, np.abs(np.linalg.det(J)) / fact


def sigma_weights(ref_nodes, degree, dim):
    """w_l = int phi_l / |K| of the Lagrange element whose nodes are ref_nodes (P1: 1/(d+1) each; P2: vertices
    0 (d = 2) or -1/20 (d = 3), edge midpoints 1/3 or 1/5; d = 1 Simpson)."""
    from . import thermal_oracle as to
    basis = to.NodalBasis(dim, degree, np.asarray(ref_nodes, dtype=np.float64))
    pts, wts = to.simplex_rule(dim, max(2, degree))
    vals = basis.values(pts)                                          # [n_pts, n_ld]
    w = wts @ vals
    return w / w.sum()


class MechanicsOracle:
    def __init__(self, x, cells, sigma_dofmap, sigma_w, fixed):
        self.x, self.cells = np.asarray(x, dtype=np.float64), np.asarray(cells)
        self.d = self.cells.shape[1] - 1
        self.nv, self.nc = self.x.shape[0], self.cells.shape[0]
        self.sdm = np.asarray(sigma_dofmap)
        self.w = np.asarray(sigma_w, dtype=np.float64)
        self.fixed = np.asarray(fixed, dtype=bool).reshape(-1)        # [nv * d]
        self.g, self.vol = p1_gradients(self.x, self.cells)
        self.n_sigma = int(self.sdm.max()) + 1
        # last-cell-wins map of the sigma nodes
        self.winner = np.zeros(self.n_sigma, dtype=np.int64)
        self.winner[self.sdm.ravel()] = np.repeat(np.arange(self.nc), self.sdm.shape[1])
        d = self.d
        # E[c, a, i] = sym(e_i (x) g_a): strain of the unit displacement of vertex a in direction i
        E = np.zeros((self.nc, d + 1, d, d, d))
        for i in range(d):
            E[:, :, i, i, :] += 0.5 * self.g
            E[:, :, i, :, i] += 0.5 * self.g
        self.E = E
        self.dofs = (self.cells[:, :, None] * d + np.arange(d)[None, None, :]).reshape(self.nc, -1)

    def cell_mean(self, nodal, bs=1):
        v = np.asarray(nodal).reshape(self.n_sigma, bs)[self.sdm]     # [nc, n_ld, bs]
        return np.einsum("l,clb->cb", self.w, v)

    def _C(self, Gc, Kc, eps):
        """2 G dev(eps) + K tr(eps) I for [..., d, d] strains with per-cell moduli."""
        d = self.d
        tr = np.trace(eps, axis1=-2, axis2=-1)
        I = np.eye(d)
        shape = (-1,) + (1,) * (eps.ndim - 1)
        return 2.0 * Gc.reshape(shape) * (eps - (tr / d)[..., None, None] * I) + Kc.reshape(shape) * tr[..., None, None] * I

    def stiffness(self, G_nodal, K_nodal):
        d = self.d
        Gc, Kc = self.cell_mean(G_nodal)[:, 0], self.cell_mean(K_nodal)[:, 0]
        CE = self._C(Gc, Kc, self.E)                                                   # [nc, a, i, d, d]
        Kloc = np.einsum("c,caipq,cbjpq->caibj", self.vol, CE, self.E).reshape(self.nc, (d + 1) * d, (d + 1) * d)
        rows = np.repeat(self.dofs[:, :, None], (d + 1) * d, axis=2)
        cols = np.repeat(self.dofs[:, None, :], (d + 1) * d, axis=1)
        n = self.nv * d
        return sp.coo_matrix((Kloc.ravel(), (rows.ravel(), cols.ravel())), shape=(n, n)).tocsr()

    def rhs(self, sigma0_nodal):
        d = self.d
        s0 = self.cell_mean(np.nan_to_num(np.asarray(sigma0_nodal, dtype=np.float64), nan=0.0), d * d).reshape(self.nc, d, d)
        s0 = 0.5 * (s0 + np.transpose(s0, (0, 2, 1)))
        floc = -np.einsum("c,cpq,caipq->cai", self.vol, s0, self.E).reshape(self.nc, -1)
        b = np.zeros(self.nv * d)
        np.add.at(b, self.dofs.ravel(), floc.ravel())
        return b

    def constrained(self, A, b):
        """P A P + (I - P), P b: rows/columns of the held components replaced by the identity."""
        keep = (~self.fixed).astype(np.float64)
        P = sp.diags(keep)
        return (P @ A @ P + sp.diags(1.0 - keep)).tocsr(), keep * b

    def apply(self, G_nodal, K_nodal, x):
        A, _ = self.constrained(self.stiffness(G_nodal, K_nodal), np.zeros(self.nv * self.d))
        return A @ np.asarray(x, dtype=np.float64)

    def solve(self, G_nodal, K_nodal, sigma0_nodal):
        A, b = self.constrained(self.stiffness(G_nodal, K_nodal), self.rhs(sigma0_nodal))
        return spla.spsolve(A.tocsc(), b)

    def cell_strain(self, du):
        u = np.asarray(du).reshape(self.nv, self.d)[self.cells]       # [nc, d+1, d]
        grad = np.einsum("cai,caj->cij", u, self.g)
        return 0.5 * (grad + np.transpose(grad, (0, 2, 1)))

    def nodal_strain(self, du):
        return self.cell_strain(du)[self.winner]                       # [n_sigma, d, d]

    def correct(self, du, G_nodal, K_nodal, sigma0_nodal):
        """sigma = sigma0 + 2 G dev(eps) + K tr(eps) I at the sigma nodes; returns (sigma, eps) flattened."""
        d = self.d
        eps = self.nodal_strain(du)
        s0 = np.nan_to_num(np.asarray(sigma0_nodal, dtype=np.float64), nan=0.0)   # 0/0 entries of the reference read as 0
        sig = s0.reshape(self.n_sigma, d, d) + self._C(np.asarray(G_nodal), np.asarray(K_nodal), eps)
        return sig.ravel(), eps.ravel()

    def residual(self, sigma_nodal):
        """Discrete equilibrium residual B^T sigma_h on the free components (0 for DG sigma spaces after correct())."""
        return -self.rhs(sigma_nodal) * (~self.fixed)


def symmetry_planes(x, tol=1e-12):
    """Default constraint: component c held on the plane x_c = min x_c (a quarter/eighth model of a symmetric plate);
    removes every rigid-body motion without restraining the thermal expansion."""
    x = np.asarray(x, dtype=np.float64)
    lo, span = x.min(axis=0), np.maximum(x.max(axis=0) - x.min(axis=0), 1e-300)
    return (np.abs(x - lo) <= tol * span).reshape(-1)
