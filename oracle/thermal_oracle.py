"""TEST INFRASTRUCTURE — CPU restatement of hot path (B): the ThermalModel heat-equation solve.

Restates, with scipy.sparse and an explicitly ASSEMBLED Jacobian, what the reference obtains from
dolfinx.fem.petsc.NonlinearProblem + dolfinx.nls.petsc.NewtonSolver + PETSc KSP
(/root/reference/ThermoViscoProblem.py, TVP):
  * weak form                 TVP:293-306  (mass, dt*(alpha*stiffness - f + 0.001*sigma*eps*(T^4-Ta^4) + 0.001*htc*(T-Ta)) on ds)
  * SIP-DG interior facets    TVP:308-325  (penalty 5.0, h = CellDiameter('+'), jump(v,n) = v+ n+ + v- n-, avg = (a+ + a-)/2)
  * Newton                    TVP:330-337  (incremental criterion, rtol 1e-12; dolfinx defaults atol 1e-10, max_it 50)
  * constants                 /root/reference/ThermalModel.py:18-27
The linear solves use a sparse direct factorisation, i.e. the exact discrete Newton step that the
reference's CG+GAMG (TVP:343-344) approximates.

PARITY UNPINNED against a reference RUN: the reference has no tests/golden vectors and dolfinx/PETSc are un-vendored
and not installable here.  What pins this module instead: tests/golden/thermal_kat.json (the reference's graded 1-D
line) and tests/golden/thermal_kat_2d3d.json (perturbed triangle / tetrahedron meshes), hand-evaluated from the weak
form TVP:293-325 in plain Python by the committed generators; tests/test_fe_tables.py holds this module to them at 1e-12.  Independence from the product: basis functions come from a Vandermonde inversion on
monomials (not the product's barycentric formulas), the '-' side of an interior facet is evaluated by
inverting the affine map at the physical quadrature point (no permutation tables), normals and measures
come from vertex coordinates.  Only the node numbering (dofmap + reference node order) is shared so that
vectors can be compared entry by entry.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this module.
"""
from __future__ import annotations

import itertools
import math

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla
from scipy.special import roots_jacobi


# ----------------------------------------------------------------------------- reference element
def _monomial_exponents(dim, degree):
    return [e for e in itertools.product(range(degree + 1), repeat=dim) if sum(e) <= degree]


class NodalBasis:
    """Lagrange basis dual to `ref_nodes`, built by inverting the monomial Vandermonde matrix."""

    def __init__(self, dim, degree, ref_nodes):
        self.dim, self.degree = dim, degree
        self.exps = _monomial_exponents(dim, degree)
        V = self._mono(np.asarray(ref_nodes, dtype=np.float64))
        assert V.shape[0] == V.shape[1], "node count must match the polynomial space"
        self.coef = np.linalg.inv(V)  # column j: monomial coefficients of phi_j
        self.n = V.shape[0]

    def _mono(self, pts):
        return np.stack([np.prod(pts ** np.array(e), axis=1) for e in self.exps], axis=1)

    def _dmono(self, pts, a):
        cols = []
        for e in self.exps:
            if e[a] == 0:
                cols.append(np.zeros(pts.shape[0]))
            else:
                e2 = list(e)
                e2[a] -= 1
                cols.append(e[a] * np.prod(pts ** np.array(e2), axis=1))
        return np.stack(cols, axis=1)

    def values(self, pts):
        return self._mono(np.atleast_2d(pts)) @ self.coef                      # [npts, n]

    def grads(self, pts):
        pts = np.atleast_2d(pts)
        return np.stack([self._dmono(pts, a) @ self.coef for a in range(self.dim)], axis=1)  # [npts, dim, n]


def _gauss01(n):
    x, w = np.polynomial.legendre.leggauss(n)
    return 0.5 * (x + 1), 0.5 * w


def simplex_rule(dim, degree):
    """Collapsed (Duffy) Gauss-Jacobi rule, exact to `degree`; weights sum to 1/dim!."""
    n = degree // 2 + 1
    if dim == 0:
        return np.zeros((1, 0)), np.ones(1)
    x0, w0 = _gauss01(n)
    if dim == 1:
        return x0[:, None], w0
    x1, w1 = roots_jacobi(n, 1, 0)
    x1, w1 = 0.5 * (x1 + 1), w1 / 4
    if dim == 2:
        P = np.array([(a * (1 - b), b) for b in x1 for a in x0])
        W = np.array([wa * wb for wb in w1 for wa in w0])
        return P, W
    x2, w2 = roots_jacobi(n, 2, 0)
    x2, w2 = 0.5 * (x2 + 1), w2 / 8
    P = np.array([(a * (1 - b) * (1 - c), b * (1 - c), c) for c in x2 for b in x1 for a in x0])
    W = np.array([wa * wb * wc for wc in w2 for wb in w1 for wa in w0])
    return P, W


# ----------------------------------------------------------------------------- the discrete problem
class ThermalOracle:
    """Assembled heat-equation operator on a simplicial mesh.

    x [nv, d], cells [nc, d+1]; dofmap [nc, n_ld] and ref_nodes [n_ld, d] fix the numbering;
    family 'CG' | 'DG'; params: the model_params dict of main.py:29-55; dt: time step."""

    PENALTY = 5.0  # TVP:313

    def __init__(self, x, cells, dofmap, ref_nodes, family, degree, params, dt, quad_degree=None):
        self.x = np.asarray(x, dtype=np.float64).reshape(len(x), -1)
        self.cells = np.asarray(cells, dtype=np.int64)
        self.dofmap = np.asarray(dofmap, dtype=np.int64)
        self.d = self.cells.shape[1] - 1
        self.family, self.degree, self.dt = family, degree, float(dt)
        self.n_dof = int(self.dofmap.max()) + 1
        self.basis = NodalBasis(self.d, degree, ref_nodes)
        self.p = params
        self.qdeg = quad_degree or max(2 * degree, 5 * degree) + 2
        self._geometry()
        self._facets()
        self.M = self._assemble_cells(mass=True)
        self.K = self._assemble_cells(mass=False)
        self.load = self._assemble_load()
        self.S = self._assemble_sip() if family == "DG" else sp.csr_matrix((self.n_dof, self.n_dof))
        a = float(params["alpha"])
        self.A_lin = (self.M + self.dt * (a * self.K) + self.S).tocsr()   # T-independent part of the Jacobian

    # -- geometry -------------------------------------------------------------------------------
    def _geometry(self):
        xv = self.x[self.cells]                                            # [nc, d+1, d]
        self.J = np.transpose(xv[:, 1:] - xv[:, :1], (0, 2, 1))            # dx/dxi
        self.detJ = np.abs(np.linalg.det(self.J)) if self.d > 1 else np.abs(self.J[:, 0, 0])
        self.Jinv = np.linalg.inv(self.J)
        h = np.zeros(len(self.cells))
        for a, b in itertools.combinations(range(self.d + 1), 2):
            h = np.maximum(h, np.linalg.norm(xv[:, a] - xv[:, b], axis=1))
        self.h = h                                                         # CellDiameter (TVP:314)

    def _facets(self):
        d = self.d
        loc = [tuple(v for v in range(d + 1) if v != f) for f in range(d + 1)]
        table = {}
        for c, cell in enumerate(self.cells):
            for f, lv in enumerate(loc):
                key = tuple(sorted(int(cell[v]) for v in lv))
                table.setdefault(key, []).append((c, f))
        self.loc_facets = loc
        self.ext = [v[0] for v in table.values() if len(v) == 1]
        self.int = [tuple(sorted(v)) for v in table.values() if len(v) == 2]   # '+' = lower cell index
        assert all(len(v) <= 2 for v in table.values())

    def _facet_frames(self, c, f):
        """Vectorised over facets: vertices [nf, d, d], measures [nf], unit normals pointing out of cell c [nf, d]."""
        d = self.d
        c, f = np.asarray(c, dtype=np.int64), np.asarray(f, dtype=np.int64)
        loc = np.array(self.loc_facets, dtype=np.int64).reshape(d + 1, d)
        fv = self.x[self.cells[c[:, None], loc[f]]]                        # [nf, d, d]
        opp = self.x[self.cells[c, f]]                                     # [nf, d]
        if d == 1:
            n = np.where(fv[:, 0, :] > opp, 1.0, -1.0)
            return fv, np.ones(len(c)), n
        if d == 2:
            t = fv[:, 1] - fv[:, 0]
            meas = np.linalg.norm(t, axis=1)
            n = np.stack([t[:, 1], -t[:, 0]], axis=1) / meas[:, None]
        else:
            cr = np.cross(fv[:, 1] - fv[:, 0], fv[:, 2] - fv[:, 0])
            nrm = np.linalg.norm(cr, axis=1)
            meas = 0.5 * nrm
            n = cr / nrm[:, None]
        flip = np.einsum("nd,nd->n", n, fv[:, 0] - opp) < 0
        n[flip] *= -1.0
        return fv, meas, n

    def _ref_coords(self, c, xp):
        """reference coordinates in cells c [nf] of physical points xp [nf, nq, d] (inverse affine map)."""
        return np.einsum("nqc,nac->nqa", xp - self.x[self.cells[c, 0]][:, None, :], self.Jinv[c])

    def _tab(self, c, xq, grads=False):
        """basis values [nf, nq, nl] (and physical gradients [nf, nq, d, nl]) at physical points xq of cells c."""
        nf, nq, d = xq.shape
        ref = self._ref_coords(c, xq).reshape(nf * nq, d)
        v = self.basis.values(ref).reshape(nf, nq, -1)
        if not grads:
            return v
        g = self.basis.grads(ref).reshape(nf, nq, d, -1)
        return v, np.einsum("nab,nqai->nqbi", self.Jinv[c], g)

    # -- cell integrals -------------------------------------------------------------------------
    def _assemble_cells(self, mass):
        P, W = simplex_rule(self.d, 2 * self.degree)
        nl = self.basis.n
        if mass:
            v = self.basis.values(P)
            ref = np.einsum("q,qi,qj->ij", W, v, v)
            Ke = self.detJ[:, None, None] * ref[None]
        else:
            g = self.basis.grads(P)                                        # [q, a, i]
            gp = np.einsum("cab,qai->cqbi", self.Jinv, g)                  # physical gradients
            Ke = np.einsum("q,c,cqbi,cqbj->cij", W, self.detJ, gp, gp)
        rows = np.repeat(self.dofmap[:, :, None], nl, axis=2)
        cols = np.repeat(self.dofmap[:, None, :], nl, axis=1)
        return sp.coo_matrix((Ke.ravel(), (rows.ravel(), cols.ravel())), shape=(self.n_dof, self.n_dof)).tocsr()

    def _assemble_load(self):
        P, W = simplex_rule(self.d, self.degree)
        ref = np.einsum("q,qi->i", W, self.basis.values(P))
        b = np.zeros(self.n_dof)
        np.add.at(b, self.dofmap.ravel(), (self.detJ[:, None] * ref[None]).ravel())
        return b

    # -- SIP-DG interior facets (TVP:318-325) -----------------------------------------------------
    def _assemble_sip(self):
        d, nl = self.d, self.basis.n
        a_plus = float(self.p["alpha"])
        if not self.int:
            return sp.csr_matrix((self.n_dof, self.n_dof))
        FP, FW = simplex_rule(d - 1, 2 * self.degree)
        FW = FW / FW.sum()
        bary = np.concatenate([1 - FP.sum(axis=1, keepdims=True), FP], axis=1) if d > 1 else np.ones((1, 1))
        pairs = np.array([[cp, fp, cm, fm] for (cp, fp), (cm, fm) in self.int], dtype=np.int64)
        rows, cols, vals = [], [], []
        for lo in range(0, len(pairs), 20000):                              # chunks bound the temporary memory
            cp, fp, cm = pairs[lo:lo + 20000, 0], pairs[lo:lo + 20000, 1], pairs[lo:lo + 20000, 2]
            fv, meas, n_p = self._facet_frames(cp, fp)
            n_m = -n_p
            xq = np.einsum("qk,nkd->nqd", bary, fv)
            vp, gp = self._tab(cp, xq, grads=True)
            vm, gm = self._tab(cm, xq, grads=True)
            # jump(w, n) = w+ n+ + w- n-  as [nf, q, d, 2nl] over the stacked (+,-) basis
            jump = np.concatenate([vp[:, :, None, :] * n_p[:, None, :, None], vm[:, :, None, :] * n_m[:, None, :, None]], axis=3)
            avg_g = 0.5 * np.concatenate([gp, gm], axis=3)
            pen = float(self.p.get("sip_penalty", self.PENALTY)) / self.h[cp]   # product knob for non-reference penalties
            wq = FW[None, :] * meas[:, None]
            out = pen[:, None, None] * np.einsum("nq,nqbi,nqbj->nij", wq, jump, jump)
            out -= np.einsum("nq,nqbi,nqbj->nij", wq, avg_g, jump)
            out -= np.einsum("nq,nqbi,nqbj->nij", wq, jump, avg_g)
            out *= self.dt * a_plus
            dofs = np.concatenate([self.dofmap[cp], self.dofmap[cm]], axis=1)      # [nf, 2nl]
            rows.append(np.repeat(dofs[:, :, None], 2 * nl, axis=2).ravel())
            cols.append(np.repeat(dofs[:, None, :], 2 * nl, axis=1).ravel())
            vals.append(out.ravel())
        return sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))),
                             shape=(self.n_dof, self.n_dof)).tocsr()

    # -- exterior facets: radiation + convection (TVP:302-304) -----------------------------------
    def _boundary(self, T, want_matrix):
        d, nl, p = self.d, self.basis.n, self.p
        se, htc, Ta = float(p["sigma"]) * float(p["epsilon"]), float(p["htc"]), float(p["T_ambient"])
        if not hasattr(self, "_bcache"):
            FP, FW = simplex_rule(d - 1, self.qdeg)
            FW = FW / FW.sum()
            bary = np.concatenate([1 - FP.sum(axis=1, keepdims=True), FP], axis=1) if d > 1 else np.ones((1, 1))
            c = np.array([e[0] for e in self.ext], dtype=np.int64)
            f = np.array([e[1] for e in self.ext], dtype=np.int64)
            fv, meas, _ = self._facet_frames(c, f)
            v = self._tab(c, np.einsum("qk,nkd->nqd", bary, fv))                   # [nf, q, nl]
            self._bcache = (c, v, FW[None, :] * meas[:, None], self.dofmap[c])
        c, v, wq, dofs = self._bcache
        Tq = np.einsum("nqj,nj->nq", v, T[dofs])
        flux = 0.001 * se * (Tq ** 4 - Ta ** 4) + 0.001 * htc * (Tq - Ta)
        vec = np.zeros(self.n_dof)
        np.add.at(vec, dofs.ravel(), (self.dt * np.einsum("nqi,nq->ni", v, wq * flux)).ravel())
        B = None
        if want_matrix:
            coef = 0.001 * (4.0 * se * Tq ** 3 + htc)
            out = self.dt * np.einsum("nq,nqi,nqj->nij", wq * coef, v, v)
            rows = np.repeat(dofs[:, :, None], nl, axis=2).ravel()
            cols = np.repeat(dofs[:, None, :], nl, axis=1).ravel()
            B = sp.coo_matrix((out.ravel(), (rows, cols)), shape=(self.n_dof, self.n_dof)).tocsr()
        return vec, B

    # -- residual / Jacobian / Newton --------------------------------------------------------------
    def residual(self, T, T_prev):
        """F(T; v) of TVP:293-325 as a vector."""
        a, f = float(self.p["alpha"]), float(self.p["f"])
        bvec, _ = self._boundary(T, False)
        return self.M @ (T - T_prev) + self.dt * (a * (self.K @ T) - f * self.load) + self.S @ T + bvec

    def jacobian(self, T):
        """dF/dT (what NonlinearProblem derives automatically, TVP:331)."""
        _, B = self._boundary(T, True)
        return (self.A_lin + B).tocsr()

    def newton(self, T0, T_prev, rtol=1e-12, atol=1e-10, max_it=50, report=False):
        """dolfinx NewtonSolver, convergence_criterion='incremental' (TVP:334-336)."""
        T = T0.copy()
        r0 = None
        for it in range(1, max_it + 1):
            b = self.residual(T, T_prev)
            dx = spla.spsolve(self.jacobian(T).tocsc(), b)
            T = T - dx
            r = np.linalg.norm(dx)
            if it == 1:
                r0 = r
                converged = False
            else:
                converged = (r / r0 < rtol) if r0 > 0 else True
                converged = converged or r < atol
            if report:
                print(f"  newton it {it}: |dx| = {r:.3e}")
            if r0 == 0.0 or converged:
                return T, it, True
        return T, max_it, False
