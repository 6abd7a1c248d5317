"""TEST INFRASTRUCTURE: numpy emulation of the table-driven algorithm csrc/thermal.cu implements
(cell-centric cells + DG interior facets via permutation tables + exterior facets).  It exists so that
the host-side tables of fem_glass_tempering_b200.fe (quadrature, facet permutations, neighbour maps) can
be validated against the assembled oracle on the CPU, where no GPU is available.  Never used by the product."""
import math

import numpy as np

from fem_glass_tempering_b200 import fe


def _dlam(d):
    g = np.zeros((d + 1, d))
    g[0, :] = -1.0
    for i in range(d):
        g[i + 1, i] = 1.0
    return g


def jac_apply(space, tabs, geo, topo, params, dt, x, T_lin=None, xm=None, residual=False, T_prev=None,
              cell_lo=0, cell_hi=None):
    """y = J(T_lin) x   (residual=False)   or   y = F(x; T_prev)   (residual=True).
    Only cells [cell_lo, cell_hi) are integrated (a rank's share of a partitioned mesh)."""
    mesh, d, nl = space.mesh, space.mesh.dim, space.n_ld
    a = float(params["alpha"])
    dm = space.dofmap
    y = np.zeros(space.n_nodes)
    xk = x[dm]                                                    # [nc, nl]
    xmass = xk - T_prev[dm] if residual else xk
    ym = geo.detJ[:, None] * (xmass @ tabs.mass.T)
    gref = np.einsum("qaj,cj->cqa", tabs.cq_grad, xk)
    gphys = np.einsum("cab,cqa->cqb", geo.Jinv, gref)
    flux = np.einsum("cab,cqb->cqa", geo.Jinv, gphys) * (tabs.cq_w[None, :, None] * geo.detJ[:, None, None])
    ys = np.einsum("qai,cqa->ci", tabs.cq_grad, flux)
    yk = ym + dt * a * ys
    if residual:
        yk -= dt * float(params["f"]) * geo.detJ[:, None] * tabs.load[None, :]
    if space.family == "DG":
        dl = _dlam(d)
        nc = mesh.n_cells
        for f in range(d + 1):
            nb = topo.neighbor[:, f]
            act = np.nonzero(nb >= 0)[0]
            if act.size == 0:
                continue
            nbc = nb[act]
            g = np.einsum("cab,a->cb", geo.Jinv[act], dl[f])                  # grad lambda_f
            gn = np.linalg.norm(g, axis=1)
            n = -g / gn[:, None]
            area = geo.detJ[act] * gn / math.factorial(d - 1)
            hplus = np.where(act < nbc, geo.h[act], geo.h[nbc])
            jn = np.einsum("cab,cb->ca", geo.Jinv[act], n)                    # (Jinv n)_a, own cell
            jnN = np.einsum("cab,cb->ca", geo.Jinv[nbc], n)
            nbf = topo.nb_facet[act, f].astype(int)
            pid = topo.nb_perm[act, f].astype(int)
            xK, xN = x[dm[act]], x[dm[nbc]]
            for q in range(tabs.fq_w.size):
                qn = tabs.fq_perm[pid, q]
                vK = tabs.fq_val[f, q] @ xK.T                                  # [na]
                phiN = tabs.fq_val[nbf, qn]                                    # [na, nl]
                vN = np.einsum("cj,cj->c", phiN, xN)
                dphiK = np.einsum("ca,ai->ci", jn, tabs.fq_grad[f, q])         # dn of own basis
                dphiN = np.einsum("ca,cai->ci", jnN, tabs.fq_grad[nbf, qn])
                dnK = np.einsum("ci,ci->c", dphiK, xK)
                dnN = np.einsum("ci,ci->c", dphiN, xN)
                jump = vK - vN
                w = dt * a * tabs.fq_w[q] * area
                contrib = w[:, None] * ((float(params.get("sip_penalty", fe_penalty())) / hplus * jump)[:, None] * tabs.fq_val[f, q][None, :]
                                        - 0.5 * dphiK * jump[:, None]
                                        - tabs.fq_val[f, q][None, :] * (0.5 * (dnK + dnN))[:, None])
                np.add.at(yk, act, contrib)
    cell_hi = mesh.n_cells if cell_hi is None else cell_hi
    yk[:cell_lo] = 0.0
    yk[cell_hi:] = 0.0
    np.add.at(y, dm.ravel(), yk.ravel())
    # exterior facets
    se, htc, Ta = float(params["sigma"]) * float(params["epsilon"]), float(params["htc"]), float(params["T_ambient"])
    area = fe.facet_measures(mesh, geo, topo.bnd_cell, topo.bnd_facet)
    for c, f, ar in zip(topo.bnd_cell, topo.bnd_facet, area):
        if c < cell_lo or c >= cell_hi:
            continue
        v = tabs.bq_val[f]                                                     # [nq, nl]
        dofs = dm[c]
        if residual:
            Tq = v @ x[dofs]
            flux = 0.001 * se * (Tq ** 4 - Ta ** 4) + 0.001 * htc * (Tq - Ta)
            y[dofs] += dt * ar * (v.T @ (tabs.bq_w * flux))
        else:
            Tq = v @ T_lin[dofs]
            coef = 0.001 * (4 * se * Tq ** 3 + htc)
            y[dofs] += dt * ar * (v.T @ (tabs.bq_w * coef * (v @ x[dofs])))
    return y


def fe_penalty():
    return 5.0


def cell_matrices(space, tabs, geo, params, dt):
    """The cell part of the Jacobian per cell, A_K = |detJ| M_ref + dt alpha K_K  (what the class tables of thermal.cu hold)."""
    g = np.einsum("cab,qai->cqbi", geo.Jinv, tabs.cq_grad)
    K = np.einsum("cqbi,cqbj,q,c->cij", g, g, tabs.cq_w, geo.detJ)
    return geo.detJ[:, None, None] * tabs.mass[None] + dt * float(params["alpha"]) * K


def row_stencil_classes(dofmap, cell_class, n_rows):
    """numpy emulation of csrc/stencil.cu: rows are equivalent when the multisets {(class of K, local row a,
    dofmap[K][.] - row)} over their (cell, local row) pairs agree.  Returns (class id per row, number of classes)."""
    nc, nl = dofmap.shape
    dm = dofmap.astype(np.int64)
    rel = (dm[:, None, :] - dm[:, :, None]).reshape(nc * nl, nl)                 # [K, a, b] -> dm[K, b] - dm[K, a]
    key = np.concatenate([np.repeat(cell_class, nl)[:, None], np.tile(np.arange(nl), nc)[:, None], rel], axis=1)
    uniq, contrib = np.unique(key, axis=0, return_inverse=True)                  # exact (no hashing) contribution ids
    rows = dm.ravel()
    order = np.lexsort((contrib.ravel(), rows))
    sig = {}
    cls = np.zeros(n_rows, dtype=np.int64)
    start = np.searchsorted(rows[order], np.arange(n_rows + 1))
    cs = contrib.ravel()[order]
    for r in range(n_rows):
        cls[r] = sig.setdefault(tuple(cs[start[r]:start[r + 1]]), len(sig))
    return cls, len(sig)


def row_stencil_apply(dofmap, cell_class, mats_by_class, row_class, n_classes, x):
    """y = A x in gather form: per class ONE (offset, coefficient) list taken from a representative row."""
    nc, nl = dofmap.shape
    dm = dofmap.astype(np.int64)
    rep = np.full(n_classes, -1, dtype=np.int64)
    rep[row_class[::-1]] = np.arange(row_class.size)[::-1]                       # first row of every class
    table = [dict() for _ in range(n_classes)]
    is_rep = np.zeros(row_class.size, dtype=bool)
    is_rep[rep] = True
    for K in np.nonzero(is_rep[dm].any(axis=1))[0]:                              # ascending cell order, like the kernel
        for a in range(nl):
            r = dm[K, a]
            if is_rep[r]:
                t = table[row_class[r]]
                for b in range(nl):
                    t[dm[K, b] - r] = t.get(dm[K, b] - r, 0.0) + mats_by_class[cell_class[K]][a, b]
    y = np.zeros(row_class.size)
    for c in range(n_classes):
        rows = np.nonzero(row_class == c)[0]
        for off, coef in sorted(table[c].items()):
            y[rows] += coef * x[rows + off]
    return y, table
