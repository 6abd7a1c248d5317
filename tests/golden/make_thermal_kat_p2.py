"""Generates tests/golden/thermal_kat_p2.json — known-answer vectors for hot path (B) with the CG2 (P2 Lagrange) element of
BASELINE configs 2 and 4, HAND-EVALUATED in plain Python on small perturbed triangle / tetrahedron meshes:

  basis        phi_i = lambda_i (2 lambda_i - 1) at the vertices, phi_ab = 4 lambda_a lambda_b on the edges
               (edge order of the reference element: 2-D (1,2),(0,2),(0,1); 3-D (2,3),(1,3),(1,2),(0,3),(0,2),(0,1))
  numbering    vertex dofs = vertex ids; edge dofs = n_vertices + rank of the edge among all edges sorted by (min, max)
  integration  Gauss-Legendre points computed here by Newton iteration on the Legendre polynomials, mapped to the simplex
               with the collapsed (Duffy) transformation — exact for the degrees that occur (mass 4, stiffness 2,
               radiation T^4 v on a facet: 10)
  weak form    TVP:293-306:  F = M (T - T_prev) + dt (alpha K T - f load) + dt*0.001*int_ds (sigma eps (T^4 - Ta^4) + htc (T - Ta)) v
               Jacobian-vector product: M x + dt alpha K x + dt*0.001*int_ds (4 sigma eps T^3 + htc) x v

No numpy, no FE library, nothing from the oracle or the product.  The meshes are stored in the file.  Floats as hex.
Run:  python tests/golden/make_thermal_kat_p2.py
"""
import itertools
import json
import math
import os
import random
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_thermal_kat_simplex import P, DT, simplex, facet_measure_normal, kuhn_mesh   # noqa: E402

EDGES = {2: [(1, 2), (0, 2), (0, 1)], 3: [(2, 3), (1, 3), (1, 2), (0, 3), (0, 2), (0, 1)]}


def gauss_legendre_01(n):
    """nodes/weights on [0, 1] (Newton iteration on P_n)."""
    xs, ws = [], []
    for i in range(n):
        x = math.cos(math.pi * (i + 0.75) / (n + 0.5))
        for _ in range(100):
            p0, p1 = 1.0, x
            for k in range(2, n + 1):
                p0, p1 = p1, ((2 * k - 1) * x * p1 - (k - 1) * p0) / k
            dp = n * (x * p1 - p0) / (x * x - 1.0)
            dx = p1 / dp
            x -= dx
            if abs(dx) < 1e-16:
                break
        p0, p1 = 1.0, x
        for k in range(2, n + 1):
            p0, p1 = p1, ((2 * k - 1) * x * p1 - (k - 1) * p0) / k
        dp = n * (x * p1 - p0) / (x * x - 1.0)
        xs.append(0.5 * (x + 1.0))
        ws.append(1.0 / ((1.0 - x * x) * dp * dp))
    return xs, ws


def simplex_rule(dim, n):
    """barycentric points + weights (sum = 1) on the dim-simplex from an n-point Gauss-Legendre rule per direction."""
    g, w = gauss_legendre_01(n)
    if dim == 1:
        return [((1.0 - a, a), wa) for a, wa in zip(g, w)]
    if dim == 2:      # x = a (1 - b), y = b, Jacobian (1 - b), area 1/2
        return [((1.0 - a * (1.0 - b) - b, a * (1.0 - b), b), 2.0 * wa * wb * (1.0 - b)) for a, wa in zip(g, w) for b, wb in zip(g, w)]
    out = []          # x = a (1-b)(1-c), y = b (1-c), z = c, Jacobian (1-b)(1-c)^2, volume 1/6
    for a, wa in zip(g, w):
        for b, wb in zip(g, w):
            for c, wc in zip(g, w):
                x, y, z = a * (1 - b) * (1 - c), b * (1 - c), c
                out.append(((1.0 - x - y - z, x, y, z), 6.0 * wa * wb * wc * (1 - b) * (1 - c) ** 2))
    return out


def p2_values(lam, d):
    v = [l * (2.0 * l - 1.0) for l in lam]
    return v + [4.0 * lam[a] * lam[b] for a, b in EDGES[d]]


def p2_grads(lam, glam, d):
    g = [[(4.0 * lam[i] - 1.0) * c for c in glam[i]] for i in range(d + 1)]
    return g + [[4.0 * (lam[a] * glam[b][c] + lam[b] * glam[a][c]) for c in range(d)] for a, b in EDGES[d]]


def evaluate(x, cells, T, Tp, xv):
    d = len(cells[0]) - 1
    nv = len(x)
    all_edges = sorted({tuple(sorted((c[a], c[b]))) for c in cells for a, b in EDGES[d]})
    eid = {e: nv + i for i, e in enumerate(all_edges)}
    dofmap = [[c[i] for i in range(d + 1)] + [eid[tuple(sorted((c[a], c[b])))] for a, b in EDGES[d]] for c in cells]
    n = nv + len(all_edges)
    assert len(T) == n
    F, Jx = [0.0] * n, [0.0] * n
    al, se, Ta, htc = P["alpha"], P["sigma"] * P["epsilon"], P["T_ambient"], P["htc"]
    cell_rule = simplex_rule(d, 4)
    for c, cell in enumerate(cells):
        vol, glam = simplex([x[v] for v in cell])
        dd = dofmap[c]
        nl = len(dd)
        for lam, w in cell_rule:
            ph = p2_values(lam, d)
            gr = p2_grads(lam, glam, d)
            dTq = sum(ph[j] * (T[dd[j]] - Tp[dd[j]]) for j in range(nl))
            xq = sum(ph[j] * xv[dd[j]] for j in range(nl))
            gT = [sum(gr[j][a] * T[dd[j]] for j in range(nl)) for a in range(d)]
            gx = [sum(gr[j][a] * xv[dd[j]] for j in range(nl)) for a in range(d)]
            for i in range(nl):
                gi = gr[i]
                F[dd[i]] += vol * w * (ph[i] * dTq + DT * (al * sum(gi[a] * gT[a] for a in range(d)) - P["f"] * ph[i]))
                Jx[dd[i]] += vol * w * (ph[i] * xq + DT * al * sum(gi[a] * gx[a] for a in range(d)))
    facets = {}
    for c, cell in enumerate(cells):
        for f in range(d + 1):
            facets.setdefault(tuple(sorted(cell[v] for v in range(d + 1) if v != f)), []).append((c, f))
    frule = simplex_rule(d - 1, 7)
    for key, owners in facets.items():
        if len(owners) != 1:
            continue
        c, f = owners[0]
        loc = [v for v in range(d + 1) if v != f]
        meas, _ = facet_measure_normal([x[cells[c][v]] for v in loc], x[cells[c][f]])
        dd = dofmap[c]
        for bary, w in frule:
            lam = [0.0] * (d + 1)
            for b, v in zip(bary, loc):
                lam[v] = b
            ph = p2_values(lam, d)
            Tq = sum(ph[j] * T[dd[j]] for j in range(len(dd)))
            xq = sum(ph[j] * xv[dd[j]] for j in range(len(dd)))
            flux = DT * 0.001 * (se * (Tq ** 4 - Ta ** 4) + htc * (Tq - Ta))
            dflux = DT * 0.001 * (4.0 * se * Tq ** 3 + htc)
            for i in range(len(dd)):
                if ph[i] != 0.0:
                    F[dd[i]] += meas * w * flux * ph[i]
                    Jx[dd[i]] += meas * w * dflux * xq * ph[i]
    return F, Jx, n, dofmap


def main():
    rng = random.Random(4242)
    out = {"about": "hand-evaluated CG2 heat-equation residual / Jacobian-vector product (see make_thermal_kat_p2.py)", "dt": DT,
           "params": P, "cases": []}
    hx = lambda v: [float(q).hex() for q in v]
    for dim, n, lengths in ((2, (3, 2), (3.0, 2.2)), (3, (2, 2, 1), (2.0, 2.4, 0.9))):
        x, cells = kuhn_mesh(dim, n, lengths, rng, 0.1)
        nv = len(x)
        ne = len({tuple(sorted((c[a], c[b]))) for c in cells for a, b in EDGES[dim]})
        nd = nv + ne
        T = [700.0 + 100.0 * rng.random() for _ in range(nd)]
        Tp = [t + rng.random() for t in T]
        xv = [rng.uniform(-1.0, 1.0) for _ in range(nd)]
        F, Jx, _, dofmap = evaluate(x, cells, T, Tp, xv)
        out["cases"].append({"dim": dim, "family": "CG", "degree": 2, "x": [hx(p) for p in x], "cells": cells, "dofmap": dofmap,
                             "T": hx(T), "T_prev": hx(Tp), "v": hx(xv), "residual": hx(F), "jac_x": hx(Jx)})
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "thermal_kat_p2.json"), "w") as fh:
        json.dump(out, fh)
    print("wrote thermal_kat_p2.json:", [(c["dim"], len(c["cells"]), len(c["T"])) for c in out["cases"]])


if __name__ == "__main__":
    main()
