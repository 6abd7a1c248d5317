"""Generates tests/golden/thermal_kat_2d3d.json — known-answer vectors for hot path (B) on small 2-D / 3-D DG1 and CG1
meshes, HAND-EVALUATED from the weak form of ThermoViscoProblem.py:293-325 in plain Python (no numpy, no FE library):

  P1 on a simplex K with vertices p_0..p_d:  grad(lambda_i) from the inverse of [p_1-p_0 ... p_d-p_0],
     mass |K| (1+delta_ij) / ((d+1)(d+2)),  stiffness |K| grad(lambda_i).grad(lambda_j),  load |K|/(d+1)
  exterior facet F (radiation + convection, TVP:302-304):  int_F g(T) v ds  with a Gauss rule exact to degree 5
  interior facet F = K+ | K-  ('+' = lower cell index), n = unit normal out of K+, h = CellDiameter(K+) = longest edge:
     dt*alpha*[ 5.0/h int jump(v).jump(T) - int avg(grad v).jump(T) - int jump(v).avg(grad T) ],
     jump(w) = (w+ - w-) n,  avg(g) = (g+ + g-)/2;  traces are linear on F, so
     int_F a b = |F| sum_kl a_k b_l (1+delta_kl) / (d (d+1)),  int_F a = |F| sum_k a_k / d   (k, l over the facet's vertices)

The meshes (coordinates + cells) are stored in the file, so the tests do not depend on the product's mesher either.
Dof numbering: CG1 dof = vertex id; DG1 dof = (d+1)*cell + local vertex.  Floats are stored as hex.
Run:  python tests/golden/make_thermal_kat_simplex.py
"""
import itertools
import json
import math
import os
import random

P = {"f": 0.0, "epsilon": 0.93, "sigma": 5.670e-8, "T_ambient": 600.0, "alpha": 1.0, "htc": 280.1}   # main.py:29-48
DT, PENALTY = 0.1, 5.0                                                                               # main.py:15, TVP:313


# ------------------------------------------------------------------------------------------------ tiny linear algebra
def sub(a, b):
    return [x - y for x, y in zip(a, b)]


def dot(a, b):
    return sum(x * y for x, y in zip(a, b))


def norm(a):
    return math.sqrt(dot(a, a))


def det(m):
    n = len(m)
    if n == 1:
        return m[0][0]
    if n == 2:
        return m[0][0] * m[1][1] - m[0][1] * m[1][0]
    return (m[0][0] * (m[1][1] * m[2][2] - m[1][2] * m[2][1]) - m[0][1] * (m[1][0] * m[2][2] - m[1][2] * m[2][0])
            + m[0][2] * (m[1][0] * m[2][1] - m[1][1] * m[2][0]))


def inverse(m):
    n = len(m)
    a = [list(map(float, row)) + [1.0 if i == j else 0.0 for j in range(n)] for i, row in enumerate(m)]
    for c in range(n):
        piv = max(range(c, n), key=lambda r: abs(a[r][c]))
        a[c], a[piv] = a[piv], a[c]
        d = a[c][c]
        a[c] = [v / d for v in a[c]]
        for r in range(n):
            if r != c:
                f = a[r][c]
                a[r] = [v - f * w for v, w in zip(a[r], a[c])]
    return [row[n:] for row in a]


# ------------------------------------------------------------------------------------------------ geometry of one simplex
def simplex(pts):
    """volume and grad(lambda_i), i = 0..d."""
    d = len(pts) - 1
    J = [[pts[a + 1][c] - pts[0][c] for a in range(d)] for c in range(d)]      # J[c][a] = d x_c / d xi_a
    vol = abs(det(J)) / math.factorial(d)
    Ji = inverse(J)                                                            # Ji[a][c] = d xi_a / d x_c
    grads = [[-sum(Ji[a][c] for a in range(d)) for c in range(d)]] + [[Ji[a][c] for c in range(d)] for a in range(d)]
    return vol, grads


def facet_measure_normal(fp, opposite):
    """measure of the facet with vertices fp and its unit normal pointing away from `opposite`."""
    d = len(fp)
    if d == 2:
        t = sub(fp[1], fp[0])
        meas = norm(t)
        n = [t[1] / meas, -t[0] / meas]
    else:
        u, v = sub(fp[1], fp[0]), sub(fp[2], fp[0])
        cr = [u[1] * v[2] - u[2] * v[1], u[2] * v[0] - u[0] * v[2], u[0] * v[1] - u[1] * v[0]]
        meas = 0.5 * norm(cr)
        n = [c / (2.0 * meas) for c in cr]
    if dot(n, sub(fp[0], opposite)) < 0.0:
        n = [-c for c in n]
    return meas, n


# Gauss rules exact to degree 5 on the reference facet, barycentric points + weights (sum 1)
_G3 = [(0.5 - math.sqrt(0.6) / 2.0, 5.0 / 18.0), (0.5, 8.0 / 18.0), (0.5 + math.sqrt(0.6) / 2.0, 5.0 / 18.0)]
SEG_RULE = [((1.0 - x, x), w) for x, w in _G3]
_a1, _a2 = (6.0 - math.sqrt(15.0)) / 21.0, (6.0 + math.sqrt(15.0)) / 21.0
_w1, _w2 = (155.0 - math.sqrt(15.0)) / 1200.0, (155.0 + math.sqrt(15.0)) / 1200.0
TRI_RULE = [((1.0 / 3.0, 1.0 / 3.0, 1.0 / 3.0), 9.0 / 40.0)]
for a_, w_ in ((_a1, _w1), (_a2, _w2)):
    b_ = 1.0 - 2.0 * a_
    TRI_RULE += [((a_, a_, b_), w_), ((a_, b_, a_), w_), ((b_, a_, a_), w_)]


def evaluate(x, cells, family, T, Tp, xv):
    d = len(cells[0]) - 1
    nl = d + 1
    dof = (lambda c, i: cells[c][i]) if family == "CG" else (lambda c, i: nl * c + i)
    n = len(x) if family == "CG" else nl * len(cells)
    F, Jx = [0.0] * n, [0.0] * n
    al, se, Ta, htc = P["alpha"], P["sigma"] * P["epsilon"], P["T_ambient"], P["htc"]
    geo = []
    for c, cell in enumerate(cells):
        pts = [x[v] for v in cell]
        vol, g = simplex(pts)
        geo.append((vol, g))
        dd = [dof(c, i) for i in range(nl)]
        for i in range(nl):
            fi = ji = 0.0
            for j in range(nl):
                m = vol * (2.0 if i == j else 1.0) / ((d + 1) * (d + 2))
                k = vol * dot(g[i], g[j])
                fi += m * (T[dd[j]] - Tp[dd[j]]) + DT * al * k * T[dd[j]]
                ji += (m + DT * al * k) * xv[dd[j]]
            F[dd[i]] += fi - DT * P["f"] * vol / (d + 1)
            Jx[dd[i]] += ji
    # facets
    facets = {}
    for c, cell in enumerate(cells):
        for f in range(nl):
            key = tuple(sorted(cell[v] for v in range(nl) if v != f))
            facets.setdefault(key, []).append((c, f))
    rule = SEG_RULE if d == 2 else TRI_RULE
    for key, owners in facets.items():
        if len(owners) == 1:                                   # exterior facet: Robin + radiation
            c, f = owners[0]
            loc = [v for v in range(nl) if v != f]
            fp = [x[cells[c][v]] for v in loc]
            meas, _ = facet_measure_normal(fp, x[cells[c][f]])
            dd = [dof(c, v) for v in loc]
            for bary, w in rule:
                Tq = sum(b * T[q] for b, q in zip(bary, dd))
                xq = sum(b * xv[q] for b, q in zip(bary, dd))
                flux = DT * 0.001 * (se * (Tq ** 4 - Ta ** 4) + htc * (Tq - Ta))
                dflux = DT * 0.001 * (4.0 * se * Tq ** 3 + htc)
                for b, q in zip(bary, dd):
                    F[q] += meas * w * flux * b
                    Jx[q] += meas * w * dflux * xq * b
        elif family == "DG":                                   # interior facet, '+' = lower cell index
            (cp, fp_), (cm, fm) = sorted(owners)
            gv = list(key)                                      # global vertices of the facet
            pts_f = [x[v] for v in gv]
            meas, nrm = facet_measure_normal(pts_f, x[cells[cp][fp_]])
            hp = max(norm(sub(x[a], x[b])) for a, b in itertools.combinations(cells[cp], 2))
            pen = PENALTY / hp
            gp, gm = geo[cp][1], geo[cm][1]
            dofs = [dof(cp, i) for i in range(nl)] + [dof(cm, i) for i in range(nl)]
            # trace of basis function q at facet vertex k (+ side positive, - side enters the jump with a minus sign)
            tr = [[(1.0 if cells[cp][i] == v else 0.0) for v in gv] for i in range(nl)] + \
                 [[(-1.0 if cells[cm][i] == v else 0.0) for v in gv] for i in range(nl)]
            dn = [0.5 * dot(gp[i], nrm) for i in range(nl)] + [0.5 * dot(gm[i], nrm) for i in range(nl)]   # avg(grad v).n
            for vec, out in ((T, F), (xv, Jx)):
                jump = [sum(tr[q][k] * vec[dofs[q]] for q in range(2 * nl)) for k in range(d)]              # at facet vertices
                avg_n = sum(dn[q] * vec[dofs[q]] for q in range(2 * nl))
                int_jump = meas * sum(jump) / d
                for q in range(2 * nl):
                    jj = meas * sum(tr[q][k] * jump[l] * (2.0 if k == l else 1.0) for k in range(d) for l in range(d)) / (d * (d + 1))
                    int_jv = meas * sum(tr[q]) / d
                    out[dofs[q]] += DT * al * (pen * jj - dn[q] * int_jump - int_jv * avg_n)
    return F, Jx


def kuhn_mesh(dim, n, lengths, rng, jitter):
    """Kuhn triangulation of a box with randomly perturbed interior vertices (so that no two cells share a shape)."""
    dims = [k + 1 for k in n]
    strides = [1] * dim
    for c in range(dim - 2, -1, -1):
        strides[c] = strides[c + 1] * dims[c + 1]
    x = []
    for idx in itertools.product(*[range(k) for k in dims]):
        p = [idx[c] * lengths[c] / n[c] for c in range(dim)]
        if all(0 < idx[c] < n[c] for c in range(dim)):
            p = [v + rng.uniform(-jitter, jitter) for v in p]
        x.append(p)
    cells = []
    for base in itertools.product(*[range(k) for k in n]):
        for perm in itertools.permutations(range(dim)):
            p = list(base)
            cell = [sum(a * b for a, b in zip(p, strides))]
            for axis in perm:
                p[axis] += 1
                cell.append(sum(a * b for a, b in zip(p, strides)))
            cells.append(cell)
    return x, cells


def main():
    rng = random.Random(777)
    out = {"about": "hand-evaluated heat-equation residual / Jacobian-vector product on small simplicial meshes (see "
                    "make_thermal_kat_simplex.py)", "dt": DT, "params": P, "cases": []}
    hx = lambda v: [float(q).hex() for q in v]
    for dim, n, lengths in ((2, (3, 2), (3.0, 2.2)), (3, (2, 2, 2), (2.0, 2.4, 1.6))):
        x, cells = kuhn_mesh(dim, n, lengths, rng, 0.12)
        for family in ("DG", "CG"):
            nd = len(x) if family == "CG" else (dim + 1) * len(cells)
            T = [700.0 + 100.0 * rng.random() for _ in range(nd)]
            Tp = [t + rng.random() for t in T]
            xv = [rng.uniform(-1.0, 1.0) for _ in range(nd)]
            F, Jx = evaluate(x, cells, family, T, Tp, xv)
            out["cases"].append({"dim": dim, "family": family, "degree": 1, "x": [hx(p) for p in x], "cells": cells,
                                 "T": hx(T), "T_prev": hx(Tp), "v": hx(xv), "residual": hx(F), "jac_x": hx(Jx)})
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "thermal_kat_2d3d.json"), "w") as fh:
        json.dump(out, fh)
    print("wrote thermal_kat_2d3d.json:", [(c["dim"], c["family"], len(c["cells"])) for c in out["cases"]])


if __name__ == "__main__":
    main()
