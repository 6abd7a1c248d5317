"""Generates tests/golden/thermal_kat.json — known-answer vectors for hot path (B) on the reference's OWN mesh.

The reference ships no tests and dolfinx cannot run here, so these vectors are HAND-EVALUATED from the weak form of
ThermoViscoProblem.py:293-325 with the constants of main.py:29-55 on the graded 1-D line of geometry.py:7-19, with
closed-form P1 element matrices and plain Python floats — independent of fem_glass_tempering_b200.fe, of
oracle/thermal_oracle.py and of the CUDA kernels:

  cell [a, b], h = b - a:   mass h/6 [[2,1],[1,2]],  stiffness 1/h [[1,-1],[-1,1]],  load h/2 [1,1]
  residual  F = M (T - T_prev) + dt (alpha K T - f load)  + dt*0.001*(sigma eps (T^4 - Ta^4) + htc (T - Ta)) at x = 0, L
  DG (TVP:308-325) at the interior vertex between cell i ('+', lower index) and i+1:
      jump(w) = w_L - w_R,  avg(w') = (w'_L + w'_R)/2,  h('+') = h_i,
      F_v += dt*alpha*( 5.0/h_i * jump(v) jump(T) - avg(v') jump(T) - jump(v) avg(T') )
  Jacobian-vector product: the same with T -> x in the linear terms and dt*0.001*(4 sigma eps T^3 + htc) x at x = 0, L.

Dof numbering: CG1 dof = vertex; DG1 dofs (2i, 2i+1) = (left, right) end of cell i.  Floats are stored as hex.
Run:  python tests/golden/make_thermal_kat.py
"""
import json
import math
import os
import random

P = {"f": 0.0, "epsilon": 0.93, "sigma": 5.670e-8, "T_ambient": 600.0, "alpha": 1.0, "htc": 280.1}   # main.py:29-48
DT = 0.1                                                                                             # main.py:15


def graded_points(points=(0.0, 5.0, 25.0, 45.0, 50.0), sizes=(0.1, 1.0, 3.0, 1.0, 0.1)):
    """gmsh's 1-D meshing of a line whose end points carry characteristic lengths (geometry.py:7-19): equidistribute the
    integral of 1/h(x), h linear along the line, element count = that integral rounded to the nearest integer."""
    xs = [points[0]]
    for a, b, ha, hb in zip(points[:-1], points[1:], sizes[:-1], sizes[1:]):
        r = (hb - ha) / (b - a)
        total = math.log(hb / ha) / r
        n = max(1, int(round(total)))
        for i in range(1, n):
            xs.append(a + ha * (math.exp(r * total * i / n) - 1.0) / r)
        xs.append(b)
    return xs


def robin(T):
    se, Ta = P["sigma"] * P["epsilon"], P["T_ambient"]
    return DT * 0.001 * (se * (T ** 4 - Ta ** 4) + P["htc"] * (T - Ta)), DT * 0.001 * (4.0 * se * T ** 3 + P["htc"])


def evaluate(xs, family, T, Tp, x):
    """Returns (residual F(T; T_prev), J(T) x) as lists."""
    nc = len(xs) - 1
    dof = (lambda i, k: i + k) if family == "CG" else (lambda i, k: 2 * i + k)
    n = len(xs) if family == "CG" else 2 * nc
    F, Jx = [0.0] * n, [0.0] * n
    a = P["alpha"]
    for i in range(nc):
        h = xs[i + 1] - xs[i]
        d = (dof(i, 0), dof(i, 1))
        dT = [T[d[0]] - Tp[d[0]], T[d[1]] - Tp[d[1]]]
        for r_ in range(2):
            m = [h / 6.0 * (2.0 if r_ == c else 1.0) for c in range(2)]
            k = [(1.0 if r_ == c else -1.0) / h for c in range(2)]
            F[d[r_]] += m[0] * dT[0] + m[1] * dT[1] + DT * (a * (k[0] * T[d[0]] + k[1] * T[d[1]]) - P["f"] * h / 2.0)
            Jx[d[r_]] += m[0] * x[d[0]] + m[1] * x[d[1]] + DT * a * (k[0] * x[d[0]] + k[1] * x[d[1]])
    for b in (dof(0, 0), dof(nc - 1, 1)):                  # the two exterior "facets" (points, weight 1)
        flux, dflux = robin(T[b])
        F[b] += flux
        Jx[b] += dflux * x[b]
    if family == "DG":
        for i in range(nc - 1):                            # interior vertex between cell i ('+') and i + 1
            hL, hR = xs[i + 1] - xs[i], xs[i + 2] - xs[i + 1]
            d = (2 * i, 2 * i + 1, 2 * i + 2, 2 * i + 3)
            jv = (0.0, 1.0, -1.0, 0.0)
            av = (-0.5 / hL, 0.5 / hL, -0.5 / hR, 0.5 / hR)
            for vec, out in ((T, F), (x, Jx)):
                jump = vec[d[1]] - vec[d[2]]
                avg = 0.5 * ((vec[d[1]] - vec[d[0]]) / hL + (vec[d[3]] - vec[d[2]]) / hR)
                for q in range(4):
                    out[d[q]] += DT * a * (5.0 / hL * jv[q] * jump - av[q] * jump - jv[q] * avg)
    return F, Jx


def main():
    rng = random.Random(20240517)
    xs = graded_points()
    cases = []
    for family in ("CG", "DG"):
        n = len(xs) if family == "CG" else 2 * (len(xs) - 1)
        T = [700.0 + 100.0 * rng.random() for _ in range(n)]
        Tp = [t + rng.random() for t in T]
        x = [rng.uniform(-1.0, 1.0) for _ in range(n)]
        F, Jx = evaluate(xs, family, T, Tp, x)
        hx = lambda v: [float(q).hex() for q in v]
        cases.append({"family": family, "degree": 1, "T": hx(T), "T_prev": hx(Tp), "x": hx(x), "residual": hx(F), "jac_x": hx(Jx)})
    out = {"about": "hand-evaluated heat-equation residual / Jacobian-vector product on the graded 1-D line (see make_thermal_kat.py)",
           "dt": DT, "params": P, "points": [float(q).hex() for q in xs], "cases": cases}
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "thermal_kat.json"), "w") as fh:
        json.dump(out, fh, indent=0)
    print(f"{len(xs) - 1} cells, wrote thermal_kat.json")


if __name__ == "__main__":
    main()
