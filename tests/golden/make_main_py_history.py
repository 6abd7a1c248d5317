"""Generates tests/golden/main_py_history.json — the first time steps of the reference's DEFAULT run (main.py: graded 1-D
line, T = DG1, sigma = CG1, dt = 0.1, T_0 = 800 K), HAND-EVALUATED in plain Python:

  heat equation   make_thermal_kat.evaluate (closed-form P1 matrices, SIP with '+' = lower cell, TVP:293-325) inside
                  dolfinx's Newton iteration (incremental criterion rtol 1e-12 / atol 1e-10, TVP:334-337) with DENSE
                  Gaussian elimination for every Newton system (the Jacobian columns are evaluate()'s action on unit vectors);
  viscoelastic    make_visco_kat.step (the 16 expressions of VM:111-242 in the call order of TVP:367-381) at every DG1 node
                  for phi / Tf_partial / Tf / xi, and at every CG1 node of the sigma space for the stress, where the
                  T-space inputs are taken from the LAST cell that touches the node (dolfinx interpolates cell by cell, later
                  cells overwrite earlier ones: vertex v < n_cells gets the left end of cell v, the last vertex the right
                  end of the last cell);
  initial state   TVP:187-233: T = T_prev = T_0, Tf = T, every Tf_partial = T.x.array[0], zero stress histories.

Nothing here uses numpy, the oracle or the product.  Floats are stored as hex.  NaN (0/0 where a node's temperature did not
change bit-wise) is stored as "nan"; which nodes are stationary depends on the rounding of the linear solver,
so the tests only compare the stress where |dT| > 1e-6 K.
Run:  python tests/golden/make_main_py_history.py
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import make_thermal_kat as tk   # noqa: E402
import make_visco_kat as vk     # noqa: E402

STEPS, T0, DT = 5, 800.0, tk.DT


def solve_dense(A, b):
    n = len(b)
    M = [row[:] + [b[i]] for i, row in enumerate(A)]
    for c in range(n):
        piv = max(range(c, n), key=lambda r: abs(M[r][c]))
        M[c], M[piv] = M[piv], M[c]
        for r in range(c + 1, n):
            f = M[r][c] / M[c][c]
            if f != 0.0:
                for k in range(c, n + 1):
                    M[r][k] -= f * M[c][k]
    x = [0.0] * n
    for r in range(n - 1, -1, -1):
        x[r] = (M[r][n] - sum(M[r][k] * x[k] for k in range(r + 1, n))) / M[r][r]
    return x


def newton(xs, T_start, T_prev):
    n = len(T_start)
    T, r0 = T_start[:], None
    for it in range(1, 51):
        F, _ = tk.evaluate(xs, "DG", T, T_prev, [0.0] * n)
        cols = []
        for j in range(n):
            e = [0.0] * n
            e[j] = 1.0
            cols.append(tk.evaluate(xs, "DG", T, T_prev, e)[1])
        J = [[cols[j][i] for j in range(n)] for i in range(n)]
        dx = solve_dense(J, F)
        T = [t - d for t, d in zip(T, dx)]
        r = sum(d * d for d in dx) ** 0.5
        if it == 1:
            r0 = r
            if r0 == 0.0:
                return T, it
        elif r / r0 < 1e-12 or r < 1e-10:
            return T, it
    raise RuntimeError("Newton did not converge")


def main():
    xs = tk.graded_points()
    nc = len(xs) - 1
    nT, nS, N = 2 * nc, nc + 1, 6
    T_cur, T_prev = [T0] * nT, [T0] * nT
    Tfp = [[T0] * N for _ in range(nT)]
    zero = [[0.0] for _ in range(N)]
    winner = [2 * v for v in range(nc)] + [2 * (nc - 1) + 1]          # T-space dof that feeds sigma node v
    hist = []
    for step in range(STEPS):
        T_cur, its = newton(xs, T_cur, T_prev)
        node = [vk.step(1, DT, T_cur[q], T_prev[q], Tfp[q], zero, zero) for q in range(nT)]
        Tfp = [r["Tf_partial"] for r in node]
        sigma = [node[w]["sigma"][0] for w in winner]
        hist.append({"newton_its": its, "T": [float(v).hex() for v in T_cur], "T_prev": [float(v).hex() for v in T_prev],
                     "Tf": [float(r["Tf"]).hex() for r in node], "phi": [float(r["phi"]).hex() for r in node],
                     "xi": [float(r["xi"]).hex() for r in node],
                     "sigma": ["nan" if s != s else float(s).hex() for s in sigma]})
        T_prev = T_cur[:]
        print(f"step {step + 1}: newton {its}, T in [{min(T_cur):.6f}, {max(T_cur):.6f}], sigma(0) = {sigma[0]:.6e}")
    out = {"about": "hand-evaluated first steps of the reference's default run (see make_main_py_history.py)", "dt": DT, "T_0": T0,
           "points": [float(v).hex() for v in xs], "winner_dof": winner, "steps": hist}
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "main_py_history.json"), "w") as fh:
        json.dump(out, fh)
    print("wrote main_py_history.json")


if __name__ == "__main__":
    main()
