"""Generator of tests/golden/mech_kat.json — hand-evaluated known answers of the mechanical equilibrium step
(an extension, the reference has no counterpart: VM:135-139).  Plain Python floats, no numpy, no
import from oracle/ or the package, so that the oracle AND the CUDA path are checked against something neither produced.

Case "bar": a 1-D bar 0 = x_0 < ... < x_n held at x_0, free at x_n, sigma space DG1 (two nodes per cell).  d = 1, so
dev(eps) = 0 and sigma = K_eff eps.  Weak equilibrium with a free end means the MEAN stress of every cell vanishes:
    mean(K)_c eps_c + mean(sigma0)_c = 0,      eps_c = (u_{c+1} - u_c)/h_c,
with mean = the average of the two nodal values (w = 1/2, 1/2).  Hence eps_c, u by summation, and the corrected nodal
stress sigma_i = sigma0_i + K_i eps_c(i).

Case "patch": any mesh, constant moduli, sigma0 = -(2 G dev(E) + K tr(E) I) for a constant diagonal strain E: the exact
solution with the symmetry-plane constraint is u_i = E_ii (x_i - min x_i), the corrected stress is zero.
"""
import json
import os


def bar_case():
    xs = [0.0, 0.1, 0.35, 0.9, 1.0, 1.6, 2.5]
    n = len(xs) - 1
    # nodal fields of the DG1 sigma space, node 2c and 2c+1 belong to cell c
    K = [30.0 + 3.0 * ((7 * i) % 5) + 0.25 * i for i in range(2 * n)]
    G = [20.0 + 2.0 * ((3 * i) % 7) for i in range(2 * n)]          # irrelevant in 1-D (dev = 0), must not matter
    s0 = [-(0.002 + 0.0003 * ((5 * i) % 11)) * K[i] for i in range(2 * n)]
    u = [0.0]
    eps = []
    for c in range(n):
        Km = 0.5 * K[2 * c] + 0.5 * K[2 * c + 1]
        sm = 0.5 * s0[2 * c] + 0.5 * s0[2 * c + 1]
        e = -sm / Km
        eps.append(e)
        u.append(u[-1] + e * (xs[c + 1] - xs[c]))
    sig = [s0[i] + K[i] * eps[i // 2] for i in range(2 * n)]
    return dict(name="bar", x=[v.hex() for v in xs], K=[v.hex() for v in K], G=[v.hex() for v in G],
                sigma0=[v.hex() for v in s0], u=[v.hex() for v in u], eps=[v.hex() for v in eps],
                sigma=[v.hex() for v in sig])


def patch_case(dim):
    E = [0.0011, -0.0004, 0.0007][:dim]
    G, K = 27.5, 41.0
    tr = sum(E)
    s0 = [[0.0] * dim for _ in range(dim)]
    for i in range(dim):
        s0[i][i] = -(2.0 * G * (E[i] - tr / dim) + K * tr)
    return dict(name=f"patch{dim}", dim=dim, E=[v.hex() for v in E], G=G.hex(), K=K.hex(),
                sigma0=[[v.hex() for v in row] for row in s0])


if __name__ == "__main__":
    out = dict(cases=[bar_case(), patch_case(2), patch_case(3)])
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "mech_kat.json"), "w") as fh:
        json.dump(out, fh, indent=1)
