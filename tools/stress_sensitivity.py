"""CPU study (no GPU): how far does the stress history move when the heat solve stops at the product's tolerances?

Two oracle runs of the same problem: A with direct (sparse LU) Newton steps, B with the tolerance policy of
csrc/pcg.cu (inexact Newton, Eisenstat-Walker forcing, Jacobi-PCG on the assembled Jacobian) emulated in numpy.
Prints per step the relative differences of T, xi and sigma (max-norm over nodes whose temperature moved) so that
the default solver settings can be chosen where the stress meets north_star's 1e-10.

    python tools/stress_sensitivity.py [1d|3d|2d] [lin_rtol] [newton_atol] [eta] [steps]
"""
import os
import sys

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from fem_glass_tempering_b200 import fe                    # noqa: E402
from fem_glass_tempering_b200 import mesh as msh           # noqa: E402
from oracle.reference_problem import OracleProblem         # noqa: E402
from oracle.visco_oracle import MAIN_PARAMS                # noqa: E402
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from helpers import stress_rounding_floor                  # noqa: E402


def pcg(J, b, tol_abs, maxit=100000):
    dinv = 1.0 / J.diagonal()
    x = np.zeros_like(b)
    r = b.copy()
    z = dinv * r
    p = z.copy()
    rz = r @ z
    it = 0
    while np.sqrt(r @ r) > tol_abs and it < maxit:
        Ap = J @ p
        a = rz / (p @ Ap)
        x += a * p
        r -= a * Ap
        z = dinv * r
        rz2 = r @ z
        p = z + (rz2 / rz) * p
        rz = rz2
        it += 1
    return x, it


def gpu_like_newton(orc, T0, T_prev, lin_rtol, lin_atol, newton_rtol, newton_atol, eta1, max_it=50):
    T = T0.copy()
    r0 = None
    F_prev, target, its = 0.0, 0.0, 0
    for it in range(1, max_it + 1):
        b = orc.residual(T, T_prev)
        nb = np.linalg.norm(b)
        if it == 1:
            target = max(lin_atol, lin_rtol * nb)
        if eta1 > 0:
            if it > 1 and nb <= target:
                dx = np.zeros_like(b)
            else:
                eta = eta1 if F_prev == 0 else min(eta1, 0.9 * (nb / F_prev) ** 2)
                dx, k = pcg(orc.jacobian(T).tocsr(), b, max(eta * nb, 0.5 * target))
                its += k
        else:
            dx, k = pcg(orc.jacobian(T).tocsr(), b, target)
            its += k
        F_prev = nb
        T = T - dx
        r = np.linalg.norm(dx)
        if it == 1:
            r0 = r
            if r0 == 0:
                return T, it, its
        elif r / r0 < newton_rtol or r < newton_atol:
            return T, it, its
    raise RuntimeError("no convergence")


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "1d"
    lin_rtol = float(sys.argv[2]) if len(sys.argv) > 2 else 1e-12
    natol = float(sys.argv[3]) if len(sys.argv) > 3 else 1e-10
    eta = float(sys.argv[4]) if len(sys.argv) > 4 else 1e-3
    steps = int(sys.argv[5]) if len(sys.argv) > 5 else 20
    params = dict(MAIN_PARAMS)
    if which == "1d":
        m, cfg = msh.graded_line_mesh(), (("DG", 1), ("CG", 1))
    elif which == "3d":
        m, cfg = msh.box_mesh(6, 6, 3, 6.0, 6.0, 3.0), (("DG", 1), ("DG", 1))
        params["sip_penalty"] = 6.0
    elif which == "3dcg2":
        m, cfg = msh.box_mesh(4, 4, 2, 4.0, 4.0, 2.0), (("CG", 2), ("CG", 2))
    else:
        m, cfg = msh.rectangle_mesh(12, 6, 6.0, 3.0), (("CG", 2), ("CG", 2))
    sT = fe.ScalarSpace(m, *cfg[0])
    sS = sT if cfg[0] == cfg[1] else fe.ScalarSpace(m, *cfg[1])
    sp_ = lambda s: dict(dofmap=s.dofmap, ref_nodes=s.element.nodes, family=s.family, degree=s.degree)
    A = OracleProblem(m.x, m.cells, sp_(sT), sp_(sS), params, 0.1)
    B = OracleProblem(m.x, m.cells, sp_(sT), sp_(sS), params, 0.1)
    B.thermal.newton = lambda T0, Tp: (lambda r: (r[0], r[1], True))(
        gpu_like_newton(B.thermal, T0, Tp, lin_rtol, 0.0, 1e-12, natol, eta))
    d = A.d
    for s in range(steps):
        A.step()
        B.step()
        fa, fb = A.f, B.f
        eT = np.max(np.abs(fa["T_cur"] - fb["T_cur"])) / np.max(np.abs(fa["T_cur"]))
        dT = np.max(np.abs(fa["T_cur"] - fa["T_prev"]))
        exi = np.nanmax(np.abs(fa["xi"] - fb["xi"])) / np.nanmax(np.abs(fa["xi"]))
        dT_at_S = np.abs(A._T_at_sigma_points(fa["T_cur"]) - A._T_at_sigma_points(fa["T_prev"]))
        node_dT = np.zeros(A.nS)
        node_dT[sS.dofmap.ravel()] = dT_at_S
        good = node_dT > 1e-6
        sa, sb = fa["sigma_next"].reshape(-1, d * d)[good], fb["sigma_next"].reshape(-1, d * d)[good]
        es = np.max(np.abs(sa - sb)) / np.max(np.abs(sa)) if good.any() else 0.0
        # conditioning floor of the reference formula: eps / |xi/lambda| per Prony term
        xi_s = np.abs(A._T_at_sigma_points(fa["xi"]))
        node_xi = np.zeros(A.nS)
        node_xi[sS.dofmap.ravel()] = xi_s
        fl = stress_rounding_floor(A.vp, node_dT[good], node_xi[good])
        floor_rel = 2 * np.max(fl) / np.max(np.abs(sa)) if good.any() else 0.0
        worst_i = np.argmax(np.max(np.abs(sa - sb), axis=1))
        print(f"      floor(2x, normwise) {floor_rel:.2e}   at worst node: err {np.max(np.abs(sa - sb)[worst_i]):.2e} vs 2*floor {2 * fl[worst_i]:.2e}")
        print(f"step {s + 1:3d}: max dT {dT:.3e}  relerr T {eT:.2e}  xi {exi:.2e}  sigma {es:.2e}  "
              f"newton {B.newton_its[-1]}  min|xi| {np.min(xi_s[xi_s > 0]) if (xi_s > 0).any() else 0:.2e} good {good.sum()}/{good.size}")
        A.end_step()
        B.end_step()


if __name__ == "__main__":
    main()
