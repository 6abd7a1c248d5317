"""Generates tests/golden/visco_kat.json — known-answer vectors for hot path (A).

The reference (pzimbrod/fem-glass-tempering) ships no tests or golden files and
dolfinx cannot be installed here, so these vectors are HAND-EVALUATED from the
reference formulas (ViscoelasticModel.py:111-242 in the call order of
ThermoViscoProblem.py:367-381) with plain Python floats — an evaluation that is
independent of both oracle/visco_oracle.c and the CUDA kernel.  Floats are stored
as hex so the comparison is bit-exact.

Run:  python tests/golden/make_visco_kat.py
"""
import json
import math
import os
import random

M = (5.523e-2, 8.205e-2, 1.215e-1, 2.286e-1, 2.860e-1, 2.265e-1)          # VM:19-26
LM = (5.965e-4, 1.077e-2, 1.362e-1, 1.505e-1, 6.747e+0, 2.963e+1)         # VM:27-34
G = (1.585, 2.354, 3.486, 6.558, 8.205, 6.498)                            # VM:35-42
LG = (6.658e-5, 1.197e-3, 1.514e-2, 1.672e-1, 7.497e-1, 3.292e+0)         # VM:43-50
K = (7.588e-1, 7.650e-1, 9.806e-1, 7.301e+0, 1.347e+1, 1.090e+1)          # VM:51-59
LK = (5.009e-5, 9.945e-4, 2.022e-3, 1.925e-2, 1.199e-1, 2.033e+0)         # VM:60-68
H, RG, TB, A_S, A_L = 627.8e3, 8.314, 869.0, 9.10e-6, 25.10e-6            # main.py:49-53


def taylor(xi, lam):                                  # VM:233-242
    a = (-1.0 * xi) / lam
    return (1.0 + a) + 0.5 * (a * a)


def phi(T):                                           # VM:156-161
    return math.exp(H / RG * (1 / TB - 1 / T))


def fdiv(a, b):
    """IEEE division (Python raises on /0)."""
    if b == 0.0:
        if a == 0.0 or a != a:
            return float("nan")
        return math.copysign(float("inf"), a) * math.copysign(1.0, b)
    return a / b


def step(d, dt, T_cur, T_prev, Tfp_prev, s_tilde, sig_tilde):
    """One point, one timestep; s_tilde/sig_tilde are [N][d*d] lists."""
    N = len(M)
    ph = phi(T_cur)
    tfp = [(LM[i] * Tfp_prev[i] + T_cur * dt * ph) / (LM[i] + dt * ph) for i in range(N)]   # VM:111-119
    tf = M[0] * tfp[0]
    for i in range(1, N):
        tf = tf + M[i] * tfp[i]                                                              # VM:122-125
    eth = A_S * (T_cur - T_prev) + (A_L - A_S) * (tf - tf)                                   # VM:128-133 (+TVP:481)
    tot = [[(-1.0 * eth) if i == j else -0.0 for j in range(d)] for i in range(d)]           # VM:136-139
    tr = tot[0][0]
    for i in range(1, d):
        tr = tr + tot[i][i]
    dev = [[tot[i][j] - 1 / d * tr if i == j else tot[i][j] for j in range(d)] for i in range(d)]  # VM:142-146
    T_next = T_cur + (T_cur - T_prev)                                                        # VM:150-153
    ph_next = phi(T_next)
    xi = dt / 2 * (ph_next - ph)                                                             # VM:170-173
    sigma = [None] * (d * d)
    s_new = [[0.0] * (d * d) for _ in range(N)]
    k_new = [[0.0] * (d * d) for _ in range(N)]
    s_part = [[0.0] * (d * d) for _ in range(N)]
    k_part = [[0.0] * (d * d) for _ in range(N)]
    for n in range(N):
        tg, tk = taylor(xi, LG[n]), taylor(xi, LK[n])
        for i in range(d):
            for j in range(d):
                c = i * d + j
                ds = fdiv(2.0 * G[n] * dev[i][j], xi) * LG[n] * (1.0 - tg)                    # VM:176-182
                dk = (fdiv(K[n] * tr, xi) * LK[n] * (1.0 - tk)) if i == j else 0.0            # VM:185-191
                s_new[n][c] = s_tilde[n][c] * tg                                             # VM:194-200
                k_new[n][c] = sig_tilde[n][c] * tk                                           # VM:203-209
                s_part[n][c] = ds + s_new[n][c]                                              # VM:212-215
                k_part[n][c] = dk + k_new[n][c]                                              # VM:218-221
                pn = s_part[n][c] + k_part[n][c]
                sigma[c] = pn if n == 0 else sigma[c] + pn                                   # VM:224-228
    return dict(phi=ph, Tf_partial=tfp, Tf=tf, thermal_strain=eth, T_next=T_next, phi_next=ph_next, xi=xi,
                s_tilde=s_new, sigma_tilde=k_new, s_partial=s_part, sigma_partial=k_part, sigma=sigma)


def hx(v):
    if isinstance(v, list):
        return [hx(x) for x in v]
    return float(v).hex()


def main():
    rng = random.Random(20261018)
    cases = []
    # the SURVEY §4(1) point: T_prev=800, T_cur=790, Tf_partial_prev=800, zero history
    for d in (1, 2, 3):
        z = [[0.0] * (d * d) for _ in range(6)]
        cases.append(dict(name=f"survey_kat_d{d}", d=d, dt=0.1, T_cur=790.0, T_prev=800.0, Tfp_prev=[800.0] * 6,
                          s_tilde=z, sigma_tilde=z))
    # random states with non-zero history so the recursions are exercised
    for d in (1, 2, 3):
        for r in range(6):
            T_cur = rng.uniform(650.0, 850.0)
            T_prev = T_cur + rng.uniform(0.05, 1.0) * (1 if r % 3 else -1)
            cases.append(dict(name=f"random_d{d}_{r}", d=d, dt=rng.choice([0.1, 0.05, 1.0]), T_cur=T_cur,
                              T_prev=T_prev, Tfp_prev=[T_prev + rng.uniform(0, 5) for _ in range(6)],
                              s_tilde=[[rng.gauss(0, 1e-3) for _ in range(d * d)] for _ in range(6)],
                              sigma_tilde=[[rng.gauss(0, 1e-3) for _ in range(d * d)] for _ in range(6)]))
    # Q5: T_cur == T_prev bit-exactly -> xi == 0 -> 0/0 = NaN in the stress
    d = 3
    cases.append(dict(name="nan_equal_T_d3", d=3, dt=0.1, T_cur=800.0, T_prev=800.0, Tfp_prev=[800.0] * 6,
                      s_tilde=[[1e-3] * 9 for _ in range(6)], sigma_tilde=[[2e-3] * 9 for _ in range(6)]))
    out = []
    for c in cases:
        res = step(c["d"], c["dt"], c["T_cur"], c["T_prev"], c["Tfp_prev"], c["s_tilde"], c["sigma_tilde"])
        out.append(dict(name=c["name"], d=c["d"],
                        inputs={k: hx(c[k]) for k in ("dt", "T_cur", "T_prev", "Tfp_prev", "s_tilde", "sigma_tilde")},
                        expected={k: hx(v) for k, v in res.items()}))
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "visco_kat.json")
    with open(path, "w") as fh:
        json.dump(dict(note="hand-evaluated from VM:111-242; see make_visco_kat.py", cases=out), fh, indent=1)
    print("wrote", path, len(out), "cases")


if __name__ == "__main__":
    main()
