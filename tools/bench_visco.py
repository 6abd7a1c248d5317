"""Micro-benchmark of the fused viscoelastic update kernel (hot path A) on one GPU.

usage: python tools/bench_visco.py [--n NODES] [--d DIM] [--terms N] [--iters K] [--full]
Prints one JSON line per configuration: node updates/s, achieved algorithmic GB/s and the
fraction of the measured HBM copy peak (MEASURED_PEAKS.json) / 8 TB/s spec.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fem_glass_tempering_b200 import _lib  # noqa: E402

PRONY = dict(m=(5.523e-2, 8.205e-2, 1.215e-1, 2.286e-1, 2.860e-1, 2.265e-1),
             lambda_m=(5.965e-4, 1.077e-2, 1.362e-1, 1.505e-1, 6.747e+0, 2.963e+1),
             g=(1.585, 2.354, 3.486, 6.558, 8.205, 6.498),
             lambda_g=(6.658e-5, 1.197e-3, 1.514e-2, 1.672e-1, 7.497e-1, 3.292e+0),
             k=(7.588e-1, 7.650e-1, 9.806e-1, 7.301e+0, 1.347e+1, 1.090e+1),
             lambda_k=(5.009e-5, 9.945e-4, 2.022e-3, 1.925e-2, 1.199e-1, 2.033e+0))


def tables(N):
    if N <= 6:
        return {k: v[:N] for k, v in PRONY.items()}
    lam = tuple(float(v) for v in np.logspace(-5, 2, N))
    w = np.linspace(1.0, 2.0, N)
    w = w / w.sum()
    return dict(m=tuple(w), lambda_m=lam, g=tuple(w * sum(PRONY["g"])), lambda_g=lam,
                k=tuple(w * sum(PRONY["k"])), lambda_k=lam)


def peak_gbs():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"], "measured"
    except Exception:
        return 6650.0, "fallback"


def run(ctx, n, d, N, iters, full, corrected=False):
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(1234)
    dd = d * d
    t = {"T_cur": 650 + 200 * torch.rand(n, dtype=torch.float64, device=dev, generator=g)}
    t["T_prev"] = t["T_cur"] + 0.05 + 0.95 * torch.rand(n, dtype=torch.float64, device=dev, generator=g)
    t["Tf_partial"] = t["T_prev"].repeat_interleave(N) + 5 * torch.rand(n * N, dtype=torch.float64, device=dev, generator=g)
    t["s_tilde"] = 1e-3 * torch.randn(n * N * dd, dtype=torch.float64, device=dev, generator=g)
    t["sigma_tilde"] = 1e-3 * torch.randn(n * N * dd, dtype=torch.float64, device=dev, generator=g)
    for name, bs in (("Tf", 1), ("phi", 1), ("xi", 1), ("sigma", dd)):
        t[name] = torch.zeros(n * bs, dtype=torch.float64, device=dev)
    if full:
        for name in ("T_next", "phi_next"):
            t[name] = torch.zeros(n, dtype=torch.float64, device=dev)
        for name in ("thermal_strain", "total_strain", "deviatoric_strain"):
            t[name] = torch.zeros(n * dd, dtype=torch.float64, device=dev)
        for name in ("ds_partial", "dsigma_partial", "s_partial", "sigma_partial"):
            t[name] = torch.zeros(n * N * dd, dtype=torch.float64, device=dev)
    plan = _lib.ViscoPlan(ctx, dim=d, dt=0.1, H=627.8e3, Rg=8.314, Tb=869.0, alpha_solid=9.10e-6,
                          alpha_liquid=25.10e-6, mode=_lib.VISCO_CORRECTED if corrected else _lib.VISCO_REFERENCE, **tables(N))
    bpn = plan.bytes_per_node(t)
    for _ in range(3):
        plan.update(n, t)
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in evs:
        a.record()
        plan.update(n, t)
        b.record()
    torch.cuda.synchronize()
    ms = sorted(a.elapsed_time(b) for a, b in evs)
    med = ms[len(ms) // 2]
    gbs = n * bpn / (med * 1e-3) / 1e9
    pk, how = peak_gbs()
    return dict(kernel="visco_fused" + ("_corrected_scheme" if corrected else ""), n_nodes=n, dim=d, terms=N, full_materialisation=full, bytes_per_node=bpn,
                ms_median=round(med, 4), ms_min=round(ms[0], 4), node_updates_per_s=n / (med * 1e-3),
                achieved_GBs=round(gbs, 1), frac_of_peak=round(gbs / pk, 3), peak=f"{pk} GB/s {how}",
                frac_of_8TBs=round(gbs / 8000, 3))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=19_660_800)
    ap.add_argument("--d", type=int, default=3)
    ap.add_argument("--terms", type=int, default=6)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--full", action="store_true")
    ap.add_argument("--sweep", action="store_true", help="BASELINE config 5: N in {3,4,6,8,10,12}, plus d=1,2")
    a = ap.parse_args()
    ctx = _lib.Context(0)
    if a.sweep:
        for N in (3, 4, 6, 8, 10, 12):
            print(json.dumps(run(ctx, a.n * 6 // N if N > 6 else a.n, 3, N, a.iters, False)), flush=True)
        print(json.dumps(run(ctx, 334_153, 2, 6, a.iters, False)), flush=True)
        print(json.dumps(run(ctx, 30_710_797, 2, 6, a.iters, False)), flush=True)
        print(json.dumps(run(ctx, 30_710_797, 1, 6, a.iters, False)), flush=True)
        print(json.dumps(run(ctx, a.n // 3, 3, 6, a.iters, True)), flush=True)
        print(json.dumps(run(ctx, a.n, 3, 6, a.iters, False, corrected=True)), flush=True)
    else:
        print(json.dumps(run(ctx, a.n, a.d, a.terms, a.iters, a.full)), flush=True)


if __name__ == "__main__":
    main()
