"""Micro-benchmark of the CG Jacobian apply on one GPU's share of config 4 (3-D CG2 plate): the gather form
(row-stencil classes, csrc/stencil.cu) against the cell-centric class kernel (cg_class_apply, RED.ADD scatter), same
process, same mesh.  Kernel times come from CUDA event pairs on the launch stream (sg_thermal_profile); the first
invocation writes --out, later ones (e.g. the same command line under ncu) leave it alone.

    python tools/cg_apply_probe.py --out gpurun_out/cg_probe.json
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fem_glass_tempering_b200 import _lib, fe  # noqa: E402
from fem_glass_tempering_b200 import mesh as msh  # noqa: E402
from fem_glass_tempering_b200.thermal_op import ThermalOperator  # noqa: E402

PARAMS = {"alpha": 1.0, "f": 0.0, "sigma": 5.670e-8, "epsilon": 0.93, "htc": 280.1, "T_ambient": 600.0}   # thermal subset of main.py:29-55


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, nargs=3, default=(96, 768, 6))
    ap.add_argument("--degree", type=int, default=2)
    ap.add_argument("--reps", type=int, default=40)
    ap.add_argument("--out", default=None)
    ap.add_argument("--variants", type=int, nargs="*", default=[], help="also time these SG_STENCIL_VARIANT kernels (csrc/stencil.cu)")
    a = ap.parse_args()
    t0 = time.time()
    m = msh.plate_mesh(3, tuple(a.n), tuple(float(k) for k in a.n))
    space = fe.ScalarSpace(m, "CG", a.degree)
    ctx = _lib.Context(0)
    L = _lib.lib()
    n = space.n_nodes
    T = torch.full((n,), 800.0, dtype=torch.float64, device="cuda:0")
    x = torch.rand(n, dtype=torch.float64, device="cuda:0")
    y = torch.empty(n, dtype=torch.float64, device="cuda:0")
    res = {"cells": int(m.n_cells), "rows": int(n), "degree": a.degree, "setup_s": None, "forms": {}}
    forms = [("row_stencil", True, "0"), ("cell_scatter", False, "0")]
    forms += [(f"row_stencil_variant{v}", True, str(v)) for v in a.variants]
    for name, st, variant in forms:
        os.environ["SG_STENCIL_VARIANT"] = variant        # read by the library when the operator is created
        op = ThermalOperator(ctx, space, PARAMS, 0.1, use_stencil=st, cheb_degree=0)
        for _ in range(3):
            op.jac_apply(T, x, y)
        _lib.check(L.sg_thermal_profile(op.handle, 1, a.reps + 8))
        for _ in range(a.reps):
            op.jac_apply(T, x, y)
        torch.cuda.synchronize()
        k, ms = C.c_int64(0), C.c_double(0.0)
        _lib.check(L.sg_thermal_profile_read_kind(op.handle, 0, C.byref(k), C.byref(ms)))
        _lib.check(L.sg_thermal_profile(op.handle, 0, 0))
        us = 1e3 * ms.value / max(1, k.value)
        res["forms"][name] = {"us_per_apply": us, "launches": int(k.value), "layout_bytes": op.apply_bytes(),
                              "layout_GBs": op.apply_bytes() / (us * 1e-6) / 1e9, "stencil": op.stencil_info(),
                              "cell_classes": op.class_info(), "checksum": float(y.double().sum().item())}
        del op
    res["setup_s"] = round(time.time() - t0, 1)
    line = json.dumps(res)
    print(line)
    if a.out and not os.path.exists(a.out):
        with open(a.out, "w") as fh:
            fh.write(line + "\n")


if __name__ == "__main__":
    main()
