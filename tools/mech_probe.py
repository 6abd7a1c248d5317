"""GPU diagnostic of the mechanical-equilibrium extension (csrc/mech.cu): time per tangent apply, PCG iterations and
time per time step with model_parameters["mechanics"] on plates of growing size.  One JSON line per plate.

    python tools/mech_probe.py [nx ny nz ...] > gpurun_out/mech_probe.jsonl
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from fem_glass_tempering_b200 import ThermoViscoProblem, _lib          # noqa: E402
from fem_glass_tempering_b200 import mesh as msh                       # noqa: E402

MAIN_PARAMS = {"f": 0.0, "epsilon": 0.93, "sigma": 5.670e-8, "T_ambient": 600.0, "T_0": 800.0, "alpha": 1.0, "htc": 280.1,
               "rho": 2500.0, "cp": 1433.0, "k": 1.0, "H": 627.8e3, "Tb": 869.0e0, "Rg": 8.314, "alpha_solid": 9.10e-6,
               "alpha_liquid": 25.10e-6, "Tf_init": 873.0}
DG1 = {"element": "DG", "degree": 1}


def run(ctx, n, steps=4, physics="reference", rtol=1e-8):
    mesh = msh.plate_mesh(3, n, tuple(float(k) for k in n))
    params = dict(MAIN_PARAMS, sip_penalty=6.0, physics=physics, mechanics={"rtol": rtol})
    prob = ThermoViscoProblem(mesh_path="", time=(0.0, 50.0), dt=0.1, config={"T": DG1, "sigma": DG1}, model_parameters=params,
                              mesh=mesh, ctx=ctx, verbose=False, materialize="minimal")
    prob.setup(dirichlet_bc=False)
    me = prob.mechanics
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    its, ms_step, ms_mech = [], [], []
    for s in range(steps):
        prob.t += prob.dt
        ev[0].record()
        prob._solve_T()
        prob._solve_viscoelastic()
        ev[1].record()
        prob._solve_mechanics()
        ev[2].record()
        prob._update_values(current=prob.functions_current["T"], previous=prob.functions_previous["T"])
        torch.cuda.synchronize()
        ms_step.append(ev[0].elapsed_time(ev[2]))
        ms_mech.append(ev[1].elapsed_time(ev[2]))
        its.append(me.last_iters)
    # one tangent apply, timed alone
    nv3 = mesh.n_vertices * 3
    x, y = torch.randn(nv3, dtype=torch.float64, device="cuda"), torch.empty(nv3, dtype=torch.float64, device="cuda")
    for _ in range(3):
        me.apply(x, y)
    ev[0].record()
    for _ in range(20):
        me.apply(x, y)
    ev[1].record()
    torch.cuda.synchronize()
    t_apply = ev[0].elapsed_time(ev[1]) / 20
    out = dict(cells=list(n), n_cells=mesh.n_cells, n_vertices=mesh.n_vertices, physics=physics, rtol=rtol, pcg_its=its,
               ms_per_step=[round(v, 3) for v in ms_step], ms_mechanics=[round(v, 3) for v in ms_mech],
               ms_apply=round(t_apply, 4), apply_bytes=me.apply_bytes(), apply_GBs=round(me.apply_bytes() / t_apply / 1e6, 1),
               max_u=float(prob.functions["displacement"].x.array.abs().max()),
               max_sigma=float(prob.functions_next["sigma"].x.array.abs().max()))
    print(json.dumps(out), flush=True)
    me.close()


if __name__ == "__main__":
    ctx = _lib.Context(0)
    args = [int(a) for a in sys.argv[1:]]
    plates = [tuple(args[i:i + 3]) for i in range(0, len(args), 3)] or [(48, 48, 8), (160, 160, 8)]
    for n in plates:
        run(ctx, n)
    run(ctx, plates[0], physics="corrected")
