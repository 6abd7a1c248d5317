"""GPU diagnostic: per-step differences between ThermoViscoProblem (CUDA) and the CPU oracle of the whole time step, for
several meshes and solver settings.  Prints one JSON line per (case, setting) with the worst relative errors of T, Tf,
xi and sigma over the run, the worst excess of |d sigma| over the reference formula's rounding floor
(tests/helpers.stress_rounding_floor) and the PCG iteration count — the data behind the default solver tolerances and the
stress assertions of tests/test_problem_gpu.py.

    python tools/parity_probe.py > gpurun_out/parity_probe.jsonl
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from fem_glass_tempering_b200 import ThermoViscoProblem, _lib          # noqa: E402
from fem_glass_tempering_b200 import mesh as msh                       # noqa: E402
from helpers import stress_rounding_floor                              # noqa: E402
from oracle.reference_problem import OracleProblem                     # noqa: E402
from oracle.visco_oracle import MAIN_PARAMS                            # noqa: E402

DG1 = {"element": "DG", "degree": 1}
CG1 = {"element": "CG", "degree": 1}
CG2 = {"element": "CG", "degree": 2}

CASES = [
    ("main_py_1d_DG1_CG1", lambda: msh.graded_line_mesh(), {"T": DG1, "sigma": CG1}, {}, 60, None),
    ("box3d_6x6x3_DG1", lambda: msh.box_mesh(6, 6, 3, 6.0, 6.0, 3.0), {"T": DG1, "sigma": DG1}, {}, 8, 3),
    ("plate3d_12x12x4_DG1_pen6_cheb4", lambda: msh.plate_mesh(3, (12, 12, 4), (12.0, 12.0, 4.0)), {"T": DG1, "sigma": DG1},
     {"sip_penalty": 6.0}, 8, 4),
    ("rect2d_12x6_CG2", lambda: msh.rectangle_mesh(12, 6, 6.0, 3.0), {"T": CG2, "sigma": CG2}, {}, 8, None),
    ("box3d_4x4x2_CG2", lambda: msh.box_mesh(4, 4, 2, 4.0, 4.0, 2.0), {"T": CG2, "sigma": CG2}, {}, 8, None),
]
SETTINGS = [
    ("default", {}),
    ("lin_rtol=1e-13,atol=1e-13", {"linear_rtol": 1e-13, "atol": 1e-13}),
    ("eta=0", {"forcing_eta": 0.0}),
]


def cpu(F):
    return F.x.array.cpu().numpy()


def rel(a, b):
    m = np.isfinite(b)
    s = np.max(np.abs(b[m])) if m.any() else 1.0
    return float(np.max(np.abs(a[m] - b[m])) / (s if s > 0 else 1.0))


def main():
    ctx = _lib.Context(0)
    for name, mk, cfg, over, steps, cheb in CASES:
        for sname, sets in SETTINGS:
            mesh = mk()
            params = dict(MAIN_PARAMS, **over)
            prob = ThermoViscoProblem(mesh_path="", time=(0.0, 50.0), dt=0.1, config=cfg, model_parameters=params, mesh=mesh,
                                      ctx=ctx, verbose=False, materialize="minimal")
            prob.setup(dirichlet_bc=False)
            if cheb is not None:
                prob._thermal_op.set_chebyshev(cheb)
            for k, v in sets.items():
                setattr(prob.solver, k, v)
            sp = lambda s: dict(dofmap=s.dofmap, ref_nodes=s.element.nodes, family=s.family, degree=s.degree)
            T, S = prob.functionSpaces["T"].scalar, prob.functionSpaces["sigma"].scalar
            orc = OracleProblem(mesh.x, mesh.cells, sp(T), sp(S), params, 0.1)
            d = orc.d
            worst = dict(T=0.0, Tf=0.0, xi=0.0, sigma=0.0, sigma_excess=0.0, sigma_step=0)
            its = newton = 0
            for step in range(steps):
                prob.solve_timestep(t=0.0)
                orc.step()
                f = orc.f
                its += prob.solver.last_stats.lin_its
                newton += prob.solver.last_stats.newton_its
                worst["T"] = max(worst["T"], rel(cpu(prob.functions_current["T"]), f["T_cur"]))
                worst["Tf"] = max(worst["Tf"], rel(cpu(prob.functions_current["Tf"]), f["Tf_cur"]))
                worst["xi"] = max(worst["xi"], rel(cpu(prob.functions["xi"]), f["xi"]))
                dT_at_S = np.abs(orc._T_at_sigma_points(f["T_cur"]) - orc._T_at_sigma_points(f["T_prev"]))
                xi_at_S = np.abs(orc._T_at_sigma_points(f["xi"]))
                node_dT, node_xi = np.zeros(S.n_nodes), np.zeros(S.n_nodes)
                node_dT[S.dofmap.ravel()] = dT_at_S
                node_xi[S.dofmap.ravel()] = xi_at_S
                good = node_dT > 1e-6
                sg = cpu(prob.functions_next["sigma"]).reshape(-1, d * d)[good]
                so = f["sigma_next"].reshape(-1, d * d)[good]
                scale = np.max(np.abs(so))
                err = np.max(np.abs(sg - so), axis=1)
                floor = stress_rounding_floor(orc.vp, node_dT[good], node_xi[good])
                e = float(np.max(err) / scale)
                if e > worst["sigma"]:
                    worst["sigma"], worst["sigma_step"] = e, step + 1
                worst["sigma_excess"] = max(worst["sigma_excess"], float(np.max(err - 2.0 * floor) / scale))
                orc.end_step()
            print(json.dumps(dict(case=name, setting=sname, steps=steps, pcg_its_per_step=its / steps,
                                  newton_its_per_step=newton / steps, cheb=prob._thermal_op.chebyshev_info(), **worst)), flush=True)
            prob._thermal_op.close()


if __name__ == "__main__":
    main()
