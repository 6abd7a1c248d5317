"""Multi-GPU parity check (run under torchrun, one rank per GPU; NCCL):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
        tests/multigpu_check.py

Every rank steps its x-slab of a plate through ThermoViscoProblem (halo exchange + all-reduced PCG scalars through
the library's NCCL communicator); rank 0 additionally steps the WHOLE plate on its GPU with a single-rank context
and compares temperature / fictive temperature / stress on every rank's owned nodes.  Tolerances: T, Tf 1e-10
relative (north_star); stress per node 1e-10 of max + twice the rounding floor of the reference's own formula
(tests/helpers.stress_rounding_floor), and 1e-8 of max norm-wise.
Prints one line 'MULTIGPU_CHECK OK ...' on rank 0 and exits non-zero on failure."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from fem_glass_tempering_b200 import ThermoViscoProblem, _lib, distributed, fe  # noqa: E402
from fem_glass_tempering_b200 import mesh as msh  # noqa: E402
from helpers import stress_rounding_floor  # noqa: E402
from oracle.visco_oracle import ViscoParams  # noqa: E402

PARAMS = {"f": 0.0, "epsilon": 0.93, "sigma": 5.670e-8, "T_ambient": 600.0, "T_0": 800.0, "alpha": 1.0, "htc": 280.1,
          "rho": 2500.0, "cp": 1433.0, "k": 1.0, "H": 627.8e3, "Tb": 869.0e0, "Rg": 8.314,
          "alpha_solid": 9.10e-6, "alpha_liquid": 25.10e-6, "Tf_init": 873.0}
# (dim, cells per axis PER RANK, cell edge [mm], fe_config).  The reference's SIP penalty 5.0/h (TVP:313) does not
# depend on the degree and is only coercive for P2 while the mass term dominates: the DG2 case uses 3 mm cells.
CASES = [
    (3, (6, 5, 3), 1.0, {"T": {"element": "DG", "degree": 1}, "sigma": {"element": "DG", "degree": 1}}),
    (3, (4, 4, 2), 1.0, {"T": {"element": "CG", "degree": 2}, "sigma": {"element": "CG", "degree": 2}}),
    (2, (9, 6), 1.0, {"T": {"element": "CG", "degree": 2}, "sigma": {"element": "CG", "degree": 2}}),
    (2, (8, 5), 3.0, {"T": {"element": "DG", "degree": 2}, "sigma": {"element": "DG", "degree": 2}}),
]
STEPS = 4


def make_problem(mesh, cfg, ctx, part):
    p = ThermoViscoProblem(mesh_path="", time=(0.0, 50.0), dt=0.1, config=cfg, model_parameters=PARAMS, mesh=mesh,
                           ctx=ctx, partition=part, materialize="minimal", verbose=False)
    p.setup(dirichlet_bc=False)
    return p


def main():
    rank, world, local = distributed.init_process_group()
    assert world >= 2, "run under torchrun with >= 2 ranks"
    torch.cuda.set_device(local)
    ctx = distributed.make_context(rank, world, local)
    ctx1 = _lib.Context(local) if rank == 0 else None
    ok = True
    report = []
    for dim, n_per, edge, cfg in CASES:
        n = (n_per[0] * world,) + tuple(n_per[1:])
        lengths = tuple(edge * k for k in n)
        fam, deg = cfg["T"]["element"], cfg["T"]["degree"]
        # CG: the 3-D case runs the gather form of the apply (row-wise share of x.Ax), the 2-D case the scatter form
        # (cell-wise share); small meshes default to the latter (thermal_op.STENCIL_MIN_ROWS)
        os.environ["SG_STENCIL"] = "1" if (fam == "CG" and dim == 3) else "0"
        m, part, info = distributed.slab_partition(dim, n, lengths, fam, deg, rank, world)
        prob = make_problem(m, cfg, ctx, part)
        if fam == "DG":      # the partitioned run uses the Chebyshev-preconditioned solver (halo exchange of every
            assert prob._thermal_op.set_chebyshev(3)      # polynomial step), the single-GPU reference the plain one
        try:
            for _ in range(STEPS):
                prob.solve_timestep(t=0.0)
        except AssertionError as e:      # non-convergence is raised on every rank alike (all-reduced scalars)
            ok = False
            report.append(f"{fam}{deg} d={dim} n={n}: partitioned solve FAILED: {e}")
            if rank == 0:
                print(report[-1], flush=True)
            continue
        space = prob.functionSpaces["T"].scalar
        own = slice(part["own_lo"], part["own_hi"])
        xl = space.tabulate_dof_coordinates()[own]
        d2 = dim * dim
        loc = {"T": prob.functions_current["T"].x.array[own].cpu().numpy(),
               "Tf": prob.functions_current["Tf"].x.array[own].cpu().numpy(),
               "sigma": prob.functions_next["sigma"].x.array.view(-1, d2)[own].cpu().numpy(),
               "x": xl, "its": (prob.solver.last_stats.newton_its, prob.solver.last_stats.lin_its)}
        gathered = [None] * world
        dist.all_gather_object(gathered, loc)
        if rank == 0:
            gm = msh.plate_mesh(dim, n, lengths)
            ref = make_problem(gm, cfg, ctx1, None)
            for k in range(STEPS):
                if k == STEPS - 1:
                    gTp = ref.functions_previous["T"].x.array.cpu().numpy().copy()
                ref.solve_timestep(t=0.0)
            gs = ref.functionSpaces["T"].scalar
            gx = gs.tabulate_dof_coordinates()
            key = lambda X: [tuple(r) for r in np.round(X * 1e6).astype(np.int64)]
            gT = ref.functions_current["T"].x.array.cpu().numpy()
            gTf = ref.functions_current["Tf"].x.array.cpu().numpy()
            gS = ref.functions_next["sigma"].x.array.view(-1, d2).cpu().numpy()
            floor = stress_rounding_floor(ViscoParams(dim=dim, dt=0.1), np.abs(gT - gTp),
                                          np.abs(ref.functions["xi"].x.array.cpu().numpy()))
            scale_S = np.nanmax(np.abs(gS))
            seen = np.zeros(gs.n_nodes, dtype=np.int64)
            worst = {"T": 0.0, "Tf": 0.0, "sigma": 0.0, "sigma_over_floor": 0.0}
            if fam == "DG":
                col = gm.n_cells // n[0] * gs.n_ld
            else:
                lut = {k: i for i, k in enumerate(key(gx))}
            for r, g in enumerate(gathered):
                if fam == "DG":
                    c0, _ = distributed.column_range(n[0], r, world)
                    idx = np.arange(g["T"].size) + c0 * col
                    assert np.allclose(gx[idx], g["x"], atol=1e-9)
                else:
                    idx = np.array([lut[k] for k in key(g["x"])])
                seen[idx] += 1
                worst["T"] = max(worst["T"], np.max(np.abs(g["T"] - gT[idx])) / np.max(np.abs(gT)))
                worst["Tf"] = max(worst["Tf"], np.max(np.abs(g["Tf"] - gTf[idx])) / np.max(np.abs(gTf)))
                a, b = g["sigma"], gS[idx]
                assert np.array_equal(np.isnan(a), np.isnan(b)), "NaN positions of the stress differ"
                fin = ~np.isnan(b)
                if fin.any():
                    worst["sigma"] = max(worst["sigma"], np.max(np.abs(a[fin] - b[fin])) / np.max(np.abs(b[fin])))
                    err = np.where(fin, np.abs(a - b), 0.0).max(axis=1)
                    worst["sigma_over_floor"] = max(worst["sigma_over_floor"], float(np.max(err - 2.0 * floor[idx]) / scale_S))
            tiles = bool((seen == 1).all())
            good = tiles and worst["T"] <= 1e-10 and worst["Tf"] <= 1e-10 and worst["sigma"] <= 1e-8 and worst["sigma_over_floor"] <= 1e-10
            ok = ok and good
            report.append(f"{fam}{deg} d={dim} n={n}: owned ranges tile={tiles} relerr T={worst['T']:.1e} Tf={worst['Tf']:.1e} "
                          f"sigma={worst['sigma']:.1e} (over floor {max(worst['sigma_over_floor'], 0.0):.1e}) its(ref)={ref.solver.last_stats.newton_its}/{ref.solver.last_stats.lin_its} "
                          f"its(part)={gathered[0]['its']} peer_memory={prob._thermal_op.peer_memory} {'ok' if good else 'FAIL'}")
        dist.barrier()
    if rank == 0:
        for line in report:
            print(line)
        print(f"MULTIGPU_CHECK {'OK' if ok else 'FAILED'} world={world}", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
