"""CPU, world_size 2 and 3 over gloo: the element partition / halo plan of fem_glass_tempering_b200.distributed.

Each rank builds its slab (with ghost columns), fills the OWNED dofs of a smooth field, runs the forward scatter
described by the halo segments (here with gloo send/recv on host arrays; on the GPU the same segments drive
sg_halo_forward's ncclSend/ncclRecv), and checks that
  * every local dof then holds the global field's value (ghost ranges are exactly the received ranges),
  * the table-driven operator applied to the rank's cell range reproduces the global operator on owned dofs,
  * the owned ranges tile the global dof set (all-reduced owned dot product == global dot product).
"""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))

CASES = [(2, (6, 3), "CG", 1), (2, (6, 3), "CG", 2), (2, (6, 3), "DG", 1), (2, (7, 2), "DG", 2),
         (3, (5, 2, 2), "DG", 1), (3, (4, 2, 2), "CG", 2), (3, (5, 2, 2), "CG", 1)]


def field(x):
    return 700.0 + 10.0 * np.sin(0.7 * x[:, 0]) + 3.0 * x[:, 1] ** 2 + (x[:, 2] if x.shape[1] > 2 else 0.0)


def _worker(rank, world, port, results, outdir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import kernel_mirror
    from fem_glass_tempering_b200 import distributed, fe
    from fem_glass_tempering_b200 import mesh as msh
    from oracle.visco_oracle import MAIN_PARAMS
    ok = True
    try:
        for dim, n, family, degree in CASES:
            lengths = tuple(float(k) for k in n)
            m, part, info = distributed.slab_partition(dim, n, lengths, family, degree, rank, world)
            space = fe.ScalarSpace(m, family, degree)
            xl = space.tabulate_dof_coordinates()
            u = field(xl)
            v = np.full(space.n_nodes, np.nan)
            v[part["own_lo"]:part["own_hi"]] = u[part["own_lo"]:part["own_hi"]]
            # forward scatter over the halo segments
            vt = torch.from_numpy(v)
            reqs = []
            for peer, so, sc, ro, rc in part["halo"]:
                if sc:
                    reqs.append(dist.isend(vt[so:so + sc].clone(), peer))
                if rc:
                    reqs.append(dist.irecv(vt[ro:ro + rc], peer))
            for r in reqs:
                r.wait()
            assert not np.isnan(v).any(), f"{family}{degree} d={dim}: unfilled ghost dofs on rank {rank}"
            assert np.allclose(v, u, rtol=0, atol=1e-12), f"{family}{degree} d={dim}: wrong ghost values on rank {rank}"
            # operator on this rank's cells == global operator on owned dofs
            gm = msh.plate_mesh(dim, n, lengths)
            gs = fe.ScalarSpace(gm, family, degree)
            gx = gs.tabulate_dof_coordinates()
            tabs = fe.operator_tables(dim, degree)
            ug = field(gx)
            rng = np.random.default_rng(5)
            pert = rng.standard_normal(gs.n_nodes)
            yg = kernel_mirror.jac_apply(gs, tabs, fe.cell_geometry(gm), fe.facet_topology(gm), MAIN_PARAMS, 0.1,
                                         pert, T_lin=ug)
            # map local dofs to global dofs by coordinates (CG) or by cell offset (DG)
            if family == "DG":
                c0 = info["columns"][0] - info["ghost_left"]
                col = m.n_cells // (info["columns"][1] - info["columns"][0] + info["ghost_left"] + info["ghost_right"])
                l2g = np.arange(space.n_nodes) + c0 * col * space.n_ld
            else:
                key = lambda X: np.round(X * 1e6).astype(np.int64)
                lut = {tuple(k): i for i, k in enumerate(key(gx))}
                l2g = np.array([lut[tuple(k)] for k in key(xl)])
            yl = kernel_mirror.jac_apply(space, tabs, fe.cell_geometry(m), fe.facet_topology(m, part["exterior_mask"]),
                                         MAIN_PARAMS, 0.1, pert[l2g], T_lin=ug[l2g], cell_lo=part["cell_lo"],
                                         cell_hi=part["cell_hi"])
            own = slice(part["own_lo"], part["own_hi"])
            assert np.max(np.abs(yl[own] - yg[l2g][own])) <= 1e-11 * np.max(np.abs(yg)), \
                f"{family}{degree} d={dim}: partitioned operator differs on rank {rank}"
            # owned ranges tile the global dof set
            loc = torch.tensor([float(np.dot(pert[l2g][own], yl[own])), float(part["own_hi"] - part["own_lo"])],
                               dtype=torch.float64)
            dist.all_reduce(loc)
            assert abs(loc[0].item() - float(np.dot(pert, yg))) <= 1e-9 * abs(float(np.dot(pert, yg)))
            assert int(loc[1].item()) == gs.n_nodes
            # the solver's fused reduction: x_K . (A_K x_K) over each rank's OWN cells (sg_thermal_desc.own_cell_lo/hi)
            # sums to the global x.Ax, i.e. every global cell is owned by exactly one rank
            yo = kernel_mirror.jac_apply(space, tabs, fe.cell_geometry(m), fe.facet_topology(m, part["exterior_mask"]),
                                         MAIN_PARAMS, 0.1, pert[l2g], T_lin=ug[l2g], cell_lo=part["own_cell_lo"],
                                         cell_hi=part["own_cell_hi"])
            cw = torch.tensor([float(np.dot(pert[l2g], yo))], dtype=torch.float64)
            dist.all_reduce(cw)
            assert abs(cw.item() - float(np.dot(pert, yg))) <= 1e-9 * abs(float(np.dot(pert, yg))), \
                f"{family}{degree} d={dim}: cell-wise x.Ax does not tile"
            pts = torch.tensor([float(info["owned_cell_points"])], dtype=torch.float64)
            dist.all_reduce(pts)
            assert int(pts.item()) == gm.n_cells * space.n_ld
            # per-rank output files (output.RankFiles): every rank writes its OWNED nodes, read_series() reassembles the
            # unpartitioned field; the ranks never write the same file
            from fem_glass_tempering_b200.output import RankFiles, read_series
            assert part["global_own_offset"] == l2g[part["own_lo"]] and part["global_n_dofs"] == gs.n_nodes
            case_dir = os.path.join(outdir, f"{family}{degree}_{dim}d")
            fields = {"T": {"block_size": 1, "name": "T", "n_nodes": space.n_nodes},
                      "sigma": {"block_size": dim * dim, "name": "sigma", "n_nodes": space.n_nodes}}
            rf = RankFiles(case_dir, rank, world, fields, (part["own_lo"], part["own_hi"]), part["global_own_offset"],
                           part["global_n_dofs"])
            rf.save_coordinates("T", xl)
            for step in range(2):
                rf.save_step(step, 0.1 * step, {"T": u + step, "sigma": np.repeat(u, dim * dim) * (1 + np.tile(np.arange(dim * dim), u.size)) + step})
            rf.close()
            dist.barrier()
            if rank == 0:
                t, Tser = read_series(case_dir, "T")
                assert np.allclose(t, [0.0, 0.1]) and Tser.shape == (2, gs.n_nodes)
                assert np.array_equal(Tser[1], ug + 1)
                _, Sser = read_series(case_dir, "sigma")
                assert np.array_equal(Sser[1], np.repeat(ug, dim * dim) * (1 + np.tile(np.arange(dim * dim), ug.size)) + 1)
                names = sorted(os.listdir(case_dir))
                assert len(names) == len(set(names)) == world * (2 * 2 + 2)      # 2 steps x 2 fields + index + coordinates, per rank
    except Exception as e:  # noqa: BLE001
        ok = False
        results[rank] = repr(e)
    if ok:
        results[rank] = "ok"
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_slab_partition_over_gloo(world, tmp_path):
    port = 29600 + world + (os.getpid() % 200)
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker, args=(world, port, results, str(tmp_path)), nprocs=world, join=True)
    assert all(results.get(r) == "ok" for r in range(world)), dict(results)
