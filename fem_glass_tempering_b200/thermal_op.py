"""Host-side builder of the CUDA heat-equation operator: uploads the mesh/element arrays of one
function space and wraps sg_thermal_op / sg_thermal_solver / sg_halo_plan.

This is the product's replacement for `NonlinearProblem(F, u)` + `NewtonSolver` (ThermoViscoProblem.py:
330-346).  Set-up is numpy on the host; every operation afterwards is a CUDA call through the C ABI.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib, fe
from ._lib_thermal import HaloSegmentC, NewtonOptsC, NewtonStatsC, ThermalDescC

PENALTY = 5.0  # ThermoViscoProblem.py:313; model_params["sip_penalty"] overrides it (the reference's 5.0 is not coercive
               # on tetrahedra: its 3-D DG time stepping is unstable, tests/test_fe_tables.py)


def _np_ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


STENCIL_MIN_ROWS = 0             # the gather form is the default wherever the library finds repeating rows (see use_stencil)


class ThermalOperator:
    """Matrix-free Jacobian/residual of the heat equation on one GPU (one rank's part of the mesh).

    space         fe.ScalarSpace of T
    params        model_params dict (main.py:29-55)
    partition     optional dict(cell_lo, cell_hi, own_lo, own_hi, exterior_mask, halo=[(peer, send_off, send_cnt,
                  recv_off, recv_cnt), ...]) describing this rank's slab; default: everything owned.
    """

    def __init__(self, ctx: _lib.Context, space: fe.ScalarSpace, params: dict, dt: float, partition: dict | None = None,
                 use_classes: bool = True, cheb_degree: int | None = None, use_stencil: bool | None = None):
        import torch
        self.ctx, self.space, self.dt = ctx, space, float(dt)
        mesh, d = space.mesh, space.mesh.dim
        dev = torch.device("cuda", ctx.device)
        part = partition or {}
        tabs = fe.operator_tables(d, space.degree)
        # set-up arithmetic (geometry, facet sort) runs on the device as torch array code; only small index arrays come back
        det_d, jinv_d, h_d = fe.cell_geometry_device(mesh, dev)
        topo = fe.facet_topology(mesh, part.get("exterior_mask"), device=dev)
        nc = mesh.n_cells
        self.tabs, self.n_dofs = tabs, space.n_nodes
        # geometry SoA [d*d + 2][nc]
        g = torch.empty((d * d + 2, nc), dtype=torch.float64, device=dev)
        g[: d * d] = jinv_d.reshape(nc, d * d).T
        g[d * d], g[d * d + 1] = det_d, h_d
        keep = {}
        keep["geom"] = g
        dg = space.family == "DG"
        if dg:
            keep["nbr"] = torch.from_numpy(np.ascontiguousarray(topo.neighbor.T.astype(np.int32))).to(dev)
            info = np.zeros(nc, dtype=np.int64)
            for f in range(d + 1):
                info |= ((topo.nb_facet[:, f].astype(np.int64) & 3) | (topo.nb_perm[:, f].astype(np.int64) << 2)) << (5 * f)
            keep["nbinfo"] = torch.from_numpy(info.astype(np.int32)).to(dev)
        else:
            keep["dofmap"] = torch.from_numpy(np.ascontiguousarray(space.dofmap.T.astype(np.int32))).to(dev)
        nbf = topo.bnd_cell.size
        if nbf:
            keep["bf_cell"] = torch.from_numpy(topo.bnd_cell.astype(np.int32)).to(dev)
            keep["bf_facet"] = torch.from_numpy(topo.bnd_facet.astype(np.int32)).to(dev)
            bc = torch.from_numpy(topo.bnd_cell.astype(np.int64)).to(dev)
            geo_b = fe.CellGeometry(det_d[bc].cpu().numpy(), jinv_d[bc].cpu().numpy(), None)
            keep["bf_area"] = torch.from_numpy(fe.facet_measures(mesh, geo_b, np.arange(nbf), topo.bnd_facet)).to(dev)
        del det_d, jinv_d, h_d
        self._keep = keep
        # host tables (copied by the library during create)
        h = {k: np.ascontiguousarray(getattr(tabs, k), dtype=np.float64)
             for k in ("mass", "load", "cq_w", "cq_grad", "fq_w", "fq_val", "fq_grad", "bq_w", "bq_val")}
        h["fq_perm"] = np.ascontiguousarray(tabs.fq_perm, dtype=np.int32)
        desc = ThermalDescC()
        desc.dim, desc.degree, desc.family = d, space.degree, 1 if dg else 0
        desc.n_cells, desc.cell_lo, desc.cell_hi = nc, part.get("cell_lo", 0), part.get("cell_hi", nc)
        desc.n_dofs, desc.own_lo, desc.own_hi = space.n_nodes, part.get("own_lo", 0), part.get("own_hi", space.n_nodes)
        desc.own_cell_lo, desc.own_cell_hi = part.get("own_cell_lo", desc.cell_lo), part.get("own_cell_hi", desc.cell_hi)
        if use_stencil is None:
            # CG: gather form of the apply where rows repeat (csrc/stencil.cu): one launch, exterior facets included, no
            # atomics; small problems (config 2) then run their whole PCG solve as one persistent kernel.
            # SG_NO_STENCIL=1 keeps the cell-centric scatter kernel (tests compare the two forms).
            import os
            env = os.environ
            use_stencil = (env.get("SG_NO_STENCIL", "0") != "1") and (env.get("SG_STENCIL", "1") != "0") and space.n_nodes >= STENCIL_MIN_ROWS
        # SG_THERMAL_NO_CLASSES, SG_THERMAL_NO_STENCIL
        desc.flags = (0 if use_classes else 1) | (0 if use_stencil else 8)
        for name in ("dofmap", "geom", "nbr", "nbinfo", "bf_cell", "bf_facet", "bf_area"):
            setattr(desc, name, _lib.ptr(keep.get(name)))
        desc.n_bfacets = nbf
        desc.n_ld, desc.nqc, desc.nqf, desc.nqb = tabs.n_ld, tabs.cq_w.size, tabs.fq_w.size, tabs.bq_w.size
        desc.n_perm = tabs.fq_perm.shape[0]
        for name, arr in h.items():
            setattr(desc, name, _np_ptr(arr))
        desc.dt, desc.alpha, desc.f = self.dt, float(params["alpha"]), float(params["f"])
        desc.sigma, desc.epsilon = float(params["sigma"]), float(params["epsilon"])
        desc.htc, desc.T_ambient, desc.penalty = float(params["htc"]), float(params["T_ambient"]), float(params.get("sip_penalty", PENALTY))
        self.own_lo, self.own_hi = desc.own_lo, desc.own_hi
        self.cell_lo, self.cell_hi = desc.cell_lo, desc.cell_hi
        self.n_bfacets = nbf
        L = _lib.lib()
        hnd = C.c_void_p()
        _lib.check(L.sg_thermal_op_create(ctx.handle, C.byref(desc), C.byref(hnd)))
        self.handle = hnd
        if ctx.nranks > 1:
            # the ranks' shares of the fused x.Ax tile the global sum only if all of them reduce the same way
            # (row-wise with the stencil form, cell-wise with the class kernels, dof-wise with the general kernel)
            import torch.distributed as dist
            mine = (self.class_info()["active"], self.stencil_info()["active"])
            paths = [None] * ctx.nranks
            dist.all_gather_object(paths, mine)
            if any(p != mine for p in paths):
                _lib.check(L.sg_thermal_op_destroy(self.handle))
                desc.flags |= 8 | (0 if all(p[0] for p in paths) else 1)
                hnd = C.c_void_p()
                _lib.check(L.sg_thermal_op_create(ctx.handle, C.byref(desc), C.byref(hnd)))
                self.handle = hnd
        # halo plan + solver
        self.halo = None
        segs = part.get("halo") or []
        if segs:
            arr = (HaloSegmentC * len(segs))()
            for i, (peer, so, sc, ro, rc) in enumerate(segs):
                arr[i].peer, arr[i].send_offset, arr[i].send_count = peer, so, sc
                arr[i].recv_offset, arr[i].recv_count = ro, rc
            hh = C.c_void_p()
            _lib.check(L.sg_halo_plan_create(ctx.handle, len(segs), arr, C.byref(hh)))
            self.halo = hh
        self.peer_memory = False
        nws = max(int(L.sg_thermal_solver_workspace_doubles(self.handle)), 1)
        ws_ptr = None
        if ctx.nranks > 1:
            ws_ptr = self._setup_peer_memory(L, nws, segs)
        if ws_ptr:
            self.workspace = None                   # the vectors live in the IPC-exported communication block (direct halo puts)
        else:
            self.workspace = torch.zeros(nws, dtype=torch.float64, device=dev)
            ws_ptr = self.workspace.data_ptr()
        self._ws_ptr, self._ws_doubles = int(ws_ptr), nws
        sh = C.c_void_p()
        _lib.check(L.sg_thermal_solver_create(self.handle, ws_ptr, self.halo, C.byref(sh)))
        self.solver = sh
        self.opts = NewtonOptsC(1e-12, 1e-10, 50, 1e-12, 0.0, 10000, 1e-3)
        self.chebyshev_degree = 0
        if cheb_degree is None:
            # polynomial preconditioning trades CG vector updates for operator applications: worth it on large DG
            # meshes, not on the 48-cell line of main.py (plain CG terminates early there)
            cheb_degree = 4 if space.mesh.n_cells >= 20000 else 0
        if cheb_degree:
            self.set_chebyshev(cheb_degree)
        self.last_stats = None

    def _setup_peer_memory(self, L, workspace_doubles: int, segs):
        """NVLink peer-memory transport of the halo / small all-reduces (sg_halo_peer_alloc/open): every rank exports one
        communication block that also holds the solver workspace; the 64-byte IPC handles and each rank's vector layout
        (stride, where the rows from below / above land) travel through torch.distributed.  Collective over all ranks;
        SG_NO_PEER=1 keeps NCCL on the data path.  Returns the workspace pointer inside the block, or None."""
        import os
        import torch.distributed as dist
        handle = C.create_string_buffer(64)
        rc = 0
        if self.halo is not None and os.environ.get("SG_NO_PEER", "0") != "1":
            rc = L.sg_halo_peer_alloc(self.halo, handle, 0 if os.environ.get("SG_NO_PEER_WS", "0") == "1" else workspace_doubles)
        below = next((ro for peer, so, sc, ro, rcn in segs if peer < self.ctx.rank), 0)
        above = next((ro for peer, so, sc, ro, rcn in segs if peer > self.ctx.rank), 0)
        mine = (handle.raw, (int(self.n_dofs), int(below), int(above))) if rc == 1 else None
        everyone = [None] * self.ctx.nranks
        dist.all_gather_object(everyone, mine)
        # every rank runs the same two collectives, also one without a halo plan (it reports False and
        # sg_thermal_solver_create then rejects the missing plan on that rank instead of the others hanging here)
        ok = False
        if self.halo is not None:
            if all(h is not None for h in everyone):
                blob = C.create_string_buffer(b"".join(h[0] for h in everyone), 64 * self.ctx.nranks)
                layout = (C.c_int64 * (3 * self.ctx.nranks))(*[v for h in everyone for v in h[1]])
                ok = L.sg_halo_peer_open(self.halo, blob, layout) == 0
            else:
                L.sg_halo_peer_open(self.halo, None, None)
        flags = [None] * self.ctx.nranks
        dist.all_gather_object(flags, bool(ok and L.sg_halo_uses_peer_memory(self.halo)))
        if self.halo is None:
            return None
        if not all(flags):                       # one rank could not map a neighbour: nobody uses the peer path
            L.sg_halo_peer_open(self.halo, None, None)
        self.peer_memory = bool(L.sg_halo_uses_peer_memory(self.halo))
        return L.sg_halo_peer_workspace(self.halo) if self.peer_memory else None

    # -- raw operator calls (asynchronous on the current stream) ---------------------------------
    def residual(self, T, T_prev, out):
        _lib.check(_lib.lib().sg_thermal_residual(self.handle, _lib.ptr(T), _lib.ptr(T_prev), _lib.ptr(out),
                                                  _lib.current_stream_ptr()))
        return out

    def jac_apply(self, T_lin, x, out):
        _lib.check(_lib.lib().sg_thermal_jac_apply(self.handle, _lib.ptr(T_lin), _lib.ptr(x), _lib.ptr(out),
                                                   _lib.current_stream_ptr()))
        return out

    def jac_diag(self, T_lin, out):
        _lib.check(_lib.lib().sg_thermal_jac_diag(self.handle, _lib.ptr(T_lin), _lib.ptr(out), _lib.current_stream_ptr()))
        return out

    def set_chebyshev(self, degree: int, lo: float = 0.0, hi: float = 0.0) -> bool:
        """Chebyshev polynomial preconditioner of the DG solver (sg_thermal_solver_set_chebyshev); False if unavailable."""
        rc = _lib.lib().sg_thermal_solver_set_chebyshev(self.solver, int(degree), float(lo), float(hi))
        if rc < 0:
            _lib.check(rc)
        self.chebyshev_degree = int(degree) if rc == 1 else 0
        return rc == 1

    def chebyshev_info(self) -> dict:
        d, lo, hi = C.c_int32(0), C.c_double(0.0), C.c_double(0.0)
        _lib.check(_lib.lib().sg_thermal_solver_get_chebyshev(self.solver, C.byref(d), C.byref(lo), C.byref(hi)))
        return dict(degree=d.value, lo=lo.value, hi=hi.value)

    def uses_graphs(self) -> bool:
        """True when the solver replays its PCG batches as CUDA graphs (plain PCG on one GPU): CUDA-event pairs cannot sit
        between graph nodes, so kernel timing then needs a separate profiled pass."""
        return self.ctx.nranks == 1 and self.chebyshev_info()["degree"] == 0

    def solver_description(self) -> str:
        ci = self.chebyshev_info()
        if ci["degree"]:
            return (f"CG preconditioned by a degree-{ci['degree']} Chebyshev polynomial in M^-1 J on [{ci['lo']:.3g}, {ci['hi']:.3g}] "
                    "(pcg its = outer iterations)")
        return "CG + " + ("element-mass blocks" if self.space.family == "DG" else "point Jacobi")

    def class_info(self) -> dict:
        """Local-matrix classes found by the library (sg_thermal_class_info)."""
        g, s_, f = C.c_int32(0), C.c_int32(0), C.c_int32(0)
        rc = _lib.lib().sg_thermal_class_info(self.handle, C.byref(g), C.byref(s_), C.byref(f))
        if rc < 0:
            _lib.check(rc)
        return dict(active=bool(rc), geometry=g.value, self=s_.value, facet=f.value)

    def stencil_info(self) -> dict:
        """Row-stencil form of the CG apply (sg_thermal_stencil_info): in use or not, table sizes."""
        n, e, m = C.c_int32(0), C.c_int32(0), C.c_int32(0)
        rc = _lib.lib().sg_thermal_stencil_info(self.handle, C.byref(n), C.byref(e), C.byref(m))
        if rc < 0:
            _lib.check(rc)
        return dict(active=bool(rc), classes=n.value, entries=e.value, max_nnz=m.value)

    def apply_bytes(self) -> int:
        return int(_lib.lib().sg_thermal_apply_bytes(self.handle))

    def halo_forward(self, vec, block_size: int = 1):
        if self.halo is not None:
            _lib.check(_lib.lib().sg_halo_forward(self.halo, _lib.ptr(vec), block_size, _lib.current_stream_ptr()))

    # -- solves (synchronous) ------------------------------------------------------------------------
    def pcg(self, T_lin, b, x, rtol=1e-10, atol=0.0, max_it=10000):
        """Solve J(T_lin) x = b; the Jacobi diagonal must have been set by `prepare_preconditioner`."""
        it, res = C.c_int32(0), C.c_double(0.0)
        _lib.check(_lib.lib().sg_pcg_solve(self.solver, _lib.ptr(T_lin), _lib.ptr(b), _lib.ptr(x), rtol, atol, max_it,
                                           C.byref(it), C.byref(res), _lib.current_stream_ptr()))
        return it.value, res.value

    def prepare_preconditioner(self, T_lin):
        n = self.n_dofs
        dinv = self.workspace_view(5)
        self.jac_diag(T_lin, dinv)
        dinv.reciprocal_()

    def workspace_view(self, k: int):
        """Vector k of the solver workspace (b, dx, r, p, Ap, dinv, zA, zB) as a torch tensor (tests, diagnostics)."""
        import torch
        n = self.n_dofs
        if self.workspace is not None:
            return self.workspace[k * n:(k + 1) * n]
        from ._lib import tensor_from_ptr
        return tensor_from_ptr(self._ws_ptr + 8 * k * n, n, torch.device("cuda", self.ctx.device))

    def timestep(self, T, T_prev) -> NewtonStatsC:
        """Newton solve of F(T) = 0 in place on T (NewtonSolver.solve, ThermoViscoProblem.py:389)."""
        st = NewtonStatsC()
        rc = _lib.lib().sg_thermal_timestep(self.solver, _lib.ptr(T), _lib.ptr(T_prev), C.byref(self.opts), C.byref(st),
                                            _lib.current_stream_ptr())
        self.last_stats = st
        _lib.check(rc)
        return st

    def close(self):
        L = _lib.lib()
        if getattr(self, "solver", None):
            L.sg_thermal_solver_destroy(self.solver)
            self.solver = None
        if getattr(self, "halo", None):
            L.sg_halo_plan_destroy(self.halo)
            self.halo = None
        if getattr(self, "handle", None):
            L.sg_thermal_op_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
