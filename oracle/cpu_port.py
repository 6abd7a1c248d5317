"""TEST INFRASTRUCTURE / CPU BASELINE — the reference's time step restated for the host cores.

This is what bench.py times as `cpu_baseline` and as the `--impl reference` arm (kind "port": the reference itself,
`mpiexec -np N python3 main.py` on dolfinx/PETSc, cannot run in this image — no dolfinx, petsc4py, mpi4py, gmsh — see
probe_reference_stack()).  It is also the CHECKER of bench.py's `parity_check`: the GPU runs the same plate and the
fields are compared after the same number of steps.  Never imported by the product.

One time step = ThermoViscoProblem.solve_timestep (TVP:367-381):
  * heat solve: Newton with dolfinx's incremental criterion (TVP:334-337) on the ASSEMBLED operator of
    oracle/thermal_oracle.py; the constant part of the Jacobian (mass + dt*alpha*stiffness + SIP) is assembled once, each
    Newton iteration adds the linearised radiation/convection facets into the same CSR pattern (what PETSc's
    MatAssembly does after the first iteration) and runs Jacobi-PCG on all host threads (oracle/cpu_pcg.c);
  * viscoelastic chain: the C restatement of the 16 expressions (oracle/visco_oracle.c), either as the reference's 17
    passes + 7 copies (vo_step_passes — the dolfinx-shaped variant) or fused into one sweep (vo_step_fused);
  * T_prev <- T_cur.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import time

import numpy as np
import scipy.sparse as sp

from . import thermal_oracle as to
from . import visco_oracle as vo

_HERE = os.path.dirname(os.path.abspath(__file__))


def probe_reference_stack() -> dict:
    """Can the UNMODIFIED reference run here?  Tries the imports /root/reference/ThermoViscoProblem.py:1-16 needs (from
    baseline/_ref first, where a reference install would live) and looks for an MPI launcher."""
    import importlib
    import shutil
    import sys
    ref = os.path.join(os.path.dirname(_HERE), "baseline", "_ref")
    if os.path.isdir(ref) and ref not in sys.path:
        sys.path.insert(0, ref)
    missing = []
    for mod in ("dolfinx", "ufl", "basix", "petsc4py", "mpi4py", "gmsh"):
        try:
            importlib.import_module(mod)
        except Exception:  # noqa: BLE001
            missing.append(mod)
    return {"baseline_ref_dir": os.path.isdir(ref), "missing_modules": missing,
            "mpiexec": shutil.which("mpiexec") or shutil.which("mpirun"), "runnable": not missing}


class _Pcg:
    def __init__(self, threads: int | None):
        subprocess.run(["make", "-C", _HERE], capture_output=True)
        self.lib = C.CDLL(os.path.join(_HERE, "_build", "libcpu_pcg_omp.so"))
        self.lib.cpu_pcg_jacobi.restype = C.c_int
        self.lib.cpu_set_threads.restype = C.c_int
        want = threads or len(os.sched_getaffinity(0))
        self.threads = int(self.lib.cpu_set_threads(C.c_int(want)))   # also governs the OpenMP visco oracle (same libgomp)

    @staticmethod
    def _p(a, t):
        return a.ctypes.data_as(C.POINTER(t))

    def solve(self, indptr, indices, data, dinv, b, rtol):
        x = np.empty_like(b)
        its = self.lib.cpu_pcg_jacobi(C.c_long(b.size), self._p(indptr, C.c_int32), self._p(indices, C.c_int32),
                                      self._p(data, C.c_double), self._p(dinv, C.c_double), self._p(b, C.c_double),
                                      self._p(x, C.c_double), C.c_double(rtol), C.c_int(20000))
        if its < 0:
            raise RuntimeError("CPU PCG did not converge")
        return x, its

    def spmv(self, A, x):
        y = np.empty(A.shape[0])
        self.lib.cpu_spmv(C.c_long(A.shape[0]), self._p(A.indptr, C.c_int32), self._p(A.indices, C.c_int32),
                          self._p(A.data, C.c_double), self._p(np.ascontiguousarray(x), C.c_double), self._p(y, C.c_double))
        return y


class CpuTimestep:
    def __init__(self, mesh, space, params: dict, dt: float, threads: int | None = None, pcg_rtol: float = 1e-10):
        t0 = time.time()
        self.pcg = _Pcg(threads)
        self.threads = self.pcg.threads
        self.dim, self.dt, self.params, self.pcg_rtol = mesh.dim, dt, params, pcg_rtol
        self.orc = to.ThermalOracle(mesh.x, mesh.cells, space.dofmap, space.element.nodes, space.family, space.degree, params, dt)
        o = self.orc
        a, f = float(params["alpha"]), float(params["f"])
        # constant operators in ONE CSR pattern
        A = o.A_lin.tocsr().astype(np.float64)
        A.sort_indices()
        self.indptr, self.indices = A.indptr.astype(np.int32), A.indices.astype(np.int32)
        self.A_data = A.data.copy()
        self.R = (o.M + dt * a * o.K + o.S).tocsr()                  # residual: R T - M T_prev - dt f load + boundary
        self.R.sort_indices()
        self.M = o.M.tocsr()
        self.rhs0 = dt * f * o.load
        # positions of the exterior-facet matrix entries inside A's pattern (found once)
        o._boundary(np.full(space.n_nodes, float(params["T_0"])), False)
        _, _, _, dofs = o._bcache
        nl = dofs.shape[1]
        rows = np.repeat(dofs[:, :, None], nl, axis=2).ravel()
        cols = np.repeat(dofs[:, None, :], nl, axis=1).ravel()
        idx = sp.csr_matrix((np.arange(1, A.nnz + 1, dtype=np.float64), A.indices, A.indptr), shape=A.shape)
        pos = np.asarray(idx[rows, cols]).ravel().astype(np.int64) - 1
        assert (pos >= 0).all(), "exterior-facet entries outside the cell pattern"
        self.bpos = pos
        diag_pos = np.asarray(idx[np.arange(A.shape[0]), np.arange(A.shape[0])]).ravel().astype(np.int64) - 1
        self.diag_pos = diag_pos
        n, d = space.n_nodes, self.dim
        self.vp = vo.ViscoParams(dim=d, dt=dt, H=params["H"], Rg=params["Rg"], Tb=params["Tb"],
                                 alpha_solid=params["alpha_solid"], alpha_liquid=params["alpha_liquid"])
        self.st = vo.new_state(self.vp, n, params["T_0"])
        N = self.vp.N
        self.fused = dict(Tfp=np.full(n * N, float(params["T_0"])), Tf=np.full(n, float(params["T_0"])), phi=np.zeros(n),
                          xi=np.zeros(n), s=np.zeros(n * N * d * d), k=np.zeros(n * N * d * d), sigma=np.zeros(n * d * d))
        try:
            vo._lib(True)
            self.omp = True
        except OSError:
            self.omp = False
        self.setup_s = time.time() - t0
        self.pcg_its = self.newton_its = 0

    def _boundary(self, T):
        """linearised exterior facets: residual vector and the Jacobian entries in self.bpos order (TVP:302-304)."""
        o, p = self.orc, self.params
        se, htc, Ta = float(p["sigma"]) * float(p["epsilon"]), float(p["htc"]), float(p["T_ambient"])
        _, v, wq, dofs = o._bcache
        Tq = np.einsum("nqj,nj->nq", v, T[dofs])
        flux = 0.001 * se * (Tq ** 4 - Ta ** 4) + 0.001 * htc * (Tq - Ta)
        vec = np.bincount(dofs.ravel(), weights=(self.dt * np.einsum("nqi,nq->ni", v, wq * flux)).ravel(), minlength=T.size)
        coef = 0.001 * (4.0 * se * Tq ** 3 + htc)
        mat = self.dt * np.einsum("nq,nqi,nqj->nij", wq * coef, v, v)
        return vec, mat.ravel()

    def newton(self, T0, T_prev, rtol=1e-12, atol=1e-10, max_it=50):
        T, r0 = T0.copy(), None
        MTp = self.pcg.spmv(self.M, T_prev)
        for it in range(1, max_it + 1):
            bvec, bmat = self._boundary(T)
            b = self.pcg.spmv(self.R, T) - MTp - self.rhs0 + bvec
            data = self.A_data + np.bincount(self.bpos, weights=bmat, minlength=self.A_data.size)
            dx, k = self.pcg.solve(self.indptr, self.indices, data, 1.0 / data[self.diag_pos], b, self.pcg_rtol)
            self.pcg_its += k
            self.newton_its += 1
            T = T - dx
            r = float(np.linalg.norm(dx))
            if it == 1:
                r0 = r
                if r0 == 0.0:
                    return T
            elif r / r0 < rtol or r < atol:
                return T
        raise RuntimeError("CPU Newton did not converge")

    def step(self, fused: bool = True):
        st = self.st
        st["T_cur"][:] = self.newton(st["T_cur"], st["T_prev"])
        if fused:
            f = self.fused
            vo.step_fused(self.vp, st["T_cur"], st["T_prev"], f["Tfp"], f["Tf"], f["phi"], f["xi"], f["s"], f["k"], f["sigma"],
                          omp=self.omp)
        else:
            vo.step_passes(self.vp, st, omp=self.omp)

    def end_step(self):
        self.st["T_prev"][:] = self.st["T_cur"]

    def fields(self, fused: bool = True) -> dict:
        """T, Tf, xi, sigma after step() (before end_step()) — the checker side of bench.py's parity_check."""
        if fused:
            f = self.fused
            return dict(T=self.st["T_cur"], T_prev=self.st["T_prev"], Tf=f["Tf"], xi=f["xi"], sigma=f["sigma"])
        st = self.st
        return dict(T=st["T_cur"], T_prev=st["T_prev"], Tf=st["Tf_cur"], xi=st["xi"], sigma=st["sigma_next"])
