"""ThermoViscoProblem — the reference's orchestrator (/root/reference/ThermoViscoProblem.py, TVP) with the
per-timestep hot path running on a B200 through libsurroglas_b200.

Same constructor arguments, attributes and method names as the reference.  What differs underneath:
  * _solve_T is a matrix-free Newton / Jacobi-PCG solve on the GPU (csrc/thermal.cu, csrc/pcg.cu) instead of
    dolfinx NonlinearProblem + PETSc cg/gamg (TVP:330-346);
  * the 17 Function.interpolate(Expression) passes and 7 copies of TVP:370-373 are ONE fused kernel
    (csrc/visco.cu); `_solve_Tf/_solve_strains/_solve_shifted_time/_solve_stress` still exist and run the
    corresponding phase of that kernel;
  * `previous`/`current` (and `current`/`next`) pairs that the reference keeps equal by copying
    (TVP:469,481,559-562,578-585) share one device buffer.
There is no CPU fallback: constructing the problem without a CUDA device raises.

One difference a caller of the SINGLE-PHASE methods must know: the reference's phases communicate through the stored
Functions (each interpolate reads what the previous one wrote, TVP:456-591), whereas every phase of the fused kernel
recomputes its inputs (xi, the strains, the Taylor factors) from T_cur / T_prev.  Editing functions["xi"] or a strain
Function between two phase calls therefore has no effect here, and with materialize="minimal" those intermediates are
not stored at all.  Code that needs the reference's data flow literally can replay it with
Function.interpolate(material_model.expressions[key]) (models.PointwiseExpression reads the stored arrays).
"""
from __future__ import annotations

import os
from math import ceil
from time import time as _wall

import numpy as np

from . import _lib, fe
from . import mesh as _mesh
from .function import FiniteElementInfo, Function, FunctionSpace
from .models import ThermalModel, ViscoelasticModel
from .thermal_op import ThermalOperator


class _KrylovStub:
    """self.solver.krylov_solver: PETSc option handling of TVP:339-346 is accepted and ignored."""

    def getOptionsPrefix(self):
        return "sg_"

    def setFromOptions(self):
        pass


class NewtonSolverGPU:
    """Stand-in for dolfinx.nls.petsc.NewtonSolver (TVP:334-337): same attributes, GPU solve."""

    def __init__(self, problem: "ThermoViscoProblem"):
        self._p = problem
        self.convergence_criterion = "incremental"
        self.rtol, self.atol, self.max_it = 1e-12, 1e-10, 50
        self.report = False
        self.krylov_solver = _KrylovStub()
        self.linear_rtol = 1e-12   # final linear-residual target of a step, relative to the first Newton residual
        self.linear_atol = 0.0     # absolute floor of that target
        self.forcing_eta = 1e-3    # inexact Newton: first forcing term (0 = every PCG solve runs to the target)
        self.last_stats = None

    def solve(self, u: Function):
        op = self._p._thermal_op
        op.opts.newton_rtol, op.opts.newton_atol, op.opts.newton_max_it = self.rtol, self.atol, self.max_it
        op.opts.lin_rtol = self.linear_rtol
        op.opts.lin_atol = self.linear_atol
        op.opts.forcing_eta = self.forcing_eta
        try:
            st = op.timestep(u.x.array, self._p.functions_previous["T"].x.array)
        except _lib.SgError as e:
            if e.code != _lib.SG_E_NOCONV:
                raise
            st = op.last_stats
        self.last_stats = st
        if self.report and self._p.mesh.comm.rank == 0:
            print(f"Newton: {st.newton_its} iterations, {st.lin_its} PCG iterations, |dx| = {st.dx_norm_last:.3e}")
        return st.newton_its, bool(st.converged)


class ThermoViscoProblem:
    def __init__(self, mesh_path: str, time: tuple, dt: float, config: dict, model_parameters: dict,
                 jit_options: (dict | None) = None, *, problem_dim: int | None = None, mesh: _mesh.Mesh | None = None,
                 materialize: str = "all", device: int | None = None, ctx: _lib.Context | None = None,
                 partition: dict | None = None, verbose: bool = True) -> None:
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError("ThermoViscoProblem needs a CUDA device (B200); there is no CPU fallback")
        self.mesh = mesh if mesh is not None else self._load_mesh(mesh_path, problem_dim)      # TVP:27-28
        self.cell_tags = self.facet_tags = None
        self.dim = self.mesh.topology.dim
        self.dt = dt
        self.time = time
        self.t = self.time[0]
        self.n_steps = ceil((self.time[1] - self.time[0]) / self.dt)                           # TVP:36
        self.jit_options = jit_options                                                         # accepted, unused
        self.verbose = verbose
        assert materialize in ("all", "minimal")
        self._materialize = materialize
        self._partition = partition
        self._ctx = ctx if ctx is not None else _lib.Context(device if device is not None else torch.cuda.current_device())
        self._device = torch.device("cuda", self._ctx.device)
        self._params = dict(model_parameters)

        self.material_model = ViscoelasticModel(mesh=self.mesh, model_parameters=model_parameters)   # TVP:38
        self.physical_model = ThermalModel(mesh=self.mesh, model_parameters=model_parameters)         # TVP:40
        self.__init_function_spaces(config=config)
        self.__init_functions()
        self.material_model.make_plan(self._ctx, self.dt)
        self.material_model._init_expressions(functionSpaces=self.functionSpaces, functions=self.functions,
                                              functions_current=self.functions_current,
                                              functions_previous=self.functions_previous,
                                              functions_next=self.functions_next, dt=self.dt,
                                              to_sigma=self._to_sigma)                            # TVP:48-54
        guess = model_parameters.get("newton_initial_guess", "previous")
        if guess not in ("previous", "extrapolate"):
            raise ValueError('model_parameters["newton_initial_guess"] must be "previous" (TVP:389) or "extrapolate"')
        self._extrapolate, self._T_pp = guess == "extrapolate", None
        self.mechanics = None
        if self._mech_opts:
            from .mechanics import MechanicalEquilibrium
            self.mechanics = MechanicalEquilibrium(self._ctx, self.mesh, self.functionSpaces["sigma"].scalar,
                                                   self.material_model.plan, self._device, self._mech_opts)
        self._thermal_op = None
        self.output_dir = None
        self.host_mirror = None           # output.HostMirror: per-step pinned host copies of T/phi/Tf/xi/sigma
        self.last_mirror_slot = None

    # ------------------------------------------------------------------------------------------ mesh
    @staticmethod
    def _load_mesh(mesh_path: str, problem_dim):
        if mesh_path:
            # TVP:27-28 reads the file or fails; a wrong path must not silently run a different problem
            if not os.path.exists(mesh_path):
                raise FileNotFoundError(f"mesh file {mesh_path!r} does not exist (geometry.create_mesh(path) writes the "
                                        "reference's graded line; mesh_path='' selects the built-in meshes)")
            from .meshio import read_msh
            return read_msh(mesh_path)
        if problem_dim in (None, 1):
            # geometry.py:3-29 — the graded 1-D line the reference meshes with gmsh ('mesh1d.msh' is not shipped)
            return _mesh.graded_line_mesh()
        if problem_dim == 2:
            return _mesh.rectangle_mesh(96, 48)
        if problem_dim == 3:
            return _mesh.box_mesh(32, 32, 8, 32.0, 32.0, 8.0)
        raise ValueError("problem_dim must be 1, 2 or 3")

    # ------------------------------------------------------------------------------------------ spaces
    def __init_function_spaces(self, config: dict) -> None:
        assert all(var["element"] in ['CG', 'DG'] for var in config.values()), \
            "Only CG and DG elements are supported"                                             # TVP:70-71
        N, d, cell = self.material_model.tableau_size, self.dim, self.mesh.ufl_cell()
        cT, cS = config["T"], config["sigma"]
        sT = fe.ScalarSpace(self.mesh, cT["element"], cT["degree"])
        same = (cT["element"], cT["degree"]) == (cS["element"], cS["degree"])
        sS = sT if same else fe.ScalarSpace(self.mesh, cS["element"], cS["degree"])
        self._same_space = same
        self.finiteElements = {
            "T": FiniteElementInfo(cT["element"], cell, cT["degree"], ()),                       # TVP:77-79
            "Tf_partial": FiniteElementInfo(cT["element"], cell, cT["degree"], (N,)),            # TVP:82-85
            "sigma": FiniteElementInfo(cS["element"], cell, cS["degree"], (d, d)),               # TVP:89-92
            "sigma_partial": FiniteElementInfo(cS["element"], cell, cS["degree"], (N, d, d)),    # TVP:97-100
        }
        self.functionSpaces = {
            "T": FunctionSpace(self.mesh, sT, ()),
            "Tf_partial": FunctionSpace(self.mesh, sT, (N,)),
            "sigma": FunctionSpace(self.mesh, sS, (d, d)),
            "sigma_partial": FunctionSpace(self.mesh, sS, (N, d, d)),
        }
        self._mech_opts = self._params.get("mechanics", False)
        if self._mech_opts:
            # extension: vector-P1 displacement on the mesh vertices, see mechanics.py
            sU = sT if (cT["element"], cT["degree"]) == ("CG", 1) else fe.ScalarSpace(self.mesh, "CG", 1)
            self.finiteElements["u"] = FiniteElementInfo("CG", cell, 1, (d,))
            self.functionSpaces["u"] = FunctionSpace(self.mesh, sU, (d,))
        self._gather = None
        if not same:
            import torch
            dofs, lp, w = fe.winner_map(sS, sT)
            dev = self._device
            self._gather = dict(n_ld=sT.n_ld, n_points=sS.n_ld, dofs=torch.from_numpy(dofs.ravel().copy()).to(dev),
                                local_point=torch.from_numpy(lp).to(dev), weights=torch.from_numpy(w.ravel().copy()).to(dev))
            self._gather_np = (dofs, lp, w)

    def _to_sigma(self, a):
        """Evaluate a T-space array at the sigma nodes (identity for equal spaces; last-cell-wins otherwise)."""
        if self._same_space:
            return a
        import torch
        g = self._gather
        idx = g["dofs"].view(-1, g["n_ld"]).long()
        w = g["weights"].view(g["n_points"], g["n_ld"])[g["local_point"].long()]
        vals = a[idx]
        acc = torch.zeros(idx.shape[0], dtype=torch.float64, device=a.device)
        for j in range(g["n_ld"]):
            acc = torch.where(w[:, j] != 0.0, acc + w[:, j] * vals[:, j], acc)
        return acc

    # ------------------------------------------------------------------------------------------ functions
    def __init_functions(self) -> None:
        V, dev = self.functionSpaces, self._device
        full = self._materialize == "all"
        sc = self._scatter
        F = lambda key, name=None, allocate=True, alias=None: Function(V[key], name, device=dev, allocate=allocate,
                                                                       alias=alias, scatter=sc)
        self.functions_previous, self.functions_current, self.functions, self.functions_next = {}, {}, {}, {}
        self.functions_current["T"] = F("T", "Temperature")                                       # TVP:124
        self.functions_previous["T"] = F("T")
        self.functions_next["T"] = F("T", allocate=full)
        self.v = None                                                                             # TestFunction placeholder
        self.functions_previous["Tf_partial"] = F("Tf_partial")
        self.functions_current["Tf_partial"] = F("Tf_partial", "Fictive_temperature",
                                                 alias=self.functions_previous["Tf_partial"])     # TVP:469
        self.functions_previous["Tf"] = F("T")
        self.functions_current["Tf"] = F("T", "Fictive_Temperature", alias=self.functions_previous["Tf"])  # TVP:481
        self.functions["phi"] = F("T")                                                            # TVP:142-143
        self.functions_next["phi"] = F("T", allocate=full)
        self.functions["xi"] = F("T", "Shifted_time")
        for k in ("thermal_strain", "total_strain", "deviatoric_strain"):
            self.functions[k] = F("sigma", k, allocate=full)
        self.functions["ds_partial"] = F("sigma_partial", "Deviatoric_stress_increment", allocate=full)
        self.functions["dsigma_partial"] = F("sigma_partial", "Hydrostatic_stress_increment", allocate=full)
        self.functions_current["s_tilde_partial"] = F("sigma_partial")
        self.functions_next["s_tilde_partial"] = F("sigma_partial", alias=self.functions_current["s_tilde_partial"])
        self.functions_current["sigma_tilde_partial"] = F("sigma_partial")
        self.functions_next["sigma_tilde_partial"] = F("sigma_partial", alias=self.functions_current["sigma_tilde_partial"])
        self.functions_current["s_partial"] = F("sigma_partial", allocate=full)
        self.functions_next["s_partial"] = F("sigma_partial", allocate=full, alias=self.functions_current["s_partial"])
        self.functions_current["sigma_partial"] = F("sigma_partial", allocate=full)
        self.functions_next["sigma_partial"] = F("sigma_partial", allocate=full,
                                                 alias=self.functions_current["sigma_partial"])
        self.functions_next["sigma"] = F("sigma", "Stress_tensor")                                # TVP:171
        if self._mech_opts:
            self.functions["displacement"] = F("u", "Displacement")
            self.functions["displacement_increment"] = F("u", "Displacement_increment")
            self.functions["mechanical_strain"] = F("sigma", "Mechanical_strain")

    def _scatter(self, array, block_size):
        if self._thermal_op is not None and self._same_space:
            self._thermal_op.halo_forward(array, block_size)

    def _visco_tensors(self) -> dict:
        a = lambda F: F._array
        fc, fp, fn, f = self.functions_current, self.functions_previous, self.functions_next, self.functions
        return {"T_cur": a(fc["T"]), "T_prev": a(fp["T"]), "Tf_partial": a(fc["Tf_partial"]), "Tf": a(fc["Tf"]),
                "phi": a(f["phi"]), "xi": a(f["xi"]), "s_tilde": a(fc["s_tilde_partial"]),
                "sigma_tilde": a(fc["sigma_tilde_partial"]), "sigma": a(fn["sigma"]),
                "T_next": a(fn["T"]), "phi_next": a(fn["phi"]), "thermal_strain": a(f["thermal_strain"]),
                "total_strain": a(f["total_strain"]), "deviatoric_strain": a(f["deviatoric_strain"]),
                "ds_partial": a(f["ds_partial"]), "dsigma_partial": a(f["dsigma_partial"]),
                "s_partial": a(fn["s_partial"]), "sigma_partial": a(fn["sigma_partial"])}

    # ------------------------------------------------------------------------------------------ setup
    def setup(self, dirichlet_bc: bool = False, outfile_name: str = "visco", outfile_name1: str = "stresses") -> None:
        self._set_initial_condition(temp_value=float(self.material_model.T_init))               # TVP:179
        if dirichlet_bc:
            self._set_dirichlet_bc(bc_value=None)
        self._write_initial_output(t=self.t)
        self._setup_weak_form()
        self._setup_solver()

    def _set_initial_condition(self, temp_value: float) -> None:
        """TVP:187-233: T_prev = T_cur = T_0, Tf = T, every partial fictive temperature = T.x.array[0]."""
        fc, fp = self.functions_current, self.functions_previous
        # TVP:193-201 interpolates `lambda x: np.full(x.shape[1], temp_value)`: a constant, so the dof coordinates (seconds
        # of host work on a 20 M-node plate) are not tabulated for it
        fp["T"].x.array.fill_(temp_value)
        fc["T"].x.array.fill_(temp_value)
        fp["Tf"].x.array.copy_(fp["T"].x.array)
        fc["Tf"].x.array.copy_(fc["T"].x.array)
        v0 = float(fc["T"].x.array[0])
        fp["Tf_partial"].x.array.fill_(v0)
        fc["Tf_partial"].x.array.fill_(v0)

    def _set_dirichlet_bc(self, bc_value) -> None:
        # The reference's Dirichlet path (TVP:236-243) references attributes that do not exist and never
        # reaches the solver; main.py passes dirichlet_bc=False.
        raise AttributeError("the reference's Dirichlet path is broken (self.fs / T_ambient do not exist); "
                             "only dirichlet_bc=False is supported")

    def _write_initial_output(self, t: float = 0.0) -> None:
        self._writer = None
        if self.output_dir:
            from .output import FieldWriter
            self._writer = FieldWriter(self.output_dir, self)
            self._writer.write(t)

    def _setup_weak_form(self) -> None:
        """TVP:280-327 — the residual is implemented in csrc/thermal.cu; F is kept as a description."""
        dg = self.finiteElements["T"].family() == 'Discontinuous Lagrange'
        self.F = ("(T - T_prev)*v*dx + dt*(alpha*inner(grad(T), grad(v))*dx - f*v*dx"
                  " + 0.001*sigma*epsilon*(T**4 - T_ambient**4)*v*ds + 0.001*htc*(T - T_ambient)*v*ds)"
                  + (" + dt*alpha('+')*((5.0/h('+'))*inner(jump(v,n),jump(T,n)) - inner(avg(grad(v)),jump(T,n))"
                     " - inner(jump(v,n),avg(grad(T))))*dS" if dg else ""))

    def _setup_solver(self) -> None:
        self._thermal_op = ThermalOperator(self._ctx, self.functionSpaces["T"].scalar, self._params, self.dt,
                                           partition=self._partition)
        self.prob = self._thermal_op                                                              # TVP:331
        self.solver = NewtonSolverGPU(self)                                                       # TVP:334-337
        self.ksp = self.solver.krylov_solver                                                      # TVP:339

    def _update_values(self, current: Function, previous: Function) -> None:
        current.x.scatter_forward()                                                               # TVP:351
        if previous._array is not None and current._array is not None and previous._array is not current._array:
            previous.x.array.copy_(current.x.array)                                               # TVP:353

    def _write_output(self) -> None:
        if getattr(self, "_writer", None) is not None:
            self._writer.write(self.t)
        if self.host_mirror is not None:
            self.last_mirror_slot = self.host_mirror.capture()

    # ------------------------------------------------------------------------------------------ time loop
    def solve_timestep(self, t) -> None:
        if self.verbose and self.mesh.comm.rank == 0:
            print(f"t={self.t}")
        self._solve_T()
        self._solve_viscoelastic()          # == _solve_Tf + _solve_strains + _solve_shifted_time + _solve_stress
        if self.mechanics is not None:
            self._solve_mechanics()
        self._write_output()
        self._update_values(current=self.functions_current["T"], previous=self.functions_previous["T"])  # TVP:378

    def _extrapolated_start(self) -> None:
        """model_parameters["newton_initial_guess"] = "extrapolate": start Newton from 2 T_n - T_{n-1} instead of T_n (what the
        reference's NewtonSolver starts from, TVP:389).  The discrete solution is the same to solver tolerance — the final
        linear-solve target stays anchored at |F(T_n)| through lin_atol — but the first correction is 5-20x smaller, which
        usually saves a Newton iteration per step."""
        import torch
        T, Tp, op = self.functions_current["T"]._array, self.functions_previous["T"]._array, self._thermal_op
        if self._T_pp is None:
            self._T_pp = Tp.clone()                       # first step: nothing to extrapolate from yet
            self._F_scratch = torch.empty_like(Tp)
            return
        # |F| at the reference's starting point T_n: keeps the absolute accuracy target of the step unchanged
        own = slice(op.own_lo, op.own_hi)
        F = op.residual(T, Tp, self._F_scratch)
        f2 = (F[own] * F[own]).sum()
        if self._ctx.nranks > 1:
            import torch.distributed as dist
            dist.all_reduce(f2)
        self.solver.linear_atol = self.solver.linear_rtol * float(f2.sqrt())
        guess = 2.0 * T - self._T_pp
        self._T_pp.copy_(T)
        T.copy_(guess)

    def _solve_T(self) -> None:
        if self._extrapolate:
            self._extrapolated_start()
        _, converged = self.solver.solve(self.functions_current["T"])                             # TVP:389
        assert (converged), "Newton solver did not converge: " + _lib.lib().sg_last_error().decode()   # TVP:390

    def _run_phases(self, phases: int) -> None:
        if self.material_model.physics == "corrected":
            if not self._same_space:
                raise NotImplementedError('physics="corrected" needs fe_config["T"] == fe_config["sigma"]')
            if phases != _lib.PHASE_ALL:
                raise NotImplementedError('physics="corrected" runs as one fused update (solve_timestep / _solve_viscoelastic): '
                                          "it reads the old fictive temperature while writing the new one")
        plan, t = self.material_model.plan, self._visco_tensors()
        nT, nS = self.functionSpaces["T"].n_nodes, self.functionSpaces["sigma"].n_nodes
        if self._same_space:
            plan.update(nT, t, phases)
        else:
            if phases & (_lib.PHASE_TF | _lib.PHASE_SHIFT):
                plan.update_scalar(nT, t, phases & (_lib.PHASE_TF | _lib.PHASE_SHIFT))
            if phases & (_lib.PHASE_STRAIN | _lib.PHASE_STRESS):
                plan.update_tensor(nS, t, self._gather, phases & (_lib.PHASE_STRAIN | _lib.PHASE_STRESS))

    def _solve_viscoelastic(self) -> None:
        """The whole of TVP:370-373 in one fused launch."""
        self._run_phases(_lib.PHASE_ALL)

    def _solve_Tf(self) -> None:
        self._run_phases(_lib.PHASE_TF)                                                           # TVP:393-407

    def _solve_strains(self) -> None:
        self._run_phases(_lib.PHASE_STRAIN)                                                       # TVP:409-423

    def _solve_shifted_time(self) -> None:
        self._run_phases(_lib.PHASE_SHIFT)                                                        # TVP:426-435

    def _solve_stress(self) -> None:
        self._run_phases(_lib.PHASE_STRESS)                                                       # TVP:438-452

    def _solve_mechanics(self) -> None:
        """Extension, not in the reference (VM:135-139 assumes total_strain = -thermal_strain): equilibrate the stress the
        phases above left in functions_next["sigma"] (mechanics.py).  Must follow _solve_stress of the same step."""
        t = self._visco_tensors()
        fields = {k: t[k] for k in ("sigma", "total_strain", "deviatoric_strain", "ds_partial", "dsigma_partial",
                                    "s_partial", "sigma_partial")}
        fields["mech_strain"] = self.functions["mechanical_strain"]._array
        if self.material_model.physics == "corrected":
            fields["s_tilde"], fields["sigma_tilde"] = t["s_tilde"], t["sigma_tilde"]
        xi_sigma = self._to_sigma(self.functions["xi"]._array)
        self.mechanics.step(xi_sigma, fields, self.functions["displacement_increment"]._array,
                            self.functions["displacement"]._array)

    def solve(self) -> None:
        import torch
        if self.mesh.comm.rank == 0:
            print("Starting solve")
            t_start = _wall()
        for _ in range(self.n_steps):
            self.t += self.dt                                                                     # TVP:603
            self.solve_timestep(t=self.t)
        torch.cuda.synchronize(self._device)
        if self.mesh.comm.rank == 0:
            t_end = _wall()
            print(f"Solve finished in {t_end - t_start} seconds.")
        self._finalize()

    def _finalize(self) -> None:
        if getattr(self, "_writer", None) is not None:
            self._writer.close()
