#!/usr/bin/env python
"""bench.py — SurroGlas per-timestep hot path on N B200s (one process per GPU).

    python bench.py --gpus 1 --steps K --warmup W
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...        # the CPU restatement of the reference, timed on the host cores

A "step" is ONE FULL TIMESTEP of ThermoViscoProblem.solve_timestep (TVP:367-381): the implicit-Euler heat solve
(inexact Newton + matrix-free Chebyshev-preconditioned CG, hot path B), the fused viscoelastic update at every
quadrature point (hot path A) and T_prev <- T_cur; file output is disabled.  Headline workload (config.workload):
BASELINE configs[2], the 3-D DG1 plate with radiative/convective Robin boundary, 320x320x8 hexahedra x 6 tetrahedra =
4 915 200 cells = 19 660 800 quadrature points PER GPU (weak scaling: the plate grows along x with N; x-slab
partition; ghost-dof halo and the solver's all-reduces over NVLink peer memory).

Printed JSON (rank 0, the only line on stdout): metric = quadrature-point updates/s over all GPUs, plus
  e2e            the same through ThermoViscoProblem with HOST buffers (H2D of the step's temperature, D2H of the five
                 fields the reference writes every step, TVP:357-362, overlapped with the next step);
  roofline       the kernel with the largest share of the step (CUDA events on its launch stream, algorithmic bytes
                 from the library), plus roofline_cheb_step / roofline_apply / roofline_visco;
  parity_check   the GPU and the CPU port run the SAME 48x48x8 cut of the plate with the SAME settings as the timed run;
                 T, Tf and the stress are compared after the last step and the run FAILS (exit 1) if they miss;
  cpu_baseline   the CPU port of the time step on the host cores (N = 1 only);
  other_configs  (N = 1) the other BASELINE configs: C1 main.py 1-D run, C2 2-D CG2, one GPU's share of C4, the Prony
                 sweep C5, the plate with the reference's own SIP penalty, a perturbed (general-mesh) plate;
  c4             BASELINE configs[3], the FULL 768x768x6 CG2 plate (212 M points) split over the N GPUs: strong scaling.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MAIN_PARAMS = {  # main.py:29-55
    "f": 0.0, "epsilon": 0.93, "sigma": 5.670e-8, "T_ambient": 600.0, "T_0": 800.0, "alpha": 1.0, "htc": 280.1,
    "rho": 2500.0, "cp": 1433.0, "k": 1.0, "H": 627.8e3, "Tb": 869.0e0, "Rg": 8.314,
    "alpha_solid": 9.10e-6, "alpha_liquid": 25.10e-6, "Tf_init": 873.0,
}

# The reference's SIP penalty 5.0/CellDiameter (TVP:313) is not coercive on tetrahedra: a 3-D DG run with it grows by
# 1.6x per step and diverges after ~15 steps (tests/test_fe_tables.py::test_stability_of_the_dg_time_stepping).  The 3-D
# DG plates therefore run with model_params["sip_penalty"] = 6.0, the smallest tested coercive value (same kernels,
# condition number within 12 % of the reference's); 1-D/2-D and CG workloads use the reference's parameters unchanged.
PARAM_OVERRIDES = {"C3_plate3d_DG1_robin_19.7M_qp": {"sip_penalty": 6.0}, "small_plate3d_DG1": {"sip_penalty": 6.0},
                   "perturbed_plate3d_DG1": {"sip_penalty": 8.0}}

REFERENCE_PENALTY = False    # --reference-penalty: run the 3-D DG plates with the reference's 5.0 (diverges after ~15 steps)

DG1 = {"T": {"element": "DG", "degree": 1}, "sigma": {"element": "DG", "degree": 1}}
CG2 = {"T": {"element": "CG", "degree": 2}, "sigma": {"element": "CG", "degree": 2}}
MAIN_CFG = {"T": {"element": "DG", "degree": 1}, "sigma": {"element": "CG", "degree": 1}}       # main.py:24-27

WORKLOADS = {
    # name: (dim, cells per axis PER GPU, cell edge [mm], fe_config)
    "C3_plate3d_DG1_robin_19.7M_qp": (3, (320, 320, 8), 1.0, DG1),
    "C2_plate2d_CG2_1M_qp": (2, (408, 204), 50.0 / 408, CG2),
    "C4_plate3d_CG2": (3, (96, 768, 6), 1.0, CG2),                   # one GPU's share of the 768x768x6 plate
    "C4_full_plate3d_CG2_212M_qp": (3, (768, 768, 6), 1.0, CG2),     # strong scaling: the whole plate over N GPUs
    "small_plate3d_DG1": (3, (48, 48, 8), 1.0, DG1),
    "perturbed_plate3d_DG1": (3, (160, 160, 8), 1.0, DG1),           # interior vertices jittered: no repeating cell shapes
    "C1_main_py_1d": (1, (48,), 1.0, MAIN_CFG),
}
STRONG = {"C4_full_plate3d_CG2_212M_qp"}
DEFAULT_WORKLOAD = "C3_plate3d_DG1_robin_19.7M_qp"
PARITY_WORKLOAD = "small_plate3d_DG1"
DT = 0.1
METRIC, UNIT = "quadrature-point updates/s (full timestep: heat solve + viscoelastic update)", "QP-updates/s"
PARITY_TOL = 1e-10


NEWTON_GUESS = None          # --newton-guess: model_parameters["newton_initial_guess"] of every GPU problem of this run


def params_of(workload: str) -> dict:
    p = dict(MAIN_PARAMS, **({} if REFERENCE_PENALTY else PARAM_OVERRIDES.get(workload, {})))
    if NEWTON_GUESS:
        p["newton_initial_guess"] = NEWTON_GUESS
    return p


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.samples, self._stop = index, [], threading.Event()
        self._t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                self.samples.append([s.strip() for s in out.strip().split(",")])
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        sm, mx, reasons = [], None, set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for s in self.samples:
            try:
                sm.append(float(s[0]))
                mx = float(s[1])
                for n, v in zip(names, s[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------ CPU port (reference arm)
def cpu_sample_dims(workload_name: str):
    """The bounded sample of a workload the CPU port runs: the same plate cross-section cut to 48 columns in the long
    axes (3-D: 48 x 48 x nz hexahedra)."""
    dim, n, a, cfg = WORKLOADS[workload_name]
    if dim == 3:
        n = (min(n[0], 48), min(n[1], 48), n[2])
    return dim, n, a, cfg


def make_cpu_port(workload_name: str, threads=None):
    from fem_glass_tempering_b200 import fe
    from fem_glass_tempering_b200 import mesh as msh
    from oracle.cpu_port import CpuTimestep
    dim, n, a, cfg = cpu_sample_dims(workload_name)
    mesh = msh.graded_line_mesh() if dim == 1 else msh.plate_mesh(dim, n, tuple(k * a for k in n))
    space = fe.ScalarSpace(mesh, cfg["T"]["element"], cfg["T"]["degree"])
    assert cfg["T"] == cfg["sigma"] or dim == 1, "the CPU port times equal T/sigma spaces"
    port = CpuTimestep(mesh, space, params_of(workload_name), DT, threads=threads)
    return port, mesh.n_cells * space.n_ld, n


def cpu_timestep_rate(workload_name: str, steps: int, warmup: int, fused: bool = True, port=None):
    """Times the CPU port (oracle/cpu_port.py) on the bounded sample; returns (QP/s, s/step, threads, description, port)."""
    if port is None:
        port, qp, n = make_cpu_port(workload_name)
    else:
        port, qp, n = port
    for _ in range(warmup):
        port.step(fused)
        port.end_step()
    its0 = port.pcg_its
    t0 = time.time()
    for _ in range(steps):
        port.step(fused)
        port.end_step()
    dt_s = (time.time() - t0) / steps
    dim = WORKLOADS[workload_name][0]
    sample = (f"{'x'.join(map(str, n))} cells x {(1, 2, 6)[dim - 1]} simplices = {qp} points, {steps} steps after {warmup} warm-up; "
              f"heat solve: Jacobian assembled once + exterior facets added per Newton iteration, Jacobi-PCG in OpenMP C on "
              f"{port.threads} threads ({(port.pcg_its - its0) / steps:.0f} iterations/step); viscoelastic chain: "
              f"{'fused C sweep' if fused else '17-pass C port'}, OpenMP on {port.threads} threads; set-up {port.setup_s:.1f}s not timed")
    return qp / dt_s, dt_s, port.threads, sample, (port, qp, n)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle.cpu_port import probe_reference_stack
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    value, dt_s, cores, sample, port = cpu_timestep_rate(args.workload, steps, warmup, fused=True)
    v17, dt17, _, _, _ = cpu_timestep_rate(args.workload, max(1, min(steps, 3)), 0, fused=False, port=port)
    dim, _, a, cfg = WORKLOADS[args.workload]
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": warmup, "ms_per_step": dt_s * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, "fe_config": cfg, "dt": DT, "model_param_overrides": PARAM_OVERRIDES.get(args.workload, {}),
                       "note": "CPU port of the reference path on a bounded sample of the workload; throughput is per "
                               "point so it is comparable with the GPU arm.  The unmodified reference cannot run here: "
                               "see reference_stack",
                       "reference_stack": probe_reference_stack()},
            "timesteps_per_s": 1.0 / dt_s,
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                             "dolfinx_shaped_17_pass_value": v17,
                             "note": "value = fused viscoelastic sweep; dolfinx_shaped_17_pass_value = the reference's 17 "
                                     "interpolation passes + 7 copies (TVP:393-595) around the same heat solve"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    OUT.emit(json.dumps(line))


# ------------------------------------------------------------------------------------------ GPU arm
class Runner:
    """One workload on this rank's GPU: problem set-up, device-resident timing, kernel roofline figures."""

    def __init__(self, workload: str, rank: int, world: int, local: int, ctx, cheb=None, eta=None, params=None, steps_1d=None):
        import torch
        from fem_glass_tempering_b200 import ThermoViscoProblem, distributed
        from fem_glass_tempering_b200 import mesh as msh
        self.torch, self.workload, self.rank, self.world, self.local = torch, workload, rank, world, local
        self.dev = torch.device("cuda", local)
        dim, n_in, a, cfg = WORKLOADS[workload]
        self.dim, self.cfg = dim, cfg
        strong = workload in STRONG
        n = tuple(n_in) if (strong or dim == 1) else (n_in[0] * world,) + tuple(n_in[1:])
        self.n, self.lengths = n, tuple(k * a for k in n)
        el = cfg["T"]
        n_ld = {1: dim + 1, 2: (dim + 1) * (dim + 2) // 2}[el["degree"]]
        t0 = time.time()
        if dim == 1:
            mesh, part = msh.graded_line_mesh(), None
            self.qp_local = mesh.n_cells * 2                          # sigma = CG1: 2 interpolation points per cell
        elif world == 1:
            mesh, part = msh.plate_mesh(dim, n, self.lengths), None
            if workload.startswith("perturbed"):
                mesh = msh.perturb_interior(mesh, 0.1 * a, seed=7)
            self.qp_local = mesh.n_cells * n_ld
        else:
            mesh, part, info = distributed.slab_partition(dim, n, self.lengths, el["element"], el["degree"], rank, world)
            self.qp_local = info["owned_cell_points"]
        self.t_mesh = time.time() - t0
        self.params = params if params is not None else params_of(workload)
        self.prob = ThermoViscoProblem(mesh_path="", time=(0.0, 50.0), dt=DT, config=cfg, model_parameters=self.params,
                                       mesh=mesh, ctx=ctx, partition=part, materialize="minimal", verbose=False)
        self.prob.setup(dirichlet_bc=False)
        torch.cuda.synchronize(self.dev)
        self.setup_s = time.time() - t0
        self.op = self.prob._thermal_op
        if cheb is not None:
            self.op.set_chebyshev(cheb)
        if eta is not None:
            self.prob.solver.forcing_eta = eta
        self.n_cells_local = mesh.n_cells

    def barrier(self):
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
        self.torch.cuda.synchronize(self.dev)

    def one_step(self):
        p = self.prob
        p.t += p.dt
        p.solve_timestep(t=p.t)

    def timed(self, steps: int, warmup: int, clocks: bool = False):
        """Device-resident timing of `steps` time steps; returns a dict of raw measurements."""
        import ctypes as C
        import torch.distributed as dist
        from fem_glass_tempering_b200 import _lib
        torch, prob, op, dev, world = self.torch, self.prob, self.op, self.dev, self.world
        L = _lib.lib()
        for _ in range(warmup):
            self.one_step()
        self.barrier()
        # Kernel timing with CUDA-event pairs runs inside the timed region, except when the solver replays its PCG batches
        # as CUDA graphs — event pairs cannot sit between graph nodes, so those workloads get a separate profiled pass.
        in_region = not op.uses_graphs()
        if in_region:
            _lib.check(L.sg_thermal_profile(op.handle, 1, 8192))
        launches0 = L.sg_launch_count()
        lin_its = newton_its = 0
        visco_ev = []
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sampler = ClockSampler(self.local) if clocks else None
        if sampler:
            sampler.__enter__()
        self.barrier()
        ev0.record()
        for _ in range(steps):
            prob.t += prob.dt
            prob._solve_T()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record()
            prob._solve_viscoelastic()
            a1.record()
            visco_ev.append((a0, a1))
            prob._update_values(current=prob.functions_current["T"], previous=prob.functions_previous["T"])
            lin_its += prob.solver.last_stats.lin_its
            newton_its += prob.solver.last_stats.newton_its
        ev1.record()
        self.barrier()
        if sampler:
            sampler.__exit__()
        ms_total = ev0.elapsed_time(ev1)
        launches = L.sg_launch_count() - launches0
        if not in_region:
            _lib.check(L.sg_thermal_profile(op.handle, 1, 8192))
            for _ in range(steps):
                self.one_step()
            self.barrier()
        n_apply, ms_apply = C.c_int64(0), C.c_double(0.0)
        n_cheb, ms_cheb = C.c_int64(0), C.c_double(0.0)
        _lib.check(L.sg_thermal_profile_read_kind(op.handle, 0, C.byref(n_apply), C.byref(ms_apply)))
        _lib.check(L.sg_thermal_profile_read_kind(op.handle, 1, C.byref(n_cheb), C.byref(ms_cheb)))
        _lib.check(L.sg_thermal_profile(op.handle, 0, 0))
        ms_visco = sum(a.elapsed_time(b) for a, b in visco_ev) / len(visco_ev)
        t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
        q = torch.tensor([float(self.qp_local)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dist.all_reduce(q, op=dist.ReduceOp.SUM)
        return dict(ms_total=float(t.item()), qp_total=float(q.item()), steps=steps, launches=int(launches), lin_its=lin_its,
                    newton_its=newton_its, n_apply=n_apply.value, ms_apply=ms_apply.value, n_cheb=n_cheb.value,
                    ms_cheb=ms_cheb.value, ms_visco=ms_visco, kernel_timing_in_region=in_region,
                    clocks=sampler.summary() if sampler else None)

    def rooflines(self, m: dict) -> dict:
        """roofline objects of the three hand-written hot kernels from the measurements of timed()."""
        from fem_glass_tempering_b200 import _lib
        L = _lib.lib()
        op, prob, dim, cfg = self.op, self.prob, self.dim, self.cfg
        peak, peak_how = measured_peak()
        cls, stc, fam = op.class_info(), op.stencil_info(), cfg["T"]["element"]
        nT, nS = prob.functionSpaces["T"].n_nodes, prob.functionSpaces["sigma"].n_nodes
        if stc["active"]:
            kname = (f"thermal k_stencil_apply (CG Jacobian apply in gather form: 16-bit row class + the class's (offset, coefficient) "
                     f"list in shared memory, plain stores, fused x.Ax reduction; {stc['classes']} row classes, {stc['entries']} entries)")
        elif cls["active"]:
            kname = (f"thermal {'dg' if fam == 'DG' else 'cg'}_class_apply (matrix-free Jacobian apply from local-matrix class tables "
                     f"in shared memory, fused x.Ax reduction; {cls['self']} cell + {cls['facet']} facet classes)")
        else:
            kname = "thermal cell_kernel<APPLY> (matrix-free Jacobian apply from per-cell geometry)"
        out = {}
        ms_total = m["ms_total"]
        if m["n_apply"]:
            apply_bytes = op.apply_bytes()
            apply_ms = m["ms_apply"] / m["n_apply"]
            gbs = apply_bytes / (apply_ms * 1e-3) / 1e9
            out["roofline_apply"] = {"kernel": kname, "bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak,
                                     "traffic": None, "peak_source": peak_how, "algorithmic_bytes_per_launch": int(apply_bytes),
                                     "launches_timed": int(m["n_apply"]), "avg_launch_ms": apply_ms,
                                     "share_of_step": m["ms_apply"] / ms_total}
            if fam == "CG":
                # SURVEY 8(d)'s layout-independent figure for a CG apply: read x, write y, per cell the symmetric geometry
                # tensor + |detJ| and the dofmap — reported next to the bytes of the layout in use
                n_ld_T = (dim + 1) if cfg["T"]["degree"] == 1 else (dim + 1) * (dim + 2) // 2
                sb = 16 * nT + (op.cell_hi - op.cell_lo) * (8 * (dim * (dim + 1) // 2 + 1) + 4 * n_ld_T)
                out["roofline_apply"].update(survey_8d_bytes_per_launch=int(sb), frac_on_survey_8d_bytes=sb / (apply_ms * 1e-3) / 1e9 / peak)
        if nS and m["ms_visco"] > 0:
            vb = prob.material_model.plan.bytes_per_node(prob._visco_tensors()) * nS
            vg = vb / (m["ms_visco"] * 1e-3) / 1e9
            out["roofline_visco"] = {"kernel": "visco_fast_kernel (fused viscoelastic update)", "bound": "hbm", "achieved": vg,
                                     "peak": peak, "unit": "GB/s", "frac": vg / peak, "traffic": None, "peak_source": peak_how,
                                     "algorithmic_bytes_per_launch": int(vb), "avg_launch_ms": m["ms_visco"],
                                     "share_of_step": m["ms_visco"] * m["steps"] / ms_total, "frac_of_8TBs_spec": vg / 8000.0}
        deg = op.chebyshev_info()["degree"]
        if m["n_cheb"]:
            # per outer iteration: one first step (reads z, r) and degree-1 later steps (also read z_prev)
            b1, b0 = L.sg_thermal_cheb_step_bytes(op.handle, 1), L.sg_thermal_cheb_step_bytes(op.handle, 0)
            cb = (b1 + (deg - 1) * b0) / deg
            cms = m["ms_cheb"] / m["n_cheb"]
            cg = cb / (cms * 1e-3) / 1e9
            out["roofline_cheb_step"] = {
                "kernel": f"thermal dg_cheb_step (operator apply from the class tables fused with one step of the degree-{deg} "
                          "Chebyshev recurrence of the polynomial preconditioner; J z never goes to memory)",
                "bound": "hbm", "achieved": cg, "peak": peak, "unit": "GB/s", "frac": cg / peak, "traffic": None,
                "peak_source": peak_how, "algorithmic_bytes_per_launch": int(cb), "launches_timed": int(m["n_cheb"]),
                "avg_launch_ms": cms, "share_of_step": m["ms_cheb"] / ms_total}
        try:      # DRAM bytes per launch from the committed ncu --set full captures of the same kernel and workload
            with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as fh:
                traffic = json.load(fh).get(self.workload, {})
        except Exception:
            traffic = {}
        for key, kerns in (("roofline_cheb_step", ("dg_cheb_step",)), ("roofline_apply", ("dg_class_apply", "k_stencil_apply")),
                           ("roofline_visco", ("visco_fast_kernel",))):
            if out.get(key):
                out[key]["traffic"] = next((traffic.get(k) for k in kerns if k in out[key]["kernel"]), None)
                if out[key]["traffic"] is not None:
                    out[key]["traffic_source"] = ("profiles/ncu_traffic.json: dram__bytes_read.sum + dram__bytes_write.sum of one ncu "
                                                  "--set full capture of this kernel on this workload (not measured in this run)")
        return out

    def brief(self, m: dict) -> dict:
        """what other_configs / c4 carry for one workload."""
        try:
            rf = self.rooflines(m)
        except Exception:  # noqa: BLE001  (cross-space 1-D run: no byte model for the split update)
            rf = {}
        cands = [rf[k] for k in rf]
        dom = max(cands, key=lambda r: r["share_of_step"]) if cands else None
        ci = self.op.chebyshev_info()
        return {"workload": self.workload, "cells": list(self.n), "qp_total": int(m["qp_total"]), "ms_per_step": m["ms_total"] / m["steps"],
                "value": m["qp_total"] * m["steps"] / (m["ms_total"] * 1e-3), "unit": UNIT, "steps": m["steps"],
                "pcg_its_per_step": m["lin_its"] / m["steps"], "newton_its_per_step": m["newton_its"] / m["steps"],
                "solver": self.op.solver_description(), "setup_s": round(self.setup_s, 2), "gpu_launches": m["launches"],
                "model_param_overrides": {k: v for k, v in self.params.items() if MAIN_PARAMS.get(k) != v},
                "dominant_kernel": None if dom is None else {"kernel": dom["kernel"].split(" (")[0], "frac": dom["frac"],
                                                              "achieved_GBs": dom["achieved"], "share_of_step": dom["share_of_step"],
                                                              "avg_launch_ms": dom["avg_launch_ms"]},
                "kernel_fracs": {k.replace("roofline_", ""): round(v["frac"], 4) for k, v in rf.items()},
                "chebyshev": ci if ci["degree"] else None}

    def close(self):
        import gc
        self.op.close()
        self.prob = self.op = None
        gc.collect()
        self.torch.cuda.empty_cache()


def e2e_measure(r: Runner, steps: int):
    """End to end through the public API with host buffers: every step H2D of the step's temperature input from pinned host
    memory, solve_timestep, D2H of the five fields the reference writes every step (TVP:357-362) through
    ThermoViscoProblem.host_mirror (device snapshot + side-stream copies into pinned buffers, overlapping the next step).  The
    next step's input is the HOST copy of this step's T; the region ends when the last step's five fields are in host memory."""
    import torch.distributed as dist
    from fem_glass_tempering_b200.output import HostMirror
    torch, prob, dev = r.torch, r.prob, r.dev
    nT = prob.functionSpaces["T"].n_nodes
    prob.host_mirror = HostMirror(prob)
    host_in = torch.empty(nT, dtype=torch.float64, pin_memory=True)
    host_in.copy_(prob.functions_previous["T"].x.array)
    h2d, d2h = host_in.numel() * 8, prob.host_mirror.bytes_per_capture
    r.one_step()                                                                         # warm the mirror's buffers
    src = prob.host_mirror.field(prob.last_mirror_slot, "T")
    # what the host side can take: every rank copies 1 GiB device -> pinned host at the same time (no compute), so the
    # e2e figure can be read against the raw concurrent D2H rate of this box
    probe_dev = torch.empty(1 << 27, dtype=torch.float64, device=dev)
    probe_host = torch.empty(1 << 27, dtype=torch.float64, pin_memory=True)
    probe_host.copy_(probe_dev, non_blocking=True)
    r.barrier()
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    for _ in range(2):
        probe_host.copy_(probe_dev, non_blocking=True)
    p1.record()
    r.barrier()
    tp = torch.tensor([p0.elapsed_time(p1)], dtype=torch.float64, device=dev)
    if r.world > 1:
        dist.all_reduce(tp, op=dist.ReduceOp.MAX)
    raw_d2h = r.world * 2 * (1 << 30) / (float(tp.item()) * 1e-3) / 1e9
    del probe_dev, probe_host
    r.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        prob.functions_previous["T"].x.array.copy_(src, non_blocking=True)              # H2D: the step's input
        r.one_step()                                                                     # ... _write_output -> capture()
        src = prob.host_mirror.field(prob.last_mirror_slot, "T")                         # host consumes T (next input)
    prob.host_mirror.wait(prob.last_mirror_slot)                                         # all five fields on the host
    e1.record()
    r.barrier()
    prob.host_mirror = None
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if r.world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item()), h2d, d2h, raw_d2h


def parity_check(ctx, local: int, args, cheb_degree: int, eta, steps: int):
    """GPU vs CPU port on the SAME plate (the CPU leg's 48x48x8 sample) with the headline's solver settings; returns the
    parity object and the CPU port (reused by the cpu_baseline timing)."""
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import stress_rounding_floor
    wl = PARITY_WORKLOAD
    port = make_cpu_port(wl)
    r = Runner(wl, 0, 1, local, ctx, cheb=cheb_degree, eta=eta)
    p, prob = port[0], r.prob
    worst = {"T": 0.0, "Tf": 0.0, "xi": 0.0, "sigma": 0.0, "sigma_over_floor": 0.0}
    for _ in range(steps):
        prob.t += prob.dt
        prob._solve_T()
        prob._solve_viscoelastic()
        p.step(fused=True)
        f = p.fields(True)
        g = {"T": prob.functions_current["T"], "Tf": prob.functions_current["Tf"], "xi": prob.functions["xi"],
             "sigma": prob.functions_next["sigma"]}
        g = {k: v.x.array.cpu().numpy() for k, v in g.items()}
        rel = lambda a, b: float(np.max(np.abs(a - b)) / np.max(np.abs(b)))
        worst["T"] = max(worst["T"], rel(g["T"], f["T"]))
        worst["Tf"] = max(worst["Tf"], rel(g["Tf"], f["Tf"]))
        worst["xi"] = max(worst["xi"], rel(g["xi"], f["xi"]))
        d = r.dim
        dT = np.abs(f["T"] - f["T_prev"])
        good = dT > 1e-6
        sg, so = g["sigma"].reshape(-1, d * d)[good], f["sigma"].reshape(-1, d * d)[good]
        scale = float(np.max(np.abs(so)))
        err = np.max(np.abs(sg - so), axis=1)
        floor = stress_rounding_floor(p.vp, dT[good], np.abs(f["xi"])[good])
        worst["sigma"] = max(worst["sigma"], float(np.max(err) / scale))
        worst["sigma_over_floor"] = max(worst["sigma_over_floor"], float(np.max(err - 2.0 * floor) / scale))
        prob._update_values(current=prob.functions_current["T"], previous=prob.functions_previous["T"])
        p.end_step()
    ci = r.op.chebyshev_info()
    ok = worst["T"] <= PARITY_TOL and worst["Tf"] <= PARITY_TOL and worst["sigma_over_floor"] <= PARITY_TOL
    out = {"workload": f"{wl}: {'x'.join(map(str, r.n))} hexahedra x 6 = {int(r.qp_local)} points, {steps} steps, same model_params / "
                       f"Chebyshev degree {ci['degree']} / forcing_eta / tolerances as the timed run",
           "checker": "oracle/cpu_port.py (assembled Jacobian, Newton to |dx| < 1e-10 with PCG rtol 1e-10; fused C chain)",
           "T": worst["T"], "Tf": worst["Tf"], "xi": worst["xi"], "sigma": worst["sigma"],
           "sigma_excess_over_rounding_floor": max(worst["sigma_over_floor"], 0.0), "tol": PARITY_TOL,
           "rule": "T, Tf: max|d|/max|.| <= tol; sigma per node: |d sigma_i| <= tol*max|sigma| + 2*floor_i, floor_i = rounding "
                   "noise of the reference's own lambda*(1 - taylor)/xi (tests/helpers.stress_rounding_floor); sigma = the "
                   "norm-wise max|d sigma|/max|sigma| for information",
           "ok": bool(ok)}
    r.close()
    return out, port


def run_gpu(args):
    import torch
    import torch.distributed as dist
    from fem_glass_tempering_b200 import distributed

    rank, world, local = distributed.init_process_group()
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    torch.cuda.set_device(local)
    numa = distributed.bind_to_gpu_numa_node(local)
    ctx = distributed.make_context(rank, world, local)

    r = Runner(args.workload, rank, world, local, ctx, cheb=args.cheb, eta=args.eta)
    m = r.timed(args.steps, args.warmup, clocks=True)
    value = m["qp_total"] * m["steps"] / (m["ms_total"] * 1e-3)
    e2e_steps = max(1, min(args.steps, 5))
    ms_e2e, h2d, d2h, raw_d2h = e2e_measure(r, e2e_steps)
    e2e_value = m["qp_total"] * e2e_steps / (ms_e2e * 1e-3)
    op, prob, dim, cfg = r.op, r.prob, r.dim, r.cfg
    rf = r.rooflines(m) if rank == 0 else {}
    head = None
    if rank == 0:
        cls, stc = op.class_info(), op.stencil_info()
        ci = op.chebyshev_info()
        head = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": m["ms_total"] / m["steps"], "higher_is_better": True,
            "scaling": "strong" if args.workload in STRONG else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, "cells_per_gpu": int(r.n_cells_local), "qp_per_gpu": int(r.qp_local),
                       "qp_total": int(m["qp_total"]), "fe_config": cfg, "dt": DT, "prony_terms": 6,
                       "model_params": "main.py:29-55" + ("".join(f", {k} = {v} (reference: 5.0 at TVP:313 is not coercive on tetrahedra; its "
                                                                  "run diverges after ~15 steps)" for k, v in
                                                                  ({} if REFERENCE_PENALTY else PARAM_OVERRIDES.get(args.workload, {})).items())),
                       "plate_mm": list(r.lengths), "partition": f"x-slabs over {world} GPU(s)",
                       "transport": ("single GPU" if world == 1 else
                                     ("NVLink peer memory (IPC-mapped workspaces + flags; no NCCL on the data path)" if op.peer_memory
                                      else "NCCL send/recv + all-reduce")),
                       "cache": "state per GPU (>17 GB) is far larger than the 126 MB L2; no L2 flush needed",
                       "newton_its_per_step": m["newton_its"] / m["steps"], "pcg_its_per_step": m["lin_its"] / m["steps"],
                       "solver_settings": {"newton_rtol": prob.solver.rtol, "newton_atol": prob.solver.atol,
                                           "linear_rtol": prob.solver.linear_rtol, "forcing_eta": prob.solver.forcing_eta},
                       "setup_s": round(r.setup_s, 1), "local_matrix_classes": cls, "row_stencil_classes": stc,
                       "kernel_timing": ("CUDA events inside the timed region" if m["kernel_timing_in_region"] else
                                         "separate profiled pass after the timed region (the timed region replays CUDA graphs)"),
                       "preconditioner": op.solver_description(), "host_numa_binding": numa},
            "timesteps_per_s": m["steps"] / (m["ms_total"] * 1e-3),
            "gpu_launches": m["launches"],
            "clocks": m["clocks"],
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "steps": e2e_steps, "ms_per_step": ms_e2e / e2e_steps,
                    "aggregate_d2h_GBs": world * d2h * e2e_steps / (ms_e2e * 1e-3) / 1e9,
                    "raw_concurrent_d2h_GBs": raw_d2h,
                    "host_limit_note": "raw_concurrent_d2h_GBs = all ranks copying 1 GiB device->pinned host at once with no "
                                       "compute: the ceiling of this box's host side (PCIe root complexes / memory / IOMMU of "
                                       "the VM) for the 2 GB of per-step output each GPU produces",
                    "what": "ThermoViscoProblem.solve_timestep with host_mirror: pinned-host T_prev in, T/phi/Tf/xi/sigma out to "
                            "pinned host buffers every step (device snapshot, D2H overlapped with the next step)"},
            "roofline": None,
        }
        head.update(rf)
        cands = [rf[k] for k in ("roofline_cheb_step", "roofline_apply", "roofline_visco") if rf.get(k)]
        head["roofline"] = max(cands, key=lambda x: x["share_of_step"]) if cands else None
    cheb_deg, eta = op.chebyshev_info()["degree"], prob.solver.forcing_eta
    r.close()

    # ---- N > 1 (and N = 1): BASELINE configs[3], the full 768x768x6 CG2 plate over the N GPUs (strong scaling) ----
    if not args.no_other_configs and args.workload == DEFAULT_WORKLOAD:
        try:
            rc4 = Runner("C4_full_plate3d_CG2_212M_qp", rank, world, local, ctx)
            m4 = rc4.timed(max(2, min(args.steps, 5)), 3)
            if rank == 0:
                head["c4"] = rc4.brief(m4)
                head["c4"]["scaling"] = "strong (the whole plate on N GPUs)"
            rc4.close()
        except Exception as e:  # noqa: BLE001
            if rank == 0:
                head["c4"] = {"error": repr(e)[:300]}
            if world > 1:
                raise

    pcheck = None
    if world > 1 and not args.no_parity and args.workload == DEFAULT_WORKLOAD:
        try:
            pcheck = partition_check(ctx, rank, world, local, cheb_deg, eta)
        except Exception as e:  # noqa: BLE001
            pcheck = {"error": repr(e)[:300], "ok": False}
    if rank != 0:
        if world > 1:
            dist.barrier()
        return

    failed = False
    if pcheck is not None:
        head["partition_check"] = pcheck
        failed = not pcheck.get("ok", False)
    if world == 1:
        # ---- parity of the timed path, then the CPU baseline on the same sample ----
        port = None
        if not args.no_parity and args.workload == DEFAULT_WORKLOAD:
            head["parity_check"], port = parity_check(ctx, local, args, cheb_deg, eta, steps=3)
            failed = not head["parity_check"]["ok"]
        if not args.no_cpu_baseline:
            v, dt_s, cores, sample, port = cpu_timestep_rate(PARITY_WORKLOAD if dim == 3 and cfg == DG1 else args.workload, 4, 1, True, port)
            v17, _, _, _, _ = cpu_timestep_rate(PARITY_WORKLOAD if dim == 3 and cfg == DG1 else args.workload, 2, 0, False, port)
            head["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                                    "dolfinx_shaped_17_pass_value": v17}
        if not args.no_other_configs and args.workload == DEFAULT_WORKLOAD:
            head["other_configs"] = other_configs(ctx, local, args)
    OUT.emit(json.dumps(head))
    if world > 1:
        dist.barrier()
    if failed:
        sys.stderr.write("bench.py: parity check FAILED: " + json.dumps(head.get("parity_check") or head.get("partition_check")) + "\n")
        sys.exit(1)


def partition_check(ctx, rank: int, world: int, local: int, cheb_degree: int, eta, steps: int = 3):
    """N > 1: the parity plate (48x48x8 hexahedra) split into N x-slabs — direct halo puts, in-kernel waits and all-reduces —
    against the CPU port of the WHOLE plate on rank 0, with the rule of parity_check.  Collective: every rank calls it;
    returns the object on rank 0, None elsewhere."""
    import numpy as np
    import torch.distributed as dist
    from fem_glass_tempering_b200 import ThermoViscoProblem, distributed
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import stress_rounding_floor
    wl = PARITY_WORKLOAD
    dim, n, a, cfg = WORKLOADS[wl]
    lengths = tuple(k * a for k in n)
    mesh, part, info = distributed.slab_partition(dim, n, lengths, "DG", 1, rank, world)
    prob = ThermoViscoProblem(mesh_path="", time=(0.0, 50.0), dt=DT, config=cfg, model_parameters=params_of(wl), mesh=mesh, ctx=ctx,
                              partition=part, materialize="minimal", verbose=False)
    prob.setup(dirichlet_bc=False)
    prob._thermal_op.set_chebyshev(cheb_degree)
    prob.solver.forcing_eta = eta
    port, err = None, None
    if rank == 0:
        try:
            port = make_cpu_port(wl)[0]
        except Exception as e:  # noqa: BLE001
            err = repr(e)[:300]
    own = slice(part["own_lo"], part["own_hi"])
    worst = {"T": 0.0, "Tf": 0.0, "sigma": 0.0, "sigma_over_floor": 0.0}
    tiles = True
    for _ in range(steps):
        prob.t += prob.dt
        prob._solve_T()
        prob._solve_viscoelastic()
        loc = {"T": prob.functions_current["T"].x.array[own].cpu().numpy(), "Tf": prob.functions_current["Tf"].x.array[own].cpu().numpy(),
               "sigma": prob.functions_next["sigma"].x.array.view(-1, dim * dim)[own].cpu().numpy()}
        gathered = [None] * world
        dist.all_gather_object(gathered, loc)
        if rank == 0 and err is None:
            try:                                             # a checker failure must not unbalance the collectives
                port.step(fused=True)
                f = port.fields(True)
                T = np.concatenate([g["T"] for g in gathered])       # x-slabs in rank order = the unpartitioned DG numbering
                Tf = np.concatenate([g["Tf"] for g in gathered])
                S = np.concatenate([g["sigma"] for g in gathered])
                tiles = tiles and T.size == f["T"].size
                if tiles:
                    rel = lambda x, y: float(np.max(np.abs(x - y)) / np.max(np.abs(y)))
                    worst["T"] = max(worst["T"], rel(T, f["T"]))
                    worst["Tf"] = max(worst["Tf"], rel(Tf, f["Tf"]))
                    dT = np.abs(f["T"] - f["T_prev"])
                    good = dT > 1e-6
                    sg, so = S[good], f["sigma"].reshape(-1, dim * dim)[good]
                    scale = float(np.max(np.abs(so)))
                    errv = np.max(np.abs(sg - so), axis=1)
                    floor = stress_rounding_floor(port.vp, dT[good], np.abs(f["xi"])[good])
                    worst["sigma"] = max(worst["sigma"], float(np.max(errv) / scale))
                    worst["sigma_over_floor"] = max(worst["sigma_over_floor"], float(np.max(errv - 2.0 * floor) / scale))
                port.end_step()
            except Exception as e:  # noqa: BLE001
                err = repr(e)[:300]
        prob._update_values(current=prob.functions_current["T"], previous=prob.functions_previous["T"])
    peer = bool(prob._thermal_op.peer_memory)
    prob._thermal_op.close()
    dist.barrier()
    if rank != 0:
        return None
    if err is not None:
        return {"error": err, "ok": False}
    ok = tiles and worst["T"] <= PARITY_TOL and worst["Tf"] <= PARITY_TOL and worst["sigma_over_floor"] <= PARITY_TOL
    return {"workload": f"{wl}: {'x'.join(map(str, n))} hexahedra split into {world} x-slabs, {steps} steps, headline solver settings",
            "checker": "oracle/cpu_port.py on the whole plate (rank 0)", "owned_ranges_tile_the_plate": bool(tiles),
            "transport": "NVLink peer memory" if peer else "NCCL", "T": worst["T"], "Tf": worst["Tf"], "sigma": worst["sigma"],
            "sigma_excess_over_rounding_floor": max(worst["sigma_over_floor"], 0.0), "tol": PARITY_TOL, "ok": bool(ok)}


def mechanics_config(ctx, local: int, n=(160, 160, 8), steps: int = 3) -> dict:  # noqa: C901
    """The equilibrium extension (model_parameters["mechanics"]; the reference has no such solve,
    VM:135-139) on a 160x160x8 DG1 plate: time per step with and without it, PCG iterations, the tangent apply against the
    HBM roofline, and the size-independent check that the equilibrated DG stress is in discrete equilibrium."""
    import numpy as np
    import torch
    from fem_glass_tempering_b200 import ThermoViscoProblem
    from fem_glass_tempering_b200 import mesh as msh
    dev = torch.device("cuda", local)
    mesh = msh.plate_mesh(3, n, tuple(float(k) for k in n))
    params = dict(MAIN_PARAMS, sip_penalty=6.0, mechanics={"rtol": 1e-8})
    prob = ThermoViscoProblem(mesh_path="", time=(0.0, 50.0), dt=DT, config=DG1, model_parameters=params, mesh=mesh, ctx=ctx,
                              materialize="minimal", verbose=False)
    prob.setup(dirichlet_bc=False)
    me = prob.mechanics
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    ms_rest, ms_mech, its, resid = [], [], [], []
    nv3 = mesh.n_vertices * 3
    b = torch.empty(nv3, dtype=torch.float64, device=dev)
    for _ in range(steps + 1):                                  # the first step (cold start of du) is not reported
        prob.t += prob.dt
        ev[0].record()
        prob._solve_T()
        prob._solve_viscoelastic()
        ev[1].record()
        prob._solve_mechanics()
        ev[2].record()
        prob._update_values(current=prob.functions_current["T"], previous=prob.functions_previous["T"])
        torch.cuda.synchronize(dev)
        ms_rest.append(ev[0].elapsed_time(ev[1]))
        ms_mech.append(ev[1].elapsed_time(ev[2]))
        its.append(me.last_iters)
        sig = prob.functions_next["sigma"].x.array
        me.rhs(sig, b)                                          # out-of-balance force of the equilibrated stress
        resid.append(float(b.abs().max() / (sig.abs().max() * float(max(n)))))
    x, y = torch.randn(nv3, dtype=torch.float64, device=dev), torch.empty(nv3, dtype=torch.float64, device=dev)
    for _ in range(3):
        me.apply(x, y)
    ev[0].record()
    for _ in range(20):
        me.apply(x, y)
    ev[1].record()
    torch.cuda.synchronize(dev)
    t_apply = ev[0].elapsed_time(ev[1]) / 20
    peak, _src = measured_peak()
    gbs = me.apply_bytes() / t_apply / 1e6
    qp = mesh.n_cells * 4
    ms = float(np.mean(ms_rest[1:]) + np.mean(ms_mech[1:]))
    return {"workload": "plate3d_DG1 + mechanical equilibrium (extension; vector-P1 displacement, Jacobi-PCG to rtol 1e-8)",
            "cells": list(n), "qp_total": qp, "displacement_dofs": nv3, "ms_per_step": ms, "value": qp / ms * 1e3, "unit": UNIT,
            "ms_thermal_and_visco": float(np.mean(ms_rest[1:])), "ms_mechanics": float(np.mean(ms_mech[1:])),
            "mech_pcg_its_per_step": float(np.mean(its[1:])), "steps": steps,
            "equilibrium_residual_rel": max(resid), "equilibrium_ok": bool(max(resid) < 1e-6),
            "dominant_kernel": {"kernel": "k_mech_apply<3> (tangent apply: gather 4x3 displacements, constant strain, RED.ADD.F64 scatter)",
                                "avg_launch_ms": t_apply, "achieved_GBs": gbs, "frac": gbs / peak, "bound": "hbm",
                                "algorithmic_bytes_per_launch": me.apply_bytes()}}


def other_configs(ctx, local: int, args) -> dict:
    """The other BASELINE configs on one GPU, each a short timed run (3 warm-up + a few steps) — parity for these shapes is
    in tests/; here they are put next to the headline so that every named config has a driver-run number."""
    out = {}

    def run(name, workload, steps, warmup=3, **kw):
        try:
            r = Runner(workload, 0, 1, local, ctx, **kw)
            b = r.brief(r.timed(steps, warmup))
            r.close()
            out[name] = b
        except Exception as e:  # noqa: BLE001
            out[name] = {"error": repr(e)[:300]}

    run("C1_main_py_1d_DG1_CG1", "C1_main_py_1d", 100, 10)
    run("C2_plate2d_CG2_1M_qp", "C2_plate2d_CG2_1M_qp", 10)
    run("C4_share_plate3d_CG2_26.5M_qp", "C4_plate3d_CG2", 5)
    global REFERENCE_PENALTY
    REFERENCE_PENALTY = True
    run("C3_with_reference_penalty_5.0", DEFAULT_WORKLOAD, 5)          # diverges after ~15 steps: 3 + 5 stay below
    REFERENCE_PENALTY = False
    run("perturbed_plate3d_DG1_general_mesh_kernels", "perturbed_plate3d_DG1", 3)
    try:
        out["mechanics_plate3d_DG1_160x160x8"] = mechanics_config(ctx, local)
    except Exception as e:  # noqa: BLE001
        out["mechanics_plate3d_DG1_160x160x8"] = {"error": repr(e)[:300]}
    try:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import bench_visco
        sweep = {}
        for N in (3, 4, 6, 8, 10, 12):
            v = bench_visco.run(ctx, 19_660_800 * 6 // N if N > 6 else 19_660_800, 3, N, 5, False)
            sweep[str(N)] = {"frac": v["frac_of_peak"], "achieved_GBs": v["achieved_GBs"], "ms": v["ms_median"], "n_nodes": v["n_nodes"],
                             "bytes_per_node": v["bytes_per_node"]}
        out["C5_prony_sweep_visco_kernel_d3"] = sweep
    except Exception as e:  # noqa: BLE001
        out["C5_prony_sweep_visco_kernel_d3"] = {"error": repr(e)[:300]}
    return out


class StdoutToStderr:
    """The contract is ONE JSON line on stdout: while the benchmark runs, file descriptor 1 points at stderr so that
    native libraries (NCCL prints its version banner on stdout) cannot interleave; emit() writes to the real stdout."""

    def __enter__(self):
        sys.stdout.flush()
        self._saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def emit(self, text: str):
        sys.stdout.flush()
        os.write(self._saved, (text + "\n").encode())

    def __exit__(self, *a):
        sys.stdout.flush()
        os.dup2(self._saved, 1)
        os.close(self._saved)


OUT = None


def main():
    global OUT
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the GPU-vs-CPU-port parity check of the timed path")
    ap.add_argument("--no-other-configs", action="store_true", help="headline workload only (no other_configs / c4 objects)")
    ap.add_argument("--cheb", type=int, default=None, help="Chebyshev preconditioner degree of the DG solver (0 = off)")
    ap.add_argument("--reference-penalty", action="store_true",
                    help="3-D DG plates with the reference's SIP penalty 5.0 instead of the coercive 6.0 (keep the run under 15 steps)")
    ap.add_argument("--newton-guess", default=None, choices=["previous", "extrapolate"],
                    help='start of the Newton iteration: "previous" = T_n like the reference (TVP:389), "extrapolate" = 2 T_n - T_{n-1}')
    ap.add_argument("--eta", type=float, default=None, help="first forcing term of the inexact Newton iteration (0 = fixed tolerance)")
    args = ap.parse_args()
    global REFERENCE_PENALTY, NEWTON_GUESS
    REFERENCE_PENALTY = args.reference_penalty
    NEWTON_GUESS = args.newton_guess
    with StdoutToStderr() as OUT:
        if args.impl == "reference":
            run_reference(args)
        else:
            run_gpu(args)


if __name__ == "__main__":
    main()
