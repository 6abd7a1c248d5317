"""ctypes structs/argtypes of the thermal operator, solver and halo entry points (include/surroglas_b200.h)."""
import ctypes as C


class ThermalDescC(C.Structure):
    _fields_ = [("dim", C.c_int32), ("degree", C.c_int32), ("family", C.c_int32),
                ("n_cells", C.c_int64), ("cell_lo", C.c_int64), ("cell_hi", C.c_int64),
                ("n_dofs", C.c_int64), ("own_lo", C.c_int64), ("own_hi", C.c_int64),
                ("dofmap", C.c_void_p), ("geom", C.c_void_p), ("nbr", C.c_void_p), ("nbinfo", C.c_void_p),
                ("n_bfacets", C.c_int64), ("bf_cell", C.c_void_p), ("bf_facet", C.c_void_p), ("bf_area", C.c_void_p),
                ("n_ld", C.c_int32), ("nqc", C.c_int32), ("nqf", C.c_int32), ("nqb", C.c_int32), ("n_perm", C.c_int32),
                ("mass", C.c_void_p), ("load", C.c_void_p), ("cq_w", C.c_void_p), ("cq_grad", C.c_void_p),
                ("fq_w", C.c_void_p), ("fq_val", C.c_void_p), ("fq_grad", C.c_void_p), ("fq_perm", C.c_void_p),
                ("bq_w", C.c_void_p), ("bq_val", C.c_void_p),
                ("dt", C.c_double), ("alpha", C.c_double), ("f", C.c_double), ("sigma", C.c_double),
                ("epsilon", C.c_double), ("htc", C.c_double), ("T_ambient", C.c_double), ("penalty", C.c_double),
                ("own_cell_lo", C.c_int64), ("own_cell_hi", C.c_int64), ("flags", C.c_int32)]


class HaloSegmentC(C.Structure):
    _fields_ = [("peer", C.c_int32), ("send_offset", C.c_int64), ("send_count", C.c_int64),
                ("recv_offset", C.c_int64), ("recv_count", C.c_int64)]


class NewtonOptsC(C.Structure):
    _fields_ = [("newton_rtol", C.c_double), ("newton_atol", C.c_double), ("newton_max_it", C.c_int32),
                ("lin_rtol", C.c_double), ("lin_atol", C.c_double), ("lin_max_it", C.c_int32),
                ("forcing_eta", C.c_double)]


class NewtonStatsC(C.Structure):
    _fields_ = [("newton_its", C.c_int32), ("lin_its", C.c_int32), ("converged", C.c_int32),
                ("dx_norm_first", C.c_double), ("dx_norm_last", C.c_double), ("lin_rel_res_last", C.c_double)]


def bind(L) -> None:
    vp = C.c_void_p
    L.sg_thermal_op_create.argtypes = [vp, C.POINTER(ThermalDescC), C.POINTER(vp)]
    L.sg_thermal_op_destroy.argtypes = [vp]
    L.sg_thermal_residual.argtypes = [vp, vp, vp, vp, vp]
    L.sg_thermal_jac_apply.argtypes = [vp, vp, vp, vp, vp]
    L.sg_thermal_jac_diag.argtypes = [vp, vp, vp, vp]
    L.sg_thermal_class_info.argtypes = [vp, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    L.sg_thermal_stencil_info.argtypes = [vp, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    L.sg_thermal_profile.argtypes = [vp, C.c_int32, C.c_int32]
    L.sg_thermal_profile_read.argtypes = [vp, C.POINTER(C.c_int64), C.POINTER(C.c_double)]
    L.sg_thermal_profile_read_kind.argtypes = [vp, C.c_int32, C.POINTER(C.c_int64), C.POINTER(C.c_double)]
    L.sg_thermal_cheb_step_bytes.argtypes = [vp, C.c_int32]
    L.sg_thermal_cheb_step_bytes.restype = C.c_int64
    L.sg_thermal_apply_bytes.argtypes = [vp]
    L.sg_thermal_apply_bytes.restype = C.c_int64
    L.sg_halo_plan_create.argtypes = [vp, C.c_int32, C.POINTER(HaloSegmentC), C.POINTER(vp)]
    L.sg_halo_plan_destroy.argtypes = [vp]
    L.sg_halo_forward.argtypes = [vp, vp, C.c_int32, vp]
    L.sg_halo_peer_alloc.argtypes = [vp, vp, C.c_int64]
    L.sg_halo_peer_open.argtypes = [vp, vp, vp]
    L.sg_halo_peer_workspace.argtypes = [vp]
    L.sg_halo_peer_workspace.restype = vp
    L.sg_halo_uses_peer_memory.argtypes = [vp]
    L.sg_thermal_solver_workspace_doubles.argtypes = [vp]
    L.sg_thermal_solver_workspace_doubles.restype = C.c_int64
    L.sg_thermal_solver_create.argtypes = [vp, vp, vp, C.POINTER(vp)]
    L.sg_thermal_solver_destroy.argtypes = [vp]
    L.sg_thermal_solver_set_chebyshev.argtypes = [vp, C.c_int32, C.c_double, C.c_double]
    L.sg_thermal_solver_get_chebyshev.argtypes = [vp, C.POINTER(C.c_int32), C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.sg_pcg_solve.argtypes = [vp, vp, vp, vp, C.c_double, C.c_double, C.c_int32, C.POINTER(C.c_int32),
                               C.POINTER(C.c_double), vp]
    L.sg_thermal_timestep.argtypes = [vp, vp, vp, C.POINTER(NewtonOptsC), C.POINTER(NewtonStatsC), vp]
