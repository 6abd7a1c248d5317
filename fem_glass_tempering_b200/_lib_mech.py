"""ctypes structs/argtypes of the mechanical-equilibrium entry points (include/surroglas_b200.h, section C)."""
import ctypes as C


class MechDescC(C.Structure):
    _fields_ = [("dim", C.c_int32), ("n_vertices", C.c_int64), ("n_cells", C.c_int64),
                ("coords", C.c_void_p), ("cells", C.c_void_p), ("fixed", C.c_void_p),
                ("n_ld_sigma", C.c_int32), ("n_sigma_nodes", C.c_int64),
                ("sigma_dofmap", C.c_void_p), ("sigma_weights", C.c_void_p), ("winner_cell", C.c_void_p)]


class MechFieldsC(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in ("sigma", "mech_strain", "total_strain", "deviatoric_strain", "ds_partial",
                                          "dsigma_partial", "s_partial", "sigma_partial", "s_tilde", "sigma_tilde")]


def bind(L) -> None:
    vp = C.c_void_p
    L.sg_mech_op_create.argtypes = [vp, C.POINTER(MechDescC), C.POINTER(vp)]
    L.sg_mech_op_destroy.argtypes = [vp]
    L.sg_mech_coefficients.argtypes = [vp, C.c_int64, vp, vp, vp, vp]
    L.sg_mech_set_moduli.argtypes = [vp, vp, vp, vp]
    L.sg_mech_apply.argtypes = [vp, vp, vp, vp]
    L.sg_mech_rhs.argtypes = [vp, vp, vp, vp]
    L.sg_mech_solve.argtypes = [vp, vp, vp, C.c_double, C.c_double, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_double), vp]
    L.sg_mech_correct.argtypes = [vp, vp, vp, vp, vp, vp, C.POINTER(MechFieldsC), vp]
    L.sg_mech_apply_bytes.argtypes = [vp]
    L.sg_mech_apply_bytes.restype = C.c_int64
