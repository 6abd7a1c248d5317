"""ctypes binding of libsurroglas_b200.so (the C ABI in include/surroglas_b200.h).

There is NO CPU fallback: if the CUDA library is missing, or a call fails, this
module raises.  PyTorch is used by callers only to own device buffers; what crosses
this boundary are raw device pointers, sizes and a stream handle.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "lib", "libsurroglas_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_PKG), "include", "surroglas_b200.h")

SG_MAX_TERMS = 16
SG_OK, SG_E_INVALID, SG_E_CUDA, SG_E_NOCONV, SG_E_NCCL, SG_E_UNSUPPORTED = 0, -1, -2, -3, -4, -5
PHASE_TF, PHASE_STRAIN, PHASE_SHIFT, PHASE_STRESS, PHASE_ALL = 1, 2, 4, 8, 15

_dbl16 = C.c_double * SG_MAX_TERMS


class SgError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libsurroglas_b200 error {code}: {msg}")
        self.code = code


class ViscoParamsC(C.Structure):
    _fields_ = [("dim", C.c_int32), ("n_terms", C.c_int32),
                ("H", C.c_double), ("Rg", C.c_double), ("Tb", C.c_double),
                ("alpha_solid", C.c_double), ("alpha_liquid", C.c_double), ("dt", C.c_double),
                ("m", _dbl16), ("lambda_m", _dbl16), ("g", _dbl16), ("lambda_g", _dbl16),
                ("k", _dbl16), ("lambda_k", _dbl16),
                ("mode", C.c_int32), ("chi", C.c_double)]


VISCO_REFERENCE, VISCO_CORRECTED = 0, 1
VISCO_FIELD_NAMES = ("T_cur", "T_prev", "Tf_partial", "Tf", "phi", "xi", "s_tilde", "sigma_tilde", "sigma",
                     "T_next", "phi_next", "thermal_strain", "total_strain", "deviatoric_strain",
                     "ds_partial", "dsigma_partial", "s_partial", "sigma_partial")


class ViscoFieldsC(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in VISCO_FIELD_NAMES]


class ViscoGatherC(C.Structure):
    _fields_ = [("n_ld", C.c_int32), ("n_points", C.c_int32), ("dofs", C.c_void_p),
                ("local_point", C.c_void_p), ("weights", C.c_void_p)]


def build(verbose: bool = False) -> str:
    """Compile the CUDA library in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    res = subprocess.run(["make", "-j", str(min(8, os.cpu_count() or 1)), "-C", os.path.join(_PKG, "csrc")],
                         capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("building libsurroglas_b200.so failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stdout)
    return LIB_PATH


_lib = None


def lib() -> C.CDLL:
    """Load the shared library; raises if it has not been built (no fallback path exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C fem_glass_tempering_b200/csrc`). There is no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        L.sg_last_error.restype = C.c_char_p
        L.sg_visco_bytes_per_node.restype = C.c_int64
        L.sg_launch_count.restype = C.c_int64
        L.sg_visco_update.argtypes = [C.c_void_p, C.c_int64, C.POINTER(ViscoFieldsC), C.c_uint32, C.c_void_p]
        L.sg_visco_update_scalar.argtypes = [C.c_void_p, C.c_int64, C.POINTER(ViscoFieldsC), C.c_uint32, C.c_void_p]
        L.sg_visco_update_tensor.argtypes = [C.c_void_p, C.c_int64, C.POINTER(ViscoFieldsC), C.POINTER(ViscoGatherC),
                                             C.c_uint32, C.c_void_p]
        L.sg_visco_plan_create.argtypes = [C.c_void_p, C.POINTER(ViscoParamsC), C.POINTER(C.c_void_p)]
        L.sg_visco_plan_destroy.argtypes = [C.c_void_p]
        L.sg_ctx_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]
        L.sg_ctx_destroy.argtypes = [C.c_void_p]
        L.sg_ctx_sm_count.argtypes = [C.c_void_p]
        L.sg_nccl_unique_id.argtypes = [C.c_void_p]
        _bind_thermal(L)
        _lib = L
    return _lib


def _bind_thermal(L) -> None:
    """argtypes of the thermal/PCG/halo entry points (present once thermal.cu is built)."""
    if not hasattr(L, "sg_thermal_op_create"):
        return
    from . import _lib_mech, _lib_thermal
    _lib_thermal.bind(L)
    _lib_mech.bind(L)


def check(rc: int) -> None:
    if rc != 0:
        raise SgError(rc, lib().sg_last_error().decode())


def exported_symbols_in_header() -> list[str]:
    """Every function name include/surroglas_b200.h declares (used by the CPU-side ABI test)."""
    import re
    src = open(HEADER_PATH).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sg_[a-z0-9_]+)\s*\(", src)))


def ptr(t) -> int | None:
    """Device pointer of a torch tensor (None -> NULL)."""
    if t is None:
        return None
    assert t.is_cuda and t.is_contiguous(), "libsurroglas_b200 takes contiguous CUDA tensors"
    return t.data_ptr()


class _DevicePtrView:
    """__cuda_array_interface__ wrapper of library-owned device memory."""

    def __init__(self, address: int, n: int):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (address, False), "version": 3, "strides": None}


def tensor_from_ptr(address: int, n: int, device):
    """float64 torch view of n doubles of device memory the library owns (e.g. the IPC-exported solver workspace)."""
    import torch
    return torch.as_tensor(_DevicePtrView(address, n), device=device)


def current_stream_ptr() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream


class Context:
    """sg_ctx wrapper: one per process/GPU."""

    def __init__(self, device: int = 0, rank: int = 0, nranks: int = 1, nccl_unique_id: bytes | None = None):
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError("fem_glass_tempering_b200 needs a CUDA device (B200, sm_100a); there is no CPU path")
        self.device = device
        self.rank, self.nranks = rank, nranks
        h = C.c_void_p()
        buf = C.create_string_buffer(nccl_unique_id, 128) if nccl_unique_id is not None else None
        check(lib().sg_ctx_create(device, rank, nranks, buf, C.byref(h)))
        self.handle = h

    @property
    def sm_count(self) -> int:
        return lib().sg_ctx_sm_count(self.handle)

    def close(self) -> None:
        if self.handle:
            lib().sg_ctx_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def nccl_unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    check(lib().sg_nccl_unique_id(buf))
    return buf.raw


class ViscoPlan:
    """sg_visco_plan wrapper (constants of ViscoelasticModel.__init__, VM:9-84)."""

    def __init__(self, ctx: Context, *, dim: int, dt: float, H: float, Rg: float, Tb: float, alpha_solid: float,
                 alpha_liquid: float, m, lambda_m, g, lambda_g, k, lambda_k, mode: int = VISCO_REFERENCE, chi: float = 0.5):
        p = ViscoParamsC()
        p.mode, p.chi = int(mode), float(chi)
        p.dim, p.n_terms = dim, len(m)
        if not (len(m) == len(lambda_m) == len(g) == len(lambda_g) == len(k) == len(lambda_k)):
            raise ValueError("Prony tables must have equal lengths")
        if len(m) > SG_MAX_TERMS:
            raise ValueError(f"at most {SG_MAX_TERMS} Prony terms are supported")
        p.H, p.Rg, p.Tb, p.alpha_solid, p.alpha_liquid, p.dt = H, Rg, Tb, alpha_solid, alpha_liquid, dt
        for name, vals in (("m", m), ("lambda_m", lambda_m), ("g", g), ("lambda_g", lambda_g), ("k", k),
                           ("lambda_k", lambda_k)):
            arr = getattr(p, name)
            for i, v in enumerate(vals):
                arr[i] = float(v)
        self.params = p
        self.ctx = ctx
        self.dim, self.N = dim, len(m)
        h = C.c_void_p()
        check(lib().sg_visco_plan_create(ctx.handle, C.byref(p), C.byref(h)))
        self.handle = h

    def _fields(self, tensors: dict) -> ViscoFieldsC:
        f = ViscoFieldsC()
        for name in VISCO_FIELD_NAMES:
            setattr(f, name, ptr(tensors.get(name)))
        return f

    def update(self, n_nodes: int, tensors: dict, phases: int = PHASE_ALL, stream: int | None = None) -> None:
        f = self._fields(tensors)
        check(lib().sg_visco_update(self.handle, n_nodes, C.byref(f), phases,
                                    stream if stream is not None else current_stream_ptr()))

    def update_scalar(self, n_nodes: int, tensors: dict, phases: int = PHASE_ALL, stream: int | None = None) -> None:
        f = self._fields(tensors)
        check(lib().sg_visco_update_scalar(self.handle, n_nodes, C.byref(f), phases,
                                           stream if stream is not None else current_stream_ptr()))

    def update_tensor(self, n_nodes: int, tensors: dict, gather: dict, phases: int = PHASE_ALL,
                      stream: int | None = None) -> None:
        f = self._fields(tensors)
        g = ViscoGatherC()
        g.n_ld, g.n_points = gather["n_ld"], gather["n_points"]
        g.dofs, g.local_point, g.weights = ptr(gather["dofs"]), ptr(gather["local_point"]), ptr(gather["weights"])
        check(lib().sg_visco_update_tensor(self.handle, n_nodes, C.byref(f), C.byref(g), phases,
                                           stream if stream is not None else current_stream_ptr()))

    def bytes_per_node(self, tensors: dict, phases: int = PHASE_ALL) -> int:
        f = self._fields(tensors)
        return lib().sg_visco_bytes_per_node(C.byref(self.params), C.byref(f), phases)

    def close(self) -> None:
        if self.handle:
            lib().sg_visco_plan_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
