"""TEST INFRASTRUCTURE — numpy/ctypes front end of oracle/visco_oracle.c.

CPU restatement of hot path (A), the ViscoelasticModel pointwise chain of
/root/reference/ViscoelasticModel.py (VM) in the call order of
/root/reference/ThermoViscoProblem.py (TVP).  PARITY UNPINNED (the reference
has no tests or golden vectors; dolfinx cannot be installed here) — see the
header of visco_oracle.c.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may
import this module.  The product package never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_MAXN = 16

# VM:19-68 — the six 6-term Prony tableaux, restated.
PRONY_M = (5.523e-2, 8.205e-2, 1.215e-1, 2.286e-1, 2.860e-1, 2.265e-1)
PRONY_LAMBDA_M = (5.965e-4, 1.077e-2, 1.362e-1, 1.505e-1, 6.747e+0, 2.963e+1)
PRONY_G = (1.585, 2.354, 3.486, 6.558, 8.205, 6.498)
PRONY_LAMBDA_G = (6.658e-5, 1.197e-3, 1.514e-2, 1.672e-1, 7.497e-1, 3.292e+0)
PRONY_K = (7.588e-1, 7.650e-1, 9.806e-1, 7.301e+0, 1.347e+1, 1.090e+1)
PRONY_LAMBDA_K = (5.009e-5, 9.945e-4, 2.022e-3, 1.925e-2, 1.199e-1, 2.033e+0)

# main.py:29-55 — default model parameters, restated.
MAIN_PARAMS = {
    "f": 0.0, "epsilon": 0.93, "sigma": 5.670e-8, "T_ambient": 600.0, "T_0": 800.0,
    "alpha": 1.0, "htc": 280.1, "rho": 2500.0, "cp": 1433.0, "k": 1.0,
    "H": 627.8e3, "Tb": 869.0e0, "Rg": 8.314,
    "alpha_solid": 9.10e-6, "alpha_liquid": 25.10e-6, "Tf_init": 873.0,
}


class _CParams(C.Structure):
    _fields_ = [("dim", C.c_int), ("N", C.c_int),
                ("H", C.c_double), ("Rg", C.c_double), ("Tb", C.c_double),
                ("alpha_solid", C.c_double), ("alpha_liquid", C.c_double), ("dt", C.c_double),
                ("m", C.c_double * _MAXN), ("lambda_m", C.c_double * _MAXN),
                ("g", C.c_double * _MAXN), ("lambda_g", C.c_double * _MAXN),
                ("k", C.c_double * _MAXN), ("lambda_k", C.c_double * _MAXN)]


_STATE_FIELDS = ["T_cur", "T_prev", "T_next", "Tf_partial_cur", "Tf_partial_prev", "Tf_cur", "Tf_prev",
                 "phi", "phi_next", "xi", "thermal_strain", "total_strain", "deviatoric_strain",
                 "ds_partial", "dsigma_partial", "s_tilde_cur", "s_tilde_next",
                 "sigma_tilde_cur", "sigma_tilde_next", "s_partial_cur", "s_partial_next",
                 "sigma_partial_cur", "sigma_partial_next", "sigma_next"]


class _CState(C.Structure):
    _fields_ = [(f, C.POINTER(C.c_double)) for f in _STATE_FIELDS]


def build(force: bool = False) -> None:
    """Compile the C oracle (gcc) into oracle/_build/."""
    out = os.path.join(_HERE, "_build", "libvisco_oracle.so")
    src = os.path.join(_HERE, "visco_oracle.c")
    if force or not os.path.exists(out) or os.path.getmtime(out) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-s"], check=True, capture_output=True)


_libs: dict = {}


def _lib(omp: bool = False):
    key = "omp" if omp else "serial"
    if key not in _libs:
        build()
        name = "libvisco_oracle_omp.so" if omp else "libvisco_oracle.so"
        path = os.path.join(_HERE, "_build", name)
        if omp and not os.path.exists(path):
            raise OSError("OpenMP oracle build unavailable")
        lib = C.CDLL(path)
        assert lib.vo_sizeof_params() == C.sizeof(_CParams)
        assert lib.vo_sizeof_state() == C.sizeof(_CState)
        _libs[key] = lib
    return _libs[key]


def _p(a: np.ndarray):
    assert a.dtype == np.float64 and a.flags.c_contiguous
    return a.ctypes.data_as(C.POINTER(C.c_double))


@dataclass
class ViscoParams:
    """Constants of VM:9-84 plus dt (VM:88)."""
    dim: int
    dt: float
    H: float = MAIN_PARAMS["H"]
    Rg: float = MAIN_PARAMS["Rg"]
    Tb: float = MAIN_PARAMS["Tb"]
    alpha_solid: float = MAIN_PARAMS["alpha_solid"]
    alpha_liquid: float = MAIN_PARAMS["alpha_liquid"]
    m: tuple = PRONY_M
    lambda_m: tuple = PRONY_LAMBDA_M
    g: tuple = PRONY_G
    lambda_g: tuple = PRONY_LAMBDA_G
    k: tuple = PRONY_K
    lambda_k: tuple = PRONY_LAMBDA_K
    _c: _CParams = field(default=None, repr=False)

    @property
    def N(self) -> int:
        return len(self.m)

    def c(self) -> _CParams:
        cp = _CParams()
        cp.dim, cp.N = self.dim, self.N
        cp.H, cp.Rg, cp.Tb = self.H, self.Rg, self.Tb
        cp.alpha_solid, cp.alpha_liquid, cp.dt = self.alpha_solid, self.alpha_liquid, self.dt
        for name in ("m", "lambda_m", "g", "lambda_g", "k", "lambda_k"):
            arr = getattr(cp, name)
            for i, v in enumerate(getattr(self, name)):
                arr[i] = v
        return cp


def prony_tables(n_terms: int) -> dict:
    """Prony tables for the N-term sweep (BASELINE config 5): the first
    N reference entries for N<=6, log-spaced relaxation times with normalised
    weights for N>6 (the reference hard-codes N=6, VM:16)."""
    if n_terms <= 6:
        s = slice(0, n_terms)
        return dict(m=PRONY_M[s], lambda_m=PRONY_LAMBDA_M[s], g=PRONY_G[s], lambda_g=PRONY_LAMBDA_G[s],
                    k=PRONY_K[s], lambda_k=PRONY_LAMBDA_K[s])
    lam = tuple(float(v) for v in np.logspace(-5, 2, n_terms))
    w = np.linspace(1.0, 2.0, n_terms)
    return dict(m=tuple(float(v) for v in w / w.sum()), lambda_m=lam,
                g=tuple(float(v) for v in w / w.sum() * sum(PRONY_G)), lambda_g=lam,
                k=tuple(float(v) for v in w / w.sum() * sum(PRONY_K)), lambda_k=lam)


# ---- one function per reference Expression (arrays are evaluation-point major) ------------

def phi(p: ViscoParams, T):
    T = np.ascontiguousarray(T, dtype=np.float64)
    out = np.empty_like(T)
    cp = p.c()
    _lib().vo_phi(C.byref(cp), C.c_long(T.size), _p(T), _p(out))
    return out


def Tf_partial(p, Tfp_prev, T_cur, phi_):
    out = np.empty_like(Tfp_prev)
    cp = p.c()
    _lib().vo_Tf_partial(C.byref(cp), C.c_long(T_cur.size), _p(Tfp_prev), _p(T_cur), _p(phi_), _p(out))
    return out


def Tf(p, Tfp):
    n = Tfp.size // p.N
    out = np.empty(n)
    cp = p.c()
    _lib().vo_Tf(C.byref(cp), C.c_long(n), _p(Tfp), _p(out))
    return out


def thermal_strain(p, T_cur, T_prev, Tf_cur, Tf_prev):
    out = np.empty(T_cur.size * p.dim * p.dim)
    cp = p.c()
    _lib().vo_thermal_strain(C.byref(cp), C.c_long(T_cur.size), _p(T_cur), _p(T_prev), _p(Tf_cur), _p(Tf_prev), _p(out))
    return out


def total_strain(p, eth):
    out = np.empty_like(eth)
    cp = p.c()
    _lib().vo_total_strain(C.byref(cp), C.c_long(eth.size // (p.dim * p.dim)), _p(eth), _p(out))
    return out


def deviatoric_strain(p, tot):
    out = np.empty_like(tot)
    cp = p.c()
    _lib().vo_deviatoric_strain(C.byref(cp), C.c_long(tot.size // (p.dim * p.dim)), _p(tot), _p(out))
    return out


def T_next(p, T_cur, T_prev):
    out = np.empty_like(T_cur)
    cp = p.c()
    _lib().vo_T_next(C.byref(cp), C.c_long(T_cur.size), _p(T_cur), _p(T_prev), _p(out))
    return out


def xi(p, phi_next, phi_):
    out = np.empty_like(phi_)
    cp = p.c()
    _lib().vo_xi(C.byref(cp), C.c_long(phi_.size), _p(phi_next), _p(phi_), _p(out))
    return out


def ds_partial(p, dev, xi_):
    out = np.empty(xi_.size * p.N * p.dim * p.dim)
    cp = p.c()
    _lib().vo_ds_partial(C.byref(cp), C.c_long(xi_.size), _p(dev), _p(xi_), _p(out))
    return out


def dsigma_partial(p, tot, xi_):
    out = np.empty(xi_.size * p.N * p.dim * p.dim)
    cp = p.c()
    _lib().vo_dsigma_partial(C.byref(cp), C.c_long(xi_.size), _p(tot), _p(xi_), _p(out))
    return out


def tilde_next(p, which: str, tilde_cur, xi_):
    lam = np.zeros(_MAXN)
    lam[:p.N] = p.lambda_g if which == "g" else p.lambda_k
    out = np.empty_like(tilde_cur)
    cp = p.c()
    _lib().vo_tilde_next(C.byref(cp), C.c_long(xi_.size), _p(lam), _p(tilde_cur), _p(xi_), _p(out))
    return out


def add(a, b):
    out = np.empty_like(a)
    _lib().vo_add(C.c_long(a.size), _p(a), _p(b), _p(out))
    return out


def sigma_next(p, s_part, sig_part):
    n = s_part.size // (p.N * p.dim * p.dim)
    out = np.empty(n * p.dim * p.dim)
    cp = p.c()
    _lib().vo_sigma_next(C.byref(cp), C.c_long(n), _p(s_part), _p(sig_part), _p(out))
    return out


# ---- whole-step drivers ----------------------------------------------------------------------

def new_state(p: ViscoParams, n: int, T0: float = 800.0) -> dict:
    """The 24 Functions of TVP:106-173 with the initial conditions of TVP:187-233."""
    N, dd = p.N, p.dim * p.dim
    bs = {"T_cur": 1, "T_prev": 1, "T_next": 1, "Tf_partial_cur": N, "Tf_partial_prev": N, "Tf_cur": 1,
          "Tf_prev": 1, "phi": 1, "phi_next": 1, "xi": 1, "thermal_strain": dd, "total_strain": dd,
          "deviatoric_strain": dd, "sigma_next": dd}
    st = {}
    for f in _STATE_FIELDS:
        st[f] = np.zeros(n * bs.get(f, N * dd))
    for f in ("T_cur", "T_prev", "Tf_cur", "Tf_prev", "Tf_partial_cur", "Tf_partial_prev"):
        st[f][:] = T0
    return st


def step_passes(p: ViscoParams, st: dict, omp: bool = False) -> None:
    """TVP:370-373 — the 17 interpolation passes + copies, in place on `st`."""
    cs = _CState()
    for f in _STATE_FIELDS:
        setattr(cs, f, _p(st[f]))
    cp = p.c()
    _lib(omp).vo_step_passes(C.byref(cp), C.c_long(st["T_cur"].size), C.byref(cs))


def step_fused(p: ViscoParams, T_cur, T_prev, Tf_partial_, Tf_, phi_, xi_, s_tilde, sigma_tilde, sigma, omp=False) -> None:
    cp = p.c()
    _lib(omp).vo_step_fused(C.byref(cp), C.c_long(T_cur.size), _p(T_cur), _p(T_prev), _p(Tf_partial_), _p(Tf_),
                            _p(phi_), _p(xi_), _p(s_tilde), _p(sigma_tilde), _p(sigma))


def step_corrected(p: ViscoParams, chi, T_cur, T_prev, Tf_partial_, Tf_, phi_, xi_, s_hist, k_hist, sigma, omp=False) -> None:
    """CPU statement of the product's OPTIONAL corrected scheme (visco_oracle.c: vo_step_corrected).  Not reference
    behaviour — the specification the CUDA kernels of model_params["physics"] = "corrected" are tested against."""
    cp = p.c()
    _lib(omp).vo_step_corrected(C.byref(cp), C.c_double(chi), C.c_long(T_cur.size), _p(T_cur), _p(T_prev), _p(Tf_partial_),
                                _p(Tf_), _p(phi_), _p(xi_), _p(s_hist), _p(k_hist), _p(sigma))
