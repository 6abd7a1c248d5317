"""CPU: the C-ABI shared library loads and exports every symbol include/surroglas_b200.h declares."""
import ctypes
import os

from fem_glass_tempering_b200 import _lib


def test_library_is_built():
    assert os.path.exists(_lib.LIB_PATH), "run __graft_entry__.build() first"


def test_exports_every_declared_symbol():
    names = _lib.exported_symbols_in_header()
    assert len(names) >= 10
    L = ctypes.CDLL(_lib.LIB_PATH)
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, f"declared in the header but not exported: {missing}"


def test_error_reporting_without_gpu():
    L = _lib.lib()
    assert L.sg_version() >= 100
    # argument validation happens before any CUDA call
    rc = L.sg_visco_plan_create(None, None, None)
    assert rc == _lib.SG_E_INVALID
    assert b"NULL" in L.sg_last_error()


def test_bytes_per_node_formula():
    """8*(5 + 2N + 4Nd^2 + d^2) B per node for the compulsory traffic."""
    L = _lib.lib()
    for d, N, expect in ((1, 6, 336), (2, 6, 936), (3, 6, 1936), (3, 3, 1024), (3, 8, 2544), (3, 12, 3760)):
        p = _lib.ViscoParamsC()
        p.dim, p.n_terms = d, N
        f = _lib.ViscoFieldsC()
        assert L.sg_visco_bytes_per_node(ctypes.byref(p), ctypes.byref(f), _lib.PHASE_ALL) == expect
