"""CPU: the C-ABI shared library loads and exports every symbol include/surroglas_b200.h declares."""
import ctypes
import os

import pytest

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


def test_ctypes_structures_match_the_header_layout(tmp_path):
    """The ctypes mirrors in _lib.py / _lib_thermal.py against the C structs of include/surroglas_b200.h: gcc prints
    sizeof and every offsetof (field names must exist in the header for this to compile), ctypes must agree."""
    import shutil
    import subprocess
    from fem_glass_tempering_b200 import _lib_mech as lm
    from fem_glass_tempering_b200 import _lib_thermal as lt
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    pairs = [("sg_visco_params", _lib.ViscoParamsC), ("sg_visco_fields", _lib.ViscoFieldsC),
             ("sg_visco_gather", _lib.ViscoGatherC), ("sg_thermal_desc", lt.ThermalDescC),
             ("sg_halo_segment", lt.HaloSegmentC), ("sg_newton_opts", lt.NewtonOptsC), ("sg_newton_stats", lt.NewtonStatsC),
             ("sg_mech_desc", lm.MechDescC), ("sg_mech_fields", lm.MechFieldsC)]
    header = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "surroglas_b200.h")
    lines = ["#include <stdio.h>", "#include <stddef.h>", f'#include "{header}"', "int main(void) {"]
    for cname, cls in pairs:
        lines.append(f'printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ["return 0;", "}"]
    src, exe = tmp_path / "layout.c", tmp_path / "layout"
    src.write_text("\n".join(lines))
    subprocess.run(["gcc", "-std=c11", "-o", str(exe), str(src)], check=True, capture_output=True)
    out = dict(l.split() for l in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    for cname, cls in pairs:
        assert int(out[cname]) == ctypes.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert int(out[f"{cname}.{fname}"]) == getattr(cls, fname).offset, f"{cname}.{fname}"


def test_mech_entry_points_validate_their_arguments_without_a_gpu():
    """The equilibrium extension's entry points reject NULL handles before any CUDA call (error code + message)."""
    L = _lib.lib()
    assert L.sg_mech_op_create(None, None, None) == _lib.SG_E_INVALID and b"NULL" in L.sg_last_error()
    assert L.sg_mech_solve(None, None, None, 1e-8, 0.0, 10, None, None, None) == _lib.SG_E_INVALID
    assert L.sg_mech_apply_bytes(None) == -1
