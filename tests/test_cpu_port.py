"""CPU: oracle/cpu_port.py (the CPU baseline of bench.py and the checker of its parity_check) against the reference-shaped
oracle of the whole time step (oracle/reference_problem.py: direct solves, cell-wise interpolation)."""
import numpy as np
import pytest

from fem_glass_tempering_b200 import fe
from fem_glass_tempering_b200 import mesh as msh
from helpers import rel_err, stress_rounding_floor
from oracle.cpu_port import CpuTimestep, probe_reference_stack
from oracle.reference_problem import OracleProblem
from oracle.visco_oracle import MAIN_PARAMS


@pytest.mark.parametrize("dim,family,degree,n,params", [
    (3, "DG", 1, (4, 4, 2), dict(MAIN_PARAMS, sip_penalty=6.0)),
    (2, "CG", 2, (8, 4), MAIN_PARAMS),
])
def test_cpu_port_lands_on_the_oracle_solution(dim, family, degree, n, params):
    m = msh.plate_mesh(dim, n, tuple(float(k) for k in n))
    s = fe.ScalarSpace(m, family, degree)
    sp = dict(dofmap=s.dofmap, ref_nodes=s.element.nodes, family=family, degree=degree)
    orc = OracleProblem(m.x, m.cells, sp, sp, params, 0.1)
    port = CpuTimestep(m, s, params, 0.1, threads=2)
    assert port.threads >= 1
    for fused in (True, False):
        for step in range(3):
            if fused:
                orc.step()
            port.step(fused)
            f = port.fields(fused)
            if fused:
                o = orc.f
                assert rel_err(f["T"], o["T_cur"]) <= 1e-12
                assert rel_err(f["Tf"], o["Tf_cur"]) <= 1e-12
                assert rel_err(f["xi"], o["xi"]) <= 1e-10
                d = dim
                dT = np.abs(o["T_cur"] - o["T_prev"])
                err = np.max(np.abs(f["sigma"].reshape(-1, d * d) - o["sigma_next"].reshape(-1, d * d)), axis=1)
                floor = stress_rounding_floor(orc.vp, dT, np.abs(o["xi"]))
                assert np.all(err <= 1e-10 * np.max(np.abs(o["sigma_next"])) + 2 * floor)
                orc.end_step()
            port.end_step()
    # the dolfinx-shaped 17-pass chain and the fused sweep are the same arithmetic
    assert np.isfinite(port.fields(False)["T"]).all()


def test_reference_stack_probe_reports_what_is_missing():
    r = probe_reference_stack()
    assert set(r) >= {"missing_modules", "runnable", "mpiexec", "baseline_ref_dir"}
    assert r["runnable"] == (not r["missing_modules"])
