"""GPU: the hot path at BASELINE.json's FULL sizes (config 3: 320x320x8 hexahedra x 6 = 4 915 200 tetrahedra, 19 660 800
DG1 points; one GPU's share of config 4: 96x768x6, CG2, 3.86 M nodes), where the CPU oracle cannot run in seconds.  Parity
is carried over from the small-mesh tests through properties that do not depend on the size:

  * viscoelastic update: the full-size launch equals, BIT FOR BIT, the same kernel run on a random sample of its nodes
    (the update is pointwise), and that sample equals the C oracle (what tests/test_visco_gpu.py asserts on small arrays);
  * heat operator: symmetry  x.Jy = y.Jx,  linearity, positivity, and the known answer  F(T_prev = T = const) summed over
    all dofs = dt * 0.001 * (sigma eps (T^4 - Ta^4) + htc (T - Ta)) * |boundary|  (a constant is in the DG/CG space, so mass,
    stiffness and interior-penalty terms drop out of the sum: TVP:293-325);
  * PCG: the returned solution's residual, recomputed with an independent apply;
  * time steps: discrete energy conservation, sum_i F_i(T_cur, T_prev) = 0 at the Newton solution, i.e. the heat that left
    through the faces is what the plate lost; temperatures stay bounded (DG is not monotone: a 1-2 % overshoot next to the faces).
"""
import numpy as np
import pytest
import torch

from fem_glass_tempering_b200 import ThermoViscoProblem, _lib, fe
from fem_glass_tempering_b200 import mesh as msh
from fem_glass_tempering_b200.thermal_op import ThermalOperator
from helpers import assert_same
from oracle import visco_oracle as vo

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")
DT = 0.1


def params(**kw):
    return dict(vo.MAIN_PARAMS, **kw)


def flux_density(T, p):
    return 0.001 * (p["sigma"] * p["epsilon"] * (T ** 4 - p["T_ambient"] ** 4) + p["htc"] * (T - p["T_ambient"]))


def test_visco_update_full_size_equals_its_own_sample_and_the_oracle(sg_ctx):
    n, d, N = 19_660_800, 3, 6
    p = vo.ViscoParams(dim=d, dt=DT)
    plan = _lib.ViscoPlan(sg_ctx, dim=d, dt=DT, H=p.H, Rg=p.Rg, Tb=p.Tb, alpha_solid=p.alpha_solid, alpha_liquid=p.alpha_liquid,
                          m=p.m, lambda_m=p.lambda_m, g=p.g, lambda_g=p.lambda_g, k=p.k, lambda_k=p.lambda_k)
    g = torch.Generator(device=DEV).manual_seed(1234)
    rnd = lambda m: torch.rand(m, dtype=torch.float64, device=DEV, generator=g)
    t = {"T_cur": 650 + 200 * rnd(n)}
    t["T_prev"] = t["T_cur"] + 0.05 + 0.95 * rnd(n)
    t["Tf_partial"] = t["T_prev"].repeat_interleave(N) + 5 * rnd(n * N)
    t["s_tilde"] = 1e-3 * torch.randn(n * N * 9, dtype=torch.float64, device=DEV, generator=g)
    t["sigma_tilde"] = 1e-3 * torch.randn(n * N * 9, dtype=torch.float64, device=DEV, generator=g)
    for name, bs in (("Tf", 1), ("phi", 1), ("xi", 1), ("sigma", 9)):
        t[name] = torch.zeros(n * bs, dtype=torch.float64, device=DEV)
    # a random sample of nodes (plus the first and last tiles), copied BEFORE the in-place update
    idx = torch.cat([torch.arange(0, 64, device=DEV), torch.randint(0, n, (8192 - 128,), device=DEV, generator=g),
                     torch.arange(n - 64, n, device=DEV)])
    bs = {"T_cur": 1, "T_prev": 1, "Tf_partial": N, "s_tilde": N * 9, "sigma_tilde": N * 9, "Tf": 1, "phi": 1, "xi": 1, "sigma": 9}
    take = lambda a, b: a.view(n, b)[idx].reshape(-1).contiguous()
    s = {k: take(t[k], b) for k, b in bs.items()}
    host_in = {k: s[k].cpu().numpy().copy() for k in ("T_cur", "T_prev", "Tf_partial", "s_tilde", "sigma_tilde")}
    plan.update(n, t)
    plan.update(idx.numel(), s)
    torch.cuda.synchronize()
    for k, b in bs.items():
        assert torch.equal(take(t[k], b), s[k]), k                 # pointwise: full launch == sample launch, bit for bit
    # ... and the sample against the C oracle (bitwise up to the 1-ulp difference of CUDA's and glibc's exp)
    m = idx.numel()
    o = dict(Tf=np.zeros(m), phi=np.zeros(m), xi=np.zeros(m), sig=np.zeros(m * 9))
    vo.step_fused(p, host_in["T_cur"], host_in["T_prev"], host_in["Tf_partial"], o["Tf"], o["phi"], o["xi"], host_in["s_tilde"],
                  host_in["sigma_tilde"], o["sig"])
    rel = lambda a, b: float(np.max(np.abs(a - b)) / np.max(np.abs(b)))
    assert rel(s["phi"].cpu().numpy(), o["phi"]) < 1e-15 and rel(s["Tf"].cpu().numpy(), o["Tf"]) < 1e-15
    assert rel(s["Tf_partial"].cpu().numpy(), host_in["Tf_partial"]) < 1e-15
    assert rel(s["s_tilde"].cpu().numpy(), host_in["s_tilde"]) < 1e-12 and rel(s["sigma"].cpu().numpy(), o["sig"]) < 1e-10
    # every node was updated exactly once: no tile skipped, none done twice (checksum over tiles of 32 nodes)
    assert bool((t["phi"] > 0).all()) and bool(torch.isfinite(t["sigma"]).all())


@pytest.fixture(scope="module")
def c3_operator(sg_ctx):
    mesh = msh.plate_mesh(3, (320, 320, 8), (320.0, 320.0, 8.0))
    space = fe.ScalarSpace(mesh, "DG", 1)
    op = ThermalOperator(sg_ctx, space, params(sip_penalty=6.0), DT)
    yield mesh, space, op
    op.close()


def test_c3_operator_properties_at_full_size(c3_operator):
    mesh, space, op = c3_operator
    n = space.n_nodes
    assert n == 19_660_800 and op.class_info()["active"]
    g = torch.Generator(device=DEV).manual_seed(7)
    T = 700 + 100 * torch.rand(n, dtype=torch.float64, device=DEV, generator=g)
    x = torch.randn(n, dtype=torch.float64, device=DEV, generator=g)
    y = torch.randn(n, dtype=torch.float64, device=DEV, generator=g)
    new = lambda: torch.empty(n, dtype=torch.float64, device=DEV)
    Jx, Jy = op.jac_apply(T, x, new()), op.jac_apply(T, y, new())
    xJy, yJx, xJx = float(x @ Jy), float(y @ Jx), float(x @ Jx)
    assert abs(xJy - yJx) <= 1e-11 * max(abs(xJy), float(torch.linalg.norm(x) * torch.linalg.norm(Jy)) * 1e-3)     # symmetric
    assert xJx > 0.0                                                                                                # positive
    Jz = op.jac_apply(T, 2.0 * x - 3.0 * y, new())
    assert float((Jz - (2.0 * Jx - 3.0 * Jy)).abs().max()) <= 1e-12 * float(Jx.abs().max() + Jy.abs().max())          # linear
    # known answer: constant temperature, F summed over all dofs = dt * flux density * boundary area
    p = params()
    Tc = torch.full((n,), 800.0, dtype=torch.float64, device=DEV)
    F = op.residual(Tc, Tc, new())
    area = 2.0 * (320.0 * 320.0 + 2 * 320.0 * 8.0)
    expect = DT * flux_density(800.0, p) * area
    assert abs(float(F.sum()) - expect) <= 1e-10 * expect
    # interior dofs see neither the faces nor a jump: their residual entries vanish for a constant field
    assert float(F.abs().median()) <= 1e-12 * float(F.abs().max())


def test_c3_pcg_residual_and_energy_balance_at_full_size(c3_operator):
    mesh, space, op = c3_operator
    n = space.n_nodes
    new = lambda: torch.empty(n, dtype=torch.float64, device=DEV)
    g = torch.Generator(device=DEV).manual_seed(11)
    T = torch.full((n,), 800.0, dtype=torch.float64, device=DEV)
    b = torch.randn(n, dtype=torch.float64, device=DEV, generator=g)
    x = torch.zeros(n, dtype=torch.float64, device=DEV)
    op.prepare_preconditioner(T)
    its, res = op.pcg(T, b, x, rtol=1e-10)
    r = b - op.jac_apply(T, x, new())
    assert float(torch.linalg.norm(r)) <= 2e-10 * float(torch.linalg.norm(b)) and 0 < its < 200
    # three implicit-Euler steps: conservation and the maximum principle
    p = params()
    T_prev, T_cur = T.clone(), T.clone()
    for step in range(3):
        st = op.timestep(T_cur, T_prev)
        assert st.converged
        F = op.residual(T_cur, T_prev, new())
        lost = float(op.residual(T_cur, T_cur, new()).sum())       # dt * heat flux through the faces at the new temperature
        assert lost > 0.0 and abs(float(F.sum())) <= 1e-9 * lost, (step, float(F.sum()), lost)
        # DG1 + interior penalty is not monotone: the first layer of cells overshoots the initial 800 K next to the
        # steep surface gradient (the CPU oracle shows the same on small plates); bounded, and the plate as a whole cools
        assert float(T_cur.max()) <= 800.0 * 1.03 and float(T_cur.min()) >= p["T_ambient"]
        assert float(T_cur.min()) < 800.0 - 1e-3 and float(T_cur.sum()) < float(T_prev.sum())
        T_prev.copy_(T_cur)


def test_c4_share_cg2_operator_properties(sg_ctx):
    """One GPU's share of config 4 (96x768x6 hexahedra, CG2, 3.86 M nodes): the gather-form apply."""
    mesh = msh.plate_mesh(3, (96, 768, 6), (96.0, 768.0, 6.0))
    space = fe.ScalarSpace(mesh, "CG", 2)
    op = ThermalOperator(sg_ctx, space, params(), DT, cheb_degree=0)
    n = space.n_nodes
    assert n == 193 * 1537 * 13 and op.stencil_info()["active"]
    g = torch.Generator(device=DEV).manual_seed(3)
    T = 700 + 100 * torch.rand(n, dtype=torch.float64, device=DEV, generator=g)
    x = torch.randn(n, dtype=torch.float64, device=DEV, generator=g)
    y = torch.randn(n, dtype=torch.float64, device=DEV, generator=g)
    new = lambda: torch.empty(n, dtype=torch.float64, device=DEV)
    Jx, Jy = op.jac_apply(T, x, new()), op.jac_apply(T, y, new())
    xJy, yJx = float(x @ Jy), float(y @ Jx)
    assert abs(xJy - yJx) <= 1e-11 * float(torch.linalg.norm(x) * torch.linalg.norm(Jy)) and float(x @ Jx) > 0.0
    p = params()
    Tc = torch.full((n,), 800.0, dtype=torch.float64, device=DEV)
    F = op.residual(Tc, Tc, new())
    area = 2.0 * (96.0 * 768.0 + 96.0 * 6.0 + 768.0 * 6.0)
    expect = DT * flux_density(800.0, p) * area
    assert abs(float(F.sum()) - expect) <= 1e-10 * expect
    # J applied to a constant: stiffness drops out, sum = |plate| + dt * d(flux)/dT * |boundary|
    one = torch.ones(n, dtype=torch.float64, device=DEV)
    J1 = op.jac_apply(Tc, one, new())
    dflux = 0.001 * (4.0 * p["sigma"] * p["epsilon"] * 800.0 ** 3 + p["htc"])
    expect = 96.0 * 768.0 * 6.0 + DT * dflux * area
    assert abs(float(J1.sum()) - expect) <= 1e-10 * expect
    op.close()
