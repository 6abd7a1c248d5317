"""GPU parity of hot path (A): the fused CUDA update (through the C ABI) against the CPU oracle.

Bar (north_star): 1e-10 relative in float64.  What is asserted here is stronger wherever IEEE
allows it:
  * every quantity that does not pass through exp() is BIT-EXACT against the oracle;
  * exp() differs between libm and CUDA by <= 1 ulp, so phi/phi_next are held to 4 ulp, and every
    downstream quantity is checked bit-exactly against the oracle EVALUATED ON THE GPU's phi / xi
    (same inputs -> same bits), plus an end-to-end relative bound against the pure oracle.
"""
import numpy as np
import pytest
import torch

from helpers import assert_same, load_visco_kat, random_visco_state, rel_err
from oracle import visco_oracle as vo

pytestmark = pytest.mark.gpu

OPTIONAL = ("T_next", "phi_next", "thermal_strain", "total_strain", "deviatoric_strain", "ds_partial",
            "dsigma_partial", "s_partial", "sigma_partial")


def make_plan(ctx, p: vo.ViscoParams):
    from fem_glass_tempering_b200 import _lib
    return _lib.ViscoPlan(ctx, dim=p.dim, dt=p.dt, H=p.H, Rg=p.Rg, Tb=p.Tb, alpha_solid=p.alpha_solid,
                          alpha_liquid=p.alpha_liquid, m=p.m, lambda_m=p.lambda_m, g=p.g, lambda_g=p.lambda_g,
                          k=p.k, lambda_k=p.lambda_k)


def gpu_state(p, T_cur, T_prev, Tfp, s, k, materialize=True):
    n, N, dd = T_cur.size, p.N, p.dim * p.dim
    dev = torch.device("cuda:0")
    t = {"T_cur": torch.from_numpy(T_cur).to(dev), "T_prev": torch.from_numpy(T_prev).to(dev),
         "Tf_partial": torch.from_numpy(Tfp.copy()).to(dev), "s_tilde": torch.from_numpy(s.copy()).to(dev),
         "sigma_tilde": torch.from_numpy(k.copy()).to(dev)}
    for name, bs in (("Tf", 1), ("phi", 1), ("xi", 1), ("sigma", dd)):
        t[name] = torch.full((n * bs,), -7.0, dtype=torch.float64, device=dev)
    if materialize:
        for name in OPTIONAL:
            bs = 1 if name in ("T_next", "phi_next") else dd if name.endswith("strain") else N * dd
            t[name] = torch.full((n * bs,), -7.0, dtype=torch.float64, device=dev)
    return t


def check_against_oracle(p, T_cur, T_prev, Tfp_prev, s_prev, k_prev, g, full=True):
    """g: dict of numpy arrays read back from the GPU after one fused update."""
    ulp4 = 4 * np.finfo(np.float64).eps
    # exp-dependent quantities: <= 4 ulp
    phi_o = vo.phi(p, T_cur)
    assert np.max(np.abs(g["phi"] - phi_o) / phi_o) <= ulp4
    # everything else: bit-exact given the GPU's own phi / phi_next
    assert_same(g["Tf_partial"], vo.Tf_partial(p, Tfp_prev, T_cur, g["phi"]), "Tf_partial")
    assert_same(g["Tf"], vo.Tf(p, g["Tf_partial"]), "Tf")
    Tn = vo.T_next(p, T_cur, T_prev)
    phin = g["phi_next"] if full else None
    if full:
        assert_same(g["T_next"], Tn, "T_next")
        phin_o = vo.phi(p, Tn)
        assert np.max(np.abs(phin - phin_o) / phin_o) <= ulp4
        assert_same(g["xi"], vo.xi(p, phin, g["phi"]), "xi")
    eth = vo.thermal_strain(p, T_cur, T_prev, g["Tf"], g["Tf"])
    tot = vo.total_strain(p, eth)
    dev = vo.deviatoric_strain(p, tot)
    if full:
        assert_same(g["thermal_strain"], eth, "thermal_strain")
        assert_same(g["total_strain"], tot, "total_strain")
        assert_same(g["deviatoric_strain"], dev, "deviatoric_strain")
    xi = g["xi"]
    ds = vo.ds_partial(p, dev, xi)
    dk = vo.dsigma_partial(p, tot, xi)
    st = vo.tilde_next(p, "g", s_prev, xi)
    kt = vo.tilde_next(p, "k", k_prev, xi)
    sp, kp = vo.add(ds, st), vo.add(dk, kt)
    assert_same(g["s_tilde"], st, "s_tilde")
    assert_same(g["sigma_tilde"], kt, "sigma_tilde")
    if full:
        assert_same(g["ds_partial"], ds, "ds_partial")
        assert_same(g["dsigma_partial"], dk, "dsigma_partial")
        assert_same(g["s_partial"], sp, "s_partial")
        assert_same(g["sigma_partial"], kp, "sigma_partial")
    assert_same(g["sigma"], vo.sigma_next(p, sp, kp), "sigma")


@pytest.mark.parametrize("d", [1, 2, 3])
@pytest.mark.parametrize("N", [3, 4, 6, 8, 10, 12])
def test_fast_path_all_term_counts(sg_ctx, d, N):
    """The persistent TMA kernel (minimal materialisation, full tiles) for every term count of the Prony sweep; from 8
    terms on the history rows are staged 6 terms at a time (per-lane bulk copies).  Bit-exact given the GPU's exp."""
    n = 32 * 300
    p = vo.ViscoParams(dim=d, dt=0.1, **vo.prony_tables(N))
    T_cur, T_prev, Tfp, s, k = random_visco_state(n, d, N, seed=7 * N + d)
    plan = make_plan(sg_ctx, p)
    t = gpu_state(p, T_cur, T_prev, Tfp, s, k, materialize=False)
    plan.update(n, t)
    torch.cuda.synchronize()
    g = {name: v.cpu().numpy() for name, v in t.items()}
    g["xi"] = t["xi"].cpu().numpy()
    # xi on the GPU comes from the GPU's own phi / phi_next: take it as given, everything downstream must be bit-exact
    check_against_oracle(p, T_cur, T_prev, Tfp, s, k, g, full=False)


@pytest.mark.parametrize("d", [1, 2, 3])
@pytest.mark.parametrize("N", [3, 6, 12])
@pytest.mark.parametrize("n", [1, 31, 32, 33, 4097])
def test_fused_update_matches_oracle(sg_ctx, d, N, n):
    p = vo.ViscoParams(dim=d, dt=0.1, **vo.prony_tables(N))
    T_cur, T_prev, Tfp, s, k = random_visco_state(n, d, N, seed=100 + n)
    if n > 8:
        T_prev[3] = T_cur[3]  # xi = 0 -> NaN stress at this node only (Q5)
    plan = make_plan(sg_ctx, p)
    t = gpu_state(p, T_cur, T_prev, Tfp, s, k)
    plan.update(n, t)
    torch.cuda.synchronize()
    g = {name: v.cpu().numpy() for name, v in t.items()}
    check_against_oracle(p, T_cur, T_prev, Tfp, s, k, g)
    if n > 8:
        sig = g["sigma"].reshape(n, d * d)
        assert np.isnan(sig[3]).all() and np.isfinite(np.delete(sig, 3, axis=0)).all()


@pytest.mark.parametrize("d", [1, 2, 3])
def test_end_to_end_bound_vs_pure_oracle(sg_ctx, d):
    """Several steps, compulsory outputs only; pure-oracle comparison at 1e-10 (T-space) and a
    conditioning-aware bound on the stress (|xi/lambda| >= 1e-6 here)."""
    n, N = 20000, 6
    p = vo.ViscoParams(dim=d, dt=0.1)
    T_cur, T_prev, Tfp, s, k = random_visco_state(n, d, N, seed=7)
    plan = make_plan(sg_ctx, p)
    t = gpu_state(p, T_cur, T_prev, Tfp, s, k, materialize=False)
    oTfp, os_, ok_ = Tfp.copy(), s.copy(), k.copy()
    oTf, oph, oxi, osig = np.zeros(n), np.zeros(n), np.zeros(n), np.zeros(n * d * d)
    oTc, oTp = T_cur.copy(), T_prev.copy()
    for step in range(5):
        plan.update(n, t)
        vo.step_fused(p, oTc, oTp, oTfp, oTf, oph, oxi, os_, ok_, osig)
        torch.cuda.synchronize()
        assert rel_err(t["phi"].cpu().numpy(), oph) < 1e-15
        assert rel_err(t["Tf_partial"].cpu().numpy(), oTfp) < 1e-13
        assert rel_err(t["Tf"].cpu().numpy(), oTf) < 1e-13
        assert np.max(np.abs(t["xi"].cpu().numpy() - oxi) / np.abs(oxi)) < 1e-12
        assert rel_err(t["sigma"].cpu().numpy(), osig) < 1e-10
        assert rel_err(t["s_tilde"].cpu().numpy(), os_) < 1e-10
        # next step: cool by a node-dependent amount (TVP:378 T_prev <- T_cur)
        newT = oTc - np.linspace(0.2, 0.9, n)
        oTp[:], oTc[:] = oTc, newT
        t["T_prev"].copy_(t["T_cur"])
        t["T_cur"].copy_(torch.from_numpy(newT))


def test_golden_kat_on_gpu(sg_ctx):
    """The hand-evaluated golden vectors, one launch per dimension."""
    cases = load_visco_kat()
    for d in (1, 2, 3):
        cs = [c for c in cases if c["d"] == d and c["inputs"]["dt"] == 0.1]
        p = vo.ViscoParams(dim=d, dt=0.1)
        T_cur = np.array([c["inputs"]["T_cur"] for c in cs])
        T_prev = np.array([c["inputs"]["T_prev"] for c in cs])
        Tfp = np.ravel([c["inputs"]["Tfp_prev"] for c in cs])
        s = np.ravel([c["inputs"]["s_tilde"] for c in cs])
        k = np.ravel([c["inputs"]["sigma_tilde"] for c in cs])
        plan = make_plan(sg_ctx, p)
        t = gpu_state(p, T_cur, T_prev, Tfp, s, k)
        plan.update(len(cs), t)
        torch.cuda.synchronize()
        exp = lambda key: np.ravel([c["expected"][key] for c in cs])
        assert rel_err(t["phi"].cpu().numpy(), exp("phi")) < 1e-15
        assert rel_err(t["Tf_partial"].cpu().numpy(), exp("Tf_partial")) < 1e-14
        assert rel_err(t["Tf"].cpu().numpy(), exp("Tf")) < 1e-14
        assert_same(t["T_next"].cpu().numpy(), exp("T_next"), "T_next")
        assert rel_err(t["xi"].cpu().numpy(), exp("xi")) < 1e-12
        assert rel_err(t["sigma"].cpu().numpy(), exp("sigma")) < 1e-9
        assert rel_err(t["s_tilde"].cpu().numpy(), exp("s_tilde")) < 1e-12


def test_phase_masks_compose(sg_ctx):
    """Calling the four reference phases one by one (TVP:370-373) equals the fused call."""
    from fem_glass_tempering_b200 import _lib
    d, N, n = 3, 6, 1000
    p = vo.ViscoParams(dim=d, dt=0.1)
    T_cur, T_prev, Tfp, s, k = random_visco_state(n, d, N, seed=3)
    plan = make_plan(sg_ctx, p)
    a = gpu_state(p, T_cur, T_prev, Tfp, s, k)
    b = gpu_state(p, T_cur, T_prev, Tfp, s, k)
    plan.update(n, a)
    for ph in (_lib.PHASE_TF, _lib.PHASE_STRAIN, _lib.PHASE_SHIFT, _lib.PHASE_STRESS):
        plan.update(n, b, phases=ph)
    torch.cuda.synchronize()
    for name in a:
        assert_same(a[name].cpu().numpy(), b[name].cpu().numpy(), name)


def test_unaligned_and_odd_tiles_use_fallback_path(sg_ctx):
    """History tensors offset by 8 B (not 16-B aligned) and odd element counts: plain-load path."""
    d, N, n = 1, 3, 77
    p = vo.ViscoParams(dim=d, dt=0.1, **vo.prony_tables(N))
    T_cur, T_prev, Tfp, s, k = random_visco_state(n, d, N, seed=5)
    plan = make_plan(sg_ctx, p)
    t = gpu_state(p, T_cur, T_prev, Tfp, s, k)
    pad = torch.zeros(n * N * d * d + 1, dtype=torch.float64, device="cuda:0")
    pad[1:] = t["s_tilde"]
    t["s_tilde"] = pad[1:]
    assert t["s_tilde"].data_ptr() % 16 == 8
    plan.update(n, t)
    torch.cuda.synchronize()
    g = {name: v.cpu().numpy() for name, v in t.items()}
    check_against_oracle(p, T_cur, T_prev, Tfp, s, k, g)


def test_errors_are_raised_not_swallowed(sg_ctx):
    from fem_glass_tempering_b200 import _lib
    p = vo.ViscoParams(dim=3, dt=0.1)
    plan = make_plan(sg_ctx, p)
    with pytest.raises(_lib.SgError):
        plan.update(10, {})  # missing required fields
    with pytest.raises(_lib.SgError):
        _lib.ViscoPlan(sg_ctx, dim=4, dt=0.1, H=1, Rg=1, Tb=1, alpha_solid=0, alpha_liquid=0, m=[1], lambda_m=[1],
                       g=[1], lambda_g=[1], k=[1], lambda_k=[1])


# ---------------------------------------------------------------------------------------------------------------
# Optional corrected scheme (model_params["physics"] = "corrected"): an extension, specified by oracle.vo_step_corrected
@pytest.mark.parametrize("d,N,n", [(3, 6, 2000 + 17), (2, 6, 777), (1, 6, 4096), (3, 4, 96), (3, 6, 9)])
def test_corrected_scheme_matches_its_cpu_statement(sg_ctx, d, N, n):
    """Fast kernel (full tiles) + general kernel (tail): three steps so that the history (= partial stresses) matters.
    exp/expm1 differ between libm and CUDA by an ulp or two: 1e-12 relative on every output."""
    from fem_glass_tempering_b200 import _lib
    p = vo.ViscoParams(dim=d, dt=0.1) if N == 6 else vo.ViscoParams(dim=d, dt=0.1, **{
        k: getattr(vo.ViscoParams(dim=d, dt=0.1), k)[:N] for k in ("m", "lambda_m", "g", "lambda_g", "k", "lambda_k")})
    chi = 0.5
    plan = _lib.ViscoPlan(sg_ctx, dim=p.dim, dt=p.dt, H=p.H, Rg=p.Rg, Tb=p.Tb, alpha_solid=p.alpha_solid,
                          alpha_liquid=p.alpha_liquid, m=p.m, lambda_m=p.lambda_m, g=p.g, lambda_g=p.lambda_g, k=p.k,
                          lambda_k=p.lambda_k, mode=_lib.VISCO_CORRECTED, chi=chi)
    T_cur, T_prev, Tfp, s, k = random_visco_state(n, d, N=p.N)
    T_cur[:3] = T_prev[:3]                       # stationary nodes: the reference scheme gives 0/0 here, this one must not
    Tf = T_prev + 1.0
    t = gpu_state(p, T_cur, T_prev, Tfp, s, k, materialize=False)
    t["Tf"] = torch.from_numpy(Tf.copy()).to("cuda:0")
    o = dict(Tfp=Tfp.copy(), Tf=Tf.copy(), s=s.copy(), k=k.copy(), phi=np.zeros(n), xi=np.zeros(n), sig=np.zeros(n * d * d))
    rng = np.random.default_rng(1)
    for step in range(3):
        plan.update(n, t)
        vo.step_corrected(p, chi, T_cur, T_prev, o["Tfp"], o["Tf"], o["phi"], o["xi"], o["s"], o["k"], o["sig"])
        g = {name: t[name].cpu().numpy() for name in ("Tf_partial", "Tf", "phi", "xi", "s_tilde", "sigma_tilde", "sigma")}
        assert np.isfinite(g["sigma"]).all() and (g["xi"] > 0).all()
        for name, ref in (("Tf_partial", o["Tfp"]), ("Tf", o["Tf"]), ("phi", o["phi"]), ("xi", o["xi"]),
                          ("s_tilde", o["s"]), ("sigma_tilde", o["k"]), ("sigma", o["sig"])):
            assert rel_err(g[name], ref) <= 1e-12, (name, step, rel_err(g[name], ref))
        # next step: cool further
        T_prev = T_cur.copy()
        T_cur = T_cur - rng.uniform(0.05, 1.0, n)
        t["T_cur"].copy_(torch.from_numpy(T_cur))
        t["T_prev"].copy_(torch.from_numpy(T_prev))


def test_corrected_scheme_full_materialisation_and_phase_rule(sg_ctx):
    from fem_glass_tempering_b200 import _lib
    d, n = 2, 130
    p = vo.ViscoParams(dim=d, dt=0.1)
    plan = _lib.ViscoPlan(sg_ctx, dim=p.dim, dt=p.dt, H=p.H, Rg=p.Rg, Tb=p.Tb, alpha_solid=p.alpha_solid,
                          alpha_liquid=p.alpha_liquid, m=p.m, lambda_m=p.lambda_m, g=p.g, lambda_g=p.lambda_g, k=p.k,
                          lambda_k=p.lambda_k, mode=_lib.VISCO_CORRECTED, chi=0.5)
    T_cur, T_prev, Tfp, s, k = random_visco_state(n, d)
    Tf = T_prev + 0.5
    t = gpu_state(p, T_cur, T_prev, Tfp, s, k, materialize=True)
    t["Tf"] = torch.from_numpy(Tf.copy()).to("cuda:0")
    plan.update(n, t)
    o = dict(Tfp=Tfp.copy(), Tf=Tf.copy(), s=s.copy(), k=k.copy(), phi=np.zeros(n), xi=np.zeros(n), sig=np.zeros(n * d * d))
    vo.step_corrected(p, 0.5, T_cur, T_prev, o["Tfp"], o["Tf"], o["phi"], o["xi"], o["s"], o["k"], o["sig"])
    assert rel_err(t["sigma"].cpu().numpy(), o["sig"]) <= 1e-12
    assert_same(t["s_partial"].cpu().numpy(), t["s_tilde"].cpu().numpy(), "history == partial stress")
    assert_same(t["sigma_partial"].cpu().numpy(), t["sigma_tilde"].cpu().numpy(), "history == partial stress")
    eth = t["thermal_strain"].cpu().numpy().reshape(n, d, d)[:, 0, 0]
    expect = p.alpha_solid * (T_cur - T_prev) + (p.alpha_liquid - p.alpha_solid) * (o["Tf"] - Tf)
    assert np.max(np.abs(eth - expect)) <= 1e-15
    with pytest.raises(_lib.SgError):
        plan.update(n, t, _lib.PHASE_TF)            # the corrected scheme only runs as the fused update
