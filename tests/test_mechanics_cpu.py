"""CPU: the mechanical-equilibrium oracle (oracle/mechanics_oracle.py — test infrastructure for the extension of SURVEY
§8(f) row 4; the reference has no counterpart, VM:135-139) against the hand-evaluated known answers of
tests/golden/mech_kat.json and against properties that hold for any mesh."""
import json
import os

import numpy as np
import pytest

from fem_glass_tempering_b200 import fe, mechanics
from fem_glass_tempering_b200 import mesh as M
from oracle import mechanics_oracle as mo

from helpers import GOLDEN, unhex


def load_mech_kat():
    with open(os.path.join(GOLDEN, "mech_kat.json")) as fh:
        return {c["name"]: c for c in json.load(fh)["cases"]}


MESHES = {1: lambda: M.interval_mesh(9, 3.0), 2: lambda: M.perturb_interior(M.rectangle_mesh(5, 4, 5.0, 4.0), 0.15, seed=3),
          3: lambda: M.perturb_interior(M.box_mesh(3, 3, 2, 3.0, 3.0, 2.0), 0.12, seed=5)}


def make_oracle(mesh, family, degree):
    S = fe.ScalarSpace(mesh, family, degree)
    w = mo.sigma_weights(S.element.nodes, degree, mesh.dim)
    return S, mo.MechanicsOracle(mesh.x, mesh.cells, S.dofmap, w, mo.symmetry_planes(mesh.x))


def test_bar_closed_form():
    c = load_mech_kat()["bar"]
    xs = np.array(unhex(c["x"]))
    mesh = M.line_mesh(xs)
    S, O = make_oracle(mesh, "DG", 1)
    K, G, s0 = (np.array(unhex(c[k])) for k in ("K", "G", "sigma0"))
    du = O.solve(G, K, s0)
    np.testing.assert_allclose(du, unhex(c["u"]), rtol=0, atol=1e-15)
    sig, eps = O.correct(du, G, K, s0)
    np.testing.assert_allclose(eps, np.repeat(unhex(c["eps"]), 2), rtol=1e-13, atol=0)
    np.testing.assert_allclose(sig, unhex(c["sigma"]), rtol=0, atol=1e-15)


@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("space", [("DG", 1), ("CG", 1), ("CG", 2)])
def test_patch_test(dim, space):
    c = load_mech_kat()[f"patch{dim}"]
    mesh = MESHES[dim]()
    S, O = make_oracle(mesh, *space)
    E = np.array(unhex(c["E"]))
    G = np.full(S.n_nodes, float.fromhex(c["G"]))
    K = np.full(S.n_nodes, float.fromhex(c["K"]))
    s0 = np.tile(np.array(unhex(c["sigma0"])).ravel(), S.n_nodes)
    du = O.solve(G, K, s0)
    exact = (mesh.x - mesh.x.min(axis=0)) * E
    np.testing.assert_allclose(du.reshape(-1, dim), exact, rtol=0, atol=1e-15)
    sig, eps = O.correct(du, G, K, s0)
    assert np.abs(sig).max() < 1e-13
    np.testing.assert_allclose(eps.reshape(-1, dim, dim), np.broadcast_to(np.diag(E), (S.n_nodes, dim, dim)), atol=1e-15)


@pytest.mark.parametrize("dim", [1, 2, 3])
def test_tangent_is_symmetric_positive_definite_and_rigid_motions_are_in_its_kernel(dim):
    mesh = MESHES[dim]()
    S, O = make_oracle(mesh, "DG", 1)
    rng = np.random.default_rng(dim)
    G, K = rng.uniform(20, 30, S.n_nodes), rng.uniform(30, 45, S.n_nodes)
    A = O.stiffness(G, K)
    assert abs(A - A.T).max() < 1e-12 * abs(A).max()
    # translations (and, d >= 2, an infinitesimal rotation) produce no strain, hence no force
    for i in range(dim):
        t = np.zeros((mesh.n_vertices, dim))
        t[:, i] = 1.0
        assert np.abs(A @ t.ravel()).max() < 1e-11
    if dim >= 2:
        r = np.zeros((mesh.n_vertices, dim))
        r[:, 0], r[:, 1] = -mesh.x[:, 1], mesh.x[:, 0]
        assert np.abs(A @ r.ravel()).max() < 1e-10
    Ac, _ = O.constrained(A, np.zeros(A.shape[0]))
    assert np.linalg.eigvalsh(Ac.toarray()).min() > 0.0


@pytest.mark.parametrize("dim", [1, 2, 3])
def test_corrected_dg_stress_is_in_discrete_equilibrium(dim):
    """For a DG sigma space every node takes its own cell's strain, so B^T sigma_h vanishes on the free components."""
    mesh = MESHES[dim]()
    S, O = make_oracle(mesh, "DG", 1)
    rng = np.random.default_rng(10 + dim)
    G, K = rng.uniform(20, 30, S.n_nodes), rng.uniform(30, 45, S.n_nodes)
    s0 = (-(K * dim * rng.uniform(1e-4, 1e-3, S.n_nodes)))[:, None, None] * np.eye(dim)
    du = O.solve(G, K, s0.ravel())
    sig, _ = O.correct(du, G, K, s0.ravel())
    scale = np.abs(O.rhs(s0.ravel())).max()
    assert np.abs(O.residual(sig)).max() < 1e-12 * scale
    assert np.abs(O.residual(s0.ravel())).max() > 1e-3 * scale    # the restrained state is NOT in equilibrium


def test_tangent_moduli_follow_the_chain():
    """G_eff/K_eff are the per-term factors of VM:176-191 summed: analytically sum_n g_n (1 - xi/(2 lambda_n))."""
    from oracle import visco_oracle as vo
    xi = np.array([1e-3, -2e-3, 5e-2, 0.0])
    G, K = mo.tangent_moduli(xi, vo.PRONY_G, vo.PRONY_LAMBDA_G, vo.PRONY_K, vo.PRONY_LAMBDA_K)
    for j, x in enumerate(xi):
        Ga = sum(g * (1.0 - x / (2.0 * l)) for g, l in zip(vo.PRONY_G, vo.PRONY_LAMBDA_G))
        Ka = sum(k * (1.0 - x / (2.0 * l)) for k, l in zip(vo.PRONY_K, vo.PRONY_LAMBDA_K))
        assert abs(G[j] - Ga) < 1e-9 * abs(Ga) and abs(K[j] - Ka) < 1e-9 * abs(Ka)
    Gc, Kc = mo.tangent_moduli(xi, vo.PRONY_G, vo.PRONY_LAMBDA_G, vo.PRONY_K, vo.PRONY_LAMBDA_K, mode="corrected")
    x = 5e-2
    Ga = sum(g * (1.0 - np.exp(-x / l)) / (x / l) for g, l in zip(vo.PRONY_G, vo.PRONY_LAMBDA_G))
    assert abs(Gc[2] - Ga) < 1e-13 * Ga and abs(Gc[3] - sum(vo.PRONY_G)) < 1e-13


def test_package_helpers_agree_with_the_oracle():
    """mechanics.py's host-side tables (cell weights, winner cells, symmetry planes) equal the oracle's."""
    for dim in (1, 2, 3):
        mesh = MESHES[dim]()
        for fam, deg in (("DG", 1), ("CG", 1), ("CG", 2)):
            S, O = make_oracle(mesh, fam, deg)
            np.testing.assert_allclose(mechanics.sigma_cell_weights(dim, deg), O.w, atol=1e-14)
            assert np.array_equal(mechanics.winner_cells(S), O.winner)
        assert np.array_equal(mechanics.symmetry_planes(mesh.x).ravel(), mo.symmetry_planes(mesh.x))
