"""Lagrange P1/P2 finite-element tables on simplices, quadrature, dof maps and mesh topology.

Host-side (numpy) counterpart of what the reference gets from basix/FFCx/dolfinx
(ThermoViscoProblem.py:61-103: FiniteElement/VectorElement/TensorElement on CG or DG).
Everything here is set-up: the arrays are built once, uploaded, and consumed by the CUDA
operator (csrc/thermal.cu).

Conventions (internal, documented in DESIGN.md):
  * reference simplex: v0 = 0, v_i = e_i; barycentric l_0 = 1 - sum(x), l_i = x_i
  * local facet f is the facet OPPOSITE local vertex f (its outward normal is -grad l_f/|grad l_f|)
  * P2 local dof order: vertices, then edges in the order of `ref_edges(dim)`
  * '+' side of an interior facet = the cell with the lower index
"""
from __future__ import annotations

import functools
import itertools
import math
from dataclasses import dataclass

import numpy as np

from .mesh import Mesh

# --------------------------------------------------------------------------- reference topology


def ref_edges(dim: int):
    return {1: [(0, 1)], 2: [(1, 2), (0, 2), (0, 1)],
            3: [(2, 3), (1, 3), (1, 2), (0, 3), (0, 2), (0, 1)]}[dim]


def ref_facets(dim: int):
    """facet f = all local vertices except f, ascending."""
    return [tuple(v for v in range(dim + 1) if v != f) for f in range(dim + 1)]


def ref_vertices(dim: int) -> np.ndarray:
    v = np.zeros((dim + 1, dim))
    for i in range(dim):
        v[i + 1, i] = 1.0
    return v


# --------------------------------------------------------------------------- quadrature


def gauss_legendre_01(n: int):
    x, w = np.polynomial.legendre.leggauss(n)
    return 0.5 * (x + 1.0), 0.5 * w


def simplex_quadrature(dim: int, degree: int):
    """Collapsed Gauss-Jacobi rule exact for polynomials of total degree <= `degree` on the reference
    simplex.  Returns (points [nq, dim], weights [nq]) with sum(weights) = 1/dim!."""
    from scipy.special import roots_jacobi
    n = degree // 2 + 1
    if dim == 0:
        return np.zeros((1, 0)), np.ones(1)
    if dim == 1:
        x, w = gauss_legendre_01(n)
        return x[:, None], w
    x0, w0 = gauss_legendre_01(n)
    x1, w1 = roots_jacobi(n, 1, 0)
    x1, w1 = 0.5 * (x1 + 1.0), w1 / 4.0
    if dim == 2:
        pts = [(a1, a0 * (1.0 - a1)) for a1 in x1 for a0 in x0]
        wts = [b1 * b0 for b1 in w1 for b0 in w0]
        pts, wts = np.array(pts), np.array(wts)
        return pts[:, ::-1].copy(), wts
    x2, w2 = roots_jacobi(n, 2, 0)
    x2, w2 = 0.5 * (x2 + 1.0), w2 / 8.0
    pts, wts = [], []
    for a2, b2 in zip(x2, w2):
        for a1, b1 in zip(x1, w1):
            for a0, b0 in zip(x0, w0):
                pts.append((a0 * (1 - a1) * (1 - a2), a1 * (1 - a2), a2))
                wts.append(b0 * b1 * b2)
    return np.array(pts), np.array(wts)


def minimal_cell_rule(dim: int, degree: int):
    """Fewest-point rules for the stiffness integrand: centroid (degree <= 1), and the classical
    degree-2 rules with dim+1 points; otherwise the collapsed rule."""
    if degree <= 1:
        return np.full((1, dim), 1.0 / (dim + 1)), np.array([1.0 / math.factorial(dim)])
    if degree == 2 and dim == 2:
        a = 1.0 / 6.0
        return np.array([[a, a], [1 - 2 * a, a], [a, 1 - 2 * a]]), np.full(3, 1.0 / 6.0)
    if degree == 2 and dim == 3:
        a, b = (5.0 - math.sqrt(5.0)) / 20.0, (5.0 + 3.0 * math.sqrt(5.0)) / 20.0
        return np.array([[a, a, a], [b, a, a], [a, b, a], [a, a, b]]), np.full(4, 1.0 / 24.0)
    return simplex_quadrature(dim, degree)


def symmetric_facet_rule(fdim: int, degree: int):
    """Rule on the reference facet simplex that is invariant under vertex permutations, in facet
    BARYCENTRIC coordinates [nq, fdim+1]; weights sum to 1.  Needed on interior facets, where the two
    cells see the facet through different vertex orders (FFCx's `quadrature_permutation`)."""
    if fdim == 0:
        return np.ones((1, 1)), np.ones(1)
    if fdim == 1:
        x, w = gauss_legendre_01(degree // 2 + 1)
        return np.stack([1.0 - x, x], axis=1), w / w.sum()
    if fdim == 2:
        if degree <= 2:
            a = 1.0 / 6.0
            orbits = [(a, 1.0 / 3.0)]
        elif degree <= 4:  # Dunavant 6-point rule
            orbits = [(0.44594849091596488631832925388305199, 0.22338158967801146569500700843312280),
                      (0.09157621350977074345957146340220151, 0.10995174365532186763832632490021053)]
        else:
            raise NotImplementedError("symmetric triangle rule of degree > 4")
        pts, wts = [], []
        for a, w in orbits:
            b = 1.0 - 2.0 * a
            pts += [(a, a, b), (a, b, a), (b, a, a)]
            wts += [w] * 3
        return np.array(pts), np.array(wts)
    raise ValueError(fdim)


def facet_permutation_table(fdim: int, bary: np.ndarray):
    """perms: list of all vertex permutations s of the facet; table[p][q] = index q' of the rule point
    whose barycentric coordinates are bary[q][s] (i.e. mu'_k = mu_{s(k)})."""
    perms = list(itertools.permutations(range(fdim + 1)))
    table = np.zeros((len(perms), bary.shape[0]), dtype=np.int32)
    for p, s in enumerate(perms):
        for q in range(bary.shape[0]):
            target = bary[q][list(s)]
            d = np.abs(bary - target).sum(axis=1)
            j = int(np.argmin(d))
            assert d[j] < 1e-12, "facet rule is not symmetric"
            table[p, q] = j
    return perms, table


# --------------------------------------------------------------------------- the element


class LagrangeElement:
    """Scalar Lagrange element of degree 1 or 2 on the reference simplex."""

    def __init__(self, dim: int, degree: int):
        if degree not in (1, 2):
            raise NotImplementedError("Lagrange degree must be 1 or 2")
        self.dim, self.degree = dim, degree
        self.edges = ref_edges(dim)
        self.n_ld = dim + 1 if degree == 1 else dim + 1 + len(self.edges)
        v = ref_vertices(dim)
        nodes = [v[i] for i in range(dim + 1)]
        if degree == 2:
            nodes += [0.5 * (v[a] + v[b]) for a, b in self.edges]
        self.nodes = np.array(nodes)  # == basix interpolation points of the element

    def interpolation_points(self) -> np.ndarray:
        return self.nodes

    def tabulate(self, pts: np.ndarray):
        """values [npts, n_ld], reference gradients [npts, dim, n_ld]."""
        pts = np.atleast_2d(np.asarray(pts, dtype=np.float64))
        npts, d = pts.shape[0], self.dim
        lam = np.empty((npts, d + 1))
        lam[:, 0] = 1.0 - pts.sum(axis=1)
        lam[:, 1:] = pts
        dlam = np.zeros((d + 1, d))
        dlam[0, :] = -1.0
        for i in range(d):
            dlam[i + 1, i] = 1.0
        vals = np.empty((npts, self.n_ld))
        grads = np.empty((npts, d, self.n_ld))
        if self.degree == 1:
            vals[:] = lam
            grads[:] = dlam.T[None, :, :]
            return vals, grads
        for i in range(d + 1):
            vals[:, i] = lam[:, i] * (2.0 * lam[:, i] - 1.0)
            grads[:, :, i] = (4.0 * lam[:, i] - 1.0)[:, None] * dlam[i][None, :]
        for e, (a, b) in enumerate(self.edges):
            j = d + 1 + e
            vals[:, j] = 4.0 * lam[:, a] * lam[:, b]
            grads[:, :, j] = 4.0 * (lam[:, a][:, None] * dlam[b][None, :] + lam[:, b][:, None] * dlam[a][None, :])
        return vals, grads

    def facet_points(self, f: int, bary: np.ndarray) -> np.ndarray:
        """Reference-cell coordinates of facet-barycentric points on local facet f."""
        fv = ref_facets(self.dim)[f]
        v = ref_vertices(self.dim)
        return bary @ v[list(fv)]

    def facet_dofs(self, f: int):
        """Local dofs lying on facet f (vertices of the facet, then its edges)."""
        fv = ref_facets(self.dim)[f]
        dofs = list(fv)
        if self.degree == 2:
            for e, (a, b) in enumerate(self.edges):
                if a in fv and b in fv:
                    dofs.append(self.dim + 1 + e)
        return dofs


@dataclass
class OperatorTables:
    """Reference tables the CUDA thermal operator needs for one (dim, degree)."""
    n_ld: int
    mass: np.ndarray       # [n_ld, n_ld]  int phi_i phi_j over the reference cell
    load: np.ndarray       # [n_ld]        int phi_i
    cq_w: np.ndarray       # [nqc]         stiffness rule weights (sum 1/d!)
    cq_grad: np.ndarray    # [nqc, dim, n_ld]
    # interior facets (DG): symmetric rule, weights sum to 1
    fq_w: np.ndarray       # [nqf]
    fq_val: np.ndarray     # [dim+1, nqf, n_ld]
    fq_grad: np.ndarray    # [dim+1, nqf, dim, n_ld]
    fq_perm: np.ndarray    # [n_perms, nqf] int32
    # exterior facets (Robin + radiation): high-degree rule, weights sum to 1
    bq_w: np.ndarray       # [nqb]
    bq_val: np.ndarray     # [dim+1, nqb, n_ld]


@functools.lru_cache(maxsize=None)
def operator_tables(dim: int, degree: int) -> OperatorTables:
    el = LagrangeElement(dim, degree)
    # exact mass / load (degree 2p)
    pts, wts = simplex_quadrature(dim, 2 * degree)
    v, _ = el.tabulate(pts)
    mass = np.einsum("q,qi,qj->ij", wts, v, v)
    load = np.einsum("q,qi->i", wts, v)
    # stiffness rule (degree 2p-2) with the minimal number of points
    cpts, cw = minimal_cell_rule(dim, max(2 * degree - 2, 0))
    _, cg = el.tabulate(cpts)
    # interior-facet symmetric rule (degree 2p)
    fb, fw = symmetric_facet_rule(dim - 1, 2 * degree)
    _, ptab = facet_permutation_table(dim - 1, fb)
    fval = np.empty((dim + 1, fb.shape[0], el.n_ld))
    fgrad = np.empty((dim + 1, fb.shape[0], dim, el.n_ld))
    for f in range(dim + 1):
        fval[f], fgrad[f] = el.tabulate(el.facet_points(f, fb))
    # exterior-facet rule: radiation integrand T^4 v has degree 5p
    bpts, bw = simplex_quadrature(dim - 1, 5 * degree)
    bw = bw / bw.sum()
    bb = np.concatenate([1.0 - bpts.sum(axis=1, keepdims=True), bpts], axis=1) if dim > 1 else np.ones((1, 1))
    bval = np.empty((dim + 1, bb.shape[0], el.n_ld))
    for f in range(dim + 1):
        bval[f], _ = el.tabulate(el.facet_points(f, bb))
    return OperatorTables(el.n_ld, mass, load, cw, cg, fw, fval, fgrad, ptab, bw, bval)


# --------------------------------------------------------------------------- mesh topology


def _facet_keys(mesh: Mesh):
    """Sorted global vertex tuples of every (cell, local facet): [n_cells*(d+1), d]."""
    d = mesh.dim
    fac = ref_facets(d)
    fv = np.stack([mesh.cells[:, list(f)] for f in fac], axis=1).astype(np.int64)  # [nc, d+1, d]
    return np.sort(fv.reshape(-1, d), axis=1)


@dataclass
class FacetTopology:
    neighbor: np.ndarray      # [n_cells, d+1] int32 neighbour cell across local facet f, -1 on the boundary
    nb_facet: np.ndarray      # [n_cells, d+1] int8  the neighbour's local facet index
    nb_perm: np.ndarray       # [n_cells, d+1] int8  vertex-permutation id (index into itertools.permutations)
    bnd_cell: np.ndarray      # [n_bf] int32 exterior facets: owning cell
    bnd_facet: np.ndarray     # [n_bf] int32 local facet index


def _setup_device(device=None):
    """Where the O(n_cells) set-up arithmetic runs: the current CUDA device when there is one (torch is buffer/array
    plumbing here, the results are identical to the numpy statements kept below for the CPU tests), else the host."""
    import torch
    if device is not None:
        return torch.device(device)
    return torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")


def facet_topology(mesh: Mesh, exterior_mask=None, device=None) -> FacetTopology:
    """Facet -> cell connectivity by sorting facet vertex tuples (one packed 64-bit key per facet, stable sort).

    exterior_mask(facet_midpoints) -> bool may be given for a partitioned mesh to tell true domain
    boundary facets from partition cuts (facets seen once locally that are interior globally).
    Runs on `device` (default: the GPU when present — 20 M facets sort in milliseconds instead of seconds)."""
    import torch
    d, nc = mesh.dim, mesh.n_cells
    nv = int(mesh.n_vertices)
    if not nv ** d < 2 ** 62:
        return _facet_topology_np(mesh, exterior_mask)
    dev = _setup_device(device)
    cells = torch.from_numpy(mesh.cells).to(dev).long()
    fac = torch.tensor(ref_facets(d), dtype=torch.long, device=dev)             # [d+1, d]
    nf = nc * (d + 1)
    keys = torch.sort(cells[:, fac].reshape(nf, d), dim=1).values
    packed = keys[:, 0].clone()
    for c in range(1, d):
        packed = packed * nv + keys[:, c]
    del keys
    sp, order = torch.sort(packed, stable=True)
    del packed
    same = sp[1:] == sp[:-1] if nf > 1 else torch.zeros(0, dtype=torch.bool, device=dev)
    del sp
    neighbor = torch.full((nf,), -1, dtype=torch.int32, device=dev)
    nb_facet = torch.zeros(nf, dtype=torch.int8, device=dev)
    a, b = order[:-1][same], order[1:][same]          # the two views of every interior facet (a < b: stable sort)
    del order, same
    ca, cb, fa, fb = a // (d + 1), b // (d + 1), a % (d + 1), b % (d + 1)
    neighbor[a], neighbor[b] = cb.to(torch.int32), ca.to(torch.int32)
    nb_facet[a], nb_facet[b] = fb.to(torch.int8), fa.to(torch.int8)
    # vertex permutation between the two cells' views of the facet: for view i with neighbour view j,
    # s(k') = position in gv_i of gv_j[k']; the permutation of view b is the inverse of the one of view a
    nb_perm = torch.zeros(nf, dtype=torch.int8, device=dev)
    if d > 1 and a.numel():
        perms = list(itertools.permutations(range(d)))
        gv_a = cells[ca[:, None], fac[fa]]                                        # [npair, d]
        gv_b = cells[cb[:, None], fac[fb]]
        s = torch.argmax((gv_b[:, :, None] == gv_a[:, None, :]).to(torch.int8), dim=2)
        code = torch.zeros(a.numel(), dtype=torch.long, device=dev)
        for c in range(d):
            code = code * d + s[:, c]
        lut = np.full(d ** d, -1, dtype=np.int8)
        inverse = np.zeros(len(perms), dtype=np.int8)
        for pid, p in enumerate(perms):
            cc = 0
            for c in range(d):
                cc = cc * d + p[c]
            lut[cc] = pid
            inverse[pid] = perms.index(tuple(int(k) for k in np.argsort(p)))
        pid_a = torch.from_numpy(lut).to(dev)[code]
        assert bool((pid_a >= 0).all())
        nb_perm[a] = pid_a
        nb_perm[b] = torch.from_numpy(inverse).to(dev)[pid_a.long()]
    bnd = torch.nonzero(neighbor < 0).ravel().cpu().numpy()
    if exterior_mask is not None and bnd.size:
        facn = np.array(ref_facets(d), dtype=np.int64)
        ci, fi = bnd // (d + 1), bnd % (d + 1)
        mid = mesh.x[mesh.cells[ci[:, None], facn[fi]]].mean(axis=1)
        bnd = bnd[exterior_mask(mid)]
    return FacetTopology(neighbor.reshape(nc, d + 1).cpu().numpy(), nb_facet.reshape(nc, d + 1).cpu().numpy(),
                         nb_perm.reshape(nc, d + 1).cpu().numpy(),
                         (bnd // (d + 1)).astype(np.int32), (bnd % (d + 1)).astype(np.int32))


def _facet_topology_np(mesh: Mesh, exterior_mask=None) -> FacetTopology:
    """numpy statement of facet_topology (fallback for vertex counts whose packed key overflows; the CPU tests hold the
    torch version to it).

    exterior_mask(facet_midpoints) -> bool may be given for a partitioned mesh to tell true domain
    boundary facets from partition cuts (facets seen once locally that are interior globally)."""
    d, nc = mesh.dim, mesh.n_cells
    keys = _facet_keys(mesh)
    nf = keys.shape[0]
    nv = int(mesh.n_vertices)
    if nv ** d < 2 ** 62:
        # one 64-bit key per facet and a stable sort: same order as the lexicographic sort, several times faster
        packed = keys[:, 0].copy()
        for c in range(1, d):
            packed = packed * nv + keys[:, c]
        order = np.argsort(packed, kind="stable")
        sp = packed[order]
        same = sp[1:] == sp[:-1] if nf > 1 else np.zeros(0, dtype=bool)
    else:
        order = np.lexsort(tuple(keys[:, c] for c in range(d - 1, -1, -1)))
        sk = keys[order]
        same = np.all(sk[1:] == sk[:-1], axis=1) if nf > 1 else np.zeros(0, dtype=bool)
    neighbor = np.full(nf, -1, dtype=np.int32)
    nb_facet = np.zeros(nf, dtype=np.int8)
    a, b = order[:-1][same], order[1:][same]          # the two views of every interior facet (a < b: stable sort)
    neighbor[a], neighbor[b] = (b // (d + 1)).astype(np.int32), (a // (d + 1)).astype(np.int32)
    nb_facet[a], nb_facet[b] = (b % (d + 1)).astype(np.int8), (a % (d + 1)).astype(np.int8)
    # vertex permutation between the two cells' views of the facet: for view i with neighbour view j,
    # s(k') = position in gv_i of gv_j[k']; the permutation of view b is the inverse of the one of view a
    fac = np.array(ref_facets(d), dtype=np.int64)  # [d+1, d]
    nb_perm = np.zeros(nf, dtype=np.int8)
    if d > 1 and a.size:
        perms = list(itertools.permutations(range(d)))
        gv_a = mesh.cells[(a // (d + 1))[:, None], fac[a % (d + 1)]]                              # [npair, d]
        gv_b = mesh.cells[(b // (d + 1))[:, None], fac[b % (d + 1)]]
        s = np.argmax(gv_b[:, :, None] == gv_a[:, None, :], axis=2)                               # [npair, d]
        code = np.zeros(a.size, dtype=np.int64)
        for c in range(d):
            code = code * d + s[:, c]
        lut = np.full(d ** d, -1, dtype=np.int8)
        inverse = np.zeros(len(perms), dtype=np.int8)
        for pid, p in enumerate(perms):
            cc = 0
            for c in range(d):
                cc = cc * d + p[c]
            lut[cc] = pid
            inverse[pid] = perms.index(tuple(int(k) for k in np.argsort(p)))
        pid_a = lut[code]
        assert (pid_a >= 0).all()
        nb_perm[a] = pid_a
        nb_perm[b] = inverse[pid_a]
    bnd = np.nonzero(neighbor < 0)[0]
    if exterior_mask is not None and bnd.size:
        ci, fi = bnd // (d + 1), bnd % (d + 1)
        mid = mesh.x[mesh.cells[ci[:, None], fac[fi]]].mean(axis=1)
        bnd = bnd[exterior_mask(mid)]
    return FacetTopology(neighbor.reshape(nc, d + 1), nb_facet.reshape(nc, d + 1), nb_perm.reshape(nc, d + 1),
                         (bnd // (d + 1)).astype(np.int32), (bnd % (d + 1)).astype(np.int32))


@dataclass
class CellGeometry:
    detJ: np.ndarray    # [n_cells]  |det J|
    Jinv: np.ndarray    # [n_cells, d, d]  Jinv[a, c] = d xi_a / d x_c
    h: np.ndarray       # [n_cells]  CellDiameter: max vertex distance (TVP:314)


def cell_geometry_device(mesh: Mesh, device=None):
    """(|det J| [nc], J^-1 [nc, d, d], CellDiameter h [nc]) as float64 torch tensors on `device` (see _setup_device):
    closed-form determinants and cofactor inverses, one fused pass instead of 5 M LAPACK calls."""
    import torch
    dev = _setup_device(device)
    d = mesh.dim
    xv = torch.from_numpy(mesh.x).to(dev)[torch.from_numpy(mesh.cells).to(dev).long()]    # [nc, d+1, d]
    J = (xv[:, 1:, :] - xv[:, :1, :]).transpose(1, 2)                                    # J[c, a] = d x_c / d xi_a
    if d == 1:
        det = J[:, 0, 0]
        Jinv = 1.0 / J
    elif d == 2:
        a, b, c, e = J[:, 0, 0], J[:, 0, 1], J[:, 1, 0], J[:, 1, 1]
        det = a * e - b * c
        Jinv = torch.stack([torch.stack([e, -b], 1), torch.stack([-c, a], 1)], 1) / det[:, None, None]
    else:
        a, b, c = J[:, 0, 0], J[:, 0, 1], J[:, 0, 2]
        e, f, g = J[:, 1, 0], J[:, 1, 1], J[:, 1, 2]
        p, q, r = J[:, 2, 0], J[:, 2, 1], J[:, 2, 2]
        c00, c01, c02 = f * r - g * q, g * p - e * r, e * q - f * p
        det = a * c00 + b * c01 + c * c02
        adj = torch.stack([torch.stack([c00, c * q - b * r, b * g - c * f], 1),
                           torch.stack([c01, a * r - c * p, c * e - a * g], 1),
                           torch.stack([c02, b * p - a * q, a * f - b * e], 1)], 1)
        Jinv = adj / det[:, None, None]
    h = torch.zeros(mesh.n_cells, dtype=torch.float64, device=dev)
    for i, j in itertools.combinations(range(d + 1), 2):
        h = torch.maximum(h, torch.sqrt(((xv[:, i] - xv[:, j]) ** 2).sum(dim=1)))
    return det.abs(), Jinv.contiguous(), h


def cell_geometry(mesh: Mesh, device=None) -> CellGeometry:
    det, Jinv, h = cell_geometry_device(mesh, device)
    return CellGeometry(det.cpu().numpy(), np.ascontiguousarray(Jinv.cpu().numpy()), h.cpu().numpy())


def _cell_geometry_np(mesh: Mesh) -> CellGeometry:
    """LAPACK statement of cell_geometry; the CPU tests hold the closed-form version to it."""
    d = mesh.dim
    xv = mesh.x[mesh.cells]                       # [nc, d+1, d]
    J = np.transpose(xv[:, 1:, :] - xv[:, :1, :], (0, 2, 1))   # J[c, a] = d x_c / d xi_a
    if d == 1:
        det = J[:, 0, 0]
        Jinv = 1.0 / J
    else:
        det = np.linalg.det(J)
        Jinv = np.linalg.inv(J)
    h = np.zeros(mesh.n_cells)
    for a, b in itertools.combinations(range(d + 1), 2):
        h = np.maximum(h, np.linalg.norm(xv[:, a] - xv[:, b], axis=1))
    return CellGeometry(np.abs(det), np.ascontiguousarray(Jinv), h)


def facet_measures(mesh: Mesh, geo: CellGeometry, cell: np.ndarray, facet: np.ndarray) -> np.ndarray:
    """|F| = detJ * |grad l_f| / (d-1)!  for the given (cell, local facet) pairs."""
    d = mesh.dim
    dlam = np.zeros((d + 1, d))
    dlam[0, :] = -1.0
    for i in range(d):
        dlam[i + 1, i] = 1.0
    g = np.einsum("nac,na->nc", geo.Jinv[cell], dlam[facet])
    return geo.detJ[cell] * np.linalg.norm(g, axis=1) / math.factorial(d - 1)


# --------------------------------------------------------------------------- function spaces


class ScalarSpace:
    """CG or DG Lagrange space of degree 1/2: the node set and the cell->node map.

    Every reference space (T, Tf_partial, sigma, sigma_partial; TVP:77-101) is this node set times a
    block size."""

    def __init__(self, mesh: Mesh, family: str, degree: int):
        assert family in ("CG", "DG"), "Only CG and DG elements are supported"   # TVP:70-71
        self.mesh, self.family, self.degree = mesh, family, degree
        self.element = LagrangeElement(mesh.dim, degree)
        self.n_ld = self.element.n_ld
        nc = mesh.n_cells
        if family == "DG":
            self.dofmap = np.arange(nc * self.n_ld, dtype=np.int32).reshape(nc, self.n_ld)
            self.n_nodes = nc * self.n_ld
        elif degree == 1:
            self.dofmap = mesh.cells.copy()
            self.n_nodes = mesh.n_vertices
        else:
            self.dofmap, self.n_nodes = _p2_dofmap(mesh)
        self.dofmap = np.ascontiguousarray(self.dofmap, dtype=np.int32)

    def key(self):
        return (self.family, self.degree)

    def tabulate_dof_coordinates(self) -> np.ndarray:
        xv = self.mesh.x[self.mesh.cells]                            # [nc, d+1, d]
        lam = np.concatenate([1.0 - self.element.nodes.sum(axis=1, keepdims=True), self.element.nodes], axis=1)
        xc = np.einsum("lv,cvd->cld", lam, xv)                       # [nc, n_ld, d]
        out = np.empty((self.n_nodes, self.mesh.dim))
        out[self.dofmap.ravel()] = xc.reshape(-1, self.mesh.dim)
        return out


def _p2_dofmap(mesh: Mesh):
    """P2 numbering: on a lattice mesh every P2 node is a point of the half-step lattice, so its id is
    pure arithmetic; otherwise edges are numbered with np.unique.

    Lattice order (d >= 2): x half-planes slowest (what the slab partition of distributed.py relies on); inside a
    plane the LONGEST remaining axis is the run axis, and along it the even half-steps come before the odd ones
    (id = ... + parity * n_even + index // 2).  Consecutive ids therefore share the node type (vertex / which edge
    midpoint) over a whole run, while `column id - row id` still only depends on the node type and the lattice
    offset - the two properties the row-stencil form of the CG apply (csrc/stencil.cu) needs: one class per warp
    and contiguous loads."""
    d, nc = mesh.dim, mesh.n_cells
    edges = ref_edges(d)
    lat = mesh.lattice
    if lat is not None and lat["kind"] in ("box", "line") and _vertex_lattice_coords(mesh) is not None:
        import torch
        dev = _setup_device()
        vc_np = _vertex_lattice_coords(mesh)             # [nv, d] local integer coordinates
        dims2 = 2 * vc_np.max(axis=0) + 1
        cv = torch.from_numpy(vc_np).to(dev)[torch.from_numpy(mesh.cells).to(dev).long()]     # [nc, d+1, d]
        loc = torch.stack([2 * cv[:, i] for i in range(d + 1)] + [cv[:, a] + cv[:, b] for a, b in edges], dim=1)
        if d == 1:                                       # [nc, n_ld, d] doubled-lattice coordinates
            return loc[:, :, 0].to(torch.int32).cpu().numpy(), int(dims2[0])
        run = 1 + int(np.argmax(dims2[1:]))              # run axis: the longest of the in-plane axes
        n_run = int(dims2[run])
        ids = (loc[:, :, run] & 1) * ((n_run + 1) // 2) + (loc[:, :, run] >> 1)
        stride = n_run
        for c in range(d - 1, -1, -1):                   # remaining axes, x last (slowest)
            if c == run:
                continue
            ids = ids + loc[:, :, c] * stride
            stride *= int(dims2[c])
        return ids.to(torch.int32).cpu().numpy(), int(np.prod(dims2))
    nv = mesh.n_vertices
    ev = np.stack([np.stack([mesh.cells[:, a], mesh.cells[:, b]], axis=1) for a, b in edges], axis=1).astype(np.int64)
    ev.sort(axis=2)
    keys = ev[:, :, 0] * nv + ev[:, :, 1]
    uniq, inv = np.unique(keys.ravel(), return_inverse=True)
    dm = np.concatenate([mesh.cells.astype(np.int64), nv + inv.reshape(nc, len(edges))], axis=1)
    return dm.astype(np.int32), nv + uniq.size


def _vertex_lattice_coords(mesh: Mesh):
    """Integer lattice coordinates of the vertices of a structured mesh (local to the slab)."""
    lat = mesh.lattice
    if lat is None:
        return None
    if lat["kind"] == "line":
        return np.arange(mesh.n_vertices, dtype=np.int64)[:, None]
    n = lat["n"]
    i0, i1 = lat["x_range"]
    dims = [i1 - i0 + 1] + [k + 1 for k in n[1:]]
    if int(np.prod(dims)) != mesh.n_vertices:
        return None
    grids = np.meshgrid(*[np.arange(k) for k in dims], indexing="ij")
    return np.stack([g.ravel() for g in grids], axis=1).astype(np.int64)


def winner_map(sigma: ScalarSpace, T: ScalarSpace):
    """Cross-space evaluation map.

    dolfinx interpolates an Expression cell by cell and scatters through the sigma dofmap, later
    cells overwriting earlier ones, so the value at sigma node s comes from the LAST cell touching
    it.  Returns (dofs [n_sigma, n_ld_T] int32, local_point [n_sigma] uint8, weights [n_pts, n_ld_T])."""
    nc = sigma.mesh.n_cells
    winner = np.full(sigma.n_nodes, -1, dtype=np.int64)
    local = np.zeros(sigma.n_nodes, dtype=np.uint8)
    flat = sigma.dofmap.ravel()
    cell_of = np.repeat(np.arange(nc, dtype=np.int64), sigma.n_ld)
    lp_of = np.tile(np.arange(sigma.n_ld, dtype=np.uint8), nc)
    winner[flat] = cell_of      # numpy fancy assignment: the last occurrence wins
    local[flat] = lp_of
    assert (winner >= 0).all()
    weights, _ = T.element.tabulate(sigma.element.nodes)
    weights = np.where(np.abs(weights) < 1e-14, 0.0, weights)
    weights = np.where(np.abs(weights - 1.0) < 1e-14, 1.0, weights)
    return T.dofmap[winner].astype(np.int32), local, np.ascontiguousarray(weights)
