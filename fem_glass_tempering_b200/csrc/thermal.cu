// Hot path (B): matrix-free heat-equation operator of ThermalModel / ThermoViscoProblem.
//
// Replaces the FFCx tabulate_tensor kernels + PETSc AIJ assembly that
// NonlinearProblem(F, u) performs in every Newton iteration (TVP:293-332):
//   residual  F(T;v) = (T - T_prev, v) + dt*[ alpha (grad T, grad v) - (f, v)
//                      + 0.001*sigma*eps <T^4 - Ta^4, v>_ds + 0.001*htc <T - Ta, v>_ds ]      TVP:293-306
//                      + dt*alpha*[ (5/h+) <[[v]],[[T]]> - <{grad v},[[T]]> - <[[v]],{grad T}> ]_dS  (DG)  TVP:318-325
//   Jacobian  J(T)[x,v] = same with T -> x in the linear terms and
//                         dt*0.001*<(4 sigma eps T^3 + htc) x, v>_ds on the boundary.
//
// Two families of kernels:
//   (i)  class kernels (dg_class_apply, dg_cheb_step, dg_class_resid, cg_class_apply, cg_class_resid): the mesh's cells
//        fall into a handful of (shape, neighbourhood) classes found at operator creation; the per-class local matrices
//        sit in shared memory and a cell reads only its class word, neighbour ids and rows of x (further down);
//   (ii) the general kernel below, used for meshes with too many classes and for the diagonal:
// one thread per cell, element tables passed BY VALUE in the kernel parameter block so that
// every table entry is a constant-bank operand of a fully unrolled DFMA (no shared-memory or LDS
// traffic for the tables).  Cell geometry is SoA ([component][cell]) so a warp reads 32 consecutive
// doubles per component.  DG: cell-centric interior facets (each cell visits its d+1 neighbours, no
// atomics, no colouring; the neighbour's dofs/geometry come through L2).  CG: gather through an SoA
// dofmap, scatter with native fp64 atomics (RED.ADD.F64).  Exterior (Robin + radiation) facets are a
// small separate kernel.
#include <math.h>

#include <algorithm>
#include <vector>

#include "sg_common.cuh"

namespace {

constexpr int TB = 128;  // threads per block for cell kernels

__host__ __device__ constexpr int nld_of(int D, int P) { return P == 1 ? D + 1 : (D + 1) * (D + 2) / 2; }
__host__ __device__ constexpr int nqc_of(int D, int P) { return P == 1 ? 1 : D + 1; }
__host__ __device__ constexpr int nqf_of(int D, int P) {
    return D == 1 ? 1 : (D == 2 ? P + 1 : (P == 1 ? 3 : 6));
}
__host__ __device__ constexpr int nperm_of(int D) { return D == 1 ? 1 : (D == 2 ? 2 : 6); }

template <int D, int P, bool DG>
struct Tab {
    static constexpr int NLD = nld_of(D, P), NQC = nqc_of(D, P), NQF = nqf_of(D, P), NPERM = nperm_of(D);
    double mass[NLD * NLD];
    double load[NLD];
    double cq_w[NQC];
    double cq_grad[NQC * D * NLD];                                  // [q][a][j]
    double fq_w[DG ? NQF : 1];
    double fq_val[DG ? (D + 1) * NQF * NLD : 1];                    // [f][q][j]
    double fq_grad[(DG && P == 2) ? (D + 1) * NQF * D * NLD : 1];   // [f][q][a][j]
    int fq_perm[DG ? NPERM * NQF : 1];                              // [perm][q]
};

struct OpDev {
    long n_cells, cell_lo, cell_hi;
    long dot_lo, dot_hi;    // cells whose x_K . y_K enters the fused dot product
    long n_dofs;
    const int32_t *dofmap;  // CG: [NLD][n_cells]
    const double *geom;     // [D*D + 2][n_cells]
    const int32_t *nbr;     // DG: [D+1][n_cells]
    const int32_t *nbinfo;  // DG: [n_cells]
    long n_bf;
    const int32_t *bf_cell, *bf_facet;
    const double *bf_area;
    const double *btab;     // [D+1][NQB][NLD] boundary basis values
    const double *bw;       // [NQB]
    int nqb;
    double dt, dt_alpha, dt_f, se, htc, Ta, penalty;
};

enum { MODE_APPLY = 0, MODE_RESID = 1, MODE_DIAG = 2 };

// reference gradient of barycentric l_i:  l_0 -> (-1,...,-1), l_{a+1} -> e_a
template <int D>
__device__ __forceinline__ double dlam_dot(const double (&v)[D], int i) {
    if (i == 0) {
        double s = 0.0;
#pragma unroll
        for (int a = 0; a < D; ++a) s -= v[a];
        return s;
    }
    return v[i - 1];
}

// Where the input vector comes from: the global array (the operator kernels) or a unit vector (the
// builder of the local-matrix class tables, which thereby shares every line of arithmetic with the
// general kernel).
struct XGlobal {
    const double *x;
    __device__ __forceinline__ double at(long dof) const { return x[dof]; }
};
struct XUnit {
    long sel;
    __device__ __forceinline__ double at(long dof) const { return dof == sel ? 1.0 : 0.0; }
};

// Element vector of cell c: yk = (cell integrals + DG interior-facet terms)(x), MODE as below.
template <int D, int P, bool DG, int MODE, class XA>
__device__ __forceinline__ void cell_compute(const Tab<D, P, DG> &tab, const OpDev &op, const long c, const XA xa,
                                             const double *__restrict__ xprev, long (&dof)[nld_of(D, P)],
                                             double (&yk)[nld_of(D, P)]) {
    using T = Tab<D, P, DG>;
    constexpr int NLD = T::NLD, NQC = T::NQC, NQF = T::NQF;
    const long nc = op.n_cells;

    double Jinv[D][D];
#pragma unroll
    for (int a = 0; a < D; ++a)
#pragma unroll
        for (int b = 0; b < D; ++b) Jinv[a][b] = op.geom[(long)(a * D + b) * nc + c];
    const double detJ = op.geom[(long)(D * D) * nc + c];

    double xk[NLD];
#pragma unroll
    for (int i = 0; i < NLD; ++i) {
        dof[i] = DG ? c * NLD + i : (long)op.dofmap[(long)i * nc + c];
        if (MODE != MODE_DIAG) xk[i] = xa.at(dof[i]);
    }

    // ---- cell integrals: |detJ| * Mhat  +  dt*alpha * sum_q w_q detJ (Jinv^T grad)^T (Jinv^T grad) ----
    if (MODE == MODE_DIAG) {
#pragma unroll
        for (int i = 0; i < NLD; ++i) {
            double s = 0.0;
#pragma unroll
            for (int q = 0; q < NQC; ++q) {
                double gp[D];
#pragma unroll
                for (int b = 0; b < D; ++b) {
                    gp[b] = 0.0;
#pragma unroll
                    for (int a = 0; a < D; ++a) gp[b] += Jinv[a][b] * tab.cq_grad[(q * D + a) * NLD + i];
                }
                double n2 = 0.0;
#pragma unroll
                for (int b = 0; b < D; ++b) n2 += gp[b] * gp[b];
                s += tab.cq_w[q] * n2;
            }
            yk[i] = detJ * (tab.mass[i * NLD + i] + op.dt_alpha * s);
        }
    } else {
        double xm[NLD];
#pragma unroll
        for (int i = 0; i < NLD; ++i) xm[i] = (MODE == MODE_RESID) ? xk[i] - xprev[dof[i]] : xk[i];
#pragma unroll
        for (int i = 0; i < NLD; ++i) {
            double s = 0.0;
#pragma unroll
            for (int j = 0; j < NLD; ++j) s += tab.mass[i * NLD + j] * xm[j];
            yk[i] = s;
            if (MODE == MODE_RESID) yk[i] -= op.dt_f * tab.load[i];
        }
        double ys[NLD];
#pragma unroll
        for (int i = 0; i < NLD; ++i) ys[i] = 0.0;
#pragma unroll
        for (int q = 0; q < NQC; ++q) {
            double gr[D], gp[D], fl[D];
#pragma unroll
            for (int a = 0; a < D; ++a) {
                gr[a] = 0.0;
#pragma unroll
                for (int j = 0; j < NLD; ++j) gr[a] += tab.cq_grad[(q * D + a) * NLD + j] * xk[j];
            }
#pragma unroll
            for (int b = 0; b < D; ++b) {
                gp[b] = 0.0;
#pragma unroll
                for (int a = 0; a < D; ++a) gp[b] += Jinv[a][b] * gr[a];
            }
#pragma unroll
            for (int a = 0; a < D; ++a) {
                fl[a] = 0.0;
#pragma unroll
                for (int b = 0; b < D; ++b) fl[a] += Jinv[a][b] * gp[b];
                fl[a] *= tab.cq_w[q];
            }
#pragma unroll
            for (int i = 0; i < NLD; ++i)
#pragma unroll
                for (int a = 0; a < D; ++a) ys[i] += tab.cq_grad[(q * D + a) * NLD + i] * fl[a];
        }
#pragma unroll
        for (int i = 0; i < NLD; ++i) yk[i] = detJ * (yk[i] + op.dt_alpha * ys[i]);
    }

    // ---- DG: symmetric interior penalty on the d+1 facets of this cell (TVP:318-325) ----
    if constexpr (DG) {
        const double hK = op.geom[(long)(D * D + 1) * nc + c];
        const int info_all = op.nbinfo[c];
        constexpr double inv_fact = (D == 3) ? 0.5 : 1.0;  // 1/(D-1)!
#pragma unroll
        for (int f = 0; f < D + 1; ++f) {
            const long nb = op.nbr[(long)f * nc + c];
            if (nb < 0) continue;
            const int info = (info_all >> (5 * f)) & 31;
            const int nbf = info & 3, pid = info >> 2;
            // outward normal and measure of facet f from grad(l_f) = Jinv^T dlam_f
            double g[D], n[D];
            double g2 = 0.0;
#pragma unroll
            for (int b = 0; b < D; ++b) {
                double col[D];
#pragma unroll
                for (int a = 0; a < D; ++a) col[a] = Jinv[a][b];
                g[b] = dlam_dot<D>(col, f);
                g2 += g[b] * g[b];
            }
            const double gn = sqrt(g2);
#pragma unroll
            for (int b = 0; b < D; ++b) n[b] = -g[b] / gn;
            const double wf = op.dt_alpha * detJ * gn * inv_fact;  // dt*alpha*|F|
            const double hplus = (c < nb) ? hK : op.geom[(long)(D * D + 1) * nc + nb];
            const double pen = op.penalty / hplus;
            double jn[D];  // (Jinv n)_a : normal derivative of reference-gradient component a
#pragma unroll
            for (int a = 0; a < D; ++a) {
                jn[a] = 0.0;
#pragma unroll
                for (int b = 0; b < D; ++b) jn[a] += Jinv[a][b] * n[b];
            }
            if (MODE == MODE_DIAG) {
#pragma unroll
                for (int q = 0; q < NQF; ++q) {
                    const double w = wf * tab.fq_w[q];
#pragma unroll
                    for (int i = 0; i < NLD; ++i) {
                        const double ph = tab.fq_val[(f * NQF + q) * NLD + i];
                        double dph;
                        if constexpr (P == 1) {
                            dph = dlam_dot<D>(jn, i);
                        } else {
                            dph = 0.0;
#pragma unroll
                            for (int a = 0; a < D; ++a) dph += jn[a] * tab.fq_grad[((f * NQF + q) * D + a) * NLD + i];
                        }
                        yk[i] += w * (pen * ph * ph - ph * dph);
                    }
                }
                continue;
            }
            double jnN[D], xn[NLD];
#pragma unroll
            for (int a = 0; a < D; ++a) {
                jnN[a] = 0.0;
#pragma unroll
                for (int b = 0; b < D; ++b) jnN[a] += op.geom[(long)(a * D + b) * nc + nb] * n[b];
            }
#pragma unroll
            for (int j = 0; j < NLD; ++j) xn[j] = xa.at(nb * NLD + j);
            double dphK[NLD], dnK = 0.0, dnN = 0.0;
            if constexpr (P == 1) {
#pragma unroll
                for (int i = 0; i < NLD; ++i) {
                    dphK[i] = dlam_dot<D>(jn, i);
                    dnK += dphK[i] * xk[i];
                    dnN += dlam_dot<D>(jnN, i) * xn[i];
                }
            }
#pragma unroll
            for (int q = 0; q < NQF; ++q) {
                const int qn = tab.fq_perm[pid * NQF + q];
                const double *phN = &tab.fq_val[(nbf * NQF + qn) * NLD];
                double vK = 0.0, vN = 0.0;
#pragma unroll
                for (int j = 0; j < NLD; ++j) {
                    vK += tab.fq_val[(f * NQF + q) * NLD + j] * xk[j];
                    vN += phN[j] * xn[j];
                }
                if constexpr (P == 2) {
                    dnK = 0.0;
                    dnN = 0.0;
#pragma unroll
                    for (int i = 0; i < NLD; ++i) {
                        double s = 0.0, sN = 0.0;
#pragma unroll
                        for (int a = 0; a < D; ++a) {
                            s += jn[a] * tab.fq_grad[((f * NQF + q) * D + a) * NLD + i];
                            sN += jnN[a] * tab.fq_grad[((nbf * NQF + qn) * D + a) * NLD + i];
                        }
                        dphK[i] = s;
                        dnK += s * xk[i];
                        dnN += sN * xn[i];
                    }
                }
                const double jump = vK - vN, avg = 0.5 * (dnK + dnN);
                const double w = wf * tab.fq_w[q];
#pragma unroll
                for (int i = 0; i < NLD; ++i) {
                    const double ph = tab.fq_val[(f * NQF + q) * NLD + i];
                    yk[i] += w * ((pen * jump - avg) * ph - 0.5 * dphK[i] * jump);
                }
            }
        }
    }

}

template <int D, int P, bool DG, int MODE>
__global__ void __launch_bounds__(TB, (P == 1) ? 6 : 1) cell_kernel(const __grid_constant__ Tab<D, P, DG> tab, const __grid_constant__ OpDev op, const double *__restrict__ x,
                                                  const double *__restrict__ xprev, double *__restrict__ y) {
    constexpr int NLD = nld_of(D, P);
    const long c = op.cell_lo + (long)blockIdx.x * TB + threadIdx.x;
    if (c >= op.cell_hi) return;
    long dof[NLD];
    double yk[NLD];
    cell_compute<D, P, DG, MODE>(tab, op, c, XGlobal{x}, xprev, dof, yk);
#pragma unroll
    for (int i = 0; i < NLD; ++i) {
        if (DG)
            y[dof[i]] = yk[i];
        else
            atomicAdd(&y[dof[i]], yk[i]);
    }
}

// Exterior facets: radiation + convection (TVP:302-304) and their linearisation.  Grid-stride; with DOT the
// kernel also reduces sum_i x_i * (its own contribution to y_i) over facets of cells in [dot_lo, dot_hi).
template <int D, int P, bool DG, int MODE, bool DOT>
__global__ void __launch_bounds__(TB) bfacet_kernel(const OpDev op, const double *__restrict__ Tlin,
                                                    const double *__restrict__ x, double *__restrict__ y, SgRed red,
                                                    double *dot_out, const int *skip) {
    constexpr int NLD = nld_of(D, P);
    if (skip && *skip) return;
    double dsum[1] = {0.0};
    for (long b = (long)blockIdx.x * TB + threadIdx.x; b < op.n_bf; b += (long)gridDim.x * TB) {
        const long c = op.bf_cell[b];
        if (c < op.cell_lo || c >= op.cell_hi) continue;
        const int f = op.bf_facet[b];
        const double area = op.bf_area[b];
        long dof[NLD];
        double Tk[NLD], xk[NLD], acc[NLD];
#pragma unroll
        for (int i = 0; i < NLD; ++i) {
            dof[i] = DG ? c * NLD + i : (long)op.dofmap[(long)i * op.n_cells + c];
            Tk[i] = Tlin[dof[i]];
            xk[i] = (MODE == MODE_APPLY) ? x[dof[i]] : 0.0;
            acc[i] = 0.0;
        }
        for (int q = 0; q < op.nqb; ++q) {
            const double *ph = op.btab + ((long)f * op.nqb + q) * NLD;
            double Tq = 0.0, xq = 0.0;
#pragma unroll
            for (int j = 0; j < NLD; ++j) {
                Tq += ph[j] * Tk[j];
                xq += ph[j] * xk[j];
            }
            const double w = area * op.bw[q] * op.dt * 0.001;
            if (MODE == MODE_RESID) {
                const double T2 = Tq * Tq, Ta2 = op.Ta * op.Ta;
                const double flux = w * (op.se * (T2 * T2 - Ta2 * Ta2) + op.htc * (Tq - op.Ta));
#pragma unroll
                for (int i = 0; i < NLD; ++i) acc[i] += flux * ph[i];
            } else {
                const double coef = w * (4.0 * op.se * Tq * Tq * Tq + op.htc);
#pragma unroll
                for (int i = 0; i < NLD; ++i) acc[i] += (MODE == MODE_APPLY) ? coef * xq * ph[i] : coef * ph[i] * ph[i];
            }
        }
        const bool counted = DOT && c >= op.dot_lo && c < op.dot_hi;
#pragma unroll
        for (int i = 0; i < NLD; ++i)
            if (acc[i] != 0.0) {
                atomicAdd(&y[dof[i]], acc[i]);
                if (counted) dsum[0] += xk[i] * acc[i];
            }
    }
    if (DOT) sg_grid_reduce<1>(dsum, red, dot_out);
}

// sum over [lo, hi) of a*b (general path of sg_thermal_apply_dot)
__global__ void __launch_bounds__(256) k_dot_range(long lo, long hi, const double *__restrict__ a, const double *__restrict__ b,
                                                   SgRed red, double *out, const int *skip) {
    if (skip && *skip) return;
    double acc[2] = {0.0, 0.0};  // out[1] = 0: the exterior-facet part is already inside a.b
    for (long i = lo + (long)blockIdx.x * 256 + threadIdx.x; i < hi; i += (long)gridDim.x * 256) acc[0] += a[i] * b[i];
    sg_grid_reduce<2>(acc, red, out);
}

// ================================================================ local-matrix classes (fast apply path)
//
// On the plates of the benchmark (and on any mesh with few distinct cell shapes) the cell part of the
// Jacobian is a handful of distinct local matrices:
//     y_K = A_self[s(K)] x_K + sum_f A_nb[u(K,f)] x_{N(K,f)}          (DG; CG: only A_self, scattered)
// s(K) = class of (cell shape, which facets are interior, the facet classes), u(K,f) = class of (shape of
// K, f, shape of N, N's local facet, vertex permutation, which side is '+').  The classes are found at
// operator creation by hashing the cell geometry rounded to 40 mantissa bits (GPU, classify.cu), the
// tables are produced by running cell_compute on unit vectors for one representative per class, and the
// apply kernel keeps them in shared memory: per cell it reads 8 B of class ids, the neighbour ids and
// x, and writes y — no geometry, no square roots, no divisions.  Meshes with too many classes keep
// using cell_kernel.

__device__ __forceinline__ uint64_t key_mix(uint64_t h, uint64_t v) {
    h ^= v + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
    h *= 0xff51afd7ed558ccdull;
    h ^= h >> 33;
    return h;
}
// round the mantissa to 40 bits; values below `cut` count as zero
__device__ __forceinline__ uint64_t key_round(double v, double cut) {
    if (!(fabs(v) >= cut)) return 0ull;
    uint64_t b = (uint64_t)__double_as_longlong(v);
    return (b + 0x800ull) & ~0xFFFull;
}

template <int D>
__device__ __forceinline__ void geom_rounded(const double *geom, long nc, long c, uint64_t (&q)[D * D + 2]) {
    double v[D * D + 2], mx = 0.0;
#pragma unroll
    for (int k = 0; k < D * D + 2; ++k) v[k] = geom[(long)k * nc + c];
#pragma unroll
    for (int k = 0; k < D * D; ++k) mx = fmax(mx, fabs(v[k]));
#pragma unroll
    for (int k = 0; k < D * D; ++k) q[k] = key_round(v[k], mx * 0x1p-38);
    q[D * D] = key_round(v[D * D], 0.0);
    q[D * D + 1] = key_round(v[D * D + 1], 0.0);
}

template <int D>
__global__ void k_geom_key(const double *geom, long nc, uint64_t *keys) {
    const long c = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nc) return;
    uint64_t q[D * D + 2], h = 0x243F6A8885A308D3ull;
    geom_rounded<D>(geom, nc, c, q);
#pragma unroll
    for (int k = 0; k < D * D + 2; ++k) h = key_mix(h, q[k]);
    keys[c] = h;
}

// counts cells whose rounded geometry differs from their class representative's (hash collisions)
template <int D>
__global__ void k_geom_verify(const double *geom, long nc, const int32_t *gcls, const int32_t *rep, unsigned *bad) {
    const long c = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nc) return;
    uint64_t q[D * D + 2], r[D * D + 2];
    geom_rounded<D>(geom, nc, c, q);
    geom_rounded<D>(geom, nc, rep[gcls[c]], r);
    bool same = true;
#pragma unroll
    for (int k = 0; k < D * D + 2; ++k) same = same && q[k] == r[k];
    if (!same) atomicAdd(bad, 1u);
}

// key of (cell c, local facet f), layout [f][c]; 0 = no neighbour
__global__ void k_facet_key(int nnb, long nc, const int32_t *nbr, const int32_t *nbinfo, const int32_t *gcls, uint64_t *keys) {
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nc * nnb) return;
    const int f = (int)(t / nc);
    const long c = t % nc;
    const long nb = nbr[t];
    if (nb < 0) {
        keys[t] = 0ull;
        return;
    }
    const int info = (nbinfo[c] >> (5 * f)) & 31;
    uint64_t h = 0x13198A2E03707344ull;
    h = key_mix(h, (uint64_t)gcls[c]);
    h = key_mix(h, (uint64_t)gcls[nb]);
    h = key_mix(h, (uint64_t)(f * 64 + info * 2 + (c < nb ? 1 : 0)));
    keys[t] = h | 1ull;
}

__global__ void k_self_key(int nnb, long nc, const int32_t *nbr, const int32_t *gcls, const int32_t *fcls, uint64_t *keys) {
    const long c = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nc) return;
    uint64_t h = key_mix(0xA4093822299F31D0ull, (uint64_t)gcls[c]);
    for (int f = 0; f < nnb; ++f) {
        const long t = (long)f * nc + c;
        h = key_mix(h, nbr[t] < 0 ? 0ull : (uint64_t)fcls[t] + 1ull);
    }
    keys[c] = h;
}

// DG: 16 bits self class | 12 bits per facet (0 = no neighbour, else facet class + 1)
__global__ void k_pack_dg(int nnb, long nc, const int32_t *nbr, const int32_t *scls, const int32_t *fcls, uint64_t *out) {
    const long c = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nc) return;
    uint64_t w = (uint64_t)scls[c];
    for (int f = 0; f < nnb; ++f) {
        const long t = (long)f * nc + c;
        const uint64_t u = nbr[t] < 0 ? 0ull : (uint64_t)fcls[t] + 1ull;
        w |= u << (16 + 12 * f);
    }
    out[c] = w;
}
__global__ void k_pack_cg(long nc, const int32_t *gcls, uint16_t *out) {
    const long c = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c < nc) out[c] = (uint16_t)gcls[c];
}

// One thread per (class, column j): column j of the class's local matrix = cell_compute(unit vector).
// Table layout: class k at tab_out + k*S, entry (i, j) at [i*NLD + j]; self classes first, then facet classes.
template <int D, int P, bool DG>
__global__ void k_build_tables(const __grid_constant__ Tab<D, P, DG> tab, const __grid_constant__ OpDev op, int n_self,
                               const int32_t *rep_self, int n_nb, const int32_t *rep_nb, int S, double *tab_out) {
    constexpr int NLD = nld_of(D, P);
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (n_self + n_nb) * NLD) return;
    const int k = t / NLD, j = t % NLD;
    long c, sel;
    if (k < n_self) {
        c = rep_self[k];
        sel = DG ? c * NLD + j : (long)op.dofmap[(long)j * op.n_cells + c];
    } else {
        const long r = rep_nb[k - n_self];
        c = r % op.n_cells;
        const long nb = op.nbr[r];
        if (nb < 0) {
            for (int i = 0; i < NLD; ++i) tab_out[(long)k * S + i * NLD + j] = 0.0;
            return;
        }
        sel = nb * NLD + j;
    }
    long dof[NLD];
    double yk[NLD];
    cell_compute<D, P, DG, MODE_APPLY>(tab, op, c, XUnit{sel}, nullptr, dof, yk);
#pragma unroll
    for (int i = 0; i < NLD; ++i) tab_out[(long)k * S + i * NLD + j] = yk[i];
}

__global__ void k_mark_exterior(long n_bf, long nc, const int32_t *bf_cell, const int32_t *bf_facet, int32_t *nbr_ext) {
    const long b = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b < n_bf) nbr_ext[(long)bf_facet[b] * nc + bf_cell[b]] = (int32_t)(-2 - b);
}

// Dofs on local facet f (fe.py LagrangeElement.facet_dofs): the vertices other than f, ascending, then (P2) the
// edges joining two of them in the order of ref_edges(dim).  Basis functions of all other dofs vanish on f.
__host__ __device__ constexpr int nfd_of(int D, int P) { return P == 1 ? D : D * (D + 1) / 2; }
__host__ __device__ constexpr int ref_edge_vertex(int D, int e, int side) {
    // ref_edges: 1-D (0,1); 2-D (1,2),(0,2),(0,1); 3-D (2,3),(1,3),(1,2),(0,3),(0,2),(0,1)
    if (D == 1) return side;
    if (D == 2) return side == 0 ? (e == 0 ? 1 : 0) : (e == 2 ? 1 : 2);
    const int a[6] = {2, 1, 1, 0, 0, 0}, b[6] = {3, 3, 2, 3, 2, 1};
    return side == 0 ? a[e] : b[e];
}
__host__ __device__ constexpr int facet_dof(int D, int P, int f, int k) {
    if (k < D) return k < f ? k : k + 1;
    if (P == 1) return -1;
    int m = D;
    const int ne = D * (D + 1) / 2;
    for (int e = 0; e < ne; ++e) {
        if (ref_edge_vertex(D, e, 0) == f || ref_edge_vertex(D, e, 1) == f) continue;
        if (m == k) return D + 1 + e;
        ++m;
    }
    return -1;
}

// Linearised Robin + radiation matrix of every exterior facet at the temperature T_lin (TVP:302-304):
// bmat[b][k <= l packed row-wise over the facet's dofs] = dt*0.001 * int (4 sigma eps T^3 + htc) phi_k phi_l ds.
template <int D, int P, bool DG>
__global__ void __launch_bounds__(TB) k_bfacet_mats(const OpDev op, const double *__restrict__ Tlin, double *__restrict__ bmat) {
    constexpr int NLD = nld_of(D, P), NFD = nfd_of(D, P), NFDP = NFD * (NFD + 1) / 2;
    const long b = (long)blockIdx.x * TB + threadIdx.x;
    if (b >= op.n_bf) return;
    const long c = op.bf_cell[b];
    const int f = op.bf_facet[b];
    const double area = op.bf_area[b];
    int fd[NFD];
#pragma unroll
    for (int k = 0; k < NFD; ++k) {
        fd[k] = 0;
#pragma unroll
        for (int ff = 0; ff < D + 1; ++ff)
            if (ff == f) fd[k] = facet_dof(D, P, ff, k);
    }
    double Tk[NFD], B[NFDP];
#pragma unroll
    for (int k = 0; k < NFD; ++k) Tk[k] = Tlin[DG ? c * NLD + fd[k] : (long)op.dofmap[(long)fd[k] * op.n_cells + c]];
#pragma unroll
    for (int k = 0; k < NFDP; ++k) B[k] = 0.0;
    for (int q = 0; q < op.nqb; ++q) {
        const double *ph = op.btab + ((long)f * op.nqb + q) * NLD;
        double pf[NFD], Tq = 0.0;
#pragma unroll
        for (int k = 0; k < NFD; ++k) {
            pf[k] = ph[fd[k]];
            Tq += pf[k] * Tk[k];
        }
        const double coef = area * op.bw[q] * op.dt * 0.001 * (4.0 * op.se * Tq * Tq * Tq + op.htc);
        int m = 0;
#pragma unroll
        for (int k = 0; k < NFD; ++k)
#pragma unroll
            for (int l = k; l < NFD; ++l) B[m++] += coef * pf[k] * pf[l];
    }
#pragma unroll
    for (int k = 0; k < NFDP; ++k) bmat[b * NFDP + k] = B[k];
}

// DG partition: owned cells with a neighbour outside [lo, hi) (a ghost cell).  They sit at the two ends of the owned
// range (x-slabs): lowB = one past the last such cell of the lower half, highB = the first one of the upper half.
__global__ void k_split_range(int nnb, long nc, long lo, long hi, const int32_t *nbr, unsigned long long *lowB, unsigned long long *highB) {
    const long c = lo + (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= hi) return;
    bool ghost = false;
    for (int f = 0; f < nnb; ++f) {
        const long nb = nbr[(long)f * nc + c];
        ghost = ghost || (nb >= 0 && (nb < lo || nb >= hi));
    }
    if (!ghost) return;
    if (c < lo + (hi - lo) / 2)
        atomicMax(lowB, (unsigned long long)(c + 1));
    else
        atomicMin(highB, (unsigned long long)c);
}

struct ClsDev {
    long n_cells, cell_lo, cell_hi, dot_lo, dot_hi;
    const int32_t *nbr;     // DG [NNB][n_cells]
    const int32_t *dofmap;  // CG [NLD][n_cells]
    const uint64_t *cls64;  // DG
    const uint16_t *cls16;  // CG
    const double *tab;      // [(n_self + n_nb) * S]
    int n_self, n_nb, S;
    // exterior facets handled inside the DG class kernel (P1): nbr holds -2 - b for exterior facet b and
    // bmat[b] the packed symmetric matrix  dt*0.001*int (4 sigma eps T^3 + htc) phi_k phi_l ds  over the facet's dofs
    const double *bmat;
    // CG: exterior facets applied by the same kernel after the cells (bmat != NULL)
    long n_bf;
    const int32_t *bf_cell, *bf_facet;
    // DG, partitioned mesh: owned cells [split_lo, split_hi) have no ghost neighbour.  The kernels process them first and
    // wait for the neighbours' ghost rows (SgHaloWait) only before the two boundary strips [cell_lo, split_lo) and
    // [split_hi, cell_hi), so the exchange hides behind the interior.  split_lo == cell_lo, split_hi == cell_hi otherwise.
    long split_lo, split_hi;
    int rev;   // sweep direction of this launch (sg_sweep_begin)
};

constexpr int CB = 256;  // threads per block of the class kernels

struct __align__(32) Row4 {
    double v[4];
};

// One cell row (NLD contiguous doubles): 256-bit accesses (LDG.E.256 / STG.E.256, sm_100) when WIDE and
// NLD is a multiple of 4, 128-bit when NLD is even, scalar otherwise.
template <int NLD, bool WIDE>
__device__ __forceinline__ void load_row(const double *__restrict__ p, double (&v)[NLD]) {
    if constexpr (WIDE && NLD % 4 == 0) {
#pragma unroll
        for (int i = 0; i < NLD / 4; ++i) {
            const Row4 t = reinterpret_cast<const Row4 *>(p)[i];
#pragma unroll
            for (int k = 0; k < 4; ++k) v[4 * i + k] = t.v[k];
        }
    } else if constexpr (NLD % 2 == 0) {
#pragma unroll
        for (int i = 0; i < NLD / 2; ++i) {
            const double2 t = reinterpret_cast<const double2 *>(p)[i];
            v[2 * i] = t.x;
            v[2 * i + 1] = t.y;
        }
    } else {
#pragma unroll
        for (int i = 0; i < NLD; ++i) v[i] = p[i];
    }
}
template <int NLD, bool WIDE>
__device__ __forceinline__ void store_row(double *__restrict__ p, const double (&v)[NLD]) {
    if constexpr (WIDE && NLD % 4 == 0) {
#pragma unroll
        for (int i = 0; i < NLD / 4; ++i) {
            Row4 t;
#pragma unroll
            for (int k = 0; k < 4; ++k) t.v[k] = v[4 * i + k];
            reinterpret_cast<Row4 *>(p)[i] = t;
        }
    } else if constexpr (NLD % 2 == 0) {
#pragma unroll
        for (int i = 0; i < NLD / 2; ++i) reinterpret_cast<double2 *>(p)[i] = make_double2(v[2 * i], v[2 * i + 1]);
    } else {
#pragma unroll
        for (int i = 0; i < NLD; ++i) p[i] = v[i];
    }
}

// acc[i] += sum_j A[i*NLD + j] * x[j] with A in shared memory (128-bit reads when the rows are 16-B aligned)
template <int NLD>
__device__ __forceinline__ void smem_matvec_acc(const double *__restrict__ A, const double (&x)[NLD], double (&acc)[NLD]) {
#pragma unroll
    for (int i = 0; i < NLD; ++i) {
        double a = acc[i];
        if constexpr (NLD % 2 == 0) {
            const double2 *row = reinterpret_cast<const double2 *>(A + i * NLD);
#pragma unroll
            for (int j = 0; j < NLD / 2; ++j) {
                const double2 t = row[j];
                a += t.x * x[2 * j];
                a += t.y * x[2 * j + 1];
            }
        } else {
#pragma unroll
            for (int j = 0; j < NLD; ++j) a += A[i * NLD + j] * x[j];
        }
        acc[i] = a;
    }
}

// Exterior facet f of a DG cell: y_K += B_f x_K on the facet's dofs (nb = -2 - facet index, see ClsDev::bmat).
template <int NLD, int NNB, int P>
__device__ __forceinline__ void dg_exterior_facet(const ClsDev &cd, const int f, const int nb, const double (&xk)[NLD], double (&yk)[NLD]) {
    constexpr int D = NNB - 1, NFD = nfd_of(D, P), NFDP = NFD * (NFD + 1) / 2;
    const double *Bp = cd.bmat + (long)(-2 - nb) * NFDP;
    double B[NFDP];
#pragma unroll
    for (int k = 0; k < NFDP; ++k) B[k] = Bp[k];
    int m = 0;
#pragma unroll
    for (int k = 0; k < NFD; ++k)
#pragma unroll
        for (int l = k; l < NFD; ++l) {
            const int ik = facet_dof(D, P, f, k), il = facet_dof(D, P, f, l);
            yk[ik] += B[m] * xk[il];
            if (l != k) yk[il] += B[m] * xk[ik];
            ++m;
        }
}

// Element vector of DG cell c from the class tables: loads the class word, the neighbour ids, the cell's row
// xk and the neighbours' rows of x, returns yk = A_self x_K + sum_f A_nb x_N (+ exterior-facet matrices).
template <int NLD, int NNB, int P, bool WIDE, bool BND>
__device__ __forceinline__ void dg_cell_apply(const ClsDev &cd, const double *s_tab, const double *s_nb, const int c,
                                              const double *__restrict__ x, double (&xk)[NLD], double (&yk)[NLD]) {
    const int nc = (int)cd.n_cells;
    const uint64_t w = cd.cls64[c];
    int nb[NNB];
#pragma unroll
    for (int f = 0; f < NNB; ++f) nb[f] = cd.nbr[(size_t)f * nc + c];
    load_row<NLD, WIDE>(x + (size_t)c * NLD, xk);
    double xn[NNB][NLD];
#pragma unroll
    for (int f = 0; f < NNB; ++f) {
        if (nb[f] >= 0) {
            load_row<NLD, WIDE>(x + (long)nb[f] * NLD, xn[f]);
        } else {
#pragma unroll
            for (int j = 0; j < NLD; ++j) xn[f][j] = 0.0;
        }
    }
#pragma unroll
    for (int i = 0; i < NLD; ++i) yk[i] = 0.0;
    smem_matvec_acc<NLD>(s_tab + (int)(w & 0xFFFFull) * cd.S, xk, yk);
#pragma unroll
    for (int f = 0; f < NNB; ++f) {
        const int u = (int)((w >> (16 + 12 * f)) & 0xFFFull);
        if (u) smem_matvec_acc<NLD>(s_nb + (u - 1) * cd.S, xn[f], yk);
        if constexpr (BND) {
            if (nb[f] < -1) dg_exterior_facet<NLD, NNB, P>(cd, f, nb[f], xk, yk);
        }
    }
}

__device__ __forceinline__ void load_class_tables(const ClsDev &cd, double *s_tab, int ntab) {
    for (int i = threadIdx.x; i < ntab; i += CB) s_tab[i] = cd.tab[i];
    __syncthreads();
}

// DG fast apply: persistent grid-stride blocks, class tables in shared memory, fused x.y reduction.
// A warp handles 32 consecutive cells; when these share their classes (the plate meshes number the cells
// in class-uniform tiles of 32, mesh.py) every table read is a shared-memory broadcast.
template <int NLD, int NNB, int P, bool WIDE, bool BND>
__global__ void __launch_bounds__(CB, 3) dg_class_apply(const ClsDev cd, const double *__restrict__ x, double *__restrict__ y,
                                                        SgRed red, double *dot_out, const int *skip, const SgHaloWait hw) {
    extern __shared__ __align__(16) double s_tab[];
    if (skip && *skip) return;
    load_class_tables(cd, s_tab, (cd.n_self + cd.n_nb) * cd.S);
    const double *s_nb = s_tab + cd.n_self * cd.S;
    double dsum[2] = {0.0, 0.0};   // [1] stays 0: slot of the separate exterior-facet kernel (overwritten by it when it runs)
    // ONE grid-stride sweep over idx in [0, cell_hi - cell_lo): the interior cells first (ascending or, rev, descending) and -
    // partitioned mesh - the two boundary strips, the only cells that read ghost rows, at the END of the index space: a
    // block waits for the neighbours' puts only in the iteration that reaches the strips, blocks that never reach them
    // retire (so this rank's own put can always be scheduled), and the strips fill the last partial wave instead of adding
    // a second tail.  Unpartitioned: split == cell range, no strips, no wait.
    const int n_tot = (int)(cd.cell_hi - cd.cell_lo), n_int = (int)(cd.split_hi - cd.split_lo), n_lo = (int)(cd.split_lo - cd.cell_lo);
    bool waited = hw.n == 0;
    for (int base = (int)blockIdx.x * CB; base < n_tot; base += (int)gridDim.x * CB) {
        if (!waited && base + CB > n_int) {
            sg_halo_wait_inline(hw);
            waited = true;
        }
        const int idx = base + (int)threadIdx.x;
        if (idx >= n_tot) continue;
        const int j = idx - n_int;
        const int c = j < 0 ? (cd.rev ? (int)cd.split_hi - 1 - idx : (int)cd.split_lo + idx)
                            : (j < n_lo ? (int)cd.cell_lo + j : (int)cd.split_hi + (j - n_lo));
        double xk[NLD], yk[NLD];
        dg_cell_apply<NLD, NNB, P, WIDE, BND>(cd, s_tab, s_nb, c, x, xk, yk);
        store_row<NLD, WIDE>(y + (size_t)c * NLD, yk);
#pragma unroll
        for (int i = 0; i < NLD; ++i) dsum[0] += xk[i] * yk[i];
    }
    sg_grid_reduce<2>(dsum, red, dot_out);
}

// Residual from the class tables: F_K = (cell + interior-facet part of J) T  -  |detJ| Mhat T_prev  -  dt f |detJ| load.
// The cell part of the Jacobian is linear in T (the nonlinearity sits on the exterior facets, added afterwards by
// bfacet_kernel<MODE_RESID>), so the tables of the apply kernel serve the residual too.
struct ResidDev {
    const double *xprev, *detJ;
    double dt_f;
    double mass[100];   // Mhat, row-major NLD x NLD
    double load[10];
};

template <int NLD>
__device__ __forceinline__ void resid_correction(const ResidDev &rd, const long c, const double (&xp)[NLD], double (&yk)[NLD]) {
    const double dj = rd.detJ[c];
#pragma unroll
    for (int i = 0; i < NLD; ++i) {
        double m = rd.dt_f * rd.load[i];
#pragma unroll
        for (int j = 0; j < NLD; ++j) m += rd.mass[i * NLD + j] * xp[j];
        yk[i] -= dj * m;
    }
}

template <int NLD, int NNB, int P, bool WIDE>
__global__ void __launch_bounds__(CB, 3) dg_class_resid(const ClsDev cd, const __grid_constant__ ResidDev rd, const double *__restrict__ x,
                                                     double *__restrict__ y) {
    extern __shared__ __align__(16) double s_tab[];
    load_class_tables(cd, s_tab, (cd.n_self + cd.n_nb) * cd.S);
    const double *s_nb = s_tab + cd.n_self * cd.S;
    for (long c = cd.cell_lo + (long)blockIdx.x * CB + threadIdx.x; c < cd.cell_hi; c += (long)gridDim.x * CB) {
        double xk[NLD], yk[NLD], xp[NLD];
        dg_cell_apply<NLD, NNB, P, WIDE, false>(cd, s_tab, s_nb, c, x, xk, yk);   // nbr < 0 (incl. encoded exterior facets): no term
        load_row<NLD, WIDE>(rd.xprev + c * NLD, xp);
        resid_correction<NLD>(rd, c, xp, yk);
        store_row<NLD, WIDE>(y + c * NLD, yk);
    }
}

template <int NLD>
__global__ void __launch_bounds__(CB, 4) cg_class_resid(const ClsDev cd, const __grid_constant__ ResidDev rd, const double *__restrict__ x,
                                                     double *__restrict__ y) {
    extern __shared__ __align__(16) double s_tab[];
    load_class_tables(cd, s_tab, cd.n_self * cd.S);
    const long nc = cd.n_cells;
    for (long c = cd.cell_lo + (long)blockIdx.x * CB + threadIdx.x; c < cd.cell_hi; c += (long)gridDim.x * CB) {
        const double *As = s_tab + (int)cd.cls16[c] * cd.S;
        int dof[NLD];
        double xk[NLD], xp[NLD], yk[NLD];
#pragma unroll
        for (int i = 0; i < NLD; ++i) dof[i] = cd.dofmap[(long)i * nc + c];
#pragma unroll
        for (int i = 0; i < NLD; ++i) {
            xk[i] = x[dof[i]];
            xp[i] = rd.xprev[dof[i]];
            yk[i] = 0.0;
        }
        smem_matvec_acc<NLD>(As, xk, yk);
        resid_correction<NLD>(rd, c, xp, yk);
#pragma unroll
        for (int i = 0; i < NLD; ++i) atomicAdd(&y[dof[i]], yk[i]);
    }
}

// One step of the Chebyshev iteration for  (M^-1 J) z = M^-1 r  (M = block-diagonal element mass matrix), the
// polynomial preconditioner of the DG solver (pcg.cu), in three-term form
//     z_new = z + a (z - z_prev) + b M^-1 (r - J z)
// fused into the operator apply: J z never goes to memory.  z_new overwrites z_prev (only the cell itself reads
// z_prev, the neighbours read z), so two buffers alternate.  FIRST: z_prev = 0 is not read and z_new goes to the
// other buffer.  LAST: r.z_new is reduced into dot_out[0].
struct ChebDev {
    const double *r, *z_prev, *detJ;
    double *z_out;
    double a, b;
    double minv[100];   // Mhat^-1, row-major NLD x NLD
};

template <int NLD, bool WIDE, bool FIRST, bool LAST>
__device__ __forceinline__ void cheb_update(const ChebDev &ch, const int c, const double (&zk)[NLD], double (&Jz)[NLD], double &dsum) {
    double rk[NLD], zp[NLD];
    load_row<NLD, WIDE>(ch.r + (size_t)c * NLD, rk);
    if (!FIRST) load_row<NLD, WIDE>(ch.z_prev + (size_t)c * NLD, zp);
    const double idet = ch.b / ch.detJ[c];
#pragma unroll
    for (int i = 0; i < NLD; ++i) Jz[i] = rk[i] - Jz[i];
#pragma unroll
    for (int i = 0; i < NLD; ++i) {
        double m = 0.0;
#pragma unroll
        for (int j = 0; j < NLD; ++j) m += ch.minv[i * NLD + j] * Jz[j];
        const double dprev = FIRST ? zk[i] : zk[i] - zp[i];
        zp[i] = zk[i] + (ch.a * dprev + idet * m);
    }
    store_row<NLD, WIDE>(ch.z_out + (size_t)c * NLD, zp);
    if (LAST) {
#pragma unroll
        for (int i = 0; i < NLD; ++i) dsum += rk[i] * zp[i];
    }
}

template <int NLD, int NNB, int P, bool WIDE, bool BND, bool FIRST, bool LAST>
__global__ void __launch_bounds__(CB, 3) dg_cheb_step(const ClsDev cd, const __grid_constant__ ChebDev ch, const double *__restrict__ z,
                                                      SgRed red, double *dot_out, const int *skip, const SgHaloWait hw) {
    extern __shared__ __align__(16) double s_tab[];
    if (skip && *skip) return;
    load_class_tables(cd, s_tab, (cd.n_self + cd.n_nb) * cd.S);
    const double *s_nb = s_tab + cd.n_self * cd.S;
    double dsum[1] = {0.0};
    // ONE grid-stride sweep over idx in [0, cell_hi - cell_lo): the interior cells first (ascending or, rev, descending) and -
    // partitioned mesh - the two boundary strips, the only cells that read ghost rows, at the END of the index space: a
    // block waits for the neighbours' puts only in the iteration that reaches the strips, blocks that never reach them
    // retire (so this rank's own put can always be scheduled), and the strips fill the last partial wave instead of adding
    // a second tail.  Unpartitioned: split == cell range, no strips, no wait.
    const int n_tot = (int)(cd.cell_hi - cd.cell_lo), n_int = (int)(cd.split_hi - cd.split_lo), n_lo = (int)(cd.split_lo - cd.cell_lo);
    bool waited = hw.n == 0;
    for (int base = (int)blockIdx.x * CB; base < n_tot; base += (int)gridDim.x * CB) {
        if (!waited && base + CB > n_int) {
            sg_halo_wait_inline(hw);
            waited = true;
        }
        const int idx = base + (int)threadIdx.x;
        if (idx >= n_tot) continue;
        const int j = idx - n_int;
        const int c = j < 0 ? (cd.rev ? (int)cd.split_hi - 1 - idx : (int)cd.split_lo + idx)
                            : (j < n_lo ? (int)cd.cell_lo + j : (int)cd.split_hi + (j - n_lo));
        double zk[NLD], Jz[NLD];
        dg_cell_apply<NLD, NNB, P, WIDE, BND>(cd, s_tab, s_nb, c, z, zk, Jz);
        cheb_update<NLD, WIDE, FIRST, LAST>(ch, c, zk, Jz, dsum[0]);
    }
    if (LAST) sg_grid_reduce<1>(dsum, red, dot_out);
}

// Tried and rejected (measured, profiles/r2t_dmma_experiment_*.json): running the local products of class-uniform 32-cell
// tiles on the FP64 tensor cores (mma.sync.m8n8k4.f64, X^T tile as the A operand straight from a coalesced 8-byte load,
// the class matrix as a 16-lane B fragment: one LDS.64 per matrix per 32 cells instead of eight LDS.128 per cell).  Results
// identical to 1e-14, but the fused Chebyshev step took 486 us instead of 163 us and the apply 286 instead of 133 us on
// config 3: DMMA issue on sm_100a is far below the FP64 FMA pipe for these 8x8x4 products, so the shared-memory table
// reads stay.  What the experiment left behind is the layer-grouped cell order of mesh.py (163 vs 171 us per step).

// Exterior facets of a CG space from their linearised matrices: y += B_F x_F (RED.ADD), dsum += x_F . (B_F x_F) over the
// facets of the cells [dot_lo, dot_hi).
template <int D, int P>
__device__ __forceinline__ void cg_bfacets(const ClsDev &cd, const double *__restrict__ x, double *__restrict__ y, double &dsum) {
    constexpr int NFD = nfd_of(D, P), NFDP = NFD * (NFD + 1) / 2;
    const long nc = cd.n_cells;
    for (long b = (long)blockIdx.x * CB + threadIdx.x; b < cd.n_bf; b += (long)gridDim.x * CB) {
        const long c = cd.bf_cell[b];
        if (c < cd.cell_lo || c >= cd.cell_hi) continue;
        const int f = cd.bf_facet[b];
        long dof[NFD];
        double xk[NFD], yk[NFD], B[NFDP];
#pragma unroll
        for (int k = 0; k < NFD; ++k) {
            int fd = 0;
#pragma unroll
            for (int ff = 0; ff < D + 1; ++ff)
                if (ff == f) fd = facet_dof(D, P, ff, k);
            dof[k] = cd.dofmap[(long)fd * nc + c];
            xk[k] = x[dof[k]];
            yk[k] = 0.0;
        }
#pragma unroll
        for (int k = 0; k < NFDP; ++k) B[k] = cd.bmat[b * NFDP + k];
        int m = 0;
#pragma unroll
        for (int k = 0; k < NFD; ++k)
#pragma unroll
            for (int l = k; l < NFD; ++l) {
                yk[k] += B[m] * xk[l];
                if (l != k) yk[l] += B[m] * xk[k];
                ++m;
            }
        const bool counted = c >= cd.dot_lo && c < cd.dot_hi;
#pragma unroll
        for (int k = 0; k < NFD; ++k) {
            atomicAdd(&y[dof[k]], yk[k]);
            if (counted) dsum += xk[k] * yk[k];
        }
    }
}

// CG fast apply: gather through the dofmap, class matrix from shared memory, scatter with RED.ADD.F64; then, in the
// same launch, the exterior facets from their linearised matrices (both parts only add into y, no ordering needed).
// x.y is reduced cell-wise as x_K . (A_K x_K) over the cells [dot_lo, dot_hi) (each global cell on one rank) plus
// x_F . (B_F x_F) over their exterior facets.
template <int D, int P>
__global__ void __launch_bounds__(CB, 4) cg_class_apply(const ClsDev cd, const double *__restrict__ x, double *__restrict__ y,
                                                     SgRed red, double *dot_out, const int *skip) {
    constexpr int NLD = nld_of(D, P);
    extern __shared__ __align__(16) double s_tab[];
    if (skip && *skip) return;
    const int ntab = cd.n_self * cd.S;
    for (int i = threadIdx.x; i < ntab; i += CB) s_tab[i] = cd.tab[i];
    __syncthreads();
    const long nc = cd.n_cells;
    double dsum[2] = {0.0, 0.0};
    for (long c = cd.cell_lo + (long)blockIdx.x * CB + threadIdx.x; c < cd.cell_hi; c += (long)gridDim.x * CB) {
        const double *As = s_tab + (int)cd.cls16[c] * cd.S;
        int dof[NLD];
        double xk[NLD];
#pragma unroll
        for (int i = 0; i < NLD; ++i) dof[i] = cd.dofmap[(long)i * nc + c];
#pragma unroll
        for (int i = 0; i < NLD; ++i) xk[i] = x[dof[i]];
        double yk[NLD], d = 0.0;
#pragma unroll
        for (int i = 0; i < NLD; ++i) yk[i] = 0.0;
        smem_matvec_acc<NLD>(As, xk, yk);
#pragma unroll
        for (int i = 0; i < NLD; ++i) {
            atomicAdd(&y[dof[i]], yk[i]);
            d += xk[i] * yk[i];
        }
        if (c >= cd.dot_lo && c < cd.dot_hi) dsum[0] += d;
    }
    if (cd.bmat) cg_bfacets<D, P>(cd, x, y, dsum[1]);
    sg_grid_reduce<2>(dsum, red, dot_out);
}

// The exterior facets on their own, after the row-stencil apply (stencil.cu) has stored the cell part of y.
template <int D, int P>
__global__ void __launch_bounds__(CB) cg_bfacet_apply(const ClsDev cd, const double *__restrict__ x, double *__restrict__ y, SgRed red,
                                                      double *dot_out, const int *skip) {
    if (skip && *skip) return;
    double dsum[1] = {0.0};
    cg_bfacets<D, P>(cd, x, y, dsum[0]);
    sg_grid_reduce<1>(dsum, red, dot_out);
}

// set-up check of the row-stencil tables: x_i = hash(i) in [-1, 1); out[0] = max |a|, out[1] = max |a - b| (as bit patterns)
__global__ void k_hash_vec(long n, double *x) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned long long h = (unsigned long long)i * 0x9E3779B97F4A7C15ull + 0x7F4A7C15ull;
    h ^= h >> 32;
    h *= 0xD6E8FEB86659FD93ull;
    h ^= h >> 32;
    x[i] = (double)(h >> 11) * (2.0 / 9007199254740992.0) - 1.0;
}
__global__ void k_max_diff(long n, const double *__restrict__ a, const double *__restrict__ b, unsigned long long *out) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    atomicMax(&out[0], (unsigned long long)__double_as_longlong(fabs(a[i])));
    atomicMax(&out[1], (unsigned long long)__double_as_longlong(fabs(a[i] - b[i])));
}

}  // namespace

// ===================================================================================== host side
struct sg_thermal_op {
    sg_ctx *ctx;
    sg_thermal_desc d;
    OpDev dev;
    void *tab_host;      // Tab<D,P,DG> instance
    size_t tab_bytes;
    double *btab_dev, *bw_dev;
    double mass_inv[100], mass[100], load[10];
    // local-matrix classes (fast apply path); cls.tab == nullptr when not available
    ClsDev cls;
    void *cls_words;       // cls64 / cls16 storage
    double *cls_tab;
    int cls_grid;
    size_t cls_smem;
    int32_t n_geom_classes;
    SgRed own_red;         // reduction scratch of sg_thermal_jac_apply (solver-less use of the fast path)
    const SgHaloWait *wait; // set by sg_thermal_apply_dot: ghost rows of x in flight (NULL: none)
    int y_is_zero;         // set by sg_thermal_apply_dot: the caller guarantees y == 0 on entry (CG scatter needs no memset)
    int32_t *nbr_ext;      // DG P1: neighbour ids with exterior facets encoded (see ClsDev::bmat)
    double *bmat;
    SgStencil *stencil;    // CG: row-stencil classes (stencil.cu); nullptr: cell-centric cg_class_apply
    int stencil_bnd;       // the stencil kernel also applies the exterior facets (gather form; no cg_bfacet_apply launch)
    SgStencil *stencil_mass;   // CG: the same row classes for |detJ| Mhat: residual in gather form, F_cells = S_J T - S_M T_prev
    double *mass_tab;          // its class tables (device)
    double *diag_cells;    // CG: the cell part of diag J (M + dt alpha K does not depend on T), computed once at creation
    int (*linearize)(const sg_thermal_op *, const double *T_lin, cudaStream_t st);
    int (*cheb_step)(const sg_thermal_op *, const SgChebStep &, SgRed red, double *dot_out, const int *skip, cudaStream_t st);
    int (*small_pcg)(sg_thermal_op *, const double *b, double *x, const SgPcgPolicy &pol, int max_it, double *rr0_out, int *ctrl_done,
                     int *ctrl_iters, double *ctrl_rr, cudaStream_t st);
    // optional profiling of the Jacobian-apply cell kernel (bench.py roofline): event pairs on the launch stream
    int prof_on, prof_n, prof_cap;
    cudaEvent_t *prof_ev;
    unsigned char *prof_kind;   // per recorded launch: 0 = Jacobian-apply cell kernel, 1 = fused Chebyshev step
    int (*launch)(const sg_thermal_op *, int mode, const double *T, const double *x, const double *xprev, double *y,
                  SgRed red, double *dot2, const int *skip, cudaStream_t st);
    int (*build_classes)(sg_thermal_op *);
};

namespace {

inline bool inkernel_wait() {
    static const bool on = [] {
        const char *e = getenv("SG_NO_INKERNEL_WAIT");
        return !(e && e[0] == '1');
    }();
    return on;
}

// rows from which the CG residual runs in gather form (SG_GATHER_RESID_MIN_ROWS overrides: the tests set it to 0)
static const long GATHER_RESID_MIN_ROWS = getenv("SG_GATHER_RESID_MIN_ROWS") ? atol(getenv("SG_GATHER_RESID_MIN_ROWS")) : 1000000;
inline bool gather_resid_off() {      // SG_NO_GATHER_RESID=1: keep cg_class_resid (measurement / test switch)
    static const bool off = [] {
        const char *e = getenv("SG_NO_GATHER_RESID");
        return e && e[0] == '1';
    }();
    return off;
}

inline unsigned capped_grid(long n, int tb) {
    long g = (n + tb - 1) / tb;
    if (g < 1) g = 1;
    return (unsigned)(g > SG_MAX_BLOCKS ? SG_MAX_BLOCKS : g);
}

struct ProfScope {  // CUDA-event pair around the apply cell kernel when profiling is on
    const sg_thermal_op *op;
    cudaStream_t st;
    bool on;
    ProfScope(const sg_thermal_op *o, int mode, cudaStream_t s, int kind = 0) : op(o), st(s) {
        on = mode == MODE_APPLY && o->prof_on && o->prof_n < o->prof_cap;
        if (on) {
            o->prof_kind[o->prof_n] = (unsigned char)kind;
            cudaEventRecord(o->prof_ev[2 * o->prof_n], st);
        }
    }
    ~ProfScope() {
        if (on) {
            cudaEventRecord(op->prof_ev[2 * op->prof_n + 1], st);
            const_cast<sg_thermal_op *>(op)->prof_n++;
        }
    }
};

// dot2 != nullptr (MODE_APPLY only): also produce dot2[0] + dot2[1] = x.y over the owned dofs.
template <int D, int P, bool DG>
int launch_op(const sg_thermal_op *op, int mode, const double *Tlin, const double *x, const double *xprev, double *y,
              SgRed red, double *dot2, const int *skip, cudaStream_t st) {
    using T = Tab<D, P, DG>;
    constexpr int NLD = T::NLD;
    const T &tab = *static_cast<const T *>(op->tab_host);
    const OpDev &dv = op->dev;
    const long ncell = dv.cell_hi - dv.cell_lo;
    const unsigned gc = (unsigned)((ncell + TB - 1) / TB), gb = capped_grid(dv.n_bf, TB);
    const bool fast = mode == MODE_APPLY && op->cls.tab != nullptr;
    const bool cached_diag = !DG && mode == MODE_DIAG && op->diag_cells != nullptr && y != op->diag_cells;
    const bool gather_resid = !DG && mode == MODE_RESID && op->cls.tab != nullptr && !(op->d.flags & SG_THERMAL_GENERAL_RESIDUAL) &&
                              op->stencil && op->stencil_mass && dv.dt_f == 0.0 && !gather_resid_off() &&
                              dv.n_dofs >= GATHER_RESID_MIN_ROWS;   // two ~row-kernel launches: below, the one scatter launch wins
    if (!DG && !cached_diag && !gather_resid && !(mode == MODE_APPLY && (op->y_is_zero || (fast && op->stencil))))
        SG_CHECK_CUDA(cudaMemsetAsync(y, 0, sizeof(double) * (size_t)dv.n_dofs, st));
    if (cached_diag) {
        // point Jacobi asks for diag J(T) in every Newton iteration; only the exterior facets (below) depend on T
        SG_CHECK_CUDA(cudaMemcpyAsync(y, op->diag_cells, sizeof(double) * (size_t)dv.n_dofs, cudaMemcpyDeviceToDevice, st));
    } else if (fast) {
        // the class kernels always reduce x.y; without a consumer it lands in a scratch slot
        double *dst = dot2 ? dot2 : red.partials + 2 * SG_MAX_BLOCKS;
        {
        int rc_wait = SG_OK;
        (void)rc_wait;
        ProfScope ps(op, mode, st);
        const bool wide = (((uintptr_t)x | (uintptr_t)y) & 31) == 0;
        SgHaloWait none{}, hw = op->wait ? *op->wait : none;
        if (hw.n && !inkernel_wait()) {     // measurement switch: one-warp wait kernel + the plain kernel
            if ((rc_wait = sg_peer_wait(nullptr, hw, st))) return rc_wait;
            hw = none;
        }
        if constexpr (DG) {
            auto k = op->bmat ? (wide ? dg_class_apply<NLD, D + 1, P, true, true> : dg_class_apply<NLD, D + 1, P, false, true>)
                              : (wide ? dg_class_apply<NLD, D + 1, P, true, false> : dg_class_apply<NLD, D + 1, P, false, false>);
            ClsDev cdv = op->cls;
            cdv.rev = dot2 ? sg_next_sweep_dir() : 0;     // solver launches alternate the sweep direction
            k<<<op->cls_grid, CB, op->cls_smem, st>>>(cdv, x, y, red, dst, skip, hw);
        } else if (op->stencil) {
            // gather form: plain stores of every row, no zeroing of y needed.  With exterior facets following in their own
            // launch, the cross-rank sum of both parts of x.Ax is done by THAT kernel (it covers dst[0] and dst[1]).
            const bool bf_follows = op->bmat && dv.n_bf > 0 && !op->stencil_bnd;
            const int rc = sg_stencil_apply(op->stencil, x, y, op->d.own_lo, op->d.own_hi, bf_follows ? sg_red_local(red) : red, dst, skip, st,
                                            &hw);
            if (rc) return rc;
        } else {
            if (hw.n && (rc_wait = sg_peer_wait(nullptr, hw, st))) return rc_wait;
            cg_class_apply<D, P><<<op->cls_grid, CB, op->cls_smem, st>>>(op->cls, x, y, red, dst, skip);
        }
        SG_CHECK_CUDA(cudaGetLastError());
        if (DG || !op->stencil) sg_count_launch();
        }
        if constexpr (!DG) {
            if (op->stencil && op->bmat && dv.n_bf > 0 && !op->stencil_bnd) {
                SgRed rb = red;
                if (rb.peer) {
                    rb.ar_ptr = dst;
                    rb.ar_count = 2;
                }
                cg_bfacet_apply<D, P><<<capped_grid(dv.n_bf, CB), CB, 0, st>>>(op->cls, x, y, rb, dst + 1, skip);
                SG_CHECK_CUDA(cudaGetLastError());
                sg_count_launch();
            }
        }
    } else if (mode == MODE_RESID && op->cls.tab != nullptr && !(op->d.flags & SG_THERMAL_GENERAL_RESIDUAL)) {
        ResidDev rd;
        rd.xprev = xprev;
        rd.detJ = op->d.geom + (int64_t)D * D * op->d.n_cells;
        rd.dt_f = dv.dt_f;
        memcpy(rd.mass, op->mass, sizeof(double) * NLD * NLD);
        memcpy(rd.load, op->load, sizeof(double) * NLD);
        if constexpr (DG) {
            const bool wide = (((uintptr_t)x | (uintptr_t)y | (uintptr_t)xprev) & 31) == 0;
            auto k = wide ? dg_class_resid<NLD, D + 1, P, true> : dg_class_resid<NLD, D + 1, P, false>;
            SG_CHECK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)op->cls_smem));
            k<<<op->cls_grid, CB, op->cls_smem, st>>>(op->cls, rd, x, y);
        } else if (gather_resid) {
            // gather form: F_cells = S_J T - S_M T_prev, plain stores, no atomics (f = 0: no load vector)
            double *scr = op->own_red.partials + 2 * SG_MAX_BLOCKS;
            int rcg = sg_stencil_apply_cells(op->stencil, x, y, 0, op->own_red, scr, st);
            if (!rcg) rcg = sg_stencil_apply_cells(op->stencil_mass, xprev, y, 1, op->own_red, scr, st);
            if (rcg) return rcg;
        } else {
            SG_CHECK_CUDA(cudaFuncSetAttribute(cg_class_resid<NLD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)op->cls_smem));
            cg_class_resid<NLD><<<op->cls_grid, CB, op->cls_smem, st>>>(op->cls, rd, x, y);
            sg_count_launch();
        }
        SG_CHECK_CUDA(cudaGetLastError());
        if (DG) sg_count_launch();
    } else if (ncell > 0) {
        if (mode == MODE_APPLY && op->wait && op->wait->n) {
            const int rcw = sg_peer_wait(nullptr, *op->wait, st);
            if (rcw) return rcw;
        }
        ProfScope ps(op, mode, st);
        if (mode == MODE_APPLY) cell_kernel<D, P, DG, MODE_APPLY><<<gc, TB, 0, st>>>(tab, dv, x, nullptr, y);
        if (mode == MODE_RESID) cell_kernel<D, P, DG, MODE_RESID><<<gc, TB, 0, st>>>(tab, dv, x, xprev, y);
        if (mode == MODE_DIAG) cell_kernel<D, P, DG, MODE_DIAG><<<gc, TB, 0, st>>>(tab, dv, nullptr, nullptr, y);
        SG_CHECK_CUDA(cudaGetLastError());
        sg_count_launch();
    }
    const bool bdot = fast && dot2;
    if (fast && op->bmat && DG) {
        // exterior facets were applied inside the class kernel
    } else if (fast && op->bmat) {
        // CG: cg_class_apply applied the exterior facets in the same launch
    } else if (dv.n_bf > 0 || bdot) {
        if (mode == MODE_APPLY && bdot) bfacet_kernel<D, P, DG, MODE_APPLY, true><<<gb, TB, 0, st>>>(dv, Tlin, x, y, sg_red_local(red), dot2 + 1, skip);
        if (mode == MODE_APPLY && !bdot) bfacet_kernel<D, P, DG, MODE_APPLY, false><<<gb, TB, 0, st>>>(dv, Tlin, x, y, red, nullptr, skip);
        if (mode == MODE_RESID) bfacet_kernel<D, P, DG, MODE_RESID, false><<<gb, TB, 0, st>>>(dv, x, nullptr, y, red, nullptr, nullptr);
        if (mode == MODE_DIAG) bfacet_kernel<D, P, DG, MODE_DIAG, false><<<gb, TB, 0, st>>>(dv, Tlin, nullptr, y, red, nullptr, nullptr);
        SG_CHECK_CUDA(cudaGetLastError());
        sg_count_launch();
    }
    if (dot2 && !fast) {
        const long lo = op->d.own_lo, hi = op->d.own_hi;
        k_dot_range<<<capped_grid(hi - lo, 256), 256, 0, st>>>(lo, hi, x, y, red, dot2, skip);
        SG_CHECK_CUDA(cudaGetLastError());
        sg_count_launch();
    }
    return SG_OK;
}

// Refresh the per-facet boundary matrices for a new linearisation point (no-op unless they are in use).
template <int D, int P, bool DG>
int linearize_t(const sg_thermal_op *op, const double *T_lin, cudaStream_t st) {
    if (op->bmat && op->dev.n_bf > 0) {
        k_bfacet_mats<D, P, DG><<<(unsigned)((op->dev.n_bf + TB - 1) / TB), TB, 0, st>>>(op->dev, T_lin, op->bmat);
        SG_CHECK_CUDA(cudaGetLastError());
        sg_count_launch();
        if (op->stencil && op->stencil_bnd) return sg_stencil_refresh_boundary(op->stencil, st);
    }
    return SG_OK;
}

template <int D, int P, bool DG>
int cheb_step_t(const sg_thermal_op *op, const SgChebStep &cs, SgRed red, double *dot_out, const int *skip, cudaStream_t st) {
    if constexpr (DG) {
        constexpr int NLD = nld_of(D, P), NNB = D + 1;
        SG_REQUIRE(op->cls.tab, "Chebyshev step needs the class tables");
        ChebDev ch;
        ch.r = cs.r;
        ch.z_prev = cs.z_prev;
        ch.z_out = cs.z_out;
        ch.detJ = op->d.geom + (int64_t)D * D * op->d.n_cells;
        ch.a = cs.a;
        ch.b = cs.b;
        for (int i = 0; i < NLD * NLD; ++i) ch.minv[i] = op->mass_inv[i];
        const uintptr_t al = (uintptr_t)cs.z_in | (uintptr_t)cs.r | (uintptr_t)cs.z_prev | (uintptr_t)cs.z_out;
        const bool wide = (al & 31) == 0, bnd = op->bmat != nullptr, first = cs.z_prev == nullptr, last = cs.last != 0;
        using K = void (*)(const ClsDev, const ChebDev, const double *, SgRed, double *, const int *, const SgHaloWait);
        // [wide][bnd][first][last]
        static const K table[2][2][2][2] = {
            {{{dg_cheb_step<NLD, NNB, P, false, false, false, false>, dg_cheb_step<NLD, NNB, P, false, false, false, true>},
              {dg_cheb_step<NLD, NNB, P, false, false, true, false>, dg_cheb_step<NLD, NNB, P, false, false, true, true>}},
             {{dg_cheb_step<NLD, NNB, P, false, true, false, false>, dg_cheb_step<NLD, NNB, P, false, true, false, true>},
              {dg_cheb_step<NLD, NNB, P, false, true, true, false>, dg_cheb_step<NLD, NNB, P, false, true, true, true>}}},
            {{{dg_cheb_step<NLD, NNB, P, true, false, false, false>, dg_cheb_step<NLD, NNB, P, true, false, false, true>},
              {dg_cheb_step<NLD, NNB, P, true, false, true, false>, dg_cheb_step<NLD, NNB, P, true, false, true, true>}},
             {{dg_cheb_step<NLD, NNB, P, true, true, false, false>, dg_cheb_step<NLD, NNB, P, true, true, false, true>},
              {dg_cheb_step<NLD, NNB, P, true, true, true, false>, dg_cheb_step<NLD, NNB, P, true, true, true, true>}}}};
        SgHaloWait none{}, hw = cs.wait ? *cs.wait : none;
        if (hw.n && !inkernel_wait()) {
            const int rcw = sg_peer_wait(nullptr, hw, st);
            if (rcw) return rcw;
            hw = none;
        }
        const K k = table[wide][bnd][first][last];
        SG_CHECK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)op->cls_smem));
        {
            ProfScope ps(op, MODE_APPLY, st, 1);
            ClsDev cdv = op->cls;
            cdv.rev = sg_next_sweep_dir();
            k<<<op->cls_grid, CB, op->cls_smem, st>>>(cdv, ch, cs.z_in, red, dot_out, skip, hw);
        }
        SG_CHECK_CUDA(cudaGetLastError());
        sg_count_launch();
        return SG_OK;
    } else {
        sg_set_error("Chebyshev step is only available for DG spaces");
        return SG_E_UNSUPPORTED;
    }
}

// ---- tiny DG problems (config 1: the 48-cell line of main.py): the WHOLE block-Jacobi PCG solve in one block ----------------
// With a few hundred cells every kernel of the multi-launch iteration is pure launch latency (354 launches per time step on
// main.py's mesh even as CUDA-graph batches).  One thread per cell keeps x, r, p of its cell in registers; p travels
// through shared memory so that dg_cell_apply (the class-table apply, unchanged) reads the neighbours' rows from there;
// the three reductions of an iteration are block-wide sums in a fixed order.  Same arithmetic per cell as k_pcg_init_blk /
// k_update_xr_blk / k_update_p_blk (csrc/pcg.cu); the tolerance policy of the inexact Newton iteration runs on the device.
constexpr int SMALL_DG_MAX_CELLS = 256;

template <int NLD>
struct SmallPcgArgs {
    const double *b, *detJ;
    double *x;
    double minv[NLD * NLD];
    SgPcgPolicy pol;
    int max_it;
    double *rr0_out;
    int *ctrl_done, *ctrl_iters;
    double *ctrl_rr;
};

// sum of v over the block, identical in every thread (fixed order: warp shuffles, then the warps' partials)
__device__ __forceinline__ double small_block_sum(double v, double *scratch) {
    v = sg_warp_sum(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    __syncthreads();                       // the previous sum's readers are done with scratch
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    double w = lane < nw ? scratch[lane] : 0.0;
    return sg_warp_sum(w);
}

template <int NLD, int NNB, int P, bool BND>
__global__ void __launch_bounds__(SMALL_DG_MAX_CELLS) k_dg_pcg_small(const ClsDev cd, const __grid_constant__ SmallPcgArgs<NLD> a) {
    extern __shared__ __align__(16) double s_tab[];
    const int ntab = (cd.n_self + cd.n_nb) * cd.S;
    for (int i = threadIdx.x; i < ntab; i += blockDim.x) s_tab[i] = cd.tab[i];
    const double *s_nb = s_tab + cd.n_self * cd.S;
    double *s_p = s_tab + ((ntab + 1) & ~1);                 // [n_cells][NLD], 16-byte aligned rows
    double *scratch = s_p + (size_t)cd.n_cells * NLD;
    const int c = threadIdx.x;
    const bool in = c < (int)cd.n_cells;
    double x[NLD], r[NLD], p[NLD], z[NLD];
    const double inv_det = in ? 1.0 / a.detJ[c] : 0.0;
    auto mass_solve = [&](const double (&rv)[NLD], double (&zv)[NLD]) {
#pragma unroll
        for (int i = 0; i < NLD; ++i) {
            double s = 0.0;
#pragma unroll
            for (int j = 0; j < NLD; ++j) s += a.minv[i * NLD + j] * rv[j];
            zv[i] = s * inv_det;
        }
    };
#pragma unroll
    for (int i = 0; i < NLD; ++i) {
        x[i] = 0.0;
        r[i] = in ? a.b[(size_t)c * NLD + i] : 0.0;
    }
    mass_solve(r, z);
    double lrz = 0.0, lrr = 0.0;
#pragma unroll
    for (int i = 0; i < NLD; ++i) {
        p[i] = z[i];
        lrz += r[i] * z[i];
        lrr += r[i] * r[i];
    }
    double rz = small_block_sum(lrz, scratch), rr = small_block_sum(lrr, scratch);
    const double rr0 = rr, tol2 = a.pol.tol2(rr0);
    int it = 0, done = 0;
    for (;;) {
        if (!(rr > tol2) || !isfinite(rr)) {
            done = isfinite(rr) ? 1 : 2;
            break;
        }
        if (it >= a.max_it) break;
        __syncthreads();                                     // everyone has read the previous p rows
        if (in) {
#pragma unroll
            for (int i = 0; i < NLD; ++i) s_p[(size_t)c * NLD + i] = p[i];
        }
        __syncthreads();
        double xk[NLD], Ap[NLD];
#pragma unroll
        for (int i = 0; i < NLD; ++i) Ap[i] = 0.0;
        if (in) dg_cell_apply<NLD, NNB, P, false, BND>(cd, s_tab, s_nb, c, s_p, xk, Ap);
        double lpAp = 0.0;
#pragma unroll
        for (int i = 0; i < NLD; ++i) lpAp += p[i] * Ap[i];
        const double pAp = small_block_sum(lpAp, scratch);   // may be negative: the reference's penalty is not coercive for every
        const double alpha = rz / pAp;                       // element; like the multi-kernel path, CG just carries on
#pragma unroll
        for (int i = 0; i < NLD; ++i) {
            x[i] += alpha * p[i];
            r[i] -= alpha * Ap[i];
        }
        mass_solve(r, z);
        lrz = 0.0;
        lrr = 0.0;
#pragma unroll
        for (int i = 0; i < NLD; ++i) {
            lrz += r[i] * z[i];
            lrr += r[i] * r[i];
        }
        const double rz_new = small_block_sum(lrz, scratch);
        rr = small_block_sum(lrr, scratch);
        const double beta = rz_new / rz;
        rz = rz_new;
#pragma unroll
        for (int i = 0; i < NLD; ++i) p[i] = z[i] + beta * p[i];
        ++it;
    }
    if (in) {
#pragma unroll
        for (int i = 0; i < NLD; ++i) a.x[(size_t)c * NLD + i] = x[i];
    }
    if (threadIdx.x == 0) {
        *a.rr0_out = rr0;
        *a.ctrl_done = done;
        *a.ctrl_iters = it;
        *a.ctrl_rr = rr;
    }
}

template <int D, int P, bool DG>
int small_pcg_t(sg_thermal_op *op, const double *b, double *x, const SgPcgPolicy &pol, int max_it, double *rr0_out, int *ctrl_done,
                int *ctrl_iters, double *ctrl_rr, cudaStream_t st) {
    if constexpr (DG) {
        constexpr int NLD = nld_of(D, P), NNB = D + 1;
        const long nc = op->d.n_cells;
        if (!op->cls.tab || nc > SMALL_DG_MAX_CELLS || op->cls.cell_lo != 0 || op->cls.cell_hi != nc) return 0;
        const int ntab = (op->cls.n_self + op->cls.n_nb) * op->cls.S;
        const size_t smem = sizeof(double) * (((size_t)ntab + 1) / 2 * 2 + (size_t)nc * NLD + 32);
        if (smem > 200 * 1024) return 0;
        SmallPcgArgs<NLD> a;
        a.b = b;
        a.detJ = op->d.geom + (int64_t)D * D * op->d.n_cells;
        a.x = x;
        for (int i = 0; i < NLD * NLD; ++i) a.minv[i] = op->mass_inv[i];
        a.pol = pol;
        a.max_it = max_it;
        a.rr0_out = rr0_out;
        a.ctrl_done = ctrl_done;
        a.ctrl_iters = ctrl_iters;
        a.ctrl_rr = ctrl_rr;
        auto k = op->bmat ? k_dg_pcg_small<NLD, NNB, P, true> : k_dg_pcg_small<NLD, NNB, P, false>;
        SG_CHECK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const int threads = (int)((nc + 31) / 32 * 32);
        ClsDev cdv = op->cls;
        cdv.rev = 0;
        k<<<1, threads, smem, st>>>(cdv, a);
        SG_CHECK_CUDA(cudaGetLastError());
        sg_count_launch();
        return 1;
    } else {
        return 0;
    }
}

struct DevBuf {  // scoped cudaMalloc
    void *p = nullptr;
    ~DevBuf() {
        if (p) cudaFree(p);
    }
    template <class T>
    T *as() { return static_cast<T *>(p); }
};

int ensure_own_red(sg_thermal_op *op) {
    if (!op->own_red.partials) {
        SG_CHECK_CUDA(cudaMalloc(&op->own_red.partials, sizeof(double) * (2 * SG_MAX_BLOCKS + 2)));
        SG_CHECK_CUDA(cudaMalloc(&op->own_red.counter, sizeof(unsigned)));
        SG_CHECK_CUDA(cudaMemset(op->own_red.counter, 0, sizeof(unsigned)));
    }
    return SG_OK;
}

// CG: try the row-stencil form of the apply (stencil.cu) and keep it only if it reproduces the cell part of
// cg_class_apply on a pseudo-random vector (guards the 64-bit row hash and the table construction).
// one representative cell per geometry class (any member: the class mass matrix only needs its |detJ|)
__global__ void k_class_rep(long nc, const uint16_t *__restrict__ cls16, int32_t *rep) {
    const long c = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c < nc) rep[cls16[c]] = (int32_t)c;
}
template <int NLD>
struct MassTabArgs {
    double mass[NLD * NLD];
};
template <int NLD>
__global__ void k_mass_tables(const __grid_constant__ MassTabArgs<NLD> ma, int n_cls, int S, const int32_t *__restrict__ rep,
                              const double *__restrict__ detJ, double *__restrict__ tab) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_cls) return;
    const double dj = detJ[rep[k]];
    for (int e = 0; e < S; ++e) tab[(long)k * S + e] = e < NLD * NLD ? dj * ma.mass[e] : 0.0;
}

template <int D, int P>
int build_stencil_t(sg_thermal_op *op) {
    constexpr int NLD = nld_of(D, P);
    const OpDev &dv = op->dev;
    SgStencil *st = nullptr;
    int rc = sg_stencil_build(op->ctx, dv.dofmap, dv.n_cells, NLD, dv.cell_lo, dv.cell_hi, op->cls.cls16, op->cls_tab, op->cls.S,
                              dv.n_dofs, &st);
    if (rc || !st) return rc;
    if ((rc = ensure_own_red(op))) {
        sg_stencil_destroy(st);
        return rc;
    }
    const long n = dv.n_dofs;
    DevBuf xb, ya, yb, mx;
    cudaError_t e = cudaMalloc(&xb.p, sizeof(double) * (size_t)n);
    if (e == cudaSuccess) e = cudaMalloc(&ya.p, sizeof(double) * (size_t)n);
    if (e == cudaSuccess) e = cudaMalloc(&yb.p, sizeof(double) * (size_t)n);
    if (e == cudaSuccess) e = cudaMalloc(&mx.p, 2 * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMemset(ya.p, 0, sizeof(double) * (size_t)n);
    if (e == cudaSuccess) e = cudaMemset(mx.p, 0, 2 * sizeof(unsigned long long));
    if (e != cudaSuccess) {
        sg_stencil_destroy(st);
        sg_set_error("build_stencil: %s", cudaGetErrorString(e));
        return SG_E_CUDA;
    }
    const unsigned g = (unsigned)((n + 255) / 256);
    k_hash_vec<<<g, 256>>>(n, xb.as<double>());
    ClsDev cells_only = op->cls;
    cells_only.bmat = nullptr;
    double *scratch = op->own_red.partials + 2 * SG_MAX_BLOCKS;
    cg_class_apply<D, P><<<op->cls_grid, CB, op->cls_smem>>>(cells_only, xb.as<double>(), ya.as<double>(), op->own_red, scratch, nullptr);
    rc = sg_stencil_apply(st, xb.as<double>(), yb.as<double>(), 0, n, op->own_red, scratch, nullptr, 0);
    k_max_diff<<<g, 256>>>(n, ya.as<double>(), yb.as<double>(), mx.as<unsigned long long>());
    unsigned long long bits[2] = {0, 0};
    e = cudaMemcpy(bits, mx.p, sizeof(bits), cudaMemcpyDeviceToHost);
    if (rc || e != cudaSuccess) {
        sg_stencil_destroy(st);
        if (!rc) sg_set_error("build_stencil: %s", cudaGetErrorString(e));
        return rc ? rc : SG_E_CUDA;
    }
    double ymax, dmax;
    memcpy(&ymax, &bits[0], 8);
    memcpy(&dmax, &bits[1], 8);
    if (!(dmax <= 1e-12 * ymax)) {   // never seen; keeps the cell-centric kernel rather than a wrong operator
        sg_stencil_destroy(st);
        return SG_OK;
    }
    // Exterior facets join the gather form where the solve is latency bound (small problems: one launch per apply, and the
    // whole PCG solve in one persistent kernel).  On large meshes the facet-by-facet gather of the boundary rows (15 % of
    // the rows of the 3-D P2 plate) costs the row kernel more than the separate 21 us cg_bfacet_apply launch it saves
    // (measured on one GPU's share of config 4: 138 vs 77 + 21 us), so those keep the two-launch form.
    static const long BND_IN_STENCIL_MAX_ROWS = getenv("SG_BND_IN_STENCIL_MAX_ROWS") ? atol(getenv("SG_BND_IN_STENCIL_MAX_ROWS")) : 1500000;
    if (op->bmat && dv.n_bf > 0 && dv.n_dofs <= BND_IN_STENCIL_MAX_ROWS) {
        constexpr int NFD = nfd_of(D, P);
        int fdt[24] = {0};
        for (int f = 0; f < D + 1; ++f)
            for (int k = 0; k < NFD; ++k) fdt[f * 6 + k] = facet_dof(D, P, f, k);
        int attached = 0;
        rc = sg_stencil_attach_boundary(st, dv.dofmap, dv.n_cells, dv.cell_lo, dv.cell_hi, dv.n_bf, dv.bf_cell, dv.bf_facet, NFD, fdt,
                                        op->bmat, &attached);
        if (rc) {
            sg_stencil_destroy(st);
            return rc;
        }
        op->stencil_bnd = attached;
    }
    op->stencil = st;
    // The residual's cell part in the same form (no RED scatter in the time loop): a second set of row classes from the
    // class MASS matrices |detJ| Mhat.  Verified like the first against a cell-centric scatter; any failure just keeps
    // cg_class_resid.
    {
        const int NS = op->cls.n_self, S = op->cls.S;
        DevBuf rep;
        if (cudaMalloc(&rep.p, sizeof(int32_t) * (size_t)NS) == cudaSuccess && cudaMalloc(&op->mass_tab, sizeof(double) * (size_t)NS * S) == cudaSuccess) {
            MassTabArgs<NLD> ma;
            memcpy(ma.mass, op->mass, sizeof(double) * NLD * NLD);
            k_class_rep<<<(unsigned)((dv.n_cells + 255) / 256), 256>>>(dv.n_cells, op->cls.cls16, rep.as<int32_t>());
            k_mass_tables<NLD><<<(NS + 63) / 64, 64>>>(ma, NS, S, rep.as<int32_t>(), op->d.geom + (int64_t)D * D * op->d.n_cells, op->mass_tab);
            SgStencil *sm = nullptr;
            if (cudaGetLastError() == cudaSuccess &&
                sg_stencil_build(op->ctx, dv.dofmap, dv.n_cells, NLD, dv.cell_lo, dv.cell_hi, op->cls.cls16, op->mass_tab, S, dv.n_dofs, &sm) == SG_OK && sm) {
                // check: S_M x against the scatter of the class mass matrices
                ClsDev mass_cells = op->cls;
                mass_cells.bmat = nullptr;
                mass_cells.tab = op->mass_tab;
                cudaMemset(ya.p, 0, sizeof(double) * (size_t)n);
                cudaMemset(mx.p, 0, 2 * sizeof(unsigned long long));
                cg_class_apply<D, P><<<op->cls_grid, CB, op->cls_smem>>>(mass_cells, xb.as<double>(), ya.as<double>(), op->own_red, scratch, nullptr);
                int rcm = sg_stencil_apply_cells(sm, xb.as<double>(), yb.as<double>(), 0, op->own_red, scratch, 0);
                k_max_diff<<<g, 256>>>(n, ya.as<double>(), yb.as<double>(), mx.as<unsigned long long>());
                unsigned long long mb[2] = {0, 0};
                if (!rcm && cudaMemcpy(mb, mx.p, sizeof(mb), cudaMemcpyDeviceToHost) == cudaSuccess) {
                    double ym, dm;
                    memcpy(&ym, &mb[0], 8);
                    memcpy(&dm, &mb[1], 8);
                    if (dm <= 1e-12 * ym) op->stencil_mass = sm;
                }
                if (!op->stencil_mass) sg_stencil_destroy(sm);
            }
        }
        cudaGetLastError();
    }
    return SG_OK;
}

// Find the local-matrix classes of this mesh and build the tables (see the comment above key_mix).
// Leaves op->cls.tab == nullptr (general kernel stays in use) when the mesh has too many classes.
template <int D, int P, bool DG>
int build_classes_t(sg_thermal_op *op) {
    using T = Tab<D, P, DG>;
    constexpr int NLD = T::NLD, NNB = DG ? D + 1 : 0;
    constexpr int S = (NLD * NLD + 1) & ~1;  // even: every class matrix is 16-B aligned in shared memory
    const OpDev &dv = op->dev;
    const long nc = dv.n_cells;
    if (nc <= 0 || dv.cell_hi <= dv.cell_lo) return SG_OK;
    const unsigned g1 = (unsigned)((nc + 255) / 256);
    DevBuf keys, gcls, grep, fcls, frep, scls, srep, bad;
    SG_CHECK_CUDA(cudaMalloc(&keys.p, sizeof(uint64_t) * (size_t)nc * (NNB > 0 ? NNB : 1)));
    k_geom_key<D><<<g1, 256>>>(dv.geom, nc, keys.as<uint64_t>());
    SG_CHECK_CUDA(cudaGetLastError());
    int32_t G = 0, NF = 0, NS = 0;
    int rc = sg_classify_u64(keys.as<uint64_t>(), nc, (int32_t **)&gcls.p, &G, (int32_t **)&grep.p);
    if (rc) return rc;
    SG_CHECK_CUDA(cudaMalloc(&bad.p, sizeof(unsigned)));
    SG_CHECK_CUDA(cudaMemset(bad.p, 0, sizeof(unsigned)));
    k_geom_verify<D><<<g1, 256>>>(dv.geom, nc, gcls.as<int32_t>(), grep.as<int32_t>(), bad.as<unsigned>());
    unsigned nbad = 0;
    SG_CHECK_CUDA(cudaMemcpy(&nbad, bad.p, sizeof(unsigned), cudaMemcpyDeviceToHost));
    op->n_geom_classes = G;
    if (nbad) return SG_OK;  // 64-bit hash collision: keep the general kernel
    const int32_t *rep_self = grep.as<int32_t>(), *rep_nb = nullptr;
    if (DG) {
        const unsigned gf = (unsigned)((nc * NNB + 255) / 256);
        k_facet_key<<<gf, 256>>>(NNB, nc, dv.nbr, dv.nbinfo, gcls.as<int32_t>(), keys.as<uint64_t>());
        SG_CHECK_CUDA(cudaGetLastError());
        if ((rc = sg_classify_u64(keys.as<uint64_t>(), nc * NNB, (int32_t **)&fcls.p, &NF, (int32_t **)&frep.p))) return rc;
        k_self_key<<<g1, 256>>>(NNB, nc, dv.nbr, gcls.as<int32_t>(), fcls.as<int32_t>(), keys.as<uint64_t>());
        SG_CHECK_CUDA(cudaGetLastError());
        if ((rc = sg_classify_u64(keys.as<uint64_t>(), nc, (int32_t **)&scls.p, &NS, (int32_t **)&srep.p))) return rc;
        rep_self = srep.as<int32_t>();
        rep_nb = frep.as<int32_t>();
        if (NS > 65535 || NF > 4094) return SG_OK;
    } else {
        NS = G;
        if (NS > 65535) return SG_OK;
    }
    const size_t smem = sizeof(double) * (size_t)(NS + NF) * S;
    if (smem > 96 * 1024) return SG_OK;
    // tables
    SG_CHECK_CUDA(cudaMalloc(&op->cls_tab, smem));
    const T &tab = *static_cast<const T *>(op->tab_host);
    const int nthreads = (NS + NF) * NLD;
    k_build_tables<D, P, DG><<<(nthreads + 63) / 64, 64>>>(tab, dv, NS, rep_self, NF, rep_nb, S, op->cls_tab);
    SG_CHECK_CUDA(cudaGetLastError());
    // per-cell class words
    if (DG) {
        SG_CHECK_CUDA(cudaMalloc(&op->cls_words, sizeof(uint64_t) * (size_t)nc));
        k_pack_dg<<<g1, 256>>>(NNB, nc, dv.nbr, scls.as<int32_t>(), fcls.as<int32_t>(), (uint64_t *)op->cls_words);
    } else {
        SG_CHECK_CUDA(cudaMalloc(&op->cls_words, sizeof(uint16_t) * (size_t)nc));
        k_pack_cg<<<g1, 256>>>(nc, gcls.as<int32_t>(), (uint16_t *)op->cls_words);
    }
    SG_CHECK_CUDA(cudaGetLastError());
    SG_CHECK_CUDA(cudaDeviceSynchronize());
    // launch geometry: persistent blocks, as many as fit per SM
    int per_sm = 0;
    if (DG) {
        constexpr int NB = D + 1;
        SG_CHECK_CUDA(cudaFuncSetAttribute(dg_class_apply<NLD, NB, P, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        SG_CHECK_CUDA(cudaFuncSetAttribute(dg_class_apply<NLD, NB, P, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        SG_CHECK_CUDA(cudaFuncSetAttribute(dg_class_apply<NLD, NB, P, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        SG_CHECK_CUDA(cudaFuncSetAttribute(dg_class_apply<NLD, NB, P, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        SG_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, dg_class_apply<NLD, NB, P, true, true>, CB, smem));
    } else {
        SG_CHECK_CUDA(cudaFuncSetAttribute(cg_class_apply<D, P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        SG_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, cg_class_apply<D, P>, CB, smem));
    }
    if (per_sm < 1) {
        cudaFree(op->cls_tab);
        cudaFree(op->cls_words);
        op->cls_tab = nullptr;
        op->cls_words = nullptr;
        return SG_OK;
    }
    long grid = (long)per_sm * op->ctx->sm_count;
    const long need = (dv.cell_hi - dv.cell_lo + CB - 1) / CB;
    if (grid > need) grid = need;
    if (grid > SG_MAX_BLOCKS) grid = SG_MAX_BLOCKS;
    op->cls_grid = (int)grid;
    op->cls_smem = smem;
    ClsDev &cd = op->cls;
    cd.n_cells = nc;
    cd.cell_lo = dv.cell_lo;
    cd.cell_hi = dv.cell_hi;
    cd.dot_lo = dv.dot_lo;
    cd.dot_hi = dv.dot_hi;
    cd.nbr = dv.nbr;
    cd.dofmap = dv.dofmap;
    cd.cls64 = DG ? (const uint64_t *)op->cls_words : nullptr;
    cd.cls16 = DG ? nullptr : (const uint16_t *)op->cls_words;
    cd.n_self = NS;
    cd.n_nb = NF;
    cd.S = S;
    cd.split_lo = dv.cell_lo;
    cd.split_hi = dv.cell_hi;
    cd.rev = 0;
    if (DG && (dv.cell_lo > 0 || dv.cell_hi < nc)) {
        DevBuf b2;
        SG_CHECK_CUDA(cudaMalloc(&b2.p, 2 * sizeof(unsigned long long)));
        const unsigned long long init[2] = {(unsigned long long)dv.cell_lo, (unsigned long long)dv.cell_hi};
        SG_CHECK_CUDA(cudaMemcpy(b2.p, init, sizeof(init), cudaMemcpyHostToDevice));
        const long nown = dv.cell_hi - dv.cell_lo;
        k_split_range<<<(unsigned)((nown + 255) / 256), 256>>>(NNB, nc, dv.cell_lo, dv.cell_hi, dv.nbr, b2.as<unsigned long long>(),
                                                               b2.as<unsigned long long>() + 1);
        unsigned long long res[2];
        SG_CHECK_CUDA(cudaMemcpy(res, b2.p, sizeof(res), cudaMemcpyDeviceToHost));
        if ((long)res[1] >= (long)res[0]) {     // interior first, strips (which wait for the ghost rows) last
            cd.split_lo = (long)res[0];
            cd.split_hi = (long)res[1];
        } else {                                 // every owned cell touches a ghost: wait before anything
            cd.split_lo = cd.split_hi = dv.cell_lo;
        }
    }
    cd.bmat = nullptr;
    cd.n_bf = dv.n_bf;
    cd.bf_cell = dv.bf_cell;
    cd.bf_facet = dv.bf_facet;
    if (dv.n_bf > 0) {
        // exterior facets from per-facet linearised boundary matrices (sg_thermal_linearize); DG applies them inside
        // the class kernel through its own copy of the neighbour ids with the exterior facets encoded
        constexpr int NFD = nfd_of(D, P), NFDP = NFD * (NFD + 1) / 2;
        if (DG) {
            SG_CHECK_CUDA(cudaMalloc(&op->nbr_ext, sizeof(int32_t) * (size_t)nc * NNB));
            SG_CHECK_CUDA(cudaMemcpy(op->nbr_ext, dv.nbr, sizeof(int32_t) * (size_t)nc * NNB, cudaMemcpyDeviceToDevice));
            k_mark_exterior<<<(unsigned)((dv.n_bf + 255) / 256), 256>>>(dv.n_bf, nc, dv.bf_cell, dv.bf_facet, op->nbr_ext);
            SG_CHECK_CUDA(cudaGetLastError());
            cd.nbr = op->nbr_ext;
        }
        SG_CHECK_CUDA(cudaMalloc(&op->bmat, sizeof(double) * (size_t)dv.n_bf * NFDP));
        SG_CHECK_CUDA(cudaMemset(op->bmat, 0, sizeof(double) * (size_t)dv.n_bf * NFDP));
        SG_CHECK_CUDA(cudaDeviceSynchronize());
        cd.bmat = op->bmat;
    }
    cd.tab = op->cls_tab;  // set last: marks the fast path as available
    if constexpr (!DG) {
        if (!(op->d.flags & SG_THERMAL_NO_STENCIL)) {
            const int rs = build_stencil_t<D, P>(op);
            if (rs) return rs;
        }
    }
    return SG_OK;
}

template <int D, int P, bool DG>
int build_tab(sg_thermal_op *op) {
    using T = Tab<D, P, DG>;
    const sg_thermal_desc &d = op->d;
    SG_REQUIRE(d.n_ld == T::NLD && d.nqc == T::NQC, "sg_thermal_op_create: table sizes (n_ld=%d nqc=%d) do not match dim=%d degree=%d",
               d.n_ld, d.nqc, D, P);
    if (DG) SG_REQUIRE(d.nqf == T::NQF && d.n_perm == T::NPERM, "sg_thermal_op_create: facet table sizes (nqf=%d n_perm=%d) do not match", d.nqf, d.n_perm);
    T *t = new T();
    memset(t, 0, sizeof(T));
    memcpy(t->mass, d.mass, sizeof(t->mass));
    memcpy(t->load, d.load, sizeof(t->load));
    memcpy(t->cq_w, d.cq_w, sizeof(t->cq_w));
    memcpy(t->cq_grad, d.cq_grad, sizeof(t->cq_grad));
    if (DG) {
        SG_REQUIRE(d.fq_w && d.fq_val && d.fq_perm && (P == 1 || d.fq_grad), "sg_thermal_op_create: DG needs the facet tables");
        memcpy(t->fq_w, d.fq_w, sizeof(t->fq_w));
        memcpy(t->fq_val, d.fq_val, sizeof(t->fq_val));
        if (P == 2) memcpy(t->fq_grad, d.fq_grad, sizeof(t->fq_grad));
        memcpy(t->fq_perm, d.fq_perm, sizeof(t->fq_perm));
    }
    op->tab_host = t;
    op->tab_bytes = sizeof(T);
    op->launch = &launch_op<D, P, DG>;
    op->build_classes = &build_classes_t<D, P, DG>;
    op->linearize = &linearize_t<D, P, DG>;
    op->cheb_step = &cheb_step_t<D, P, DG>;
    op->small_pcg = &small_pcg_t<D, P, DG>;
    return SG_OK;
}

template <int D>
int build_tab_d(sg_thermal_op *op) {
    const bool dg = op->d.family == 1;
    if (op->d.degree == 1) return dg ? build_tab<D, 1, true>(op) : build_tab<D, 1, false>(op);
    if (op->d.degree == 2) return dg ? build_tab<D, 2, true>(op) : build_tab<D, 2, false>(op);
    sg_set_error("sg_thermal_op_create: degree must be 1 or 2 (got %d)", op->d.degree);
    return SG_E_UNSUPPORTED;
}

}  // namespace

void sg_op_info(const sg_thermal_op *op, SgOpInfo *o) {
    o->ctx = op->ctx;
    o->dim = op->d.dim;
    o->family = op->d.family;
    o->n_ld = op->d.n_ld;
    o->n_dofs = op->d.n_dofs;
    o->own_lo = op->d.own_lo;
    o->own_hi = op->d.own_hi;
    o->n_cells = op->d.n_cells;
    o->cell_lo = op->d.cell_lo;
    o->cell_hi = op->d.cell_hi;
    o->detJ = op->d.geom + (int64_t)op->d.dim * op->d.dim * op->d.n_cells;
    o->mass_inv = op->mass_inv;
}

int sg_thermal_linearize(sg_thermal_op *op, const double *T_lin, cudaStream_t st) { return op->linearize(op, T_lin, st); }

bool sg_thermal_profiling(const sg_thermal_op *op) { return op->prof_on != 0; }

bool sg_thermal_has_cheb(const sg_thermal_op *op) {
    // the exterior facets must be inside the class kernel (bmat) or absent, so that J z is complete per cell
    return op->d.family == 1 && op->cls.tab != nullptr && (op->bmat != nullptr || op->d.n_bfacets == 0);
}

int sg_thermal_cheb_step(sg_thermal_op *op, const SgChebStep &cs, SgRed red, double *dot_out, const int *skip, cudaStream_t st) {
    return op->cheb_step(op, cs, red, dot_out, skip, st);
}

int sg_thermal_pcg_persistent(sg_thermal_op *op, const double *T_lin, const double *b, const double *dinv, double *x, double *work,
                              const SgPcgPolicy &pol, int max_it, double *rr0_out, int *ctrl_done, int *ctrl_iters, double *ctrl_rr,
                              cudaStream_t st) {
    // CG space in gather form with the exterior facets inside the stencil kernel (or none), one rank
    if (op->d.family != 0 || !op->stencil || op->ctx->nranks != 1) return 0;
    if (op->dev.n_bf > 0 && !(op->bmat && op->stencil_bnd)) return 0;
    static const bool off = [] {
        const char *e = getenv("SG_NO_PERSISTENT");
        return e && e[0] == '1';
    }();
    if (off) return 0;
    const int rc = sg_thermal_linearize(op, T_lin, st);
    if (rc) return rc;
    return sg_stencil_pcg(op->stencil, op->ctx->sm_count, b, dinv, x, work, pol, max_it, rr0_out, ctrl_done, ctrl_iters, ctrl_rr, st);
}

int sg_thermal_pcg_small_dg(sg_thermal_op *op, const double *T_lin, const double *b, double *x, const SgPcgPolicy &pol, int max_it,
                            double *rr0_out, int *ctrl_done, int *ctrl_iters, double *ctrl_rr, cudaStream_t st) {
    if (op->d.family != 1 || op->ctx->nranks != 1 || !op->cls.tab || op->d.n_cells > SMALL_DG_MAX_CELLS) return 0;
    if (op->dev.n_bf > 0 && !op->bmat) return 0;        // exterior facets must be inside the class kernel
    static const bool off = [] {
        const char *e = getenv("SG_NO_PERSISTENT");
        return e && e[0] == '1';
    }();
    if (off) return 0;
    const int rc = sg_thermal_linearize(op, T_lin, st);
    if (rc) return rc;
    return op->small_pcg(op, b, x, pol, max_it, rr0_out, ctrl_done, ctrl_iters, ctrl_rr, st);
}

int sg_thermal_apply_dot(sg_thermal_op *op, const double *T_lin, const double *x, double *y, SgRed red, double *dot2,
                         const int *skip, cudaStream_t st, int y_is_zero, const SgHaloWait *wait) {
    op->y_is_zero = y_is_zero;
    op->wait = wait;
    const int rc = op->launch(op, MODE_APPLY, T_lin, x, nullptr, y, red, dot2, skip, st);
    op->y_is_zero = 0;
    op->wait = nullptr;
    return rc;
}

// Gauss-Jordan inverse of the (SPD, n <= 10) reference mass matrix
static void invert_small(const double *a, int n, double *inv) {
    double m[10][20];
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
            m[i][j] = a[i * n + j];
            m[i][n + j] = (i == j) ? 1.0 : 0.0;
        }
    for (int k = 0; k < n; ++k) {
        int piv = k;
        for (int i = k + 1; i < n; ++i)
            if (fabs(m[i][k]) > fabs(m[piv][k])) piv = i;
        for (int j = 0; j < 2 * n; ++j) {
            const double tmp = m[k][j];
            m[k][j] = m[piv][j];
            m[piv][j] = tmp;
        }
        const double d = m[k][k];
        for (int j = 0; j < 2 * n; ++j) m[k][j] /= d;
        for (int i = 0; i < n; ++i)
            if (i != k) {
                const double f = m[i][k];
                for (int j = 0; j < 2 * n; ++j) m[i][j] -= f * m[k][j];
            }
    }
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) inv[i * n + j] = m[i][n + j];
}

extern "C" {

int sg_thermal_op_create(sg_ctx *ctx, const sg_thermal_desc *d, sg_thermal_op **out) {
    SG_REQUIRE(ctx && d && out, "sg_thermal_op_create: NULL argument");
    SG_REQUIRE(d->dim >= 1 && d->dim <= 3, "sg_thermal_op_create: dim must be 1..3");
    SG_REQUIRE(d->family == 0 || d->family == 1, "sg_thermal_op_create: family must be 0 (CG) or 1 (DG)");
    SG_REQUIRE(d->n_cells >= 0 && d->cell_lo >= 0 && d->cell_lo <= d->cell_hi && d->cell_hi <= d->n_cells,
               "sg_thermal_op_create: bad cell range");
    SG_REQUIRE(d->n_cells < (int64_t)0x7fffffff, "sg_thermal_op_create: cell ids are 32-bit (int32 neighbour ids / dofmap)");
    SG_REQUIRE(d->own_lo >= 0 && d->own_lo <= d->own_hi && d->own_hi <= d->n_dofs, "sg_thermal_op_create: bad owned dof range");
    SG_REQUIRE(d->own_cell_lo >= 0 && d->own_cell_lo <= d->own_cell_hi && d->own_cell_hi <= d->n_cells,
               "sg_thermal_op_create: bad owned cell range");
    SG_REQUIRE(d->geom && d->mass && d->load && d->cq_w && d->cq_grad, "sg_thermal_op_create: missing geometry/tables");
    SG_REQUIRE(d->family == 1 || d->dofmap, "sg_thermal_op_create: CG needs a dofmap");
    SG_REQUIRE(d->family == 0 || (d->nbr && d->nbinfo), "sg_thermal_op_create: DG needs neighbour maps");
    SG_REQUIRE(d->n_bfacets == 0 || (d->bf_cell && d->bf_facet && d->bf_area && d->bq_w && d->bq_val && d->nqb > 0),
               "sg_thermal_op_create: missing exterior-facet data");
    SG_CHECK_CUDA(cudaSetDevice(ctx->device));
    sg_thermal_op *op = new sg_thermal_op();
    memset(op, 0, sizeof(*op));
    op->ctx = ctx;
    op->d = *d;
    int rc = SG_E_INVALID;
    if (d->dim == 1) rc = build_tab_d<1>(op);
    if (d->dim == 2) rc = build_tab_d<2>(op);
    if (d->dim == 3) rc = build_tab_d<3>(op);
    if (rc != SG_OK) {
        delete op;
        return rc;
    }
    invert_small(d->mass, d->n_ld, op->mass_inv);
    memcpy(op->mass, d->mass, sizeof(double) * d->n_ld * d->n_ld);
    memcpy(op->load, d->load, sizeof(double) * d->n_ld);
    OpDev &dv = op->dev;
    memset(&dv, 0, sizeof(dv));
    dv.n_cells = d->n_cells;
    dv.cell_lo = d->cell_lo;
    dv.cell_hi = d->cell_hi;
    const bool own_given = d->own_cell_hi > d->own_cell_lo;
    dv.dot_lo = own_given ? d->own_cell_lo : d->cell_lo;
    dv.dot_hi = own_given ? d->own_cell_hi : d->cell_hi;
    dv.n_dofs = d->n_dofs;
    dv.dofmap = d->dofmap;
    dv.geom = d->geom;
    dv.nbr = d->nbr;
    dv.nbinfo = d->nbinfo;
    dv.n_bf = d->n_bfacets;
    dv.bf_cell = d->bf_cell;
    dv.bf_facet = d->bf_facet;
    dv.bf_area = d->bf_area;
    dv.nqb = d->nqb;
    dv.dt = d->dt;
    dv.dt_alpha = d->dt * d->alpha;
    dv.dt_f = d->dt * d->f;
    dv.se = d->sigma * d->epsilon;
    dv.htc = d->htc;
    dv.Ta = d->T_ambient;
    dv.penalty = d->penalty;
    if (d->n_bfacets > 0) {
        const size_t nb = sizeof(double) * (size_t)(d->dim + 1) * d->nqb * d->n_ld;
        cudaError_t e = cudaMalloc(&op->btab_dev, nb);
        if (e == cudaSuccess) e = cudaMalloc(&op->bw_dev, sizeof(double) * d->nqb);
        if (e == cudaSuccess) e = cudaMemcpy(op->btab_dev, d->bq_val, nb, cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(op->bw_dev, d->bq_w, sizeof(double) * d->nqb, cudaMemcpyHostToDevice);
        if (e != cudaSuccess) {
            sg_set_error("sg_thermal_op_create: uploading boundary tables failed: %s", cudaGetErrorString(e));
            sg_thermal_op_destroy(op);
            return SG_E_CUDA;
        }
        dv.btab = op->btab_dev;
        dv.bw = op->bw_dev;
    }
    // host table pointers are not retained
    op->d.mass = op->d.load = op->d.cq_w = op->d.cq_grad = op->d.fq_w = op->d.fq_val = op->d.fq_grad = nullptr;
    op->d.fq_perm = nullptr;
    op->d.bq_w = op->d.bq_val = nullptr;
    if (!(d->flags & SG_THERMAL_NO_CLASSES)) {
        rc = op->build_classes(op);
        if (rc != SG_OK) {
            sg_thermal_op_destroy(op);
            return rc;
        }
    }
    if (d->family == 0 && d->n_dofs > 0) {
        // diag of the cell part with the general kernel (into the cache itself: launch_op then takes the uncached route
        // with no exterior facets), once
        double *cache = nullptr;
        cudaError_t e = cudaMalloc(&cache, sizeof(double) * (size_t)d->n_dofs);
        if (e != cudaSuccess) {
            sg_set_error("sg_thermal_op_create: %s", cudaGetErrorString(e));
            sg_thermal_op_destroy(op);
            return SG_E_CUDA;
        }
        const int64_t n_bf = op->dev.n_bf;
        op->dev.n_bf = 0;
        op->diag_cells = cache;
        rc = op->launch(op, MODE_DIAG, nullptr, nullptr, nullptr, cache, SgRed{nullptr, nullptr}, nullptr, nullptr, 0);
        op->dev.n_bf = n_bf;
        if (rc == SG_OK && cudaDeviceSynchronize() != cudaSuccess) {
            sg_set_error("sg_thermal_op_create: diagonal of the cell part failed");
            rc = SG_E_CUDA;
        }
        if (rc != SG_OK) {
            sg_thermal_op_destroy(op);
            return rc;
        }
    }
    *out = op;
    return SG_OK;
}

int sg_thermal_op_destroy(sg_thermal_op *op) {
    if (!op) return SG_OK;
    if (op->btab_dev) cudaFree(op->btab_dev);
    if (op->bw_dev) cudaFree(op->bw_dev);
    if (op->cls_tab) cudaFree(op->cls_tab);
    if (op->cls_words) cudaFree(op->cls_words);
    if (op->nbr_ext) cudaFree(op->nbr_ext);
    if (op->bmat) cudaFree(op->bmat);
    sg_stencil_destroy(op->stencil);
    sg_stencil_destroy(op->stencil_mass);
    cudaFree(op->mass_tab);
    if (op->diag_cells) cudaFree(op->diag_cells);
    if (op->own_red.partials) cudaFree(op->own_red.partials);
    if (op->own_red.counter) cudaFree(op->own_red.counter);
    for (int i = 0; i < 2 * op->prof_cap; ++i) cudaEventDestroy(op->prof_ev[i]);
    delete[] op->prof_ev;
    delete[] op->prof_kind;
    ::operator delete(op->tab_host);
    delete op;
    return SG_OK;
}

static SgRed no_red() { return SgRed{nullptr, nullptr}; }

int sg_thermal_residual(sg_thermal_op *op, const double *T, const double *T_prev, double *F, void *stream) {
    SG_REQUIRE(op && T && T_prev && F, "sg_thermal_residual: NULL argument");
    return op->launch(op, MODE_RESID, T, T, T_prev, F, no_red(), nullptr, nullptr, (cudaStream_t)stream);
}

int sg_thermal_jac_apply(sg_thermal_op *op, const double *T_lin, const double *x, double *y, void *stream) {
    SG_REQUIRE(op && T_lin && x && y, "sg_thermal_jac_apply: NULL argument");
    // the class kernels reduce x.y as they go and need scratch for it even when nobody reads the result
    int rc = op->cls.tab ? ensure_own_red(op) : SG_OK;
    if (rc) return rc;
    rc = op->linearize(op, T_lin, (cudaStream_t)stream);
    if (rc) return rc;
    return op->launch(op, MODE_APPLY, T_lin, x, nullptr, y, op->own_red, nullptr, nullptr, (cudaStream_t)stream);
}

int sg_thermal_jac_diag(sg_thermal_op *op, const double *T_lin, double *diag, void *stream) {
    SG_REQUIRE(op && T_lin && diag, "sg_thermal_jac_diag: NULL argument");
    return op->launch(op, MODE_DIAG, T_lin, nullptr, nullptr, diag, no_red(), nullptr, nullptr, (cudaStream_t)stream);
}

int sg_thermal_class_info(const sg_thermal_op *op, int32_t *n_geometry, int32_t *n_self, int32_t *n_facet) {
    SG_REQUIRE(op, "sg_thermal_class_info: NULL operator");
    if (n_geometry) *n_geometry = op->n_geom_classes;
    if (n_self) *n_self = op->cls.tab ? op->cls.n_self : 0;
    if (n_facet) *n_facet = op->cls.tab ? op->cls.n_nb : 0;
    return op->cls.tab ? 1 : 0;
}

int sg_thermal_stencil_info(const sg_thermal_op *op, int32_t *n_classes, int32_t *n_entries, int32_t *max_nnz) {
    SG_REQUIRE(op, "sg_thermal_stencil_info: NULL operator");
    if (!op->stencil) return 0;
    sg_stencil_info(op->stencil, n_classes, n_entries, max_nnz);
    return 1;
}

int sg_thermal_profile(sg_thermal_op *op, int32_t enable, int32_t capacity) {
    SG_REQUIRE(op, "sg_thermal_profile: NULL operator");
    if (enable && op->prof_cap < capacity) {
        for (int i = 0; i < 2 * op->prof_cap; ++i) cudaEventDestroy(op->prof_ev[i]);
        delete[] op->prof_ev;
        delete[] op->prof_kind;
        op->prof_kind = new unsigned char[capacity];
        op->prof_ev = new cudaEvent_t[2 * capacity];
        for (int i = 0; i < 2 * capacity; ++i) SG_CHECK_CUDA(cudaEventCreate(&op->prof_ev[i]));
        op->prof_cap = capacity;
    }
    op->prof_on = enable;
    op->prof_n = 0;
    return SG_OK;
}

// Launches of kind `kind` recorded since the last enable.  Launches that returned at once because the solve had
// already converged (device-side skip flag) are recognised by their duration (< 1/4 of the median) and left out.
int sg_thermal_profile_read_kind(sg_thermal_op *op, int32_t kind, int64_t *n_launches, double *ms_total) {
    SG_REQUIRE(op && n_launches && ms_total, "sg_thermal_profile_read: NULL argument");
    std::vector<float> t;
    for (int i = 0; i < op->prof_n; ++i) {
        if (op->prof_kind[i] != kind) continue;
        float ms = 0.f;
        SG_CHECK_CUDA(cudaEventSynchronize(op->prof_ev[2 * i + 1]));
        SG_CHECK_CUDA(cudaEventElapsedTime(&ms, op->prof_ev[2 * i], op->prof_ev[2 * i + 1]));
        t.push_back(ms);
    }
    double tot = 0.0;
    int64_t n = 0;
    if (!t.empty()) {
        std::vector<float> sorted(t);
        std::sort(sorted.begin(), sorted.end());
        const float cut = 0.25f * sorted[sorted.size() / 2];
        for (float v : t)
            if (v >= cut) {
                tot += v;
                ++n;
            }
    }
    *n_launches = n;
    *ms_total = tot;
    return SG_OK;
}

int sg_thermal_profile_read(sg_thermal_op *op, int64_t *n_launches, double *ms_total) {
    return sg_thermal_profile_read_kind(op, 0, n_launches, ms_total);
}

// Algorithmic HBM bytes of one fused Chebyshev step (dg_cheb_step): the apply's traffic + r, z_prev (not in the
// first step), |detJ| per cell; z_out replaces y.
int64_t sg_thermal_cheb_step_bytes(const sg_thermal_op *op, int32_t first) {
    if (!op || !op->cls.tab || op->d.family != 1) return -1;
    const sg_thermal_desc &d = op->d;
    const int64_t ncell = d.cell_hi - d.cell_lo;
    return sg_thermal_apply_bytes(op) + ncell * (8 + 8 * d.n_ld * (first ? 1 : 2));
}

// Algorithmic HBM bytes of the cell kernel of one Jacobian apply, for the layout actually in use.
int64_t sg_thermal_apply_bytes(const sg_thermal_op *op) {
    if (!op) return -1;
    const sg_thermal_desc &d = op->d;
    const int64_t ncell = d.cell_hi - d.cell_lo;
    const int64_t ndof = d.family == 1 ? ncell * d.n_ld : d.n_dofs;
    if (op->stencil) return 18 * d.n_dofs;   // row-stencil form: class id + x + y per row; the class lists live in shared memory
    int64_t per_cell;
    if (op->cls.tab) {
        per_cell = d.family == 1 ? 8 + 4 * (d.dim + 1)   // class word + neighbour ids
                                 : 2 + 4 * d.n_ld;        // class id + dofmap
    } else {
        per_cell = 8 * (d.dim * d.dim + 1);                     // Jinv + detJ
        if (d.family == 1) per_cell += 8 + 4 * (d.dim + 1) + 4;  // h, neighbour ids, packed facet info
        else per_cell += 4 * d.n_ld;                            // dofmap
    }
    const int nfd = d.degree == 1 ? d.dim : d.dim * (d.dim + 1) / 2;
    const int64_t bnd = (op->bmat && d.family == 1) ? d.n_bfacets * 8 * (nfd * (nfd + 1) / 2) : 0;   // packed facet matrices (DG: same kernel)
    return ncell * per_cell + 16 * ndof + bnd;                  // + read x, write y
}

}  // extern "C"
