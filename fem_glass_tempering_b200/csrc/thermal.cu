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
// Design: one thread per cell, element tables passed BY VALUE in the kernel parameter block so that
// every table entry is a constant-bank operand of a fully unrolled DFMA (no shared-memory or LDS
// traffic for the tables).  Cell geometry is SoA ([component][cell]) so a warp reads 32 consecutive
// doubles per component.  DG: cell-centric interior facets (each cell visits its d+1 neighbours, no
// atomics, no colouring; the neighbour's dofs/geometry come through L2).  CG: gather through an SoA
// dofmap, scatter with native fp64 atomics (RED.ADD.F64).  Exterior (Robin + radiation) facets are a
// small separate kernel.
#include <math.h>

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

template <int D, int P, bool DG, int MODE>
__global__ void __launch_bounds__(TB) cell_kernel(const __grid_constant__ Tab<D, P, DG> tab, const __grid_constant__ OpDev op, const double *__restrict__ x,
                                                  const double *__restrict__ xprev, double *__restrict__ y) {
    using T = Tab<D, P, DG>;
    constexpr int NLD = T::NLD, NQC = T::NQC, NQF = T::NQF;
    const long c = op.cell_lo + (long)blockIdx.x * TB + threadIdx.x;
    if (c >= op.cell_hi) return;
    const long nc = op.n_cells;

    double Jinv[D][D];
#pragma unroll
    for (int a = 0; a < D; ++a)
#pragma unroll
        for (int b = 0; b < D; ++b) Jinv[a][b] = op.geom[(long)(a * D + b) * nc + c];
    const double detJ = op.geom[(long)(D * D) * nc + c];

    long dof[NLD];
    double xk[NLD], yk[NLD];
#pragma unroll
    for (int i = 0; i < NLD; ++i) {
        dof[i] = DG ? c * NLD + i : (long)op.dofmap[(long)i * nc + c];
        if (MODE != MODE_DIAG) xk[i] = x[dof[i]];
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
            for (int j = 0; j < NLD; ++j) xn[j] = x[nb * NLD + j];
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

    // ---- write / scatter ----
#pragma unroll
    for (int i = 0; i < NLD; ++i) {
        if (DG)
            y[dof[i]] = yk[i];
        else
            atomicAdd(&y[dof[i]], yk[i]);
    }
}

// Exterior facets: radiation + convection (TVP:302-304) and their linearisation.
template <int D, int P, bool DG, int MODE>
__global__ void __launch_bounds__(TB) bfacet_kernel(const OpDev op, const double *__restrict__ Tlin,
                                                    const double *__restrict__ x, double *__restrict__ y) {
    constexpr int NLD = nld_of(D, P);
    const long b = (long)blockIdx.x * TB + threadIdx.x;
    if (b >= op.n_bf) return;
    const long c = op.bf_cell[b];
    if (c < op.cell_lo || c >= op.cell_hi) return;
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
#pragma unroll
    for (int i = 0; i < NLD; ++i)
        if (acc[i] != 0.0) atomicAdd(&y[dof[i]], acc[i]);
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
    // optional profiling of the Jacobian-apply cell kernel (bench.py roofline): event pairs on the launch stream
    double mass_inv[100];
    int prof_on, prof_n, prof_cap;
    cudaEvent_t *prof_ev;
    int (*launch)(const sg_thermal_op *, int mode, const double *T, const double *x, const double *xprev, double *y,
                  cudaStream_t st);
};

namespace {

template <int D, int P, bool DG>
int launch_op(const sg_thermal_op *op, int mode, const double *Tlin, const double *x, const double *xprev, double *y,
              cudaStream_t st) {
    using T = Tab<D, P, DG>;
    const T &tab = *static_cast<const T *>(op->tab_host);
    const OpDev &dv = op->dev;
    const long ncell = dv.cell_hi - dv.cell_lo;
    const unsigned gc = (unsigned)((ncell + TB - 1) / TB), gb = (unsigned)((dv.n_bf + TB - 1) / TB);
    if (!DG) SG_CHECK_CUDA(cudaMemsetAsync(y, 0, sizeof(double) * (size_t)dv.n_dofs, st));
    if (ncell > 0) {
        const bool prof = mode == MODE_APPLY && op->prof_on && op->prof_n < op->prof_cap;
        sg_thermal_op *mop = const_cast<sg_thermal_op *>(op);
        if (prof) SG_CHECK_CUDA(cudaEventRecord(op->prof_ev[2 * op->prof_n], st));
        if (mode == MODE_APPLY) cell_kernel<D, P, DG, MODE_APPLY><<<gc, TB, 0, st>>>(tab, dv, x, nullptr, y);
        if (prof) {
            SG_CHECK_CUDA(cudaEventRecord(op->prof_ev[2 * op->prof_n + 1], st));
            mop->prof_n++;
        }
        sg_count_launch();
        if (mode == MODE_RESID) cell_kernel<D, P, DG, MODE_RESID><<<gc, TB, 0, st>>>(tab, dv, x, xprev, y);
        if (mode == MODE_DIAG) cell_kernel<D, P, DG, MODE_DIAG><<<gc, TB, 0, st>>>(tab, dv, nullptr, nullptr, y);
        SG_CHECK_CUDA(cudaGetLastError());
    }
    if (dv.n_bf > 0) {
        if (mode == MODE_APPLY) bfacet_kernel<D, P, DG, MODE_APPLY><<<gb, TB, 0, st>>>(dv, Tlin, x, y);
        if (mode == MODE_RESID) bfacet_kernel<D, P, DG, MODE_RESID><<<gb, TB, 0, st>>>(dv, x, nullptr, y);
        if (mode == MODE_DIAG) bfacet_kernel<D, P, DG, MODE_DIAG><<<gb, TB, 0, st>>>(dv, Tlin, nullptr, y);
        SG_CHECK_CUDA(cudaGetLastError());
        sg_count_launch();
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
    SG_REQUIRE(d->own_lo >= 0 && d->own_lo <= d->own_hi && d->own_hi <= d->n_dofs, "sg_thermal_op_create: bad owned dof range");
    SG_REQUIRE(d->geom && d->mass && d->load && d->cq_w && d->cq_grad, "sg_thermal_op_create: missing geometry/tables");
    SG_REQUIRE(d->family == 1 || d->dofmap, "sg_thermal_op_create: CG needs a dofmap");
    SG_REQUIRE(d->family == 0 || (d->nbr && d->nbinfo), "sg_thermal_op_create: DG needs neighbour maps");
    SG_REQUIRE(d->n_bfacets == 0 || (d->bf_cell && d->bf_facet && d->bf_area && d->bq_w && d->bq_val && d->nqb > 0),
               "sg_thermal_op_create: missing exterior-facet data");
    sg_thermal_op *op = new sg_thermal_op();
    op->ctx = ctx;
    op->d = *d;
    op->tab_host = nullptr;
    op->btab_dev = op->bw_dev = nullptr;
    op->prof_on = op->prof_n = op->prof_cap = 0;
    op->prof_ev = nullptr;
    int rc = SG_E_INVALID;
    if (d->dim == 1) rc = build_tab_d<1>(op);
    if (d->dim == 2) rc = build_tab_d<2>(op);
    if (d->dim == 3) rc = build_tab_d<3>(op);
    if (rc != SG_OK) {
        delete op;
        return rc;
    }
    invert_small(d->mass, d->n_ld, op->mass_inv);
    OpDev &dv = op->dev;
    memset(&dv, 0, sizeof(dv));
    dv.n_cells = d->n_cells;
    dv.cell_lo = d->cell_lo;
    dv.cell_hi = d->cell_hi;
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
    *out = op;
    return SG_OK;
}

int sg_thermal_op_destroy(sg_thermal_op *op) {
    if (!op) return SG_OK;
    if (op->btab_dev) cudaFree(op->btab_dev);
    if (op->bw_dev) cudaFree(op->bw_dev);
    for (int i = 0; i < 2 * op->prof_cap; ++i) cudaEventDestroy(op->prof_ev[i]);
    delete[] op->prof_ev;
    ::operator delete(op->tab_host);
    delete op;
    return SG_OK;
}

int sg_thermal_residual(sg_thermal_op *op, const double *T, const double *T_prev, double *F, void *stream) {
    SG_REQUIRE(op && T && T_prev && F, "sg_thermal_residual: NULL argument");
    return op->launch(op, MODE_RESID, T, T, T_prev, F, (cudaStream_t)stream);
}

int sg_thermal_jac_apply(sg_thermal_op *op, const double *T_lin, const double *x, double *y, void *stream) {
    SG_REQUIRE(op && T_lin && x && y, "sg_thermal_jac_apply: NULL argument");
    return op->launch(op, MODE_APPLY, T_lin, x, nullptr, y, (cudaStream_t)stream);
}

int sg_thermal_jac_diag(sg_thermal_op *op, const double *T_lin, double *diag, void *stream) {
    SG_REQUIRE(op && T_lin && diag, "sg_thermal_jac_diag: NULL argument");
    return op->launch(op, MODE_DIAG, T_lin, nullptr, nullptr, diag, (cudaStream_t)stream);
}

int sg_thermal_profile(sg_thermal_op *op, int32_t enable, int32_t capacity) {
    SG_REQUIRE(op, "sg_thermal_profile: NULL operator");
    if (enable && op->prof_cap < capacity) {
        for (int i = 0; i < 2 * op->prof_cap; ++i) cudaEventDestroy(op->prof_ev[i]);
        delete[] op->prof_ev;
        op->prof_ev = new cudaEvent_t[2 * capacity];
        for (int i = 0; i < 2 * capacity; ++i) SG_CHECK_CUDA(cudaEventCreate(&op->prof_ev[i]));
        op->prof_cap = capacity;
    }
    op->prof_on = enable;
    op->prof_n = 0;
    return SG_OK;
}

int sg_thermal_profile_read(sg_thermal_op *op, int64_t *n_launches, double *ms_total) {
    SG_REQUIRE(op && n_launches && ms_total, "sg_thermal_profile_read: NULL argument");
    double tot = 0.0;
    for (int i = 0; i < op->prof_n; ++i) {
        float ms = 0.f;
        SG_CHECK_CUDA(cudaEventSynchronize(op->prof_ev[2 * i + 1]));
        SG_CHECK_CUDA(cudaEventElapsedTime(&ms, op->prof_ev[2 * i], op->prof_ev[2 * i + 1]));
        tot += ms;
    }
    *n_launches = op->prof_n;
    *ms_total = tot;
    return SG_OK;
}

int64_t sg_thermal_apply_bytes(const sg_thermal_op *op) {
    if (!op) return -1;
    const sg_thermal_desc &d = op->d;
    const int64_t ncell = d.cell_hi - d.cell_lo;
    int64_t per_cell = 8 * (d.dim * d.dim + 1);            // Jinv + detJ
    if (d.family == 1) per_cell += 8 + 4 * (d.dim + 1) + 4;  // h, neighbour ids, packed facet info
    else per_cell += 4 * d.n_ld;                            // dofmap
    const int64_t ndof = d.family == 1 ? ncell * d.n_ld : d.n_dofs;
    return ncell * per_cell + 16 * ndof;                    // + read x, write y
}

}  // extern "C"
