// (C) Mechanical equilibrium of the plate — SURVEY §8(f) row 4, an extension: the reference sets
// total_strain = -thermal_strain (VM:135-139) and never solves for a displacement.
//
// Once per time step, after sg_visco_update has produced the reference's stress sigma0, find the displacement increment
// du (vector P1 on the mesh vertices) with
//     sum_K |K| [ 2 G_K dev(eps_K(du)) + K_K tr(eps_K(du)) I + sigma0_K ] : eps_K(v) = 0      for every free v,
// G_K, K_K, sigma0_K = cell means (weights w_l = int phi_l / |K| of the sigma element) of nodal fields; the nodal tangent
// moduli are the Prony chain's own factors (VM:176-191) summed over the terms, so that adding
//     2 G_eff dev(eps(du)) + K_eff tr(eps(du)) I
// to sigma0 at every sigma node is exactly what the chain would have produced with total_strain = eps(du) - thermal_strain.
//
// Kernels: matrix-free tangent apply, one thread per cell: gather (d+1) x d displacements, constant strain, stress,
// scatter with RED.ADD.F64; Jacobi-PCG whose scalars stay on the device (the host reads a control block every
// PCG_BATCH iterations); node-wise correction kernel.  Compiled with -fmad=false like visco.cu: the per-term factors
// must come out of the same operation sequence as the stress they correct.
#include "sg_common.cuh"
#include "visco_common.cuh"

namespace {

constexpr int MB = 256;          // threads per block
constexpr int PCG_BATCH = 32;    // iterations between two host reads of the control block

struct MechCtrl {
    int done;      // set by the device when |r|^2 <= tol2
    int iters;     // iterations completed
    double rr;     // |r|^2 of the last completed iteration
    double bb;     // |b|^2
};

// device scalars: rz[2] (double-buffered by iteration parity), pAp, rr, bb
enum { S_RZ0 = 0, S_RZ1 = 1, S_PAP = 2, S_RR = 3, S_BB = 4, S_COUNT = 8 };

__device__ __forceinline__ double tol2_of(double bb, double rtol, double atol) { return fmax(rtol * rtol * bb, atol * atol); }

// Where T_cur == T_prev bitwise the reference's stress formula is 0/0 = NaN although the restrained stress
// of a zero strain increment is 0: the equilibrium step reads such entries as 0.
__device__ __forceinline__ double finite_or_zero(double v) { return (v == v) ? v : 0.0; }

// Factor A(lambda) that multiplies 2 g_n dev / k_n tr in the chain's stress increment.
//   reference scheme (VM:176-191 with VM:233-242):  lambda (1 - (1 + a + a^2/2)) / xi,  a = -xi/lambda,  which is
//     1 + a/2 identically; the reference's own sequence cancels catastrophically (relative noise ulp/|a|, 0/0 at xi = 0:
//     SURVEY Q5, H2) — harmless for a stress that is only written out, not for a tangent that must stay positive and
//     equal on both sides of the solve, so the equilibrium step evaluates the closed form;
//   corrected scheme: (1 - exp(-x))/x, x = xi/lambda (decay_fac of visco.cu).
__device__ __forceinline__ double term_factor(const VKParams &P, double xi, double lambda) {
    if (P.mode == SG_VISCO_CORRECTED) {
        double decay, fac;
        decay_fac(xi, lambda, decay, fac);
        return fac;
    }
    return 1.0 + 0.5 * ((-1.0 * xi) / lambda);
}

template <int D>
struct CellGeom {
    double g[D + 1][D];   // gradients of the P1 hat functions
    double vol;
};

template <int D>
__device__ __forceinline__ void load_geom(const double *__restrict__ ginv, const double *__restrict__ vol, long c, CellGeom<D> &G) {
#pragma unroll
    for (int j = 0; j < D; ++j) G.g[0][j] = 0.0;
#pragma unroll
    for (int a = 0; a < D; ++a)
#pragma unroll
        for (int j = 0; j < D; ++j) {
            const double v = ginv[c * D * D + a * D + j];
            G.g[a + 1][j] = v;
            G.g[0][j] -= v;
        }
    G.vol = vol[c];
}

// eps = sym(sum_a u_a (x) g_a)
template <int D>
__device__ __forceinline__ void cell_strain(const CellGeom<D> &G, const double (&u)[D + 1][D], double (&eps)[D][D]) {
    double gr[D][D];
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
        for (int j = 0; j < D; ++j) {
            double s = 0.0;
#pragma unroll
            for (int a = 0; a <= D; ++a) s += u[a][i] * G.g[a][j];
            gr[i][j] = s;
        }
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
        for (int j = 0; j < D; ++j) eps[i][j] = 0.5 * (gr[i][j] + gr[j][i]);
}

template <int D>
__device__ __forceinline__ void hooke(double Gm, double Km, const double (&eps)[D][D], double (&sig)[D][D]) {
    double tr = 0.0;
#pragma unroll
    for (int i = 0; i < D; ++i) tr += eps[i][i];
    const double vol_part = Km * tr - 2.0 * Gm * (tr / (double)D);
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
        for (int j = 0; j < D; ++j) sig[i][j] = 2.0 * Gm * eps[i][j] + (i == j ? vol_part : 0.0);
}

template <int D>
__global__ void __launch_bounds__(MB) k_mech_geom(long nc, const double *__restrict__ x, const int32_t *__restrict__ cells,
                                                  double *__restrict__ ginv, double *__restrict__ vol, int *bad) {
    const long c = (long)blockIdx.x * MB + threadIdx.x;
    if (c >= nc) return;
    double J[D][D];   // J[i][a] = x_{a+1}[i] - x_0[i]
    const long v0 = cells[c * (D + 1)];
#pragma unroll
    for (int a = 0; a < D; ++a) {
        const long va = cells[c * (D + 1) + a + 1];
#pragma unroll
        for (int i = 0; i < D; ++i) J[i][a] = x[va * D + i] - x[v0 * D + i];
    }
    double det, inv[D][D];
    if constexpr (D == 1) {
        det = J[0][0];
        inv[0][0] = 1.0 / det;
    } else if constexpr (D == 2) {
        det = J[0][0] * J[1][1] - J[0][1] * J[1][0];
        inv[0][0] = J[1][1] / det;
        inv[0][1] = -J[0][1] / det;
        inv[1][0] = -J[1][0] / det;
        inv[1][1] = J[0][0] / det;
    } else {
        const double c00 = J[1][1] * J[2][2] - J[1][2] * J[2][1];
        const double c01 = J[1][2] * J[2][0] - J[1][0] * J[2][2];
        const double c02 = J[1][0] * J[2][1] - J[1][1] * J[2][0];
        det = J[0][0] * c00 + J[0][1] * c01 + J[0][2] * c02;
        inv[0][0] = c00 / det;
        inv[1][0] = c01 / det;
        inv[2][0] = c02 / det;
        inv[0][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) / det;
        inv[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) / det;
        inv[2][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) / det;
        inv[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) / det;
        inv[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) / det;
        inv[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) / det;
    }
    if (!(fabs(det) > 0.0)) *bad = 1;
    // row a of J^-1 is the gradient of lambda_{a+1}
#pragma unroll
    for (int a = 0; a < D; ++a)
#pragma unroll
        for (int j = 0; j < D; ++j) ginv[c * D * D + a * D + j] = inv[a][j];
    constexpr double fact = (D == 3) ? 6.0 : (D == 2 ? 2.0 : 1.0);
    vol[c] = fabs(det) / fact;
}

struct SigmaW {
    double w[16];
};

__global__ void __launch_bounds__(MB) k_mech_moduli(long nc, int n_ld, const __grid_constant__ SigmaW W, const int32_t *__restrict__ sdm,
                                                    const double *__restrict__ Gn, const double *__restrict__ Kn,
                                                    double *__restrict__ Gc, double *__restrict__ Kc) {
    const long c = (long)blockIdx.x * MB + threadIdx.x;
    if (c >= nc) return;
    double g = 0.0, k = 0.0;
    for (int l = 0; l < n_ld; ++l) {
        const long s = sdm[c * n_ld + l];
        g += W.w[l] * Gn[s];
        k += W.w[l] * Kn[s];
    }
    Gc[c] = g;
    Kc[c] = k;
}

template <int D>
__global__ void __launch_bounds__(MB) k_mech_apply(long nc, const int32_t *__restrict__ cells, const double *__restrict__ ginv,
                                                   const double *__restrict__ vol, const double *__restrict__ Gc,
                                                   const double *__restrict__ Kc, const uint8_t *__restrict__ fixed,
                                                   const double *__restrict__ x, double *__restrict__ y, const int *skip) {
    if (skip && *skip) return;
    const long c = (long)blockIdx.x * MB + threadIdx.x;
    if (c >= nc) return;
    CellGeom<D> G;
    load_geom<D>(ginv, vol, c, G);
    long v[D + 1];
    double u[D + 1][D];
    bool fx[D + 1][D];
#pragma unroll
    for (int a = 0; a <= D; ++a) {
        v[a] = cells[c * (D + 1) + a];
#pragma unroll
        for (int i = 0; i < D; ++i) {
            fx[a][i] = fixed[v[a] * D + i] != 0;
            u[a][i] = fx[a][i] ? 0.0 : x[v[a] * D + i];
        }
    }
    double eps[D][D], sig[D][D];
    cell_strain<D>(G, u, eps);
    hooke<D>(Gc[c], Kc[c], eps, sig);
#pragma unroll
    for (int a = 0; a <= D; ++a)
#pragma unroll
        for (int i = 0; i < D; ++i) {
            double f = 0.0;
#pragma unroll
            for (int j = 0; j < D; ++j) f += sig[i][j] * G.g[a][j];
            if (!fx[a][i]) atomicAdd(y + v[a] * D + i, G.vol * f);
        }
}

template <int D>
__global__ void __launch_bounds__(MB) k_mech_rhs(long nc, int n_ld, const __grid_constant__ SigmaW W, const int32_t *__restrict__ cells,
                                                 const int32_t *__restrict__ sdm, const double *__restrict__ ginv,
                                                 const double *__restrict__ vol, const uint8_t *__restrict__ fixed,
                                                 const double *__restrict__ sigma0, double *__restrict__ b) {
    const long c = (long)blockIdx.x * MB + threadIdx.x;
    if (c >= nc) return;
    CellGeom<D> G;
    load_geom<D>(ginv, vol, c, G);
    double s0[D][D];
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
        for (int j = 0; j < D; ++j) s0[i][j] = 0.0;
    for (int l = 0; l < n_ld; ++l) {
        const long s = sdm[c * n_ld + l];
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
            for (int j = 0; j < D; ++j) s0[i][j] += W.w[l] * finite_or_zero(sigma0[s * D * D + i * D + j]);
    }
#pragma unroll
    for (int a = 0; a <= D; ++a) {
        const long va = cells[c * (D + 1) + a];
#pragma unroll
        for (int i = 0; i < D; ++i) {
            double f = 0.0;
#pragma unroll
            for (int j = 0; j < D; ++j) f += 0.5 * (s0[i][j] + s0[j][i]) * G.g[a][j];
            if (!fixed[va * D + i]) atomicAdd(b + va * D + i, -(G.vol * f));
        }
    }
}

template <int D>
__global__ void __launch_bounds__(MB) k_mech_diag(long nc, const int32_t *__restrict__ cells, const double *__restrict__ ginv,
                                                  const double *__restrict__ vol, const double *__restrict__ Gc,
                                                  const double *__restrict__ Kc, double *__restrict__ diag) {
    const long c = (long)blockIdx.x * MB + threadIdx.x;
    if (c >= nc) return;
    CellGeom<D> G;
    load_geom<D>(ginv, vol, c, G);
    const double Gm = Gc[c], lam = Kc[c] - 2.0 * Gm / (double)D;
#pragma unroll
    for (int a = 0; a <= D; ++a) {
        const long va = cells[c * (D + 1) + a];
        double g2 = 0.0;
#pragma unroll
        for (int j = 0; j < D; ++j) g2 += G.g[a][j] * G.g[a][j];
#pragma unroll
        for (int i = 0; i < D; ++i)
            atomicAdd(diag + va * D + i, G.vol * (Gm * (g2 + G.g[a][i] * G.g[a][i]) + lam * G.g[a][i] * G.g[a][i]));
    }
}

__global__ void __launch_bounds__(MB) k_mech_dinv(long n, const uint8_t *__restrict__ fixed, double *__restrict__ d) {
    for (long i = (long)blockIdx.x * MB + threadIdx.x; i < n; i += (long)gridDim.x * MB) d[i] = fixed[i] ? 1.0 : 1.0 / d[i];
}

// rows of the held components: y = x (the identity block of P A P + I - P)
__global__ void __launch_bounds__(MB) k_mech_fix(long n, const uint8_t *__restrict__ fixed, const double *__restrict__ x, double *__restrict__ y) {
    for (long i = (long)blockIdx.x * MB + threadIdx.x; i < n; i += (long)gridDim.x * MB)
        if (fixed[i]) y[i] = x[i];
}

// x <- P x;  r = b - A x (Ax given);  p = D^-1 r;  rz, rr, bb
__global__ void __launch_bounds__(MB) k_mech_init(long n, const uint8_t *__restrict__ fixed, const double *__restrict__ b,
                                                  const double *__restrict__ Ax, const double *__restrict__ dinv, double *__restrict__ x,
                                                  double *__restrict__ r, double *__restrict__ p, SgRed red, double *S) {
    double acc[3] = {0.0, 0.0, 0.0};
    for (long i = (long)blockIdx.x * MB + threadIdx.x; i < n; i += (long)gridDim.x * MB) {
        double ri = 0.0, bi = 0.0;
        if (fixed[i]) {
            x[i] = 0.0;
        } else {
            bi = b[i];
            ri = bi - Ax[i];
        }
        const double zi = dinv[i] * ri;
        r[i] = ri;
        p[i] = zi;
        acc[0] += ri * zi;
        acc[1] += ri * ri;
        acc[2] += bi * bi;
    }
    sg_grid_reduce<3>(acc, red, S + S_COUNT);   // totals land in S[S_COUNT..S_COUNT+2]; k_mech_begin files them
}

// thread 0 of one block: file the totals of k_mech_init and open the iteration
__global__ void k_mech_begin(double *S, MechCtrl *ctrl, double rtol, double atol) {
    S[S_RZ0] = S[S_COUNT + 0];
    S[S_RR] = S[S_COUNT + 1];
    S[S_BB] = S[S_COUNT + 2];
    ctrl->iters = 0;
    ctrl->rr = S[S_RR];
    ctrl->bb = S[S_BB];
    ctrl->done = (S[S_RR] <= tol2_of(S[S_BB], rtol, atol)) ? 1 : 0;
}

// Ap on the held rows = p;  pAp
__global__ void __launch_bounds__(MB) k_mech_dot(long n, const uint8_t *__restrict__ fixed, const double *__restrict__ p,
                                                 double *__restrict__ Ap, SgRed red, double *S, const MechCtrl *ctrl) {
    if (ctrl->done) return;
    double acc[1] = {0.0};
    for (long i = (long)blockIdx.x * MB + threadIdx.x; i < n; i += (long)gridDim.x * MB) {
        const double pi = p[i];
        double ai = Ap[i];
        if (fixed[i]) {
            ai = pi;
            Ap[i] = ai;
        }
        acc[0] += pi * ai;
    }
    sg_grid_reduce<1>(acc, red, S + S_PAP);
}

// x += alpha p;  r -= alpha Ap;  rz_new = r . D^-1 r;  rr = r . r
__global__ void __launch_bounds__(MB) k_mech_update_xr(long n, int par, const double *__restrict__ p, const double *__restrict__ Ap,
                                                       const double *__restrict__ dinv, double *__restrict__ x, double *__restrict__ r,
                                                       SgRed red, double *S, const MechCtrl *ctrl) {
    if (ctrl->done) return;
    const double alpha = S[S_RZ0 + par] / S[S_PAP];
    double acc[2] = {0.0, 0.0};
    for (long i = (long)blockIdx.x * MB + threadIdx.x; i < n; i += (long)gridDim.x * MB) {
        x[i] += alpha * p[i];
        const double ri = r[i] - alpha * Ap[i];
        r[i] = ri;
        acc[0] += ri * (dinv[i] * ri);
        acc[1] += ri * ri;
    }
    sg_grid_reduce<2>(acc, red, S + S_COUNT);
}

// p = D^-1 r + beta p; block 0 / thread 0 files the new scalars and evaluates the stopping rule
__global__ void __launch_bounds__(MB) k_mech_update_p(long n, int par, const double *__restrict__ r, const double *__restrict__ dinv,
                                                      double *__restrict__ p, double *S, MechCtrl *ctrl, double rtol, double atol) {
    if (ctrl->done) return;
    const double rz_new = S[S_COUNT + 0], rr = S[S_COUNT + 1];
    const double beta = rz_new / S[S_RZ0 + par];
    for (long i = (long)blockIdx.x * MB + threadIdx.x; i < n; i += (long)gridDim.x * MB) p[i] = dinv[i] * r[i] + beta * p[i];
    // every block has read S[S_RZ0 + par] and S[S_COUNT..] before anyone overwrites them: the other parity slot is
    // written here, S_COUNT.. only by the next k_mech_update_xr
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        S[S_RZ0 + (par ^ 1)] = rz_new;
        S[S_RR] = rr;
        ctrl->iters += 1;
        ctrl->rr = rr;
        if (rr <= tol2_of(S[S_BB], rtol, atol) || !(rr == rr)) ctrl->done = 1;
        if (!(S[S_PAP] > 0.0)) ctrl->done = 2;   // p.Ap <= 0: the tangent is not positive definite
    }
}

// G_eff, K_eff per sigma node: the factors that multiply 2 g_n dev / k_n tr in VM:176-191, same operation order as visco.cu
__global__ void __launch_bounds__(MB) k_mech_coef(const __grid_constant__ VKParams P, long n, const double *__restrict__ xi_s,
                                                  double *__restrict__ Ge, double *__restrict__ Ke) {
    const long i = (long)blockIdx.x * MB + threadIdx.x;
    if (i >= n) return;
    const double xi = xi_s[i];
    double g = 0.0, k = 0.0;
    for (int t = 0; t < P.N; ++t) {
        g = g + P.g2[t] * term_factor(P, xi, P.lg[t]);
        k = k + P.k[t] * term_factor(P, xi, P.lk[t]);
    }
    Ge[i] = 0.5 * g;   // g2 = 2 g_n
    Ke[i] = k;
}

template <int D>
__global__ void __launch_bounds__(MB) k_mech_correct(const __grid_constant__ VKParams P, long ns, const int32_t *__restrict__ winner,
                                                     const int32_t *__restrict__ cells, const double *__restrict__ ginv,
                                                     const double *__restrict__ vol, const double *__restrict__ du,
                                                     const double *__restrict__ xi_s, const double *__restrict__ Ge,
                                                     const double *__restrict__ Ke, const sg_mech_fields f) {
    constexpr int DD = D * D;
    const long s = (long)blockIdx.x * MB + threadIdx.x;
    if (s >= ns) return;
    const long c = winner[s];
    CellGeom<D> G;
    load_geom<D>(ginv, vol, c, G);
    double u[D + 1][D];
#pragma unroll
    for (int a = 0; a <= D; ++a) {
        const long va = cells[c * (D + 1) + a];
#pragma unroll
        for (int i = 0; i < D; ++i) u[a][i] = du[va * D + i];
    }
    double eps[D][D], sig[D][D];
    cell_strain<D>(G, u, eps);
    hooke<D>(Ge[s], Ke[s], eps, sig);
    double tr = 0.0;
#pragma unroll
    for (int i = 0; i < D; ++i) tr += eps[i][i];
    double dev[D][D];
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
        for (int j = 0; j < D; ++j) {
            dev[i][j] = eps[i][j] - (i == j ? P.inv_d * tr : 0.0);
            const long e = s * DD + i * D + j;
            f.sigma[e] = finite_or_zero(f.sigma[e]) + sig[i][j];
            f.mech_strain[e] = eps[i][j];
            if (f.total_strain) f.total_strain[e] += eps[i][j];
            if (f.deviatoric_strain) f.deviatoric_strain[e] += dev[i][j];
        }
    const bool corrected = P.mode == SG_VISCO_CORRECTED;
    const bool any_partial = f.ds_partial || f.dsigma_partial || f.s_partial || f.sigma_partial || (corrected && f.s_tilde);
    if (!any_partial) return;
    const double xi = xi_s[s];
    for (int t = 0; t < P.N; ++t) {
        const double fg = term_factor(P, xi, P.lg[t]), fk = term_factor(P, xi, P.lk[t]);
        const double dk_t = (P.k[t] * tr) * fk;
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
            for (int j = 0; j < D; ++j) {
                const double ds = (P.g2[t] * dev[i][j]) * fg;
                const double dk = (i == j) ? dk_t : 0.0;
                const long e = (s * P.N + t) * DD + i * D + j;
                if (f.ds_partial) f.ds_partial[e] = finite_or_zero(f.ds_partial[e]) + ds;
                if (f.dsigma_partial) f.dsigma_partial[e] = finite_or_zero(f.dsigma_partial[e]) + dk;
                if (f.s_partial) f.s_partial[e] = finite_or_zero(f.s_partial[e]) + ds;
                if (f.sigma_partial) f.sigma_partial[e] = finite_or_zero(f.sigma_partial[e]) + dk;
                if (corrected && f.s_tilde) {
                    f.s_tilde[e] += ds;
                    f.sigma_tilde[e] += dk;
                }
            }
    }
}

inline unsigned cgrid(long n) { return (unsigned)((n + MB - 1) / MB); }
inline unsigned vgrid(long n) {
    const long g = (n + MB - 1) / MB;
    return (unsigned)(g < 1 ? 1 : (g > SG_MAX_BLOCKS ? SG_MAX_BLOCKS : g));
}

}  // namespace

struct sg_mech_op {
    sg_ctx *ctx;
    int dim, n_ld;
    int64_t nv, nc, ns, n;           // n = nv * dim
    const double *coords;
    const int32_t *cells, *sdm, *winner;
    const uint8_t *fixed;
    SigmaW W;
    double *ginv, *vol, *Gc, *Kc;    // owned
    double *work;                    // r, p, Ap, dinv, b
    double *S;                       // device scalars
    double *red_partials;
    unsigned *red_counter;
    MechCtrl *ctrl, *ctrl_host;      // device / pinned host
    bool have_moduli;
    SgRed red() const {
        SgRed r;
        r.partials = red_partials;
        r.counter = red_counter;
        r.peer = nullptr;
        r.ar_ptr = nullptr;
        r.ar_count = 0;
        return r;
    }
    double *r_() const { return work; }
    double *p_() const { return work + n; }
    double *Ap_() const { return work + 2 * n; }
    double *dinv_() const { return work + 3 * n; }
    double *b_() const { return work + 4 * n; }
};

namespace {

template <int D>
int apply_d(sg_mech_op *op, const double *x, double *y, const int *skip, cudaStream_t st) {
    SG_CHECK_CUDA(cudaMemsetAsync(y, 0, sizeof(double) * op->n, st));
    k_mech_apply<D><<<cgrid(op->nc), MB, 0, st>>>(op->nc, op->cells, op->ginv, op->vol, op->Gc, op->Kc, op->fixed, x, y, skip);
    SG_CHECK_CUDA(cudaGetLastError());
    sg_count_launch();
    return SG_OK;
}

int apply_raw(sg_mech_op *op, const double *x, double *y, const int *skip, cudaStream_t st) {
    switch (op->dim) {
        case 1: return apply_d<1>(op, x, y, skip, st);
        case 2: return apply_d<2>(op, x, y, skip, st);
        default: return apply_d<3>(op, x, y, skip, st);
    }
}

int rhs_raw(sg_mech_op *op, const double *sigma0, double *b, cudaStream_t st) {
    SG_CHECK_CUDA(cudaMemsetAsync(b, 0, sizeof(double) * op->n, st));
    const unsigned g = cgrid(op->nc);
    switch (op->dim) {
        case 1: k_mech_rhs<1><<<g, MB, 0, st>>>(op->nc, op->n_ld, op->W, op->cells, op->sdm, op->ginv, op->vol, op->fixed, sigma0, b); break;
        case 2: k_mech_rhs<2><<<g, MB, 0, st>>>(op->nc, op->n_ld, op->W, op->cells, op->sdm, op->ginv, op->vol, op->fixed, sigma0, b); break;
        default: k_mech_rhs<3><<<g, MB, 0, st>>>(op->nc, op->n_ld, op->W, op->cells, op->sdm, op->ginv, op->vol, op->fixed, sigma0, b); break;
    }
    SG_CHECK_CUDA(cudaGetLastError());
    sg_count_launch();
    return SG_OK;
}

}  // namespace

extern "C" {

int sg_mech_op_create(sg_ctx *ctx, const sg_mech_desc *d, sg_mech_op **out) {
    SG_REQUIRE(ctx && d && out, "sg_mech_op_create: NULL argument");
    SG_REQUIRE(d->dim >= 1 && d->dim <= 3, "sg_mech_op_create: dim must be 1..3 (got %d)", d->dim);
    SG_REQUIRE(ctx->nranks == 1, "sg_mech_op_create: the equilibrium solve runs on one GPU (nranks = %d)", ctx->nranks);
    SG_REQUIRE(d->n_vertices > 0 && d->n_cells > 0 && d->n_sigma_nodes > 0, "sg_mech_op_create: empty mesh");
    SG_REQUIRE(d->coords && d->cells && d->fixed && d->sigma_dofmap && d->sigma_weights && d->winner_cell,
               "sg_mech_op_create: NULL array in the descriptor");
    SG_REQUIRE(d->n_ld_sigma >= 1 && d->n_ld_sigma <= 16, "sg_mech_op_create: n_ld_sigma must be 1..16 (got %d)", d->n_ld_sigma);
    SG_REQUIRE(d->n_cells < (int64_t)2147483647 * MB / 2, "sg_mech_op_create: too many cells");
    sg_mech_op *op = new sg_mech_op();
    memset(op, 0, sizeof(*op));
    op->ctx = ctx;
    op->dim = d->dim;
    op->n_ld = d->n_ld_sigma;
    op->nv = d->n_vertices;
    op->nc = d->n_cells;
    op->ns = d->n_sigma_nodes;
    op->n = d->n_vertices * d->dim;
    op->coords = d->coords;
    op->cells = d->cells;
    op->sdm = d->sigma_dofmap;
    op->winner = d->winner_cell;
    op->fixed = d->fixed;
    for (int l = 0; l < d->n_ld_sigma; ++l) op->W.w[l] = d->sigma_weights[l];
    const int DD = d->dim * d->dim;
    int rc = SG_OK;
    auto fail = [&](cudaError_t e, const char *what) {
        sg_set_error("sg_mech_op_create: %s -> %s", what, cudaGetErrorString(e));
        rc = SG_E_CUDA;
    };
    cudaError_t e;
    int *bad = nullptr;
    if ((e = cudaMalloc(&op->ginv, sizeof(double) * op->nc * DD)) != cudaSuccess) fail(e, "cudaMalloc(ginv)");
    if (!rc && (e = cudaMalloc(&op->vol, sizeof(double) * op->nc)) != cudaSuccess) fail(e, "cudaMalloc(vol)");
    if (!rc && (e = cudaMalloc(&op->Gc, sizeof(double) * op->nc)) != cudaSuccess) fail(e, "cudaMalloc(Gc)");
    if (!rc && (e = cudaMalloc(&op->Kc, sizeof(double) * op->nc)) != cudaSuccess) fail(e, "cudaMalloc(Kc)");
    if (!rc && (e = cudaMalloc(&op->work, sizeof(double) * op->n * 5)) != cudaSuccess) fail(e, "cudaMalloc(work)");
    if (!rc && (e = cudaMalloc(&op->S, sizeof(double) * 2 * S_COUNT)) != cudaSuccess) fail(e, "cudaMalloc(S)");
    if (!rc && (e = cudaMalloc(&op->red_partials, sizeof(double) * SG_MAX_BLOCKS * 3)) != cudaSuccess) fail(e, "cudaMalloc(partials)");
    if (!rc && (e = cudaMalloc(&op->red_counter, sizeof(unsigned))) != cudaSuccess) fail(e, "cudaMalloc(counter)");
    if (!rc && (e = cudaMalloc(&op->ctrl, sizeof(MechCtrl))) != cudaSuccess) fail(e, "cudaMalloc(ctrl)");
    if (!rc && (e = cudaMallocHost(&op->ctrl_host, sizeof(MechCtrl))) != cudaSuccess) fail(e, "cudaMallocHost(ctrl)");
    if (!rc && (e = cudaMalloc(&bad, sizeof(int))) != cudaSuccess) fail(e, "cudaMalloc(flag)");
    if (!rc) {
        cudaMemset(op->red_counter, 0, sizeof(unsigned));
        cudaMemset(op->S, 0, sizeof(double) * 2 * S_COUNT);
        cudaMemset(op->ctrl, 0, sizeof(MechCtrl));
        cudaMemset(bad, 0, sizeof(int));
        const unsigned g = cgrid(op->nc);
        switch (op->dim) {
            case 1: k_mech_geom<1><<<g, MB>>>(op->nc, op->coords, op->cells, op->ginv, op->vol, bad); break;
            case 2: k_mech_geom<2><<<g, MB>>>(op->nc, op->coords, op->cells, op->ginv, op->vol, bad); break;
            default: k_mech_geom<3><<<g, MB>>>(op->nc, op->coords, op->cells, op->ginv, op->vol, bad); break;
        }
        sg_count_launch();
        int hbad = 0;
        if ((e = cudaMemcpy(&hbad, bad, sizeof(int), cudaMemcpyDeviceToHost)) != cudaSuccess) fail(e, "geometry kernel");
        else if (hbad) {
            sg_set_error("sg_mech_op_create: degenerate cell (zero Jacobian determinant)");
            rc = SG_E_INVALID;
        }
    }
    if (bad) cudaFree(bad);
    if (rc) {
        sg_mech_op_destroy(op);
        return rc;
    }
    *out = op;
    return SG_OK;
}

int sg_mech_op_destroy(sg_mech_op *op) {
    if (!op) return SG_OK;
    cudaFree(op->ginv);
    cudaFree(op->vol);
    cudaFree(op->Gc);
    cudaFree(op->Kc);
    cudaFree(op->work);
    cudaFree(op->S);
    cudaFree(op->red_partials);
    cudaFree(op->red_counter);
    cudaFree(op->ctrl);
    if (op->ctrl_host) cudaFreeHost(op->ctrl_host);
    delete op;
    return SG_OK;
}

int sg_mech_coefficients(sg_visco_plan *plan, int64_t n, const double *xi_sigma, double *G_eff, double *K_eff, void *stream) {
    SG_REQUIRE(plan && xi_sigma && G_eff && K_eff, "sg_mech_coefficients: NULL argument");
    SG_REQUIRE(n >= 0, "sg_mech_coefficients: negative node count");
    if (n == 0) return SG_OK;
    k_mech_coef<<<cgrid(n), MB, 0, (cudaStream_t)stream>>>(plan->k, (long)n, xi_sigma, G_eff, K_eff);
    SG_CHECK_CUDA(cudaGetLastError());
    sg_count_launch();
    return SG_OK;
}

int sg_mech_set_moduli(sg_mech_op *op, const double *G_eff, const double *K_eff, void *stream) {
    SG_REQUIRE(op && G_eff && K_eff, "sg_mech_set_moduli: NULL argument");
    cudaStream_t st = (cudaStream_t)stream;
    k_mech_moduli<<<cgrid(op->nc), MB, 0, st>>>(op->nc, op->n_ld, op->W, op->sdm, G_eff, K_eff, op->Gc, op->Kc);
    SG_CHECK_CUDA(cudaGetLastError());
    sg_count_launch();
    // point-Jacobi preconditioner of the new tangent
    SG_CHECK_CUDA(cudaMemsetAsync(op->dinv_(), 0, sizeof(double) * op->n, st));
    const unsigned g = cgrid(op->nc);
    switch (op->dim) {
        case 1: k_mech_diag<1><<<g, MB, 0, st>>>(op->nc, op->cells, op->ginv, op->vol, op->Gc, op->Kc, op->dinv_()); break;
        case 2: k_mech_diag<2><<<g, MB, 0, st>>>(op->nc, op->cells, op->ginv, op->vol, op->Gc, op->Kc, op->dinv_()); break;
        default: k_mech_diag<3><<<g, MB, 0, st>>>(op->nc, op->cells, op->ginv, op->vol, op->Gc, op->Kc, op->dinv_()); break;
    }
    SG_CHECK_CUDA(cudaGetLastError());
    k_mech_dinv<<<vgrid(op->n), MB, 0, st>>>(op->n, op->fixed, op->dinv_());
    SG_CHECK_CUDA(cudaGetLastError());
    sg_count_launch(2);
    op->have_moduli = true;
    return SG_OK;
}

int sg_mech_apply(sg_mech_op *op, const double *x, double *y, void *stream) {
    SG_REQUIRE(op && x && y, "sg_mech_apply: NULL argument");
    SG_REQUIRE(op->have_moduli, "sg_mech_apply: call sg_mech_set_moduli first");
    cudaStream_t st = (cudaStream_t)stream;
    int rc = apply_raw(op, x, y, nullptr, st);
    if (rc) return rc;
    k_mech_fix<<<vgrid(op->n), MB, 0, st>>>(op->n, op->fixed, x, y);
    SG_CHECK_CUDA(cudaGetLastError());
    sg_count_launch();
    return SG_OK;
}

int sg_mech_rhs(sg_mech_op *op, const double *sigma0, double *b, void *stream) {
    SG_REQUIRE(op && sigma0 && b, "sg_mech_rhs: NULL argument");
    return rhs_raw(op, sigma0, b, (cudaStream_t)stream);
}

int sg_mech_solve(sg_mech_op *op, const double *sigma0, double *du, double rtol, double atol, int32_t max_it, int32_t *iters,
                  double *rel_res, void *stream) {
    SG_REQUIRE(op && sigma0 && du, "sg_mech_solve: NULL argument");
    SG_REQUIRE(op->have_moduli, "sg_mech_solve: call sg_mech_set_moduli first");
    SG_REQUIRE(max_it > 0, "sg_mech_solve: max_it must be positive");
    cudaStream_t st = (cudaStream_t)stream;
    const long n = op->n;
    const unsigned vg = vgrid(n);
    double *r = op->r_(), *p = op->p_(), *Ap = op->Ap_(), *dinv = op->dinv_(), *b = op->b_();
    int rc = rhs_raw(op, sigma0, b, st);
    if (rc) return rc;
    // r = b - A P du
    rc = apply_raw(op, du, Ap, nullptr, st);
    if (rc) return rc;
    k_mech_init<<<vg, MB, 0, st>>>(n, op->fixed, b, Ap, dinv, du, r, p, op->red(), op->S);
    k_mech_begin<<<1, 1, 0, st>>>(op->S, op->ctrl, rtol, atol);
    SG_CHECK_CUDA(cudaGetLastError());
    sg_count_launch(2);
    int it = 0;
    MechCtrl *h = op->ctrl_host;
    auto read_ctrl = [&]() -> int {
        SG_CHECK_CUDA(cudaMemcpyAsync(h, op->ctrl, sizeof(MechCtrl), cudaMemcpyDeviceToHost, st));
        SG_CHECK_CUDA(cudaStreamSynchronize(st));
        return SG_OK;
    };
    rc = read_ctrl();
    if (rc) return rc;
    while (!h->done && it < max_it) {
        const int batch = (max_it - it < PCG_BATCH) ? max_it - it : PCG_BATCH;
        for (int k = 0; k < batch; ++k, ++it) {
            const int par = it & 1;
            rc = apply_raw(op, p, Ap, &op->ctrl->done, st);
            if (rc) return rc;
            k_mech_dot<<<vg, MB, 0, st>>>(n, op->fixed, p, Ap, op->red(), op->S, op->ctrl);
            k_mech_update_xr<<<vg, MB, 0, st>>>(n, par, p, Ap, dinv, du, r, op->red(), op->S, op->ctrl);
            k_mech_update_p<<<vg, MB, 0, st>>>(n, par, r, dinv, p, op->S, op->ctrl, rtol, atol);
            sg_count_launch(3);
        }
        SG_CHECK_CUDA(cudaGetLastError());
        rc = read_ctrl();
        if (rc) return rc;
    }
    if (iters) *iters = h->iters;
    if (rel_res) *rel_res = h->bb > 0.0 ? sqrt(h->rr / h->bb) : 0.0;
    if (h->done == 2) {
        sg_set_error("sg_mech_solve: the tangent is not positive definite (p.Ap <= 0 at iteration %d): a Prony factor is "
                     "negative, i.e. xi > 2 lambda_n in the reference's Taylor form (VM:233-242) - heating step or dt too large",
                     h->iters);
        return SG_E_NOCONV;
    }
    if (!h->done || !(h->rr == h->rr)) {
        sg_set_error("sg_mech_solve: PCG did not converge in %d iterations (|r|/|b| = %.3e)", h->iters,
                     h->bb > 0.0 ? sqrt(h->rr / h->bb) : 0.0);
        return SG_E_NOCONV;
    }
    return SG_OK;
}

int sg_mech_correct(sg_mech_op *op, sg_visco_plan *plan, const double *du, const double *xi_sigma, const double *G_eff,
                    const double *K_eff, const sg_mech_fields *f, void *stream) {
    SG_REQUIRE(op && plan && du && xi_sigma && G_eff && K_eff && f, "sg_mech_correct: NULL argument");
    SG_REQUIRE(f->sigma && f->mech_strain, "sg_mech_correct: sigma and mech_strain are required");
    SG_REQUIRE(plan->p.dim == op->dim, "sg_mech_correct: plan and operator have different dimensions");
    if (plan->k.mode == SG_VISCO_CORRECTED)
        SG_REQUIRE(f->s_tilde && f->sigma_tilde, "sg_mech_correct: the corrected scheme needs s_tilde and sigma_tilde (its history is the partial stress)");
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned g = cgrid(op->ns);
    switch (op->dim) {
        case 1: k_mech_correct<1><<<g, MB, 0, st>>>(plan->k, op->ns, op->winner, op->cells, op->ginv, op->vol, du, xi_sigma, G_eff, K_eff, *f); break;
        case 2: k_mech_correct<2><<<g, MB, 0, st>>>(plan->k, op->ns, op->winner, op->cells, op->ginv, op->vol, du, xi_sigma, G_eff, K_eff, *f); break;
        default: k_mech_correct<3><<<g, MB, 0, st>>>(plan->k, op->ns, op->winner, op->cells, op->ginv, op->vol, du, xi_sigma, G_eff, K_eff, *f); break;
    }
    SG_CHECK_CUDA(cudaGetLastError());
    sg_count_launch();
    return SG_OK;
}

int64_t sg_mech_apply_bytes(const sg_mech_op *op) {
    if (!op) return -1;
    const int64_t d = op->dim;
    // per cell: d+1 vertex ids, d*d gradients + volume, two moduli; per vector entry: read x, add into y, the fixed flag
    return op->nc * (4 * (d + 1) + 8 * (d * d + 1) + 16) + op->n * 17;
}

}  // extern "C"
