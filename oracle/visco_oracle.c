/*
 * TEST INFRASTRUCTURE — NOT PRODUCT CODE.
 *
 * CPU restatement (plain C, float64, no FMA contraction) of the SurroGlas
 * ViscoelasticModel pointwise chain.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this library.
 *
 * PARITY UNPINNED: the reference ships no tests or golden vectors and its
 * arithmetic lives in un-vendored dolfinx/FFCx/UFL (requirements.txt:17-20),
 * which cannot be installed here.  The formulas below follow the reference's
 * own expression definitions line by line; the known-answer vectors in
 * tests/golden/ were hand-evaluated from those formulas, not produced by a
 * reference run.
 *
 * Each vo_<name> function restates one dolfinx Expression of
 * /root/reference/ViscoelasticModel.py (VM) evaluated at n points, in the
 * operation order the Python source builds (UFL keeps binary products/sums as
 * written; FFCx's default cffi flags have no -march=native, hence no FMA).
 * vo_step_passes() replays the 17 interpolations + 7 copies of
 * /root/reference/ThermoViscoProblem.py (TVP) solve_timestep in order.
 * vo_step_fused() is the same arithmetic in one sweep (CPU baseline "fused").
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off, optional OpenMP).
 */
#include <math.h>
#include <stddef.h>
#include <string.h>

#define VO_MAXN 16

typedef struct {
    int dim;          /* VM:17  mesh.topology.dim */
    int N;            /* VM:16  tableau_size */
    double H, Rg, Tb; /* VM:75-79 */
    double alpha_solid, alpha_liquid; /* VM:81-83 */
    double dt;        /* VM:88 */
    double m[VO_MAXN], lambda_m[VO_MAXN]; /* VM:19-34 */
    double g[VO_MAXN], lambda_g[VO_MAXN]; /* VM:35-50 */
    double k[VO_MAXN], lambda_k[VO_MAXN]; /* VM:51-68 */
} vo_params;

/* VM:233-242  sum_{k=0..2} 1/k! * (-xi/lambda)^k  ->  (1.0 + a) + 0.5*a^2 */
static inline double vo_taylor(double xi, double lambda)
{
    double a = (-1.0 * xi) / lambda;
    return (1.0 + a) + 0.5 * (a * a);
}

/* VM:156-161 (the live definition; VM:100-108 is overwritten) and VM:162-167 */
void vo_phi(const vo_params *p, long n, const double *T, double *phi)
{
    const double c = p->H / p->Rg;
    const double inv_Tb = 1.0 / p->Tb;
#pragma omp parallel for schedule(static)
    for (long q = 0; q < n; ++q)
        phi[q] = exp(c * (inv_Tb - 1.0 / T[q]));
}

/* VM:111-119 */
void vo_Tf_partial(const vo_params *p, long n, const double *Tfp_prev,
                   const double *T_cur, const double *phi, double *Tfp_out)
{
    const int N = p->N;
#pragma omp parallel for schedule(static)
    for (long q = 0; q < n; ++q)
        for (int i = 0; i < N; ++i) {
            double num = p->lambda_m[i] * Tfp_prev[q * N + i] + (T_cur[q] * p->dt) * phi[q];
            double den = p->lambda_m[i] + p->dt * phi[q];
            Tfp_out[q * N + i] = num / den;
        }
}

/* VM:122-125  inner(m, Tf_partial) */
void vo_Tf(const vo_params *p, long n, const double *Tfp, double *Tf)
{
    const int N = p->N;
#pragma omp parallel for schedule(static)
    for (long q = 0; q < n; ++q) {
        double acc = p->m[0] * Tfp[q * N];
        for (int i = 1; i < N; ++i)
            acc = acc + p->m[i] * Tfp[q * N + i];
        Tf[q] = acc;
    }
}

/* VM:128-133  I*(alpha_s*(T-Tprev) + (alpha_l-alpha_s)*(Tf-Tfprev)) */
void vo_thermal_strain(const vo_params *p, long n, const double *T_cur,
                       const double *T_prev, const double *Tf_cur,
                       const double *Tf_prev, double *eps)
{
    const int d = p->dim;
    const double da = p->alpha_liquid - p->alpha_solid;
#pragma omp parallel for schedule(static)
    for (long q = 0; q < n; ++q) {
        double s = p->alpha_solid * (T_cur[q] - T_prev[q]) + da * (Tf_cur[q] - Tf_prev[q]);
        for (int i = 0; i < d; ++i)
            for (int j = 0; j < d; ++j)
                eps[(q * d + i) * d + j] = (i == j) ? s : 0.0;
    }
}

/* VM:136-139  -thermal_strain  (Expr.__neg__ is -1*x) */
void vo_total_strain(const vo_params *p, long n, const double *eps_th, double *eps)
{
    const long dd = (long)p->dim * p->dim;
#pragma omp parallel for schedule(static)
    for (long e = 0; e < n * dd; ++e)
        eps[e] = -1.0 * eps_th[e];
}

/* VM:142-146  total - 1/dim * I * tr(total) */
void vo_deviatoric_strain(const vo_params *p, long n, const double *tot, double *dev)
{
    const int d = p->dim;
    const double inv_d = 1.0 / (double)d;
#pragma omp parallel for schedule(static)
    for (long q = 0; q < n; ++q) {
        const double *t = tot + q * d * d;
        double tr = t[0];
        for (int i = 1; i < d; ++i)
            tr = tr + t[i * d + i];
        for (int i = 0; i < d; ++i)
            for (int j = 0; j < d; ++j)
                dev[(q * d + i) * d + j] = (i == j) ? t[i * d + j] - inv_d * tr : t[i * d + j];
    }
}

/* VM:150-153 */
void vo_T_next(const vo_params *p, long n, const double *T_cur, const double *T_prev, double *T_next)
{
    (void)p;
#pragma omp parallel for schedule(static)
    for (long q = 0; q < n; ++q)
        T_next[q] = T_cur[q] + (T_cur[q] - T_prev[q]);
}

/* VM:170-173  dt/2*(phi_next - phi) */
void vo_xi(const vo_params *p, long n, const double *phi_next, const double *phi, double *xi)
{
    const double h = p->dt / 2;
#pragma omp parallel for schedule(static)
    for (long q = 0; q < n; ++q)
        xi[q] = h * (phi_next[q] - phi[q]);
}

/* VM:176-182  2.0*g_n*dev/xi*lambda_g_n*(1.0 - taylor(lambda_g_n)) */
void vo_ds_partial(const vo_params *p, long n, const double *dev, const double *xi, double *ds)
{
    const int N = p->N;
    const long dd = (long)p->dim * p->dim;
#pragma omp parallel for schedule(static)
    for (long q = 0; q < n; ++q)
        for (int t = 0; t < N; ++t) {
            double two_g = 2.0 * p->g[t];
            double one_m = 1.0 - vo_taylor(xi[q], p->lambda_g[t]);
            for (long c = 0; c < dd; ++c)
                ds[(q * N + t) * dd + c] = ((two_g * dev[q * dd + c]) / xi[q]) * p->lambda_g[t] * one_m;
        }
}

/* VM:185-191  k_n*(tr(total)*I)/xi*lambda_k_n*(1.0 - taylor(lambda_k_n)) */
void vo_dsigma_partial(const vo_params *p, long n, const double *tot, const double *xi, double *dsig)
{
    const int N = p->N, d = p->dim;
    const long dd = (long)d * d;
#pragma omp parallel for schedule(static)
    for (long q = 0; q < n; ++q) {
        const double *t = tot + q * dd;
        double tr = t[0];
        for (int i = 1; i < d; ++i)
            tr = tr + t[i * d + i];
        for (int s = 0; s < N; ++s) {
            double one_m = 1.0 - vo_taylor(xi[q], p->lambda_k[s]);
            double v = ((p->k[s] * tr) / xi[q]) * p->lambda_k[s] * one_m;
            for (int i = 0; i < d; ++i)
                for (int j = 0; j < d; ++j)
                    dsig[(q * N + s) * dd + i * d + j] = (i == j) ? v : 0.0;
        }
    }
}

/* VM:194-200 (lambda = lambda_g) and VM:203-209 (lambda = lambda_k):
 * tilde_cur[n,i,j] * taylor(lambda[n]) */
void vo_tilde_next(const vo_params *p, long n, const double *lambda, const double *tilde_cur,
                   const double *xi, double *tilde_next)
{
    const int N = p->N;
    const long dd = (long)p->dim * p->dim;
#pragma omp parallel for schedule(static)
    for (long q = 0; q < n; ++q)
        for (int t = 0; t < N; ++t) {
            double ty = vo_taylor(xi[q], lambda[t]);
            for (long c = 0; c < dd; ++c)
                tilde_next[(q * N + t) * dd + c] = tilde_cur[(q * N + t) * dd + c] * ty;
        }
}

/* VM:212-215 and VM:218-221 */
void vo_add(long count, const double *a, const double *b, double *out)
{
#pragma omp parallel for schedule(static)
    for (long e = 0; e < count; ++e)
        out[e] = a[e] + b[e];
}

/* VM:224-228  np.sum([s[n]+sig[n] for n]) -> left-nested sum */
void vo_sigma_next(const vo_params *p, long n, const double *s_part, const double *sig_part, double *sigma)
{
    const int N = p->N;
    const long dd = (long)p->dim * p->dim;
#pragma omp parallel for schedule(static)
    for (long q = 0; q < n; ++q)
        for (long c = 0; c < dd; ++c) {
            double acc = s_part[(q * N) * dd + c] + sig_part[(q * N) * dd + c];
            for (int t = 1; t < N; ++t)
                acc = acc + (s_part[(q * N + t) * dd + c] + sig_part[(q * N + t) * dd + c]);
            sigma[q * dd + c] = acc;
        }
}

static void vo_copy(long count, const double *src, double *dst)
{
#pragma omp parallel for schedule(static)
    for (long e = 0; e < count; ++e)
        dst[e] = src[e];
}

/* The 24 dolfinx Functions of TVP:106-173 for the same-space case (T space ==
 * sigma space node set): every array is indexed by node in dolfinx blocked
 * layout array[node*bs + comp]. */
typedef struct {
    double *T_cur, *T_prev, *T_next;                         /* bs 1 */
    double *Tf_partial_cur, *Tf_partial_prev;                /* bs N */
    double *Tf_cur, *Tf_prev;                                /* bs 1 */
    double *phi, *phi_next, *xi;                             /* bs 1 */
    double *thermal_strain, *total_strain, *deviatoric_strain; /* bs d*d */
    double *ds_partial, *dsigma_partial;                     /* bs N*d*d */
    double *s_tilde_cur, *s_tilde_next;
    double *sigma_tilde_cur, *sigma_tilde_next;
    double *s_partial_cur, *s_partial_next;
    double *sigma_partial_cur, *sigma_partial_next;
    double *sigma_next;                                      /* bs d*d */
} vo_state;

/* TVP:367-381 minus _solve_T and _write_output: the 17 interpolate passes and
 * the copies, in reference order ("dolfinx-shaped" CPU variant). */
void vo_step_passes(const vo_params *p, long n, vo_state *s)
{
    const long N = p->N, dd = (long)p->dim * p->dim;
    /* _solve_Tf  TVP:403-405 */
    vo_phi(p, n, s->T_cur, s->phi);                                             /* TVP:456 */
    vo_Tf_partial(p, n, s->Tf_partial_prev, s->T_cur, s->phi, s->Tf_partial_cur); /* TVP:466 */
    vo_copy(n * N, s->Tf_partial_cur, s->Tf_partial_prev);                      /* TVP:469 */
    vo_Tf(p, n, s->Tf_partial_cur, s->Tf_cur);                                  /* TVP:480 */
    vo_copy(n, s->Tf_cur, s->Tf_prev);                                          /* TVP:481 */
    /* _solve_strains  TVP:419-421 */
    vo_thermal_strain(p, n, s->T_cur, s->T_prev, s->Tf_cur, s->Tf_prev, s->thermal_strain);
    vo_total_strain(p, n, s->thermal_strain, s->total_strain);
    vo_deviatoric_strain(p, n, s->total_strain, s->deviatoric_strain);
    /* _solve_shifted_time  TVP:431-433 */
    vo_T_next(p, n, s->T_cur, s->T_prev, s->T_next);                            /* TVP:524 */
    vo_phi(p, n, s->T_cur, s->phi);                                             /* TVP:531 */
    vo_phi(p, n, s->T_next, s->phi_next);                                       /* TVP:533 */
    vo_xi(p, n, s->phi_next, s->phi, s->xi);                                    /* TVP:541 */
    /* _solve_stress  TVP:448-450 */
    vo_ds_partial(p, n, s->deviatoric_strain, s->xi, s->ds_partial);            /* TVP:549 */
    vo_tilde_next(p, n, p->lambda_g, s->s_tilde_cur, s->xi, s->s_tilde_next);   /* TVP:552 */
    vo_add(n * N * dd, s->ds_partial, s->s_tilde_next, s->s_partial_next);      /* TVP:555 */
    vo_copy(n * N * dd, s->s_tilde_next, s->s_tilde_cur);                       /* TVP:559 */
    vo_copy(n * N * dd, s->s_partial_next, s->s_partial_cur);                   /* TVP:561 */
    vo_dsigma_partial(p, n, s->total_strain, s->xi, s->dsigma_partial);         /* TVP:568 */
    vo_tilde_next(p, n, p->lambda_k, s->sigma_tilde_cur, s->xi, s->sigma_tilde_next); /* TVP:571 */
    vo_add(n * N * dd, s->dsigma_partial, s->sigma_tilde_next, s->sigma_partial_next); /* TVP:574 */
    vo_copy(n * N * dd, s->sigma_tilde_next, s->sigma_tilde_cur);               /* TVP:578 */
    vo_copy(n * N * dd, s->sigma_partial_next, s->sigma_partial_cur);           /* TVP:582 */
    vo_sigma_next(p, n, s->s_partial_next, s->sigma_partial_next, s->sigma_next); /* TVP:591 */
}

/* Same arithmetic, one sweep, minimal state (the layout the CUDA path keeps):
 * Tf_partial, s_tilde, sigma_tilde updated in place; writes phi, xi, Tf, sigma.
 * Used as the "fused" CPU baseline and to cross-check vo_step_passes. */
void vo_step_fused(const vo_params *p, long n, const double *T_cur, const double *T_prev,
                   double *Tf_partial, double *Tf, double *phi_out, double *xi_out,
                   double *s_tilde, double *sigma_tilde, double *sigma)
{
    const int N = p->N, d = p->dim;
    const long dd = (long)d * d;
    const double c = p->H / p->Rg, inv_Tb = 1.0 / p->Tb, inv_d = 1.0 / (double)d;
    const double da = p->alpha_liquid - p->alpha_solid, half_dt = p->dt / 2;
#pragma omp parallel for schedule(static)
    for (long q = 0; q < n; ++q) {
        const double Tc = T_cur[q], Tp = T_prev[q];
        const double phi = exp(c * (inv_Tb - 1.0 / Tc));
        double tf = 0.0;
        for (int i = 0; i < N; ++i) {
            double num = p->lambda_m[i] * Tf_partial[q * N + i] + (Tc * p->dt) * phi;
            double den = p->lambda_m[i] + p->dt * phi;
            double v = num / den;
            Tf_partial[q * N + i] = v;
            tf = (i == 0) ? p->m[0] * v : tf + p->m[i] * v;
        }
        Tf[q] = tf;
        const double eth = p->alpha_solid * (Tc - Tp) + da * (tf - tf);
        const double tot_d = -1.0 * eth, tot_o = -1.0 * 0.0;
        double tr = tot_d;
        for (int i = 1; i < d; ++i)
            tr = tr + tot_d;
        const double dev_d = tot_d - inv_d * tr, dev_o = tot_o;
        const double Tn = Tc + (Tc - Tp);
        const double phin = exp(c * (inv_Tb - 1.0 / Tn));
        const double xi = half_dt * (phin - phi);
        phi_out[q] = phi;
        xi_out[q] = xi;
        double acc[9];
        for (int t = 0; t < N; ++t) {
            const double tg = vo_taylor(xi, p->lambda_g[t]), tk = vo_taylor(xi, p->lambda_k[t]);
            const double two_g = 2.0 * p->g[t];
            const double ds_d = ((two_g * dev_d) / xi) * p->lambda_g[t] * (1.0 - tg);
            const double ds_o = ((two_g * dev_o) / xi) * p->lambda_g[t] * (1.0 - tg);
            const double dk_d = ((p->k[t] * tr) / xi) * p->lambda_k[t] * (1.0 - tk);
            for (int i = 0; i < d; ++i)
                for (int j = 0; j < d; ++j) {
                    const long e = (q * N + t) * dd + i * d + j;
                    const double st = s_tilde[e] * tg, sg = sigma_tilde[e] * tk;
                    s_tilde[e] = st;
                    sigma_tilde[e] = sg;
                    const double pn = ((i == j ? ds_d : ds_o) + st) + ((i == j ? dk_d : 0.0) + sg);
                    acc[i * d + j] = (t == 0) ? pn : acc[i * d + j] + pn;
                }
        }
        for (long cidx = 0; cidx < dd; ++cidx)
            sigma[q * dd + cidx] = acc[cidx];
    }
}

int vo_sizeof_params(void) { return (int)sizeof(vo_params); }
int vo_sizeof_state(void) { return (int)sizeof(vo_state); }

/* ------------------------------------------------------------------------------------------------
 * SPECIFICATION of the product's OPTIONAL "corrected" scheme (model_params["physics"] = "corrected").
 * This is NOT a restatement of reference behaviour: the reference executes the quirky chain above.  It states, in
 * plain C, the scheme the reference's comments cite (Nielsen et al. 2010) with SURVEY quirks Q1-Q4 and Q14 removed,
 * so that the CUDA kernels of that mode have an independent CPU statement to be tested against:
 *   phi      = exp(H/Rg (1/Tb - chi/T_cur - (1-chi)/Tf_prev))      (VM:100-108, dead in the reference)
 *   Tf_partial, Tf as VM:111-125 with this phi
 *   d eps_th = alpha_s (T_cur - T_prev) + (alpha_l - alpha_s)(Tf_cur - Tf_prev)   (VM:128-133 with the old Tf_prev)
 *   xi       = dt/2 (phi(T_prev, Tf_prev) + phi(T_cur, Tf_cur))
 *   x = xi/lambda: decay = 1 + expm1(-x), fac = -expm1(-x)/x (1 if x == 0)
 *   s_n <- s_n decay + 2 g_n dev fac ; sigma_n <- sigma_n decay + k_n tr I fac ; sigma = sum_n (s_n + sigma_n)
 * s_hist / k_hist hold the partial stresses themselves. */
void vo_step_corrected(const vo_params *p, double chi, long n, const double *T_cur, const double *T_prev,
                       double *Tf_partial, double *Tf, double *phi_out, double *xi_out,
                       double *s_hist, double *k_hist, double *sigma)
{
    const int N = p->N, d = p->dim;
    const long dd = (long)d * d;
    const double c = p->H / p->Rg, inv_Tb = 1.0 / p->Tb, inv_d = 1.0 / (double)d;
    const double da = p->alpha_liquid - p->alpha_solid, half_dt = p->dt / 2;
#pragma omp parallel for schedule(static)
    for (long q = 0; q < n; ++q) {
        const double Tc = T_cur[q], Tp = T_prev[q], Tfo = Tf[q];
        const double phi = exp(c * (inv_Tb - chi / Tc - (1.0 - chi) / Tfo));
        const double phi_old = exp(c * (inv_Tb - chi / Tp - (1.0 - chi) / Tfo));
        double tf = 0.0;
        for (int i = 0; i < N; ++i) {
            double v = (p->lambda_m[i] * Tf_partial[q * N + i] + (Tc * p->dt) * phi) / (p->lambda_m[i] + p->dt * phi);
            Tf_partial[q * N + i] = v;
            tf = (i == 0) ? p->m[0] * v : tf + p->m[i] * v;
        }
        Tf[q] = tf;
        const double xi = half_dt * (phi_old + exp(c * (inv_Tb - chi / Tc - (1.0 - chi) / tf)));
        phi_out[q] = phi;
        xi_out[q] = xi;
        const double eth = p->alpha_solid * (Tc - Tp) + da * (tf - Tfo);
        const double tot_d = -1.0 * eth, tot_o = -1.0 * 0.0;
        double tr = tot_d;
        for (int i = 1; i < d; ++i)
            tr = tr + tot_d;
        const double dev_d = tot_d - inv_d * tr, dev_o = tot_o;
        double acc[9];
        for (int t = 0; t < N; ++t) {
            const double xg = xi / p->lambda_g[t], xk = xi / p->lambda_k[t];
            const double eg = expm1(-xg), ek = expm1(-xk);
            const double dg = 1.0 + eg, dk = 1.0 + ek;
            const double fg = (xg != 0.0) ? (-eg) / xg : 1.0, fk = (xk != 0.0) ? (-ek) / xk : 1.0;
            const double two_g = 2.0 * p->g[t];
            for (int i = 0; i < d; ++i)
                for (int j = 0; j < d; ++j) {
                    const long e = (q * N + t) * dd + (long)i * d + j;
                    const double ds = (two_g * (i == j ? dev_d : dev_o)) * fg;
                    const double dks = (i == j) ? (p->k[t] * tr) * fk : 0.0;
                    const double sp = ds + s_hist[e] * dg, kp = dks + k_hist[e] * dk;
                    s_hist[e] = sp;
                    k_hist[e] = kp;
                    const double pn = sp + kp;
                    acc[i * d + j] = (t == 0) ? pn : acc[i * d + j] + pn;
                }
        }
        for (long e = 0; e < dd; ++e)
            sigma[q * dd + e] = acc[e];
    }
}
