/* TEST INFRASTRUCTURE / CPU BASELINE ONLY — never linked into the product.
 *
 * Jacobi-preconditioned conjugate gradients on an assembled CSR matrix with OpenMP: the CPU stand-in for the
 * PETSc KSP 'cg' solve the reference runs inside NewtonSolver (ThermoViscoProblem.py:339-346) when bench.py times
 * the reference-shaped CPU port on all host cores (scipy.sparse.linalg.cg is single-threaded).
 * Only bench.py's cpu_baseline / --impl reference legs call this. */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

static void spmv(long n, const int32_t *indptr, const int32_t *indices, const double *data, const double *x, double *y) {
#pragma omp parallel for schedule(static)
    for (long i = 0; i < n; ++i) {
        double s = 0.0;
        for (int32_t k = indptr[i]; k < indptr[i + 1]; ++k) s += data[k] * x[indices[k]];
        y[i] = s;
    }
}

/* Solves A x = b from x = 0 to |r| <= rtol |b|.  Returns the iteration count, -1 if maxit was reached. */
int cpu_pcg_jacobi(long n, const int32_t *indptr, const int32_t *indices, const double *data, const double *dinv,
                   const double *b, double *x, double rtol, int maxit) {
    double *r = (double *)malloc(sizeof(double) * n), *p = (double *)malloc(sizeof(double) * n),
           *Ap = (double *)malloc(sizeof(double) * n);
    double rz = 0.0, bb = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : rz, bb)
    for (long i = 0; i < n; ++i) {
        x[i] = 0.0;
        r[i] = b[i];
        p[i] = dinv[i] * b[i];
        rz += r[i] * p[i];
        bb += b[i] * b[i];
    }
    const double tol2 = rtol * rtol * bb;
    int it = 0;
    double rr = bb;
    while (rr > tol2 && it < maxit) {
        spmv(n, indptr, indices, data, p, Ap);
        double pAp = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : pAp)
        for (long i = 0; i < n; ++i) pAp += p[i] * Ap[i];
        const double alpha = rz / pAp;
        double rz_new = 0.0;
        rr = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : rz_new, rr)
        for (long i = 0; i < n; ++i) {
            x[i] += alpha * p[i];
            r[i] -= alpha * Ap[i];
            rz_new += r[i] * r[i] * dinv[i];
            rr += r[i] * r[i];
        }
        const double beta = rz_new / rz;
        rz = rz_new;
#pragma omp parallel for schedule(static)
        for (long i = 0; i < n; ++i) p[i] = dinv[i] * r[i] + beta * p[i];
        ++it;
    }
    free(r);
    free(p);
    free(Ap);
    return rr > tol2 ? -1 : it;
}

/* y = A x on all threads (residual evaluation of the CPU port). */
void cpu_spmv(long n, const int32_t *indptr, const int32_t *indices, const double *data, const double *x, double *y) {
    spmv(n, indptr, indices, data, x, y);
}

#ifdef _OPENMP
#include <omp.h>
/* torchrun exports OMP_NUM_THREADS=1; the CPU baseline sets its thread count explicitly and reports what it got. */
int cpu_set_threads(int n) {
    if (n > 0) omp_set_num_threads(n);
    return omp_get_max_threads();
}
#else
int cpu_set_threads(int n) {
    (void)n;
    return 1;
}
#endif
