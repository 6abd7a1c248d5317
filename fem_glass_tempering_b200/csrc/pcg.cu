// Preconditioned conjugate gradients + inexact-Newton time-step driver + ghost-dof halo for hot path (B).
//
// Replaces dolfinx.nls.petsc.NewtonSolver (incremental criterion, TVP:334-337) and its PETSc KSP 'cg'
// (TVP:343) — GAMG is not reproduced: the Jacobian M + dt*(alpha*K + boundary) is mass-dominated.
//
//   * plain iteration (CG spaces: point Jacobi; DG: element-mass blocks): 3 kernels per iteration, batches of 8
//     replayed as a CUDA graph on one GPU;
//   * DG with class tables: CG preconditioned by a Chebyshev polynomial in M^-1 J whose steps are fused with the
//     operator apply (thermal.cu dg_cheb_step), interval from Lanczos Ritz values (estimate_spectrum);
//   * all CG scalars (r.z, p.Ap, |r|^2), the tolerance and the iteration count live on the device: every fused vector
//     kernel ends with warp shuffles -> per-block partials -> "last block sums in fixed order" (deterministic), the
//     kernels test convergence themselves and later launches of a converged solve return at once; the host reads
//     the scalars through a mapped pinned mirror (k_mirror), never with a D2H memcpy;
//   * Newton: dolfinx's incremental criterion with Eisenstat-Walker forcing terms for the linear solves;
//   * the solver runs on its own stream, ordered against the caller's stream with events (StreamScope);
//   * across GPUs the 1-2 doubles are all-reduced and the ghost dofs refreshed over NVLink peer memory (peer.cu),
//     with NCCL (in-place ncclAllReduce, grouped ncclSend/ncclRecv of contiguous ranges) as the fallback.
#include <math.h>
#include <stdlib.h>

#include "sg_common.cuh"
#include "sg_nccl.h"

namespace {

constexpr int VB = 256;          // threads per block of the vector kernels
constexpr int BATCH = 8;         // PCG iterations enqueued between two host checks of the convergence flag

// Device scalars S (doubles): S[0..1] / S[2..3] = {r.z, r.r} ping-pong by iteration parity,
// S[4] + S[5] = p.Ap (cells + exterior facets), S[6] = |dx|^2.
// Device control block: the kernels themselves detect convergence, later launches of the same solve
// return immediately, and the host looks at the flag once per BATCH iterations.
struct PcgCtrl {
    int done;      // 0 running, 1 converged, 2 non-finite residual
    int iters;     // iterations completed when `done` was set
    double rr;     // |r|^2 at that point
    double tol2;   // |r|^2 target of the running solve (k_pcg_begin); lives on the device so that a captured
    int it_count;  // CUDA graph of a batch of iterations is valid for every solve; iterations executed so far
};

using Red = SgRed;

template <int NR>
__device__ __forceinline__ void grid_reduce(double (&v)[NR], const Red red, double *out) {
    sg_grid_reduce<NR>(v, red, out);
}

__device__ __forceinline__ bool owned(long i, long lo, long hi) { return i >= lo && i < hi; }

// x = 0, r = b, p = z = dinv*r ; S[0] = r.z, S[1] = r.r ; resets the control block
__global__ void __launch_bounds__(VB) k_pcg_init(long i0, long i1, long lo, long hi, const double *__restrict__ b,
                                                 const double *__restrict__ dinv, double *__restrict__ x,
                                                 double *__restrict__ r, double *__restrict__ p, Red red, double *S,
                                                 PcgCtrl *ctrl) {
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        ctrl->done = 0;
        ctrl->iters = 0;
        ctrl->rr = 0.0;
    }
    double acc[2] = {0.0, 0.0};
    for (long i = i0 + (long)blockIdx.x * VB + threadIdx.x; i < i1; i += (long)gridDim.x * VB) {
        const double ri = b[i], zi = dinv[i] * ri;
        x[i] = 0.0;
        r[i] = ri;
        p[i] = zi;
        if (owned(i, lo, hi)) {
            acc[0] += ri * zi;
            acc[1] += ri * ri;
        }
    }
    grid_reduce<2>(acc, red, S);
}

// alpha = rz/pAp ; x += alpha p ; r -= alpha Ap ; Snext = {r.(dinv r), r.r}
__global__ void __launch_bounds__(VB) k_update_xr(long i0, long i1, long lo, long hi, const double *__restrict__ p,
                                                  const double *__restrict__ Ap, const double *__restrict__ dinv,
                                                  double *__restrict__ x, double *__restrict__ r, Red red,
                                                  const double *Scur, const double *SpAp, double *Snext,
                                                  const PcgCtrl *ctrl) {
    if (ctrl->done) return;
    const double alpha = Scur[0] / (SpAp[0] + SpAp[1]);
    double acc[2] = {0.0, 0.0};
    for (long i = i0 + (long)blockIdx.x * VB + threadIdx.x; i < i1; i += (long)gridDim.x * VB) {
        x[i] += alpha * p[i];
        const double ri = r[i] - alpha * Ap[i];
        r[i] = ri;
        if (owned(i, lo, hi)) {
            acc[0] += ri * ri * dinv[i];
            acc[1] += ri * ri;
        }
    }
    grid_reduce<2>(acc, red, Snext);
}

// Convergence test by one thread with the target and the iteration count kept in the control block.
__device__ __forceinline__ void pcg_check_counted(PcgCtrl *ctrl, const double *Snext) {
    const int it = ctrl->it_count;
    const double rr = Snext[1];
    if (!(rr > ctrl->tol2) || !isfinite(rr)) {
        ctrl->done = isfinite(rr) ? 1 : 2;
        ctrl->iters = it + 1;
        ctrl->rr = rr;
    }
    ctrl->it_count = it + 1;
}

__global__ void k_pcg_begin(PcgCtrl *ctrl, double tol2) {
    ctrl->tol2 = tol2;
    ctrl->it_count = 0;
}

// Convergence test of iteration `it` (0-based), by one thread, on the (all-reduced) new residual norm.
__device__ __forceinline__ void pcg_check(PcgCtrl *ctrl, const double *Snext, double tol2, int it) {
    const double rr = Snext[1];
    if (!(rr > tol2) || !isfinite(rr)) {
        ctrl->done = isfinite(rr) ? 1 : 2;
        ctrl->iters = it + 1;
        ctrl->rr = rr;
    }
}

// beta = rz_new/rz ; p = dinv r + beta p ; flags convergence for the launches that follow
__global__ void __launch_bounds__(VB) k_update_p(long n, long i0, long i1, const double *__restrict__ r, const double *__restrict__ dinv,
                                                 double *__restrict__ p, double *__restrict__ Ap, const double *Scur,
                                                 const double *Snext, PcgCtrl *ctrl) {
    if (ctrl->done) return;
    const double beta = Snext[0] / Scur[0];
    for (long i = (long)blockIdx.x * VB + threadIdx.x; i < n; i += (long)gridDim.x * VB) {
        if (i >= i0 && i < i1) p[i] = dinv[i] * r[i] + beta * p[i];   // ghost rows of p belong to the neighbour's put
        Ap[i] = 0.0;     // the next operator apply scatters into Ap (RED.ADD): zero it here instead of a memset launch
    }
    // every block has read `done` before block 0 can change it?  No ordering is needed: a block that
    // sees done != 0 set by THIS kernel merely skips a p update nobody will use.
    if (blockIdx.x == 0 && threadIdx.x == 0) pcg_check_counted(ctrl, Snext);
}

__global__ void __launch_bounds__(VB) k_invert(long n, double *__restrict__ d) {
    for (long i = (long)blockIdx.x * VB + threadIdx.x; i < n; i += (long)gridDim.x * VB) d[i] = 1.0 / d[i];
}

// T -= dx ; out = |dx|^2 over owned dofs
__global__ void __launch_bounds__(VB) k_newton_update(long i0, long i1, long lo, long hi, double *__restrict__ T,
                                                      const double *__restrict__ dx, Red red, double *out) {
    double acc[1] = {0.0};
    for (long i = i0 + (long)blockIdx.x * VB + threadIdx.x; i < i1; i += (long)gridDim.x * VB) {
        const double d = dx[i];
        T[i] -= d;
        if (owned(i, lo, hi)) acc[0] += d * d;
    }
    grid_reduce<1>(acc, red, out);
}

// ---------------------------------------------------------------------------------------------
// DG: the mass matrix is block diagonal, M_K = |detJ_K| * Mhat, so z_K = Mhat^-1 r_K / |detJ_K| is an
// exact, set-up-free preconditioner for the mass-dominated Jacobian (about 1.7x fewer CG iterations
// than point Jacobi).  One thread per cell; the NLD dofs of a cell are contiguous (128-bit accesses
// when NLD is even).
template <int NLD>
struct MassInv {
    double a[NLD * NLD];  // Mhat^-1
};

template <int NLD>
__device__ __forceinline__ void mass_solve(const MassInv<NLD> &mi, double inv_det, const double (&r)[NLD], double (&z)[NLD]) {
#pragma unroll
    for (int i = 0; i < NLD; ++i) {
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < NLD; ++j) s += mi.a[i * NLD + j] * r[j];
        z[i] = s * inv_det;
    }
}

template <int NLD>
__device__ __forceinline__ void ld_row(const double *__restrict__ p, double (&v)[NLD]) {
    if constexpr (NLD % 2 == 0) {
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
template <int NLD>
__device__ __forceinline__ void st_row(double *__restrict__ p, const double (&v)[NLD]) {
    if constexpr (NLD % 2 == 0) {
#pragma unroll
        for (int i = 0; i < NLD / 2; ++i) reinterpret_cast<double2 *>(p)[i] = make_double2(v[2 * i], v[2 * i + 1]);
    } else {
#pragma unroll
        for (int i = 0; i < NLD; ++i) p[i] = v[i];
    }
}

template <int NLD>
__global__ void __launch_bounds__(VB) k_pcg_init_blk(const __grid_constant__ MassInv<NLD> mi, int rev, long c0, long c1, long clo, long chi,
                                                     const double *__restrict__ detJ, const double *__restrict__ b,
                                                     double *__restrict__ x, double *__restrict__ r,
                                                     double *__restrict__ p, Red red, double *S, PcgCtrl *ctrl) {
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        ctrl->done = 0;
        ctrl->iters = 0;
        ctrl->rr = 0.0;
    }
    double acc[2] = {0.0, 0.0};
    long c, cstep;
    sg_sweep_begin(c0, c1, VB, rev, c, cstep);
    for (; c >= c0 && c < c1; c += cstep) {
        double rk[NLD], zk[NLD], zero[NLD];
        ld_row<NLD>(b + c * NLD, rk);
        mass_solve<NLD>(mi, 1.0 / detJ[c], rk, zk);
#pragma unroll
        for (int i = 0; i < NLD; ++i) zero[i] = 0.0;
        st_row<NLD>(x + c * NLD, zero);
        st_row<NLD>(r + c * NLD, rk);
        st_row<NLD>(p + c * NLD, zk);
        if (c >= clo && c < chi) {
#pragma unroll
            for (int i = 0; i < NLD; ++i) {
                acc[0] += rk[i] * zk[i];
                acc[1] += rk[i] * rk[i];
            }
        }
    }
    grid_reduce<2>(acc, red, S);
}

template <int NLD>
__global__ void __launch_bounds__(VB) k_update_xr_blk(const __grid_constant__ MassInv<NLD> mi, int rev, long c0, long c1, long clo, long chi,
                                                      const double *__restrict__ detJ, const double *__restrict__ p,
                                                      const double *__restrict__ Ap, double *__restrict__ x,
                                                      double *__restrict__ r, Red red, const double *Scur,
                                                      const double *SpAp, double *Snext, const PcgCtrl *ctrl) {
    if (ctrl->done) return;
    const double alpha = Scur[0] / (SpAp[0] + SpAp[1]);
    double acc[2] = {0.0, 0.0};
    long c, cstep;
    sg_sweep_begin(c0, c1, VB, rev, c, cstep);
    for (; c >= c0 && c < c1; c += cstep) {
        double pk[NLD], ak[NLD], xk[NLD], rk[NLD], zk[NLD];
        ld_row<NLD>(p + c * NLD, pk);
        ld_row<NLD>(Ap + c * NLD, ak);
        ld_row<NLD>(x + c * NLD, xk);
        ld_row<NLD>(r + c * NLD, rk);
        const double idet = 1.0 / detJ[c];
#pragma unroll
        for (int i = 0; i < NLD; ++i) {
            xk[i] += alpha * pk[i];
            rk[i] -= alpha * ak[i];
        }
        st_row<NLD>(x + c * NLD, xk);
        st_row<NLD>(r + c * NLD, rk);
        if (c >= clo && c < chi) {
            mass_solve<NLD>(mi, idet, rk, zk);
#pragma unroll
            for (int i = 0; i < NLD; ++i) {
                acc[0] += rk[i] * zk[i];
                acc[1] += rk[i] * rk[i];
            }
        }
    }
    grid_reduce<2>(acc, red, Snext);
}

template <int NLD>
__global__ void __launch_bounds__(VB) k_update_p_blk(const __grid_constant__ MassInv<NLD> mi, int rev, long c0, long c1,
                                                     const double *__restrict__ detJ, const double *__restrict__ r,
                                                     double *__restrict__ p, const double *Scur, const double *Snext,
                                                     PcgCtrl *ctrl) {
    if (ctrl->done) return;
    const double beta = Snext[0] / Scur[0];
    long c, cstep;
    sg_sweep_begin(c0, c1, VB, rev, c, cstep);
    for (; c >= c0 && c < c1; c += cstep) {
        double rk[NLD], zk[NLD], pk[NLD];
        ld_row<NLD>(r + c * NLD, rk);
        ld_row<NLD>(p + c * NLD, pk);
        mass_solve<NLD>(mi, 1.0 / detJ[c], rk, zk);
#pragma unroll
        for (int i = 0; i < NLD; ++i) pk[i] = zk[i] + beta * pk[i];
        st_row<NLD>(p + c * NLD, pk);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) pcg_check_counted(ctrl, Snext);
}

// ---------------------------------------------------------------------------------------------
// Chebyshev-polynomial preconditioned CG (DG): z = q_k(M^-1 J) M^-1 r from k fused operator steps
// (thermal.cu dg_cheb_step).  The vector kernels around it:
//   init:   x = 0, r = b, z1 = M^-1 b / theta, |b|^2
//   x,r:    x += alpha p, r -= alpha Ap, z1 = M^-1 r / theta, |r|^2
//   p:      p = z + beta p
template <int NLD>
__global__ void __launch_bounds__(VB) k_pcg_init_cheb(const __grid_constant__ MassInv<NLD> mi, int rev, long c0, long c1, long clo, long chi,
                                                      const double *__restrict__ detJ, const double *__restrict__ b,
                                                      double *__restrict__ x, double *__restrict__ r, double *__restrict__ z1,
                                                      double inv_theta, Red red, double *S1, PcgCtrl *ctrl) {
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        ctrl->done = 0;
        ctrl->iters = 0;
        ctrl->rr = 0.0;
    }
    double acc[1] = {0.0};
    long c, cstep;
    sg_sweep_begin(c0, c1, VB, rev, c, cstep);
    for (; c >= c0 && c < c1; c += cstep) {
        double rk[NLD], zk[NLD], zero[NLD];
        ld_row<NLD>(b + c * NLD, rk);
        mass_solve<NLD>(mi, inv_theta / detJ[c], rk, zk);
#pragma unroll
        for (int i = 0; i < NLD; ++i) zero[i] = 0.0;
        st_row<NLD>(x + c * NLD, zero);
        st_row<NLD>(r + c * NLD, rk);
        st_row<NLD>(z1 + c * NLD, zk);
        if (c >= clo && c < chi) {
#pragma unroll
            for (int i = 0; i < NLD; ++i) acc[0] += rk[i] * rk[i];
        }
    }
    grid_reduce<1>(acc, red, S1);
}

template <int NLD>
__global__ void __launch_bounds__(VB) k_update_xr_cheb(const __grid_constant__ MassInv<NLD> mi, int rev, long c0, long c1, long clo, long chi,
                                                       const double *__restrict__ detJ, const double *__restrict__ p,
                                                       const double *__restrict__ Ap, double *__restrict__ x,
                                                       double *__restrict__ r, double *__restrict__ z1, double inv_theta,
                                                       Red red, const double *Scur, const double *SpAp, double *Snext1,
                                                       const PcgCtrl *ctrl) {
    if (ctrl->done) return;
    const double alpha = Scur[0] / (SpAp[0] + SpAp[1]);
    double acc[1] = {0.0};
    long c, cstep;
    sg_sweep_begin(c0, c1, VB, rev, c, cstep);
    for (; c >= c0 && c < c1; c += cstep) {
        double pk[NLD], ak[NLD], xk[NLD], rk[NLD], zk[NLD];
        ld_row<NLD>(p + c * NLD, pk);
        ld_row<NLD>(Ap + c * NLD, ak);
        ld_row<NLD>(x + c * NLD, xk);
        ld_row<NLD>(r + c * NLD, rk);
#pragma unroll
        for (int i = 0; i < NLD; ++i) {
            xk[i] += alpha * pk[i];
            rk[i] -= alpha * ak[i];
        }
        mass_solve<NLD>(mi, inv_theta / detJ[c], rk, zk);
        st_row<NLD>(x + c * NLD, xk);
        st_row<NLD>(r + c * NLD, rk);
        st_row<NLD>(z1 + c * NLD, zk);
        if (c >= clo && c < chi) {
#pragma unroll
            for (int i = 0; i < NLD; ++i) acc[0] += rk[i] * rk[i];
        }
    }
    grid_reduce<1>(acc, red, Snext1);
}

__global__ void k_pcg_check(PcgCtrl *ctrl, const double *Snext, double tol2, int it) {
    if (!ctrl->done) pcg_check(ctrl, Snext, tol2, it);
}

// p = z + beta p, beta = Snext[0]/Scur[0]  (first == 1: p = z)
__global__ void __launch_bounds__(VB) k_axpy_p(long n, int rev, const double *__restrict__ z, double *__restrict__ p, const double *Scur,
                                               const double *Snext, int first, const PcgCtrl *ctrl) {   // z, p: first row of the range
    if (ctrl->done) return;
    const double beta = first ? 0.0 : Snext[0] / Scur[0];
    const long n2 = ((((uintptr_t)z | (uintptr_t)p) & 15) == 0) ? n >> 1 : 0;   // workspace slices start at odd dofs when n is odd
    const double2 *z2 = reinterpret_cast<const double2 *>(z);
    double2 *p2 = reinterpret_cast<double2 *>(p);
    long i, istep;
    sg_sweep_begin(0, n2, VB, rev, i, istep);
    for (; i >= 0 && i < n2; i += istep) {
        const double2 a = z2[i];
        double2 q = first ? make_double2(0.0, 0.0) : p2[i];
        q.x = a.x + beta * q.x;
        q.y = a.y + beta * q.y;
        p2[i] = q;
    }
    for (long j = 2 * n2 + (long)blockIdx.x * VB + threadIdx.x; j < n; j += (long)gridDim.x * VB)
        p[j] = z[j] + (first ? 0.0 : beta * p[j]);
}

// deterministic pseudo-random start vector in (-1, 1) from the GLOBAL dof index (same field for every partition)
__global__ void __launch_bounds__(VB) k_hash_fill(long n, long global_offset, long lo, long hi, double *__restrict__ v, Red red,
                                                  double *out) {
    double acc[1] = {0.0};
    for (long i = (long)blockIdx.x * VB + threadIdx.x; i < n; i += (long)gridDim.x * VB) {
        unsigned long long h = (unsigned long long)(i + global_offset) * 0x9E3779B97F4A7C15ull;
        h ^= h >> 29;
        h *= 0xBF58476D1CE4E5B9ull;
        h ^= h >> 32;
        const double val = (double)(h >> 11) * (2.0 / 9007199254740992.0) - 1.0;
        v[i] = val;
        if (owned(i, lo, hi)) acc[0] += val * val;
    }
    grid_reduce<1>(acc, red, out);
}

// The host reads solver scalars from a MAPPED pinned mirror that this kernel fills with plain stores: no
// cudaMemcpy, so the reads never queue behind a large device->host output copy on the DMA engine
// (output.HostMirror streams 2 GB per step on a side stream).
__global__ void k_mirror(const double *S, const PcgCtrl *ctrl, double *S_host, PcgCtrl *ctrl_host) {
    if (threadIdx.x < 8) S_host[threadIdx.x] = S[threadIdx.x];
    if (threadIdx.x == 8) *ctrl_host = *ctrl;
}

inline unsigned vgrid(long n) {
    long g = (n + VB - 1) / VB;
    if (g < 1) g = 1;
    return (unsigned)(g > SG_MAX_BLOCKS ? SG_MAX_BLOCKS : g);
}

}  // namespace

using PcgTol = SgPcgPolicy;

struct sg_halo_plan {
    sg_ctx *ctx;
    int n;
    sg_halo_segment *seg;
    SgPeer *peer;      // NVLink peer-memory path (sg_halo_peer_alloc/open); NULL: NCCL send/recv
};

struct sg_thermal_solver {
    sg_thermal_op *op;
    sg_ctx *ctx;
    sg_halo_plan *halo;
    long n, lo, hi;
    double *b, *dx, *r, *p, *Ap, *dinv;   // workspace views
    Red red;
    double *S;        // device scalars [8]
    double *S_host;   // pinned mirror
    PcgCtrl *ctrl, *ctrl_host;            // device control block + pinned mirror
    // DG element-mass preconditioner
    int blk_nld;          // 0 = point Jacobi (CG), else dofs per cell
    long n_cells, clo, chi;
    const double *detJ;
    double mass_inv[100];
    double last_rhs_norm;
    // Chebyshev polynomial preconditioner (DG + class tables): degree = operator applications inside it
    int cheb_degree;            // 0 = off
    double cheb_lo, cheb_hi;    // spectrum bounds of M^-1 J in use (hi <= 0: estimate at first use)
    double *zA, *zB;            // workspace views
    int cheb_failed;
    // The solver runs on its own non-blocking stream (ordered against the caller's stream with events): batches of
    // plain PCG iterations are replayed as a CUDA graph, which cannot be captured on the legacy default stream.
    cudaStream_t own;
    cudaEvent_t ev_in, ev_out;
    cudaGraphExec_t batch_graph;
    const double *graph_T, *graph_x;
    int use_graphs;
    // Direct halo puts run on a side stream: the put of exchange e waits (event) for the kernel that produced its rows and
    // the consumer - which only needs the NEIGHBOUR's put, awaited inside the kernel - starts at once, so neither the put's
    // launch nor its transfer sits on the solver's critical path.  The next exchange first makes the solver stream wait for
    // put e (long finished by then): a later kernel may overwrite the rows it read.
    cudaStream_t side;
    cudaEvent_t ev_produced, ev_put;
    int put_pending;
};

namespace {

// Cross-rank sum of `count` doubles the kernel launched just before has reduced.  On the peer-memory path that kernel's
// last block already did it (SgRed::peer, sg_grid_reduce), so this is a no-op; NCCL otherwise.
int allreduce(sg_thermal_solver *s, double *ptr, int count, cudaStream_t st) {
    if (s->ctx->nranks == 1 || s->red.peer) return SG_OK;
    if (s->halo && sg_peer_ready(s->halo->peer)) return sg_peer_allreduce(s->halo->peer, ptr, count, st);
    SG_CHECK_NCCL(sg_nccl()->AllReduce(ptr, ptr, (size_t)count, ncclDouble, ncclSum, s->ctx->comm, st));
    return SG_OK;
}

int halo_drain(sg_thermal_solver *s, cudaStream_t st);

int read_scalars(sg_thermal_solver *s, int first, int count, cudaStream_t st) {
    (void)first;
    (void)count;
    {
        const int rcd = halo_drain(s, st);
        if (rcd) return rcd;
    }
    k_mirror<<<1, 32, 0, st>>>(s->S, s->ctrl, s->S_host, s->ctrl_host);
    SG_CHECK_CUDA(cudaGetLastError());
    SG_CHECK_CUDA(cudaStreamSynchronize(st));
    if (s->halo && s->halo->peer) return sg_peer_check(s->halo->peer);
    return SG_OK;
}

// Start the ghost-row exchange of a solver vector.  Peer path: one put kernel, *wait describes what the consumer has to
// wait for (inside its kernel where supported).  NCCL path: grouped send/recv, complete in stream order.
int halo_start(sg_thermal_solver *s, double *vec, SgHaloWait *wait, cudaStream_t st) {
    memset(wait, 0, sizeof(*wait));
    if (!s->halo) return SG_OK;
    if (!sg_peer_ready(s->halo->peer)) return sg_halo_forward(s->halo, vec, 1, st);
    if (!s->side || !sg_peer_in_workspace(s->halo->peer, vec))
        return sg_peer_put(s->halo->peer, s->halo->n, s->halo->seg, vec, wait, st);
    if (s->put_pending) SG_CHECK_CUDA(cudaStreamWaitEvent(st, s->ev_put, 0));      // the previous put has read its rows
    SG_CHECK_CUDA(cudaEventRecord(s->ev_produced, st));
    SG_CHECK_CUDA(cudaStreamWaitEvent(s->side, s->ev_produced, 0));
    const int rc = sg_peer_put(s->halo->peer, s->halo->n, s->halo->seg, vec, wait, s->side);
    if (rc) return rc;
    SG_CHECK_CUDA(cudaEventRecord(s->ev_put, s->side));
    s->put_pending = 1;
    return SG_OK;
}

// Before anything outside the iteration touches the workspace: the last side-stream put must have finished.
int halo_drain(sg_thermal_solver *s, cudaStream_t st) {
    if (s->put_pending) {
        SG_CHECK_CUDA(cudaStreamWaitEvent(st, s->ev_put, 0));
        s->put_pending = 0;
    }
    return SG_OK;
}

template <int NLD>
MassInv<NLD> mass_inv_of(const sg_thermal_solver *s) {
    MassInv<NLD> mi;
    for (int i = 0; i < NLD * NLD; ++i) mi.a[i] = s->mass_inv[i];
    return mi;
}

template <int NLD>
int blk_init(sg_thermal_solver *s, const double *b, double *x, cudaStream_t st) {
    k_pcg_init_blk<NLD><<<vgrid(s->chi - s->clo), VB, 0, st>>>(mass_inv_of<NLD>(s), sg_next_sweep_dir(), s->clo, s->chi, s->clo, s->chi, s->detJ, b, x, s->r,
                                                         s->p, s->red, s->S, s->ctrl);
    SG_CHECK_CUDA(cudaGetLastError());
    sg_count_launch();
    return SG_OK;
}
template <int NLD>
int blk_update_xr(sg_thermal_solver *s, double *x, const double *Scur, double *Snext, cudaStream_t st) {
    k_update_xr_blk<NLD><<<vgrid(s->chi - s->clo), VB, 0, st>>>(mass_inv_of<NLD>(s), sg_next_sweep_dir(), s->clo, s->chi, s->clo, s->chi, s->detJ, s->p, s->Ap,
                                                          x, s->r, s->red, Scur, s->S + 4, Snext, s->ctrl);
    SG_CHECK_CUDA(cudaGetLastError());
    sg_count_launch();
    return SG_OK;
}
template <int NLD>
int blk_update_p(sg_thermal_solver *s, const double *Scur, const double *Snext, cudaStream_t st) {
    k_update_p_blk<NLD><<<vgrid(s->chi - s->clo), VB, 0, st>>>(mass_inv_of<NLD>(s), sg_next_sweep_dir(), s->clo, s->chi, s->detJ, s->r, s->p, Scur, Snext,
                                                         s->ctrl);
    SG_CHECK_CUDA(cudaGetLastError());
    sg_count_launch();
    return SG_OK;
}

#define SG_BLK_DISPATCH(fn, ...)                                   \
    switch (s->blk_nld) {                                          \
        case 2: rc = fn<2>(__VA_ARGS__); break;                    \
        case 3: rc = fn<3>(__VA_ARGS__); break;                    \
        case 4: rc = fn<4>(__VA_ARGS__); break;                    \
        case 6: rc = fn<6>(__VA_ARGS__); break;                    \
        case 10: rc = fn<10>(__VA_ARGS__); break;                  \
        default: sg_set_error("unsupported dofs per cell %d", s->blk_nld); rc = SG_E_UNSUPPORTED; \
    }

template <int NLD>
int blk_init_cheb(sg_thermal_solver *s, const double *b, double *x, double inv_theta, cudaStream_t st) {
    k_pcg_init_cheb<NLD><<<vgrid(s->chi - s->clo), VB, 0, st>>>(mass_inv_of<NLD>(s), sg_next_sweep_dir(), s->clo, s->chi, s->clo, s->chi, s->detJ, b, x, s->r,
                                                          s->zA, inv_theta, s->red, s->S + 1, s->ctrl);
    SG_CHECK_CUDA(cudaGetLastError());
    sg_count_launch();
    return SG_OK;
}
template <int NLD>
int blk_update_xr_cheb(sg_thermal_solver *s, double *x, double inv_theta, const double *Scur, double *Snext, cudaStream_t st) {
    k_update_xr_cheb<NLD><<<vgrid(s->chi - s->clo), VB, 0, st>>>(mass_inv_of<NLD>(s), sg_next_sweep_dir(), s->clo, s->chi, s->clo, s->chi, s->detJ, s->p, s->Ap,
                                                           x, s->r, s->zA, inv_theta, s->red, Scur, s->S + 4, Snext + 1, s->ctrl);
    SG_CHECK_CUDA(cudaGetLastError());
    sg_count_launch();
    return SG_OK;
}
// Number of eigenvalues of the symmetric tridiagonal matrix (diagonal a, off-diagonal o) below x (Sturm sequence).
static int sturm_count(const double *a, const double *o, int m, double x) {
    int cnt = 0;
    double d = 1.0;
    for (int i = 0; i < m; ++i) {
        const double off2 = i ? o[i - 1] * o[i - 1] : 0.0;
        d = a[i] - x - (i ? off2 / (d != 0.0 ? d : 1e-300) : 0.0);
        if (d < 0.0) ++cnt;
    }
    return cnt;
}

// Extreme eigenvalues of M^-1 J(T_lin) from the Lanczos tridiagonal matrix that CG builds implicitly: LANCZOS steps of
// the element-mass preconditioned iteration on a pseudo-random right-hand side, alpha/beta read back every step
// (set-up only), largest/smallest Ritz value by bisection.  Ritz values converge to the extreme eigenvalues from
// inside within a fraction of a per cent in ~20 steps (20 power iterations are still 4 % short), so the Chebyshev
// interval can be tight: hi = 1.05 * ritz_max.  An UNDER-estimated upper bound makes the polynomial preconditioner
// indefinite; pcg_run widens the interval and retries when that is detected.
int estimate_spectrum(sg_thermal_solver *s, const double *T_lin, cudaStream_t st) {
    constexpr int LANCZOS = 24;
    int rc;
    double *S = s->S;
    double *rhs = s->zA, *xs = s->zB;
    k_hash_fill<<<vgrid(s->n), VB, 0, st>>>(s->n, 0, s->lo, s->hi, rhs, sg_red_local(s->red), S + 7);
    SG_CHECK_CUDA(cudaGetLastError());
    sg_count_launch();
    SG_BLK_DISPATCH(blk_init, s, rhs, xs, st);
    if (rc) return rc;
    if ((rc = allreduce(s, S, 2, st))) return rc;
    k_pcg_begin<<<1, 1, 0, st>>>(s->ctrl, 0.0);     // target 0: the iteration never stops by itself
    SG_CHECK_CUDA(cudaGetLastError());
    if ((rc = sg_thermal_linearize(s->op, T_lin, st))) return rc;
    if ((rc = read_scalars(s, 0, 8, st))) return rc;
    double rz = s->S_host[0], alpha_prev = 0.0, beta_prev = 0.0;
    double diag[LANCZOS], off[LANCZOS];
    int m = 0;
    for (int it = 0; it < LANCZOS; ++it) {
        double *Scur = S + 2 * (it & 1), *Snext = S + 2 * ((it + 1) & 1);
        SgHaloWait hw;
        if ((rc = halo_start(s, s->p, &hw, st))) return rc;
        if ((rc = sg_thermal_apply_dot(s->op, T_lin, s->p, s->Ap, s->red, S + 4, nullptr, st, 0, &hw))) return rc;
        if ((rc = allreduce(s, S + 4, 2, st))) return rc;
        SG_BLK_DISPATCH(blk_update_xr, s, xs, Scur, Snext, st);
        if (rc) return rc;
        if ((rc = allreduce(s, Snext, 2, st))) return rc;
        SG_BLK_DISPATCH(blk_update_p, s, Scur, Snext, st);
        if (rc) return rc;
        if ((rc = read_scalars(s, 0, 8, st))) return rc;
        const double pAp = s->S_host[4] + s->S_host[5], rz_next = s->S_host[2 * ((it + 1) & 1)];
        if (!(pAp > 0.0) || !(rz > 0.0) || !isfinite(rz_next)) break;
        const double alpha = rz / pAp, beta = rz_next / rz;
        diag[m] = 1.0 / alpha + (m ? beta_prev / alpha_prev : 0.0);
        if (m) off[m - 1] = sqrt(beta_prev) / alpha_prev;
        ++m;
        alpha_prev = alpha;
        beta_prev = beta;
        rz = rz_next;
        if (!(rz_next > 0.0)) break;
    }
    if (m < 3) {
        sg_set_error("estimate_spectrum: the Lanczos iteration broke down after %d steps", m);
        return SG_E_NOCONV;
    }
    double glo = diag[0], ghi = diag[0];   // Gershgorin bracket
    for (int i = 0; i < m; ++i) {
        const double rad = (i ? fabs(off[i - 1]) : 0.0) + (i + 1 < m ? fabs(off[i]) : 0.0);
        glo = fmin(glo, diag[i] - rad);
        ghi = fmax(ghi, diag[i] + rad);
    }
    double lo = glo, hi = ghi;             // largest eigenvalue: smallest x with count(x) == m
    for (int i = 0; i < 80; ++i) {
        const double mid = 0.5 * (lo + hi);
        if (sturm_count(diag, off, m, mid) >= m) hi = mid; else lo = mid;
    }
    const double ritz_max = hi;
    lo = glo;
    hi = ghi;                              // smallest eigenvalue: largest x with count(x) == 0
    for (int i = 0; i < 80; ++i) {
        const double mid = 0.5 * (lo + hi);
        if (sturm_count(diag, off, m, mid) >= 1) hi = mid; else lo = mid;
    }
    const double ritz_min = lo;
    if (!(ritz_max > 0.0) || !isfinite(ritz_max)) {
        sg_set_error("estimate_spectrum: bad Ritz value %g", ritz_max);
        return SG_E_NOCONV;
    }
    s->cheb_hi = 1.05 * ritz_max;
    s->cheb_lo = fmin(0.9 * fmax(ritz_min, 0.0) + 0.0, s->cheb_hi / 4.0);
    if (!(s->cheb_lo > 0.0)) s->cheb_lo = s->cheb_hi / 30.0;
    return SG_OK;
}

// Runs a solver entry point on the solver's own stream, ordered after everything the caller has enqueued on ITS
// stream so far, and makes the caller's stream wait for the solver's work on exit (also on error paths).
struct StreamScope {
    sg_thermal_solver *s;
    cudaStream_t caller;
    StreamScope(sg_thermal_solver *sv, cudaStream_t c) : s(sv), caller(c) {
        cudaEventRecord(s->ev_in, caller);
        cudaStreamWaitEvent(s->own, s->ev_in, 0);
    }
    ~StreamScope() {
        cudaEventRecord(s->ev_out, s->own);
        cudaStreamWaitEvent(caller, s->ev_out, 0);
    }
};

constexpr int CHEB_MAX = 8;
constexpr int CHEB_BATCH = 4;

// PCG with the Chebyshev polynomial preconditioner z = q_k(M^-1 J) M^-1 r.  One outer iteration costs k + 1
// operator applications but only ONE set of CG vector updates, which is what the plain iteration spends most
// of its time on.  Same device-side convergence logic as pcg_run.
int pcg_run_cheb(sg_thermal_solver *s, const double *T_lin, const double *b, double *x, const PcgTol &tp, int32_t max_it,
                 int32_t *iters, double *rel_res, cudaStream_t st) {
    int rc;
    const int k = s->cheb_degree;
    if (s->cheb_hi <= 0.0 && (rc = estimate_spectrum(s, T_lin, st))) return rc;
    const double theta = 0.5 * (s->cheb_hi + s->cheb_lo), delta = 0.5 * (s->cheb_hi - s->cheb_lo), sigma1 = theta / delta;
    double ca[CHEB_MAX], cb[CHEB_MAX], rho = 1.0 / sigma1;
    for (int j = 0; j < k; ++j) {
        const double rho_new = 1.0 / (2.0 * sigma1 - rho);
        ca[j] = rho_new * rho;
        cb[j] = 2.0 * rho_new / delta;
        rho = rho_new;
    }
    const double inv_theta = 1.0 / theta;
    double *S = s->S;
    const unsigned g = vgrid(s->n);
    SG_BLK_DISPATCH(blk_init_cheb, s, b, x, inv_theta, st);
    if (rc) return rc;
    if ((rc = allreduce(s, S + 1, 1, st))) return rc;
    if ((rc = read_scalars(s, 1, 1, st))) return rc;
    const double rr0 = s->S_host[1];
    s->last_rhs_norm = sqrt(rr0 > 0.0 ? rr0 : 0.0);
    if (!(rr0 >= 0.0) || !isfinite(rr0)) {
        sg_set_error("sg_pcg_solve: right-hand side is not finite");
        return SG_E_NOCONV;
    }
    const double tol2 = tp.tol2(rr0);
    int it = 0, done = rr0 > tol2 ? 0 : 1;
    double rr = rr0;
    const int *skip = &s->ctrl->done;
    // z = q_k(M^-1 J) M^-1 r from z1 (in zA); the result's address depends on the parity of k
    double *zfinal = (k % 2 == 0) ? s->zA : s->zB;
    // Partitioned mesh: the ghost rows of every input vector are put straight into the neighbours' vectors (one small
    // kernel) and the consuming kernel waits for them only before its two boundary strips (peer.cu).
    auto precondition = [&](double *Srz) -> int {
        double *zin = s->zA, *zother = s->zB;
        for (int j = 0; j < k; ++j) {
            // z_{j+1} overwrites z_{j-1} (which lives in `zother`; for j = 0 there is no z_0 and zother is free)
            SgHaloWait hw;
            int r2 = halo_start(s, zin, &hw, st);
            if (r2) return r2;
            SgChebStep cs{zin, s->r, j == 0 ? nullptr : zother, zother, ca[j], cb[j], j == k - 1 ? 1 : 0, &hw};
            if ((r2 = sg_thermal_cheb_step(s->op, cs, s->red, Srz, skip, st))) return r2;
            double *t = zin;
            zin = zother;
            zother = t;
        }
        return allreduce(s, Srz, 1, st);
    };
    if (!done) {
        if ((rc = sg_thermal_linearize(s->op, T_lin, st))) return rc;
        if ((rc = precondition(S))) return rc;
        k_axpy_p<<<g, VB, 0, st>>>(s->hi - s->lo, sg_next_sweep_dir(), zfinal + s->lo, s->p + s->lo, S, S, 1, s->ctrl);
        SG_CHECK_CUDA(cudaGetLastError());
        sg_count_launch();
    }
    while (!done && it < max_it) {
        const int nb = (max_it - it < CHEB_BATCH) ? max_it - it : CHEB_BATCH;
        for (int q = 0; q < nb; ++q, ++it) {
            double *Scur = S + 2 * (it & 1), *Snext = S + 2 * ((it + 1) & 1);
            SgHaloWait hw;
            if ((rc = halo_start(s, s->p, &hw, st))) return rc;
            if ((rc = sg_thermal_apply_dot(s->op, T_lin, s->p, s->Ap, s->red, S + 4, skip, st, 0, &hw))) return rc;
            if ((rc = allreduce(s, S + 4, 2, st))) return rc;
            SG_BLK_DISPATCH(blk_update_xr_cheb, s, x, inv_theta, Scur, Snext, st);
            if (rc) return rc;
            if ((rc = allreduce(s, Snext + 1, 1, st))) return rc;
            k_pcg_check<<<1, 1, 0, st>>>(s->ctrl, Snext, tol2, it);
            SG_CHECK_CUDA(cudaGetLastError());
            if ((rc = precondition(Snext))) return rc;
            k_axpy_p<<<g, VB, 0, st>>>(s->hi - s->lo, sg_next_sweep_dir(), zfinal + s->lo, s->p + s->lo, Scur, Snext, 0, s->ctrl);
            SG_CHECK_CUDA(cudaGetLastError());
            sg_count_launch(2);
        }
        if ((rc = read_scalars(s, 0, 8, st))) return rc;
        done = s->ctrl_host->done;
        rr = done ? s->ctrl_host->rr : s->S_host[2 * (it & 1) + 1];
        if (done) it = s->ctrl_host->iters;
        if (!done && !(s->S_host[2 * (it & 1)] > 0.0)) {   // r.z <= 0: the polynomial is not positive on the spectrum
            sg_set_error("Chebyshev preconditioner is not positive definite (r.z = %g): eigenvalue bound %g too small",
                         s->S_host[2 * (it & 1)], s->cheb_hi);
            return SG_E_NOCONV;
        }
    }
    if (iters) *iters = it;
    if (rel_res) *rel_res = rr0 > 0.0 ? sqrt(rr / rr0) : 0.0;
    if (done == 2 || !isfinite(rr)) {
        sg_set_error("sg_pcg_solve (Chebyshev): residual became non-finite at iteration %d", it);
        return SG_E_NOCONV;
    }
    if (!done) {
        sg_set_error("sg_pcg_solve (Chebyshev): no convergence in %d iterations (relative residual %.3e)", it,
                     rr0 > 0 ? sqrt(rr / rr0) : 0.0);
        return SG_E_NOCONV;
    }
    return SG_OK;
}

}  // namespace

extern "C" {

int sg_halo_plan_create(sg_ctx *ctx, int32_t n_segments, const sg_halo_segment *segments, sg_halo_plan **out) {
    SG_REQUIRE(ctx && out && n_segments >= 0 && (n_segments == 0 || segments), "sg_halo_plan_create: bad argument");
    for (int i = 0; i < n_segments; ++i) {
        SG_REQUIRE(segments[i].peer >= 0 && segments[i].peer < ctx->nranks && segments[i].peer != ctx->rank,
                   "sg_halo_plan_create: segment %d has bad peer %d", i, segments[i].peer);
        SG_REQUIRE(segments[i].send_count >= 0 && segments[i].recv_count >= 0, "sg_halo_plan_create: negative count");
    }
    sg_halo_plan *p = new sg_halo_plan();
    p->ctx = ctx;
    p->n = n_segments;
    p->seg = new sg_halo_segment[n_segments > 0 ? n_segments : 1];
    for (int i = 0; i < n_segments; ++i) p->seg[i] = segments[i];
    p->peer = nullptr;
    *out = p;
    return SG_OK;
}

int sg_halo_plan_destroy(sg_halo_plan *plan) {
    if (!plan) return SG_OK;
    if (plan->peer) sg_peer_destroy(plan->peer);
    delete[] plan->seg;
    delete plan;
    return SG_OK;
}

// Peer-memory path: returns 1 and fills handle64 when this plan can use it (at most one neighbour below and one
// above this rank), 0 otherwise.  The caller gathers the 64-byte handles of all ranks and calls sg_halo_peer_open.
int sg_halo_peer_alloc(sg_halo_plan *plan, void *handle64, int64_t workspace_doubles) {
    SG_REQUIRE(plan && handle64 && workspace_doubles >= 0, "sg_halo_peer_alloc: bad argument");
    if (plan->ctx->nranks < 2 || plan->n > 2) return 0;
    int below = 0, above = 0;
    int64_t mx = 0;
    for (int i = 0; i < plan->n; ++i) {
        (plan->seg[i].peer < plan->ctx->rank ? below : above)++;
        if (plan->seg[i].send_count > mx) mx = plan->seg[i].send_count;
        if (plan->seg[i].recv_count > mx) mx = plan->seg[i].recv_count;
    }
    if (below > 1 || above > 1) return 0;
    if (plan->peer) return 0;
    const int rc = sg_peer_create(plan->ctx, (size_t)mx, (size_t)workspace_doubles, &plan->peer, handle64);
    if (rc != SG_OK) {
        plan->peer = nullptr;
        return rc;
    }
    return 1;
}

// handles == NULL: some rank could not allocate; drop the peer path everywhere (NCCL stays in use).
int sg_halo_peer_open(sg_halo_plan *plan, const void *handles, const int64_t *layout3) {
    SG_REQUIRE(plan, "sg_halo_peer_open: NULL plan");
    if (!plan->peer) return SG_OK;
    if (!handles || !layout3) {
        sg_peer_destroy(plan->peer);
        plan->peer = nullptr;
        return SG_OK;
    }
    const int rc = sg_peer_open(plan->peer, handles, layout3);
    if (rc != SG_OK) {
        sg_peer_destroy(plan->peer);
        plan->peer = nullptr;
    }
    return rc;
}

int sg_halo_uses_peer_memory(const sg_halo_plan *plan) { return plan && sg_peer_ready(plan->peer) ? 1 : 0; }

double *sg_halo_peer_workspace(const sg_halo_plan *plan) { return plan ? sg_peer_workspace(plan->peer) : nullptr; }

int sg_halo_forward(sg_halo_plan *plan, double *vec, int32_t bs, void *stream) {
    SG_REQUIRE(plan && vec && bs >= 1, "sg_halo_forward: bad argument");
    if (plan->n == 0) return SG_OK;
    if (bs == 1 && sg_peer_ready(plan->peer)) return sg_peer_halo_forward(plan->peer, plan->n, plan->seg, vec, (cudaStream_t)stream);
    const SgNccl *n = sg_nccl();
    if (!n || !plan->ctx->comm) {
        sg_set_error("sg_halo_forward: context has no NCCL communicator");
        return SG_E_NCCL;
    }
    cudaStream_t st = (cudaStream_t)stream;
    SG_CHECK_NCCL(n->GroupStart());
    for (int i = 0; i < plan->n; ++i) {
        const sg_halo_segment &g = plan->seg[i];
        if (g.send_count > 0)
            SG_CHECK_NCCL(n->Send(vec + g.send_offset * bs, (size_t)(g.send_count * bs), ncclDouble, g.peer, plan->ctx->comm, st));
        if (g.recv_count > 0)
            SG_CHECK_NCCL(n->Recv(vec + g.recv_offset * bs, (size_t)(g.recv_count * bs), ncclDouble, g.peer, plan->ctx->comm, st));
    }
    SG_CHECK_NCCL(n->GroupEnd());
    return SG_OK;
}

int64_t sg_thermal_solver_workspace_doubles(const sg_thermal_op *op) {
    if (!op) return -1;
    SgOpInfo oi;
    sg_op_info(op, &oi);
    return 8 * oi.n_dofs;
}

int sg_thermal_solver_create(sg_thermal_op *op, double *workspace, sg_halo_plan *halo, sg_thermal_solver **out) {
    SG_REQUIRE(op && workspace && out, "sg_thermal_solver_create: NULL argument");
    sg_thermal_solver *s = new sg_thermal_solver();
    s->op = op;
    SgOpInfo oi;
    sg_op_info(op, &oi);
    s->ctx = oi.ctx;
    s->halo = halo;
    s->n = (long)oi.n_dofs;
    s->lo = (long)oi.own_lo;
    s->hi = (long)oi.own_hi;
    s->blk_nld = oi.family == 1 ? oi.n_ld : 0;
    s->n_cells = (long)oi.n_cells;
    s->clo = (long)oi.cell_lo;
    s->chi = (long)oi.cell_hi;
    s->detJ = oi.detJ;
    for (int i = 0; i < oi.n_ld * oi.n_ld; ++i) s->mass_inv[i] = oi.mass_inv[i];
    SG_REQUIRE(s->ctx->nranks == 1 || halo, "sg_thermal_solver_create: a multi-GPU context needs a halo plan");
    double *w = workspace;
    s->b = w;
    s->dx = w + s->n;
    s->r = w + 2 * s->n;
    s->p = w + 3 * s->n;
    s->Ap = w + 4 * s->n;
    s->dinv = w + 5 * s->n;
    s->zA = w + 6 * s->n;
    s->zB = w + 7 * s->n;
    s->cheb_degree = 0;
    s->cheb_lo = s->cheb_hi = 0.0;
    s->cheb_failed = 0;
    s->red.partials = nullptr;
    s->red.counter = nullptr;
    s->red.peer = nullptr;
    s->red.ar_ptr = nullptr;
    s->red.ar_count = 0;
    s->S = nullptr;
    s->S_host = nullptr;
    s->ctrl = s->ctrl_host = nullptr;
    s->own = nullptr;
    s->side = nullptr;
    s->ev_produced = s->ev_put = nullptr;
    s->put_pending = 0;
    s->ev_in = s->ev_out = nullptr;
    s->batch_graph = nullptr;
    s->graph_T = s->graph_x = nullptr;
    {
        const char *ng = getenv("SG_NO_GRAPHS");
        s->use_graphs = !(ng && ng[0] == '1');
    }
    cudaError_t e = cudaMalloc(&s->red.partials, sizeof(double) * (SG_MAX_BLOCKS * 2 + 2));
    if (e == cudaSuccess) e = cudaMalloc(&s->red.counter, sizeof(unsigned));
    if (e == cudaSuccess) e = cudaMemset(s->red.counter, 0, sizeof(unsigned));
    if (e == cudaSuccess) e = cudaMalloc(&s->S, sizeof(double) * 8);
    if (e == cudaSuccess) e = cudaMemset(s->S, 0, sizeof(double) * 8);
    if (e == cudaSuccess) e = cudaHostAlloc(&s->S_host, sizeof(double) * 8, cudaHostAllocMapped);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&s->own, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->ev_in, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->ev_out, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaMalloc(&s->ctrl, sizeof(PcgCtrl));
    if (e == cudaSuccess) e = cudaMemset(s->ctrl, 0, sizeof(PcgCtrl));
    if (e == cudaSuccess) e = cudaHostAlloc(&s->ctrl_host, sizeof(PcgCtrl), cudaHostAllocMapped);
    if (e != cudaSuccess) {
        sg_set_error("sg_thermal_solver_create: %s", cudaGetErrorString(e));
        sg_thermal_solver_destroy(s);
        return SG_E_CUDA;
    }
    // peer-memory path: all-reduces run inside the reducing kernels; the direct halo put needs the workspace to be the
    // IPC-visible one (sg_halo_peer_workspace) — any other workspace still works through the mailboxes
    if (halo && sg_peer_ready(halo->peer)) {
        s->red.peer = sg_peer_red_dev(halo->peer);
        const char *ns = getenv("SG_NO_SIDE_STREAM");
        if (!(ns && ns[0] == '1')) {
            // highest priority: when a producer kernel retires, the pending put's blocks are dispatched before the blocks of
            // the (persistent, SM-filling) consumer kernel launched right behind it on the solver stream
            int prio_lo = 0, prio_hi = 0;
            cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
            cudaError_t e2 = cudaStreamCreateWithPriority(&s->side, cudaStreamNonBlocking, prio_hi);
            if (e2 == cudaSuccess) e2 = cudaEventCreateWithFlags(&s->ev_produced, cudaEventDisableTiming);
            if (e2 == cudaSuccess) e2 = cudaEventCreateWithFlags(&s->ev_put, cudaEventDisableTiming);
            if (e2 != cudaSuccess) {
                sg_set_error("sg_thermal_solver_create: %s", cudaGetErrorString(e2));
                sg_thermal_solver_destroy(s);
                return SG_E_CUDA;
            }
        }
    }
    *out = s;
    return SG_OK;
}

int sg_thermal_solver_destroy(sg_thermal_solver *s) {
    if (!s) return SG_OK;
    if (s->red.partials) cudaFree(s->red.partials);
    if (s->red.counter) cudaFree(s->red.counter);
    if (s->S) cudaFree(s->S);
    if (s->S_host) cudaFreeHost(s->S_host);
    if (s->batch_graph) cudaGraphExecDestroy(s->batch_graph);
    if (s->own) {
        cudaStreamSynchronize(s->own);
        cudaStreamDestroy(s->own);
    }
    if (s->side) {
        cudaStreamSynchronize(s->side);
        cudaStreamDestroy(s->side);
    }
    if (s->ev_produced) cudaEventDestroy(s->ev_produced);
    if (s->ev_put) cudaEventDestroy(s->ev_put);
    if (s->ev_in) cudaEventDestroy(s->ev_in);
    if (s->ev_out) cudaEventDestroy(s->ev_out);
    if (s->ctrl) cudaFree(s->ctrl);
    if (s->ctrl_host) cudaFreeHost(s->ctrl_host);
    delete s;
    return SG_OK;
}

static int pcg_run(sg_thermal_solver *s, const double *T_lin, const double *b, double *x, const PcgTol &tp,
                   int32_t max_it, int32_t *iters, double *rel_res, cudaStream_t st);

int sg_pcg_solve(sg_thermal_solver *s, const double *T_lin, const double *b, double *x, double rtol, double atol,
                 int32_t max_it, int32_t *iters, double *rel_res, void *stream) {
    SG_REQUIRE(s && T_lin && b && x, "sg_pcg_solve: NULL argument");
    StreamScope scope(s, (cudaStream_t)stream);
    return pcg_run(s, T_lin, b, x, PcgTol{rtol, atol, false, 0.0, 0.0, 0.0, 0.0}, max_it, iters, rel_res, s->own);
}

int sg_thermal_solver_set_chebyshev(sg_thermal_solver *s, int32_t degree, double lo, double hi) {
    SG_REQUIRE(s && degree >= 0 && degree <= CHEB_MAX, "sg_thermal_solver_set_chebyshev: degree must be 0..%d", CHEB_MAX);
    if (degree > 0 && !(s->blk_nld && sg_thermal_has_cheb(s->op))) return 0;   // not available: plain PCG stays in use
    s->cheb_degree = degree;
    s->cheb_failed = 0;
    if (hi > 0.0) {
        SG_REQUIRE(lo > 0.0 && lo < hi, "sg_thermal_solver_set_chebyshev: need 0 < lo < hi");
        s->cheb_lo = lo;
        s->cheb_hi = hi;
    } else {
        s->cheb_lo = s->cheb_hi = 0.0;
    }
    return degree > 0 ? 1 : 0;
}

int sg_thermal_solver_get_chebyshev(const sg_thermal_solver *s, int32_t *degree, double *lo, double *hi) {
    SG_REQUIRE(s, "sg_thermal_solver_get_chebyshev: NULL solver");
    if (degree) *degree = s->cheb_failed ? 0 : s->cheb_degree;
    if (lo) *lo = s->cheb_lo;
    if (hi) *hi = s->cheb_hi;
    return SG_OK;
}

static int pcg_run(sg_thermal_solver *s, const double *T_lin, const double *b, double *x, const PcgTol &tp,
                   int32_t max_it, int32_t *iters, double *rel_res, cudaStream_t st) {
    if (s->cheb_degree > 0 && !s->cheb_failed) {
        int rc_c = pcg_run_cheb(s, T_lin, b, x, tp, max_it, iters, rel_res, st);
        for (int retry = 0; rc_c == SG_E_NOCONV && retry < 2 && s->cheb_hi > 0.0; ++retry) {
            s->cheb_hi *= 1.3;   // the upper bound was too small (r.z <= 0 or no convergence): widen the interval
            rc_c = pcg_run_cheb(s, T_lin, b, x, tp, max_it, iters, rel_res, st);
        }
        if (rc_c != SG_E_NOCONV) return rc_c;
        s->cheb_failed = 1;   // still failing: fall back to the block-Jacobi iteration for good
    }
    const long n = s->n, lo = s->lo, hi = s->hi;
    const unsigned g = vgrid(n);
    double *S = s->S;
    int rc;
    if (s->ctx->nranks == 1 && !sg_thermal_profiling(s->op)) {
        // small CG problem in gather form: the whole solve is ONE persistent cooperative kernel (stencil.cu k_cg_persistent);
        // tiny DG problem: ONE block (thermal.cu k_dg_pcg_small)
        if (s->blk_nld)
            rc = sg_thermal_pcg_small_dg(s->op, T_lin, b, x, tp, max_it, S + 1, &s->ctrl->done, &s->ctrl->iters, &s->ctrl->rr, st);
        else
            rc = sg_thermal_pcg_persistent(s->op, T_lin, b, s->dinv, x, s->r, tp, max_it, S + 1, &s->ctrl->done, &s->ctrl->iters, &s->ctrl->rr, st);
        if (rc < 0) return rc;
        if (rc == 1) {
            if ((rc = read_scalars(s, 0, 8, st))) return rc;
            const double rr0 = s->S_host[1], rr = s->ctrl_host->rr;
            const int done = s->ctrl_host->done;
            s->last_rhs_norm = sqrt(rr0 > 0.0 ? rr0 : 0.0);
            if (iters) *iters = s->ctrl_host->iters;
            if (rel_res) *rel_res = rr0 > 0.0 ? sqrt(rr / rr0) : 0.0;
            if (done == 2 || !isfinite(rr) || !isfinite(rr0)) {
                sg_set_error("sg_pcg_solve (persistent): residual became non-finite at iteration %d", s->ctrl_host->iters);
                return SG_E_NOCONV;
            }
            if (!done) {
                sg_set_error("sg_pcg_solve (persistent): no convergence in %d iterations (relative residual %.3e)", s->ctrl_host->iters,
                             rr0 > 0 ? sqrt(rr / rr0) : 0.0);
                return SG_E_NOCONV;
            }
            return SG_OK;
        }
    }
    if (s->blk_nld) {
        SG_BLK_DISPATCH(blk_init, s, b, x, st);
        if (rc) return rc;
    } else {
        k_pcg_init<<<g, VB, 0, st>>>(lo, hi, lo, hi, b, s->dinv, x, s->r, s->p, s->red, S, s->ctrl);
        SG_CHECK_CUDA(cudaGetLastError());
        sg_count_launch();
    }
    if ((rc = allreduce(s, S, 2, st))) return rc;
    if ((rc = read_scalars(s, 0, 2, st))) return rc;
    const double rr0 = s->S_host[1];
    s->last_rhs_norm = sqrt(rr0 > 0.0 ? rr0 : 0.0);
    const double tol2 = tp.tol2(rr0);
    if (!(rr0 >= 0.0) || !isfinite(rr0)) {
        sg_set_error("sg_pcg_solve: right-hand side is not finite");
        return SG_E_NOCONV;
    }
    int it = 0, done = rr0 > tol2 ? 0 : 1;
    double rr = rr0;
    if (!done) {
        if ((rc = sg_thermal_linearize(s->op, T_lin, st))) return rc;
        k_pcg_begin<<<1, 1, 0, st>>>(s->ctrl, tol2);
        SG_CHECK_CUDA(cudaGetLastError());
        // CG spaces scatter into Ap: zero it once, afterwards every p update leaves it zeroed
        if (!s->blk_nld) SG_CHECK_CUDA(cudaMemsetAsync(s->Ap, 0, sizeof(double) * (size_t)n, st));
    }
    // One PCG iteration = 3 launches.  Iterations are enqueued BATCH at a time without host synchronisation: the
    // kernels test convergence themselves and the launches after the converged iteration return immediately.
    auto enqueue_iteration = [&](int parity) -> int {
        double *Scur = S + 2 * parity, *Snext = S + 2 * (1 - parity);
        int r2;
        SgHaloWait hw;
        if ((r2 = halo_start(s, s->p, &hw, st))) return r2;
        if ((r2 = sg_thermal_apply_dot(s->op, T_lin, s->p, s->Ap, s->red, S + 4, &s->ctrl->done, st, !s->blk_nld, &hw))) return r2;
        if ((r2 = allreduce(s, S + 4, 2, st))) return r2;
        if (s->blk_nld) {
            int rc;
            SG_BLK_DISPATCH(blk_update_xr, s, x, Scur, Snext, st);
            if (rc) return rc;
        } else {
            k_update_xr<<<g, VB, 0, st>>>(lo, hi, lo, hi, s->p, s->Ap, s->dinv, x, s->r, s->red, Scur, S + 4, Snext, s->ctrl);
            SG_CHECK_CUDA(cudaGetLastError());
            sg_count_launch();
        }
        if ((r2 = allreduce(s, Snext, 2, st))) return r2;
        if (s->blk_nld) {
            int rc;
            SG_BLK_DISPATCH(blk_update_p, s, Scur, Snext, st);
            if (rc) return rc;
        } else {
            k_update_p<<<g, VB, 0, st>>>(n, lo, hi, s->r, s->dinv, s->p, s->Ap, Scur, Snext, s->ctrl);
            SG_CHECK_CUDA(cudaGetLastError());
            sg_count_launch();
        }
        return SG_OK;
    };
    // Single GPU: a full batch is captured once into a CUDA graph (every argument that changes between solves lives
    // in device memory) and replayed — the kernels of small meshes take a few microseconds, less than their launches.
    const bool graphs = s->use_graphs && s->ctx->nranks == 1 && !sg_thermal_profiling(s->op) && (BATCH % 2 == 0);
    while (!done && it < max_it) {
        const int nb = (max_it - it < BATCH) ? max_it - it : BATCH;
        if (graphs && nb == BATCH) {
            if (!s->batch_graph || s->graph_T != T_lin || s->graph_x != x) {
                if (s->batch_graph) cudaGraphExecDestroy(s->batch_graph);
                s->batch_graph = nullptr;
                cudaGraph_t graph = nullptr;
                SG_CHECK_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
                int rcap = SG_OK;
                for (int k = 0; k < BATCH && rcap == SG_OK; ++k) rcap = enqueue_iteration(k & 1);
                const cudaError_t ec = cudaStreamEndCapture(st, &graph);
                sg_count_launch(-3 * BATCH);   // the captured launches did not execute
                if (rcap != SG_OK || ec != cudaSuccess || !graph) {
                    if (graph) cudaGraphDestroy(graph);
                    cudaGetLastError();
                    s->use_graphs = 0;      // capture not possible here: plain launches from now on
                    if (rcap != SG_OK) return rcap;
                    continue;
                }
                const cudaError_t ei = cudaGraphInstantiate(&s->batch_graph, graph, 0);
                cudaGraphDestroy(graph);
                if (ei != cudaSuccess) {
                    cudaGetLastError();
                    s->batch_graph = nullptr;
                    s->use_graphs = 0;
                    continue;
                }
                s->graph_T = T_lin;
                s->graph_x = x;
            }
            SG_CHECK_CUDA(cudaGraphLaunch(s->batch_graph, st));
            sg_count_launch(3 * BATCH);
            it += BATCH;
        } else {
            for (int k = 0; k < nb; ++k, ++it)
                if ((rc = enqueue_iteration(it & 1))) return rc;
        }
        if ((rc = read_scalars(s, 0, 8, st))) return rc;
        done = s->ctrl_host->done;
        rr = done ? s->ctrl_host->rr : s->S_host[2 * (it & 1) + 1];
        if (done) it = s->ctrl_host->iters;
    }
    if (iters) *iters = it;
    if (rel_res) *rel_res = rr0 > 0.0 ? sqrt(rr / rr0) : 0.0;
    if (done == 2 || !isfinite(rr)) {
        sg_set_error("sg_pcg_solve: residual became non-finite at iteration %d (operator not SPD?)", it);
        return SG_E_NOCONV;
    }
    if (!done) {
        sg_set_error("sg_pcg_solve: no convergence in %d iterations (relative residual %.3e)", it,
                     rr0 > 0 ? sqrt(rr / rr0) : 0.0);
        return SG_E_NOCONV;
    }
    return SG_OK;
}

int sg_thermal_timestep(sg_thermal_solver *s, double *T, const double *T_prev, const sg_newton_opts *o,
                        sg_newton_stats *stats, void *stream) {
    SG_REQUIRE(s && T && T_prev && o, "sg_thermal_timestep: NULL argument");
    StreamScope scope(s, (cudaStream_t)stream);
    cudaStream_t st = s->own;
    const long n = s->n;
    const unsigned g = vgrid(n);
    int rc, lin_total = 0;
    double r0 = 0.0, r = 0.0, lin_res = 0.0, lin_target = 0.0, F_prev = 0.0;
    int it = 0, converged = 0;
    for (it = 1; it <= o->newton_max_it; ++it) {
        if (s->halo && (rc = sg_halo_forward(s->halo, T, 1, st))) return rc;
        if ((rc = sg_thermal_residual(s->op, T, T_prev, s->b, st))) return rc;          // b = F(T)
        if (!s->blk_nld) {                                                              // point Jacobi needs diag J(T)
            if ((rc = sg_thermal_jac_diag(s->op, T, s->dinv, st))) return rc;
            k_invert<<<g, VB, 0, st>>>(n, s->dinv);
            SG_CHECK_CUDA(cudaGetLastError());
            sg_count_launch();
        }
        // Inexact Newton (Eisenstat-Walker, choice 2): the boundary radiation makes the linear model of
        // iteration k wrong at the 1e-3..1e-4 level, so solve k only needs |r| <= eta_k |F(T_k)| with
        // eta_1 = forcing_eta, eta_k = min(eta_1, 0.9 (|F_k|/|F_{k-1}|)^2), never below half the final target
        // lin_rtol*|F(T_0)| (floor lin_atol).  Once |F(T_k)| meets that target the solve returns dx = 0 and the
        // incremental criterion below holds with |dx| = 0.  forcing_eta = 0: every solve runs to the target.
        int lin_it = 0;
        PcgTol tp;
        if (o->forcing_eta > 0.0)
            tp = PcgTol{o->lin_rtol, o->lin_atol, true, o->forcing_eta, 0.9, F_prev, lin_target};
        else
            tp = PcgTol{it == 1 ? o->lin_rtol : 0.0, it == 1 ? o->lin_atol : lin_target, false, 0.0, 0.0, 0.0, 0.0};
        rc = pcg_run(s, T, s->b, s->dx, tp, o->lin_max_it, &lin_it, &lin_res, st);
        if (it == 1) lin_target = fmax(o->lin_atol, o->lin_rtol * s->last_rhs_norm);
        F_prev = s->last_rhs_norm;
        lin_total += lin_it;
        if (rc) return rc;
        k_newton_update<<<g, VB, 0, st>>>(s->lo, s->hi, s->lo, s->hi, T, s->dx, s->red, s->S + 6);  // T <- T - dx (owned rows)
        SG_CHECK_CUDA(cudaGetLastError());
        sg_count_launch();
        if ((rc = allreduce(s, s->S + 6, 1, st))) return rc;
        if ((rc = read_scalars(s, 6, 1, st))) return rc;
        r = sqrt(s->S_host[6]);
        // dolfinx NewtonSolver, convergence_criterion = "incremental": iteration 1 only records r0
        if (it == 1) {
            r0 = r;
            if (r0 == 0.0) {
                converged = 1;
                break;
            }
        } else if (r / r0 < o->newton_rtol || r < o->newton_atol) {
            converged = 1;
            break;
        }
    }
    if (s->halo && (rc = sg_halo_forward(s->halo, T, 1, st))) return rc;
    if (stats) {
        stats->newton_its = it > o->newton_max_it ? o->newton_max_it : it;
        stats->lin_its = lin_total;
        stats->converged = converged;
        stats->dx_norm_first = r0;
        stats->dx_norm_last = r;
        stats->lin_rel_res_last = lin_res;
    }
    if (!converged) {
        sg_set_error("sg_thermal_timestep: Newton did not converge in %d iterations (|dx| = %.3e, |dx0| = %.3e)",
                     o->newton_max_it, r, r0);
        return SG_E_NOCONV;
    }
    return SG_OK;
}

}  // extern "C"
