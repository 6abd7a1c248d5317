// Shared internals of libsurroglas_b200: error plumbing, context, PTX helpers
// (mbarrier + 1-D TMA bulk copies), warp/block reductions.
#pragma once

#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/surroglas_b200.h"

struct ncclComm;

struct sg_ctx {
    int device;
    int rank, nranks;
    int sm_count;
    ncclComm *comm;  // nullptr when nranks == 1
};

void sg_set_error(const char *fmt, ...);

// number of kernels this library has launched in this process (bench.py reports it as gpu_launches)
void sg_count_launch(int n = 1);

// internal accessors of sg_thermal_op (defined in thermal.cu) for the solver in pcg.cu
struct SgOpInfo {
    sg_ctx *ctx;
    int dim, family, n_ld;
    int64_t n_dofs, own_lo, own_hi, n_cells, cell_lo, cell_hi;
    const double *detJ;        // device, [n_cells]
    const double *mass_inv;    // host, [n_ld * n_ld]: inverse of the reference mass matrix
};
void sg_op_info(const sg_thermal_op *op, SgOpInfo *info);

// NVLink peer-memory all-reduce performed INSIDE the reducing kernel (peer.cu owns the memory): the last block of
// sg_grid_reduce writes its <= 4 totals + a sequence tag into slot [rank] of every rank's communication block, waits for
// all tags of this round, and sums the slots in rank order — deterministic and bit-identical on every rank.  The
// sequence number lives in device memory (every rank runs the same reducing kernels in the same order), slots are
// double-buffered by its parity (a rank can be at most one reduction ahead of the slowest one).
constexpr int SG_PEER_MAX_RANKS = 16;
constexpr int SG_PEER_RED_VALS = 4;
struct SgPeerRedDev {
    char *base[SG_PEER_MAX_RANKS];   // every rank's block (base[rank] is local)
    size_t tags_off, vals_off;        // [2 parities][SG_PEER_MAX_RANKS] u64 / [2][SG_PEER_MAX_RANKS][SG_PEER_RED_VALS] f64
    int rank, nranks;
    unsigned long long *seq;          // local device counter of the reductions done so far
    int *err;                         // mapped host memory: set when a wait timed out
};

// Grid-wide deterministic reduction scratch: per-block partials + a completion counter (see sg_grid_reduce).
constexpr int SG_MAX_BLOCKS = 1184;  // 8 * 148 resident blocks; kernels that reduce are grid-stride beyond that
struct SgRed {
    double *partials;   // [SG_MAX_BLOCKS * 2]
    unsigned *counter;  // zero between kernels
    // multi-GPU: when set, the last block all-reduces ar_count doubles at ar_ptr over the ranks after storing its totals
    // (ar_ptr == NULL: the totals just stored).  ar_ptr may cover a value an EARLIER kernel of the stream left next to them.
    const SgPeerRedDev *peer;
    double *ar_ptr;
    int ar_count;
};
inline SgRed sg_red_local(SgRed r) {   // the same scratch without the cross-rank step
    r.peer = nullptr;
    r.ar_ptr = nullptr;
    r.ar_count = 0;
    return r;
}

// Ghost rows written directly into this rank's vector by its slab neighbours (peer.cu sg_peer_put): the consumer kernel
// waits until flag[i] >= seq before it touches cells/rows that read ghost values.  n == 0: nothing to wait for.
struct SgHaloWait {
    const unsigned long long *flag[2];
    unsigned long long seq;
    int *err;      // mapped host memory
    int n;
};

// y = J(T_lin) x fused with the owned-dof dot product: dot2[0] + dot2[1] = sum over owned dofs of x*y
// (cells + exterior facets).  When *skip != 0 (device flag, may be NULL) every kernel returns at once.
// Must precede sg_thermal_apply_dot whenever T_lin changed (refreshes the linearised boundary matrices).
int sg_thermal_linearize(sg_thermal_op *op, const double *T_lin, cudaStream_t st);
// y_is_zero: the caller guarantees y == 0 on entry (CG spaces scatter into y; saves the memset launch).
// wait (partitioned mesh): ghost rows of x are still travelling; the kernels that support it process the cells/rows that
// read no ghost value first and wait inside the kernel, the others are preceded by a one-warp wait kernel.
int sg_thermal_apply_dot(sg_thermal_op *op, const double *T_lin, const double *x, double *y, SgRed red, double *dot2,
                         const int *skip, cudaStream_t st, int y_is_zero = 0, const SgHaloWait *wait = nullptr);

// One fused Chebyshev step of the polynomial preconditioner (DG + class tables only, see dg_cheb_step):
// z_out = z_in + a (z_in - z_prev) + b M^-1 (r - J z_in); z_prev == NULL means 0 (first step); z_out may alias
// z_prev but not z_in; last: *dot_out = r.z_out.
struct SgChebStep {
    const double *z_in, *r, *z_prev;
    double *z_out;
    double a, b;
    int last;
    const SgHaloWait *wait;   // ghost rows of z_in in flight (NULL: none)
};
bool sg_thermal_has_cheb(const sg_thermal_op *op);
bool sg_thermal_profiling(const sg_thermal_op *op);   // event pairs around the kernels are being recorded
int sg_thermal_cheb_step(sg_thermal_op *op, const SgChebStep &cs, SgRed red, double *dot_out, const int *skip, cudaStream_t st);

// Tolerance policy of one PCG solve: the |r|^2 target as a function of |b|^2, which is only known after the first
// reduction.  Returning >= |b|^2 means "nothing to do": zero iterations, x = 0.
struct SgPcgPolicy {
    double rtol, atol;   // plain solve: |r| <= max(rtol |b|, atol).  Inexact Newton: the FINAL target of the time step
    bool forcing;        // inexact Newton (Eisenstat-Walker choice 2)
    double eta1, gamma;  // eta_k = min(eta1, gamma (|F_k| / |F_{k-1}|)^2)
    double F_prev;       // |F_{k-1}| (0 for the first Newton iteration)
    double target;       // final absolute target fixed by the first iteration (0 while unknown)
    __host__ __device__ double tol2(double rr0) const {
        if (!forcing) return fmax(rtol * rtol * rr0, atol * atol);
        const double nb = sqrt(rr0);
        const double tgt = target > 0.0 ? target : fmax(atol, rtol * nb);
        if (target > 0.0 && nb <= tgt) return rr0;   // the nonlinear residual already meets the target: dx = 0
        double eta = eta1;
        if (F_prev > 0.0) eta = fmin(eta1, gamma * (nb / F_prev) * (nb / F_prev));
        const double tol = fmax(eta * nb, 0.5 * tgt);
        return tol * tol;
    }
};

// stencil.cu / thermal.cu: the whole Jacobi-PCG solve of a small CG problem in ONE persistent cooperative kernel (see
// k_cg_persistent).  sg_thermal_pcg_persistent returns 1 when it ran (results in *rr0, ctrl fields, x), 0 when this
// operator / size does not qualify (caller runs the multi-kernel iteration), < 0 on error.
struct SgStencil;
int sg_stencil_pcg(SgStencil *s, int sm_count, const double *b, const double *dinv, double *x, double *work, const SgPcgPolicy &pol, int max_it,
                   double *rr0_out, int *ctrl_done, int *ctrl_iters, double *ctrl_rr, cudaStream_t st);
// thermal.cu: the whole block-Jacobi PCG solve of a tiny DG problem (<= 256 cells, class tables) in ONE block; same
// return convention as sg_thermal_pcg_persistent
int sg_thermal_pcg_small_dg(sg_thermal_op *op, const double *T_lin, const double *b, double *x, const SgPcgPolicy &pol, int max_it,
                            double *rr0_out, int *ctrl_done, int *ctrl_iters, double *ctrl_rr, cudaStream_t st);
int sg_thermal_pcg_persistent(sg_thermal_op *op, const double *T_lin, const double *b, const double *dinv, double *x, double *work,
                              const SgPcgPolicy &pol, int max_it, double *rr0_out, int *ctrl_done, int *ctrl_iters, double *ctrl_rr,
                              cudaStream_t st);

// peer.cu: NVLink peer-memory halo exchange and small all-reduce (replaces NCCL on the solver's data path)
struct SgPeer;
int sg_peer_create(sg_ctx *ctx, size_t mailbox_doubles, size_t workspace_doubles, SgPeer **out, void *handle64);
// handles: nranks x 64 bytes; layout: per rank {vector stride (dofs), ghost offset of the rows received from below, from above}
int sg_peer_open(SgPeer *p, const void *handles, const int64_t *layout3);
int sg_peer_destroy(SgPeer *p);
bool sg_peer_ready(const SgPeer *p);
size_t sg_peer_mailbox_doubles(const SgPeer *p);
double *sg_peer_workspace(const SgPeer *p);          // IPC-visible solver workspace inside the communication block
const SgPeerRedDev *sg_peer_red_dev(const SgPeer *p); // device copy of the all-reduce descriptor (for SgRed::peer)
int sg_peer_check(SgPeer *p);
// mailbox path (any vector, e.g. the caller's temperature): push into the neighbour's mailbox, pull into the ghost rows
int sg_peer_halo_forward(SgPeer *p, int n_seg, const sg_halo_segment *seg, double *vec, cudaStream_t st);
// direct path (vectors inside the IPC workspace): ONE kernel stores the boundary rows straight into the neighbour's ghost
// rows of the same vector and raises its flag; *wait tells the consumer what to wait for.  Returns SG_OK and wait->n = 0
// after falling back to the mailbox path (vector outside the workspace).
bool sg_peer_in_workspace(const SgPeer *p, const double *vec);
int sg_peer_put(SgPeer *p, int n_seg, const sg_halo_segment *seg, double *vec, SgHaloWait *wait, cudaStream_t st);
int sg_peer_wait(SgPeer *p, const SgHaloWait &wait, cudaStream_t st);   // one-warp kernel for consumers without an in-kernel wait
int sg_peer_allreduce(SgPeer *p, double *vals, int count, cudaStream_t st);

// classify.cu: equivalence classes of 64-bit keys.  cls_out[i] = class of keys[i] in [0, *n_cls),
// rep_out[k] = index of one member of class k; both are cudaMalloc'ed here and freed by the caller.
int sg_classify_u64(const uint64_t *keys_dev, int64_t n, int32_t **cls_out, int32_t *n_cls, int32_t **rep_out);

int sg_exclusive_scan_i32(int32_t *data_dev, int64_t n, int64_t *total);

// stencil.cu: row-stencil classes of a CG operator (gather form of the Jacobian apply).  *out stays NULL (and SG_OK is
// returned) when the mesh has no small set of repeating rows.  dot2[0] = x.y over the rows [own_lo, own_hi), dot2[1] = 0.
struct SgStencil;
int sg_stencil_build(sg_ctx *ctx, const int32_t *dofmap, int64_t n_cells, int n_ld, int64_t cell_lo, int64_t cell_hi,
                     const uint16_t *cls16, const double *tab, int S, int64_t n_rows, SgStencil **out);
int sg_stencil_apply(const SgStencil *s, const double *x, double *y, int64_t own_lo, int64_t own_hi, SgRed red, double *dot2,
                     const int *skip, cudaStream_t st, const SgHaloWait *wait = nullptr);
int sg_stencil_apply_cells(const SgStencil *s, const double *x, double *y, int subtract, SgRed red, double *dot2, cudaStream_t st);
int sg_stencil_attach_boundary(SgStencil *s, const int32_t *dofmap, int64_t n_cells, int64_t cell_lo, int64_t cell_hi, int64_t n_bf,
                               const int32_t *bf_cell, const int32_t *bf_facet, int nfd, const int *facet_dofs, const double *bmat,
                               int *attached);
int sg_stencil_refresh_boundary(const SgStencil *s, cudaStream_t st);
void sg_stencil_destroy(SgStencil *s);
void sg_stencil_info(const SgStencil *s, int32_t *n_classes, int32_t *n_entries, int32_t *max_nnz);

#define SG_CHECK_CUDA(expr)                                                                  \
    do {                                                                                     \
        cudaError_t _e = (expr);                                                             \
        if (_e != cudaSuccess) {                                                             \
            sg_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return SG_E_CUDA;                                                                \
        }                                                                                    \
    } while (0)

#define SG_REQUIRE(cond, ...)          \
    do {                               \
        if (!(cond)) {                 \
            sg_set_error(__VA_ARGS__); \
            return SG_E_INVALID;       \
        }                              \
    } while (0)

// ---------------------------------------------------------------- PTX helpers
namespace sgptx {

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

// make the barrier init visible to the async (TMA) proxy
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// 1-D TMA bulk copy global -> shared, completion signalled on an mbarrier
// (SASS: UBLKCP).  dst/src 16-B aligned, bytes a multiple of 16.
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// 1-D TMA bulk copy shared -> global (bulk async-group completion).
__device__ __forceinline__ void bulk_s2g(void *gmem_dst, const void *smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst),
                 "r"(smem_u32(smem_src)), "r"(bytes)
                 : "memory");
}
// 2-D tensor-map TMA (SASS: UTMALDG / UTMASTG): box of the map at element coordinates (c0 = inner, c1 = row).
// smem 128-B aligned; elements outside the tensor are zero-filled on load and skipped on store.
__device__ __forceinline__ void tensor_g2s_2d(void *smem_dst, const void *tmap, int c0, int c1, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(tmap), "r"(c0), "r"(c1), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tensor_s2g_2d(const void *tmap, int c0, int c1, const void *smem_src) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];" ::"l"(tmap), "r"(c0), "r"(c1),
                 "r"(smem_u32(smem_src))
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// wait until the shared-memory SOURCE of all committed bulk stores has been read
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// generic-proxy shared writes -> visible to the async proxy
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

}  // namespace sgptx

// ------------------------------------------------------------------ sweep direction
// Consecutive streaming kernels of the solver alternate the direction in which they sweep their arrays: each then starts
// on the part the previous kernel touched last, which is what the 126 MB L2 still holds (the vectors are 157 MB each on
// config 3).  Grid-stride loop over [lo, hi): ascending, or - rev - the same (block, thread) -> element map walked from
// the last grid-stride iteration down to the first, so per-block partial sums cover the same elements either way.
__device__ __forceinline__ void sg_sweep_begin(long lo, long hi, int tb, int rev, long &c, long &step) {
    step = (long)gridDim.x * tb;
    c = lo + (long)blockIdx.x * tb + threadIdx.x;
    if (rev && c < hi) {
        c += ((hi - 1 - c) / step) * step;
        step = -step;
    }
}
__device__ __forceinline__ void sg_sweep_begin_i(int lo, int hi, int tb, int rev, int &c, int &step) {
    step = (int)gridDim.x * tb;
    c = lo + (int)blockIdx.x * tb + (int)threadIdx.x;
    if (rev && c < hi) {
        c += ((hi - 1 - c) / step) * step;
        step = -step;
    }
}
// host: direction of the next streaming launch (alternates; SG_SERPENTINE=0 keeps every sweep ascending)
int sg_next_sweep_dir();

// ------------------------------------------------------------------ reductions
__device__ __forceinline__ double sg_warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block-wide sum, result valid in thread 0 (warp 0).  `scratch` holds >= 32 doubles.
__device__ __forceinline__ double sg_block_sum(double v, double *scratch) {
    v = sg_warp_sum(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    const int nw = (blockDim.x + 31) >> 5;
    v = (threadIdx.x < nw) ? scratch[threadIdx.x] : 0.0;
    if (warp == 0) v = sg_warp_sum(v);
    __syncthreads();
    return v;
}

// Bounded spin on a flag another GPU raises after fencing its data stores; false on expiry (about 20 s).  The flag is read
// with ld.acquire.sys: everything the writer published before the flag is visible to loads that follow — no
// system-scope fence (MEMBAR.SC.SYS costs microseconds per block when 444 blocks issue it at the end of a kernel).
__device__ __forceinline__ unsigned long long sg_ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ bool sg_spin_until(const unsigned long long *flag, unsigned long long seq) {
    const long long t0 = clock64();
    while (sg_ld_acquire_sys(flag) < seq) {
        if (clock64() - t0 > 40000000000ll) return false;
        __nanosleep(32);
    }
    return true;
}

// Block-level wait for ghost rows (SgHaloWait); every thread of the block must call it.  Not inlined: it runs once per
// block and must not cost the hot loops of the operator kernels any registers.  The acquire by threads 0/1 reaches the
// rest of the block through the barrier; the block has not read a ghost row before (interior cells come first) and L1 is
// invalid at kernel start, so no stale line can be hit.
static __device__ __noinline__ void sg_halo_wait_block(const SgHaloWait &w) {
    if (w.n == 0) return;
    if (threadIdx.x < (unsigned)w.n) {
        const unsigned long long *f = threadIdx.x == 0 ? w.flag[0] : w.flag[1];   // no dynamic indexing of the parameter struct
        if (!sg_spin_until(f, w.seq)) *w.err = 1;
    }
    __syncthreads();
}

// The same wait inlined (a handful of instructions after the main sweep of the class kernels).
__device__ __forceinline__ void sg_halo_wait_inline(const SgHaloWait &w) {
    if (threadIdx.x == 0) {
        bool ok = sg_spin_until(w.flag[0], w.seq);
        if (w.n > 1) ok = sg_spin_until(w.flag[1], w.seq) && ok;
        if (!ok) *w.err = 1;
    }
    __syncthreads();
}

// All-reduce of count <= SG_PEER_RED_VALS doubles at vals over the ranks, executed by warp 0 of ONE block per rank.
__device__ __forceinline__ void sg_peer_allreduce_warp(const SgPeerRedDev &a, double *vals, int count) {
    const int t = threadIdx.x;
    unsigned long long seq = 0;
    if (t == 0) seq = *a.seq + 1ull;
    seq = __shfl_sync(0xffffffffu, seq, 0);
    const int par = (int)(seq & 1ull);
    if (t < a.nranks) {
        double *v = reinterpret_cast<double *>(a.base[t] + a.vals_off) + ((size_t)par * SG_PEER_MAX_RANKS + a.rank) * SG_PEER_RED_VALS;
        for (int k = 0; k < count; ++k) reinterpret_cast<volatile double *>(v)[k] = vals[k];
        __threadfence_system();
        unsigned long long *tag = reinterpret_cast<unsigned long long *>(a.base[t] + a.tags_off) + (size_t)par * SG_PEER_MAX_RANKS + a.rank;
        *reinterpret_cast<volatile unsigned long long *>(tag) = seq;
    }
    bool ok = true;
    if (t < a.nranks) {
        const unsigned long long *tag =
            reinterpret_cast<const unsigned long long *>(a.base[a.rank] + a.tags_off) + (size_t)par * SG_PEER_MAX_RANKS + t;
        ok = sg_spin_until(tag, seq);
    }
    ok = __all_sync(0xffffffffu, ok);
    // every lane reads the slot whose tag IT acquired; lane 0 adds them in rank order (bit-identical on every rank)
    const volatile double *v =
        reinterpret_cast<const volatile double *>(a.base[a.rank] + a.vals_off) + ((size_t)par * SG_PEER_MAX_RANKS + (t < a.nranks ? t : 0)) * SG_PEER_RED_VALS;
    for (int k = 0; k < count; ++k) {
        const double mine = (ok && t < a.nranks) ? v[k] : 0.0;
        double s = 0.0;
        for (int r = 0; r < a.nranks; ++r) s += __shfl_sync(0xffffffffu, mine, r);
        if (t == 0 && ok) vals[k] = s;
    }
    if (t == 0) {
        if (!ok) *a.err = 1;
        *a.seq = seq;
        __threadfence_system();
    }
}

// Sum NR per-thread values over the grid; the LAST block to finish adds the per-block partials in block
// order (deterministic) and stores the totals in out[0..NR); with red.peer set it then all-reduces them over the ranks
// (see SgRed).  Must be reached by every thread of every block; gridDim.x <= SG_MAX_BLOCKS.
template <int NR>
__device__ __forceinline__ void sg_grid_reduce(double (&v)[NR], const SgRed red, double *out) {
    __shared__ double scratch[32];
    __shared__ bool is_last;
#pragma unroll
    for (int k = 0; k < NR; ++k) {
        const double s = sg_block_sum(v[k], scratch);
        if (threadIdx.x == 0) red.partials[blockIdx.x * NR + k] = s;
    }
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned done = atomicAdd(red.counter, 1u);
        is_last = (done == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
#pragma unroll
        for (int k = 0; k < NR; ++k) {
            double s = 0.0;
            for (unsigned b = threadIdx.x; b < gridDim.x; b += blockDim.x) s += red.partials[b * NR + k];
            s = sg_block_sum(s, scratch);
            if (threadIdx.x == 0) out[k] = s;
        }
        if (threadIdx.x == 0) *red.counter = 0u;
        if (red.peer) {
            __syncthreads();   // out[] stored by thread 0 is visible to warp 0
            if (threadIdx.x < 32) sg_peer_allreduce_warp(*red.peer, red.ar_ptr ? red.ar_ptr : out, red.ar_count ? red.ar_count : NR);
        }
    }
}
