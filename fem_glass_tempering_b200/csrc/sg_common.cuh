// Shared internals of libsurroglas_b200: error plumbing, context, PTX helpers
// (mbarrier + 1-D TMA bulk copies), warp/block reductions.
#pragma once

#include <cuda_runtime.h>
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

// Grid-wide deterministic reduction scratch: per-block partials + a completion counter (see sg_grid_reduce).
constexpr int SG_MAX_BLOCKS = 1184;  // 8 * 148 resident blocks; kernels that reduce are grid-stride beyond that
struct SgRed {
    double *partials;   // [SG_MAX_BLOCKS * 2]
    unsigned *counter;  // zero between kernels
};

// y = J(T_lin) x fused with the owned-dof dot product: dot2[0] + dot2[1] = sum over owned dofs of x*y
// (cells + exterior facets).  When *skip != 0 (device flag, may be NULL) every kernel returns at once.
// Must precede sg_thermal_apply_dot whenever T_lin changed (refreshes the linearised boundary matrices).
int sg_thermal_linearize(sg_thermal_op *op, const double *T_lin, cudaStream_t st);
// y_is_zero: the caller guarantees y == 0 on entry (CG spaces scatter into y; saves the memset launch).
// part (DG class kernels on a partitioned mesh, see sg_thermal_can_split): SG_PART_ALL, or SG_PART_INTERIOR (cells
// without ghost neighbours: may run before the halo has arrived) followed by SG_PART_BOUNDARY (the rest; ADDS its share
// of the reduction to dot2).
enum { SG_PART_ALL = 0, SG_PART_INTERIOR = 1, SG_PART_BOUNDARY = 2 };
bool sg_thermal_can_split(const sg_thermal_op *op);
int sg_thermal_apply_dot(sg_thermal_op *op, const double *T_lin, const double *x, double *y, SgRed red, double *dot2,
                         const int *skip, cudaStream_t st, int y_is_zero = 0, int part = SG_PART_ALL);

// One fused Chebyshev step of the polynomial preconditioner (DG + class tables only, see dg_cheb_step):
// z_out = z_in + a (z_in - z_prev) + b M^-1 (r - J z_in); z_prev == NULL means 0 (first step); z_out may alias
// z_prev but not z_in; last: *dot_out = r.z_out.
struct SgChebStep {
    const double *z_in, *r, *z_prev;
    double *z_out;
    double a, b;
    int last;
    int part;   // SG_PART_*
};
bool sg_thermal_has_cheb(const sg_thermal_op *op);
bool sg_thermal_profiling(const sg_thermal_op *op);   // event pairs around the kernels are being recorded
int sg_thermal_cheb_step(sg_thermal_op *op, const SgChebStep &cs, SgRed red, double *dot_out, const int *skip, cudaStream_t st);

// peer.cu: NVLink peer-memory halo exchange and small all-reduce (replaces NCCL on the solver's data path)
struct SgPeer;
int sg_peer_create(sg_ctx *ctx, size_t mailbox_doubles, SgPeer **out, void *handle64);
int sg_peer_open(SgPeer *p, const void *handles /* nranks x 64 bytes */);
int sg_peer_destroy(SgPeer *p);
bool sg_peer_ready(const SgPeer *p);
size_t sg_peer_mailbox_doubles(const SgPeer *p);
int sg_peer_check(SgPeer *p);
int sg_peer_halo_forward(SgPeer *p, int n_seg, const sg_halo_segment *seg, double *vec, cudaStream_t st);
// the two halves of a forward exchange, so that work that needs no ghost values can run between them
int sg_peer_halo_push(SgPeer *p, int n_seg, const sg_halo_segment *seg, const double *vec, cudaStream_t st);
int sg_peer_halo_pull(SgPeer *p, int n_seg, const sg_halo_segment *seg, double *vec, cudaStream_t st);
int sg_peer_allreduce(SgPeer *p, double *vals, int count, cudaStream_t st);

// classify.cu: equivalence classes of 64-bit keys.  cls_out[i] = class of keys[i] in [0, *n_cls),
// rep_out[k] = index of one member of class k; both are cudaMalloc'ed here and freed by the caller.
int sg_classify_u64(const uint64_t *keys_dev, int64_t n, int32_t **cls_out, int32_t *n_cls, int32_t **rep_out);

// stencil.cu: row-stencil classes of a CG operator (gather form of the Jacobian apply).  *out stays NULL (and SG_OK is
// returned) when the mesh has no small set of repeating rows.  dot2[0] = x.y over the rows [own_lo, own_hi), dot2[1] = 0.
struct SgStencil;
int sg_stencil_build(sg_ctx *ctx, const int32_t *dofmap, int64_t n_cells, int n_ld, int64_t cell_lo, int64_t cell_hi,
                     const uint16_t *cls16, const double *tab, int S, int64_t n_rows, SgStencil **out);
int sg_stencil_apply(const SgStencil *s, const double *x, double *y, int64_t own_lo, int64_t own_hi, SgRed red, double *dot2,
                     const int *skip, cudaStream_t st);
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
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// wait until the shared-memory SOURCE of all committed bulk stores has been read
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// generic-proxy shared writes -> visible to the async proxy
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

}  // namespace sgptx

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

// Sum NR per-thread values over the grid; the LAST block to finish adds the per-block partials in block
// order (deterministic) and stores the totals in out[0..NR).  Must be reached by every thread of every
// block; gridDim.x <= SG_MAX_BLOCKS.
template <int NR>
__device__ __forceinline__ void sg_grid_reduce(double (&v)[NR], const SgRed red, double *out, const bool accumulate = false) {
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
            if (threadIdx.x == 0) out[k] = accumulate ? out[k] + s : s;   // accumulate: second launch over another cell range
        }
        if (threadIdx.x == 0) *red.counter = 0u;
    }
}
