// Row-stencil classes of a CG operator: the gather form of the Jacobian apply.
//
// cg_class_apply (thermal.cu) is cell-centric: every cell multiplies its class matrix with its gathered dofs and
// scatters with RED.ADD.F64, and the L2 atomic path bounds it at ~15 % of the HBM roofline.  The
// assembled row i of the same operator is  y_i = sum over (cell K, local row a) with dofmap[K][a] == i of
// sum_b A_cls(K)[a][b] x[dofmap[K][b]].  On a mesh whose cells repeat their local matrices AND whose numbering is
// translation invariant (the plate meshes: lattice-numbered P1/P2 nodes), the multiset
//     { (cls(K), a, dofmap[K][.] - i) }
// takes only a handful of values over all rows (3-D P2 Kuhn plate: 64), so a row is fully described by a 16-bit class
// id and the class's list of (column offset, coefficient) pairs:
//     y_i = sum_k coef[cls_i][k] * x[i + off[cls_i][k]]          (plain store, no atomics, deterministic)
// with the class lists staged in shared memory.  Compulsory traffic per row: class id + x + y = 18 B (the cell-centric
// form moves 42 B per P2 tetrahedron on top of that), neighbours of x come from L1/L2.
//
// Rows are classified by a 64-bit commutative hash of that multiset (sum of per-contribution hashes), the tables are
// accumulated from one representative row per class in a fixed order, and thermal.cu verifies the whole construction
// against cg_class_apply on a pseudo-random vector before it switches the operator over; a mesh with too many classes
// (unstructured numbering) simply keeps the cell-centric kernel.
// Reference: the assembled PETSc MatMult inside KSP cg of TVP:340-346, on the Jacobian of TVP:293-306.
#include <stdlib.h>

#include <algorithm>
#include <vector>

#include "sg_common.cuh"

namespace {

constexpr int STB = 256;          // threads per block of the apply
constexpr int MAX_NLD = 10;       // P2 tetrahedron
constexpr int MAX_CONTRIB = 128;  // (cell, local row) pairs of one representative row
constexpr int MAX_NNZ = 192;      // entries of one class row
constexpr int MAX_CLASSES = 8192;   // < 2^15: bit 15 of the 16-bit row class flags rows on an exterior facet
constexpr size_t MAX_SMEM_TABLE = 48 * 1024;

struct __align__(16) Entry {
    double coef;
    long long off;
};

struct StDev {
    long n_rows, own_lo, own_hi;
    long safe_lo, safe_hi;   // rows [safe_lo, safe_hi) read no ghost row: processed first, the rest after the halo wait
    const uint16_t *rcls;
    const int32_t *ptr;   // [n_classes + 1]
    const Entry *ent;     // [n_entries]
    int n_classes, n_entries;
    // Exterior (Robin + radiation) facets in gather form, so that the apply is ONE launch with no atomics: rows that lie on
    // an exterior facet carry bit 15 in rcls; brow_of[row] indexes their list of (facet b, local facet dof k) pairs, and
    // the row adds  sum_l B_b[k][l] x[dof(b, l)]  from the facets' linearised matrices bmat (refreshed once per Newton
    // iteration by sg_thermal_linearize).  bmat == NULL: no boundary part.
    const int32_t *brow_of, *blist, *bcnt, *bf_cell, *bf_facet, *dofmap;
    const int32_t *bcols, *bncol;   // per boundary row: its distinct columns [BND_MAXC] and their number
    const double *bvals;            // per boundary row: the assembled entries, refreshed from bmat once per linearisation
    long n_brows;                   // bcols / bvals are ENTRY-major: [BND_MAXC][n_brows]
    const double *bmat;
    long nc;
    int nfd, maxf;
    const int32_t *fd;    // [4][6] local dofs of the facets (thermal.cu facet_dof), device memory
    int sub;              // 1: y -= (stencil row) instead of y = (stencil row)  (residual in gather form, sg_stencil_apply_cells)
};

constexpr int ROW_BND = 0x8000;
constexpr int BND_MAXF = 16;      // exterior facets per row (Kuhn P2 plate corner: 6)
constexpr int BND_MAXC = 32;      // distinct columns of a boundary row's exterior-facet part (P2 surface vertex: 19)

// boundary part of row `row` (see StDev): the row's assembled exterior-facet entries; CG_LOADS: read x through L2
// (persistent kernel: x changes during the launch)
template <bool CG_LOADS = false>
__device__ __forceinline__ double stencil_boundary_row(const StDev &sd, const long row, const double *x) {
    const int j = sd.brow_of[row];
    const int n = sd.bncol[j];
    const long nb = sd.n_brows;
    double acc = 0.0;
    for (int i0 = 0; i0 < n; i0 += 8) {            // 8 loads in flight (L2 round trips in the persistent kernel); sum in entry order
        const int m = n - i0 < 8 ? n - i0 : 8;
        int32_t c[8];
        double v[8], xv[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            c[u] = u < m ? sd.bcols[(long)(i0 + u) * nb + j] : 0;
            v[u] = u < m ? sd.bvals[(long)(i0 + u) * nb + j] : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) xv[u] = u < m ? (CG_LOADS ? __ldcg(x + c[u]) : x[c[u]]) : 0.0;
#pragma unroll
        for (int u = 0; u < 8; ++u)
            if (u < m) acc = fma(v[u], xv[u], acc);
    }
    return acc;
}

__device__ __forceinline__ uint64_t mix64(uint64_t h, uint64_t v) {
    h ^= v + 0x9E3779B97F4A7C15ull;
    h *= 0xBF58476D1CE4E5B9ull;
    h ^= h >> 31;
    h *= 0x94D049BB133111EBull;
    h ^= h >> 29;
    return h;
}

// sig[row] += hash(cls(K), a, dofmap[K][.] - row) for every (K, a) with dofmap[K][a] == row
__global__ void k_row_sig(const int32_t *__restrict__ dofmap, long nc, int nld, long cell_lo, long cell_hi,
                          const uint16_t *__restrict__ cls16, unsigned long long *sig) {
    const long c = cell_lo + (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cell_hi) return;
    long dof[MAX_NLD];
    for (int j = 0; j < nld; ++j) dof[j] = dofmap[(long)j * nc + c];
    const uint64_t seed = mix64(0x5157454E43494Cull, (uint64_t)cls16[c]);
    for (int a = 0; a < nld; ++a) {
        uint64_t h = mix64(seed, (uint64_t)a);
        for (int b = 0; b < nld; ++b) h = mix64(h, (uint64_t)(long long)(dof[b] - dof[a]));
        atomicAdd(&sig[dof[a]], (unsigned long long)h);
    }
}

// the (cell, local row) pairs that make up each class's representative row
__global__ void k_rep_collect(const int32_t *__restrict__ dofmap, long nc, int nld, long cell_lo, long cell_hi,
                              const int32_t *__restrict__ rcls, const int32_t *__restrict__ rep, int *cnt, long long *list,
                              unsigned *overflow) {
    const long c = cell_lo + (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cell_hi) return;
    for (int a = 0; a < nld; ++a) {
        const int32_t row = dofmap[(long)a * nc + c];
        const int32_t rc = rcls[row];
        if (rep[rc] != row) continue;
        const int slot = atomicAdd(&cnt[rc], 1);
        if (slot < MAX_CONTRIB) list[(long)rc * MAX_CONTRIB + slot] = (long long)c * 16 + a;
        else atomicExch(overflow, 1u);
    }
}

// one thread per class: accumulate the representative row in ascending (cell, local row) order, sort by offset
__global__ void k_rep_build(int n_classes, const int32_t *__restrict__ dofmap, long nc, int nld, const uint16_t *__restrict__ cls16,
                            const double *__restrict__ tab, int S, const int32_t *__restrict__ rep, const int *__restrict__ cnt,
                            long long *list, Entry *out, int *out_nnz, unsigned *overflow) {
    const int rc = blockIdx.x * blockDim.x + threadIdx.x;
    if (rc >= n_classes) return;
    const int n = cnt[rc] < MAX_CONTRIB ? cnt[rc] : MAX_CONTRIB;
    long long *l = list + (long)rc * MAX_CONTRIB;
    for (int i = 1; i < n; ++i) {
        const long long v = l[i];
        int j = i - 1;
        for (; j >= 0 && l[j] > v; --j) l[j + 1] = l[j];
        l[j + 1] = v;
    }
    const long row = rep[rc];
    Entry *e = out + (long)rc * MAX_NNZ;
    int nnz = 0;
    for (int i = 0; i < n; ++i) {
        const long c = (long)(l[i] >> 4);
        const int a = (int)(l[i] & 15);
        const double *A = tab + (long)cls16[c] * S + a * nld;
        for (int b = 0; b < nld; ++b) {
            const long long off = (long long)dofmap[(long)b * nc + c] - row;
            int k = 0;
            while (k < nnz && e[k].off != off) ++k;
            if (k == nnz) {
                if (nnz == MAX_NNZ) {
                    atomicExch(overflow, 1u);
                    continue;
                }
                e[k].off = off;
                e[k].coef = 0.0;
                ++nnz;
            }
            e[k].coef += A[b];
        }
    }
    for (int i = 1; i < nnz; ++i) {
        const Entry v = e[i];
        int j = i - 1;
        for (; j >= 0 && e[j].off > v.off; --j) e[j + 1] = e[j];
        e[j + 1] = v;
    }
    out_nnz[rc] = nnz;
}

__global__ void k_narrow16(long n, const int32_t *__restrict__ in, uint16_t *__restrict__ out) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (uint16_t)in[i];
}

// y_i = sum_k coef[cls_i][k] x[i + off[cls_i][k]];  dot_out[0] = sum over owned rows of x_i y_i, dot_out[1] = 0.
// The x loads of U entries are issued before the first FMA (they are independent; two thirds of them miss L1 and
// come from L2, so the loads in flight per warp set the pace); the sum itself stays in entry order for every U.
template <bool SMEM, int U, int MINB>
__global__ void __launch_bounds__(STB, MINB) k_stencil_apply(const StDev sd, const double *__restrict__ x, double *__restrict__ y,
                                                            SgRed red, double *dot_out, const int *skip, const SgHaloWait hw) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    if (skip && *skip) return;
    const Entry *ent = sd.ent;
    const int32_t *ptr = sd.ptr;
    if constexpr (SMEM) {
        Entry *s_ent = reinterpret_cast<Entry *>(s_raw);
        int32_t *s_ptr = reinterpret_cast<int32_t *>(s_ent + sd.n_entries);
        for (int i = threadIdx.x; i < sd.n_entries; i += STB) s_ent[i] = sd.ent[i];
        for (int i = threadIdx.x; i <= sd.n_classes; i += STB) s_ptr[i] = sd.ptr[i];
        __syncthreads();
        ent = s_ent;
        ptr = s_ptr;
    }
    double dsum[2] = {0.0, 0.0};
    // one sweep over idx in [0, n_rows): the rows that read no ghost value first, the rows next to the slab faces at the end
    // of the index space (only blocks in their last iteration wait for the neighbours' puts)
    const long n_safe = sd.safe_hi - sd.safe_lo;
    bool waited = hw.n == 0;
    for (long base = (long)blockIdx.x * STB; base < sd.n_rows; base += (long)gridDim.x * STB) {
        if (!waited && base + STB > n_safe) {     // block-uniform; blocks whose rows are all safe never wait and retire
            sg_halo_wait_inline(hw);
            waited = true;
        }
        const long idx = base + threadIdx.x;
        if (idx >= sd.n_rows) continue;
        const long j = idx - n_safe;
        const long row = idx < n_safe ? sd.safe_lo + idx : (j < sd.safe_lo ? j : sd.safe_hi + (j - sd.safe_lo));
        const int cw = sd.rcls[row], c = cw & (ROW_BND - 1);
        const int p1 = ptr[c + 1];
        int k = ptr[c];
        const double *xr = x + row;
        double acc = 0.0;
        for (; k + U <= p1; k += U) {
            Entry e[U];
            double xv[U];
#pragma unroll
            for (int u = 0; u < U; ++u) e[u] = ent[k + u];
#pragma unroll
            for (int u = 0; u < U; ++u) xv[u] = __ldg(xr + e[u].off);
#pragma unroll
            for (int u = 0; u < U; ++u) acc = fma(e[u].coef, xv[u], acc);
        }
        for (; k < p1; ++k) {
            const Entry e = ent[k];
            acc = fma(e.coef, __ldg(xr + e.off), acc);
        }
        if ((cw & ROW_BND) && sd.bmat) acc += stencil_boundary_row(sd, row, x);
        y[row] = sd.sub ? y[row] - acc : acc;
        if (row >= sd.own_lo && row < sd.own_hi) dsum[0] += __ldg(xr) * acc;
    }
    sg_grid_reduce<2>(dsum, red, dot_out);
}

// ---- boundary rows: mark, number, collect
__global__ void k_bnd_mark(long n_bf, const int32_t *__restrict__ bf_cell, const int32_t *__restrict__ bf_facet,
                           const int32_t *__restrict__ dofmap, long nc, long cell_lo, long cell_hi, int nfd, const StDev sd, int32_t *flag) {
    const long b = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_bf) return;
    const long cell = bf_cell[b];
    if (cell < cell_lo || cell >= cell_hi) return;
    const int f = bf_facet[b];
    for (int k = 0; k < nfd; ++k) flag[dofmap[(long)sd.fd[f * 6 + k] * nc + cell]] = 1;
}
// pos = exclusive scan of flag: boundary rows are numbered in ROW ORDER, so that consecutive boundary rows (one warp of
// the apply) read consecutive entries of the entry-major bcols/bvals arrays
__global__ void k_bnd_number(long n, const int32_t *__restrict__ flag, const int32_t *__restrict__ pos, int32_t *brow_of) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) brow_of[i] = flag[i] ? pos[i] : -1;
}
__global__ void k_bnd_collect(long n_bf, const int32_t *__restrict__ bf_cell, const int32_t *__restrict__ bf_facet,
                              const int32_t *__restrict__ dofmap, long nc, long cell_lo, long cell_hi, int nfd, const StDev sd,
                              const int32_t *__restrict__ brow_of, int32_t *bcnt, int32_t *blist, unsigned *overflow) {
    const long b = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_bf) return;
    const long cell = bf_cell[b];
    if (cell < cell_lo || cell >= cell_hi) return;
    const int f = bf_facet[b];
    for (int k = 0; k < nfd; ++k) {
        const int j = brow_of[dofmap[(long)sd.fd[f * 6 + k] * nc + cell]];
        const int slot = atomicAdd(&bcnt[j], 1);
        if (slot < BND_MAXF) blist[(long)j * BND_MAXF + slot] = (int)(b * 8 + k);
        else atomicExch(overflow, 1u);
    }
}
__global__ void k_bnd_sort_flag(long n_brows, int32_t *bcnt, int32_t *blist) {      // fixed summation order per row
    const long j = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_brows) return;
    const int n = bcnt[j] < BND_MAXF ? bcnt[j] : BND_MAXF;
    int32_t *l = blist + j * BND_MAXF;
    for (int i = 1; i < n; ++i) {
        const int32_t v = l[i];
        int q = i - 1;
        for (; q >= 0 && l[q] > v; --q) l[q + 1] = l[q];
        l[q + 1] = v;
    }
}
// distinct columns of every boundary row, in order of first appearance over its (sorted) facet list
__global__ void k_bnd_columns(long n_brows, const StDev sd, int32_t *bcols, int32_t *bncol, unsigned *overflow) {
    const long j = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_brows) return;
    const int n = sd.bcnt[j];
    int32_t cols[BND_MAXC];
    int nc = 0;
    for (int e = 0; e < n; ++e) {
        const int pk = sd.blist[j * BND_MAXF + e];
        const int b = pk >> 3;
        const long cell = sd.bf_cell[b];
        const int f = sd.bf_facet[b];
        for (int l = 0; l < sd.nfd; ++l) {
            const int32_t col = sd.dofmap[(long)sd.fd[f * 6 + l] * sd.nc + cell];
            int q = 0;
            while (q < nc && cols[q] != col) ++q;
            if (q == nc) {
                if (nc == BND_MAXC) {
                    atomicExch(overflow, 1u);
                    continue;
                }
                cols[nc++] = col;
            }
        }
    }
    bncol[j] = nc;
    for (int i = 0; i < nc; ++i) bcols[(long)i * n_brows + j] = cols[i];
}
// bvals[j][i] = sum over the row's (facet b, local dof k) pairs of B_b[k][l] with dof(b, l) == column i   (fixed order)
__global__ void k_bnd_refresh(long n_brows, const StDev sd, double *bvals) {
    const long j = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_brows) return;
    const int n = sd.bcnt[j], nc = sd.bncol[j], nfd = sd.nfd;
    int32_t cols[BND_MAXC];
    double v[BND_MAXC];
    for (int i = 0; i < nc; ++i) {
        v[i] = 0.0;
        cols[i] = sd.bcols[(long)i * n_brows + j];
    }
    for (int e = 0; e < n; ++e) {
        const int pk = sd.blist[j * BND_MAXF + e];
        const int b = pk >> 3, k = pk & 7;
        const long cell = sd.bf_cell[b];
        const int f = sd.bf_facet[b];
        const double *B = sd.bmat + (long)b * (nfd * (nfd + 1) / 2);
        for (int l = 0; l < nfd; ++l) {
            const int lo = k < l ? k : l, hi = k < l ? l : k;
            const double bkl = B[lo * nfd - lo * (lo - 1) / 2 + (hi - lo)];   // packed upper triangle, row-wise
            const int32_t col = sd.dofmap[(long)sd.fd[f * 6 + l] * sd.nc + cell];
            for (int i = 0; i < nc; ++i)
                if (cols[i] == col) v[i] += bkl;
        }
    }
    for (int i = 0; i < nc; ++i) bvals[(long)i * n_brows + j] = v[i];
}
__global__ void k_bnd_flag_rows(long n, const int32_t *__restrict__ brow_of, uint16_t *rcls) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && brow_of[i] >= 0) rcls[i] |= (uint16_t)ROW_BND;
}

// ================================================================ persistent PCG (small CG problems)
// Config 2 (334 k rows) is latency bound: a Jacobi-PCG iteration is three ~4 us kernels and the time step needs ~240 of
// them.  Here the WHOLE linear solve is one cooperative launch: every thread keeps its rows of x, r, p, s = Ap and 1/diag in
// registers for the entire solve, the only vector that goes through memory (L2) is u = M^-1 r, which the neighbours' rows
// gather, and an iteration costs two grid-wide barriers instead of three launches.  Single-reduction CG (Chronopoulos &
// Gear 1989): with w = A u,  gamma = r.u,  delta = w.u  reduced TOGETHER,
//     beta = gamma/gamma_old,  alpha = gamma / (delta - beta gamma / alpha_old),
//     p = u + beta p,  s = w + beta s,  x += alpha p,  r -= alpha s,  u = M^-1 r.
// Reductions are deterministic (per-block partials summed in block order by every block).  The tolerance policy (plain
// rtol/atol or the Eisenstat-Walker forcing term of the inexact Newton iteration) is evaluated on the device from |b|.
constexpr int PTB = 512;   // threads per block of the persistent kernel: ONE block per SM (a grid barrier has <= 148 participants),
                           // 128 registers per thread: the rows' state never spills

// three block-wide sums with one shared-memory stage; result valid in every thread
__device__ __forceinline__ void block_sum3(double (&v)[3], double (*scratch)[32], double (&tot)[3]) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        const double w = sg_warp_sum(v[q]);
        if (lane == 0) scratch[q][warp] = w;
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        double w = lane < nw ? scratch[q][lane] : 0.0;
        tot[q] = sg_warp_sum(w);      // every warp adds the same nw values in the same order
    }
    __syncthreads();
}

struct PersistArgs {
    StDev sd;
    const double *b, *dinv;
    double *x, *u;            // u: work vector [n_rows] in global memory
    double *partials;         // [2][gridDim.x][4]
    unsigned *bar;            // grid barrier counter, zero at launch
    SgPcgPolicy pol;
    int max_it;
    long win;                 // > 0: rows are block-contiguous and u is gathered from a shared-memory window of the block's
                              // rows +- win (= largest |column offset|); 0: rows are grid-strided and u is gathered from L2
    int *ctrl_done, *ctrl_iters;
    double *ctrl_rr, *rr0_out;
    long long *dbg;           // SG_PERSIST_TIMING=1: clock cycles per phase of block 0 (measurement only)
    int bcap;                 // boundary rows per block whose assembled entries are staged in shared memory for the solve
};
constexpr int PERSIST_BS = 10;    // ... with at most this many entries each (2-D P2 boundary vertex: 5); others read global memory

// Grid-wide barrier of the persistent kernel (all blocks co-resident: cooperative launch).  Arrival is a RELEASE reduction
// (the block's earlier stores are visible at GPU scope before the count is), the poll is a RELAXED load: an acquire load
// makes the compiler invalidate the whole L1 (CCTL.IVALL) after every barrier, which turned every constant the rows read
// (class ids, boundary entries, spilled registers) into an L2 round trip per iteration.  Everything another block wrote
// is read with ld.cg (L2), and bar.sync orders the poll before the block's following loads.
__device__ __forceinline__ void grid_barrier(unsigned *bar, unsigned &gen) {
    __syncthreads();
    if (threadIdx.x == 0) {
        gen += gridDim.x;
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
        unsigned v;
        do {
            asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
        } while (v < gen);
    }
    __syncthreads();
}

// WIN: x is the block's shared-memory window (xw[j] = u[win_lo + j]); the boundary part still reads the global vector xg
template <bool WIN>
__device__ __forceinline__ double stencil_row(const StDev &sd, const Entry *ent, const int32_t *ptr, const long row, const int cw,
                                              const double *x, const double *xg) {
    const int c = cw & (ROW_BND - 1);
    const int p1 = ptr[c + 1];
    int k = ptr[c];
    const double *xr = x + row;
    double acc = 0.0;
    // up to 16 loads in flight: the vector comes from L2 (it was written by other SMs since the last barrier, so L1 must be
    // bypassed) and the iteration is bound by L2 round trips, not by bandwidth; the sum stays in entry order
    for (; k < p1; k += 16) {
        const int nloc = p1 - k < 16 ? p1 - k : 16;
        double xv[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) xv[u] = u < nloc ? (WIN ? xr[ent[k + u].off] : __ldcg(xr + ent[k + u].off)) : 0.0;
#pragma unroll
        for (int u = 0; u < 16; ++u)
            if (u < nloc) acc = fma(ent[k + u].coef, xv[u], acc);
    }
    // the boundary columns share a cell with the row: inside the window as well (WIN: x - absolute index - is the window)
    if ((cw & ROW_BND) && sd.bmat) acc += WIN ? stencil_boundary_row<false>(sd, row, x) : stencil_boundary_row<true>(sd, row, xg);
    return acc;
}

// Boundary part of a row whose index j into the boundary-row arrays is already known (persistent kernel: looked up once
// per solve), gathering from the block's shared-memory window: one level of (L1-resident) global loads per iteration
// instead of the chain brow_of -> bncol -> bcols/bvals.
__device__ __forceinline__ double stencil_boundary_row_window(const StDev &sd, const int j, const double *s_win, const int win_lo) {
    const int n = sd.bncol[j];
    const long nb = sd.n_brows;
    double acc = 0.0;
    for (int i0 = 0; i0 < n; i0 += 8) {
        const int m = n - i0 < 8 ? n - i0 : 8;
        int32_t c[8];
        double v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            c[u] = u < m ? sd.bcols[(long)(i0 + u) * nb + j] : win_lo;
            v[u] = u < m ? sd.bvals[(long)(i0 + u) * nb + j] : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u)
            if (u < m) acc = fma(v[u], s_win[c[u] - win_lo], acc);
    }
    return acc;
}

template <int R, bool WIN, bool BST>
__global__ void __launch_bounds__(PTB, 1) k_cg_persistent(const __grid_constant__ PersistArgs a) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    __shared__ double scratch[3][32];
    const StDev &sd = a.sd;
    Entry *s_ent = reinterpret_cast<Entry *>(s_raw);
    int32_t *s_ptr = reinterpret_cast<int32_t *>(s_ent + sd.n_entries);
    for (int i = threadIdx.x; i < sd.n_entries; i += PTB) s_ent[i] = sd.ent[i];
    for (int i = threadIdx.x; i <= sd.n_classes; i += PTB) s_ptr[i] = sd.ptr[i];
    __syncthreads();
    const long n = sd.n_rows;
    // WIN: block b owns the contiguous rows [b R PTB, (b+1) R PTB); otherwise rows are grid-strided (coalesced either way)
    const long T = WIN ? (long)PTB : (long)gridDim.x * PTB;
    const long t0 = WIN ? (long)blockIdx.x * R * PTB + threadIdx.x : (long)blockIdx.x * PTB + threadIdx.x;
    // shared-memory window of u: the block's rows and a.win rows on either side (the stencil's reach)
    double *s_win = reinterpret_cast<double *>(s_raw + ((sizeof(Entry) * (size_t)sd.n_entries + sizeof(int32_t) * (size_t)(sd.n_classes + 1) + 15) & ~(size_t)15));
    const long blk_lo = (long)blockIdx.x * R * PTB, win_lo = blk_lo - a.win, win_n = (long)R * PTB + 2 * a.win;
    unsigned gen = 0;
    // per-row state of the whole solve: r, p, s in registers; WIN: x, 1/diag and the boundary-row index in shared memory
    // ([k][thread], conflict-free) — with all seven vectors in registers the kernel spilled (R = 5: 112 bytes per thread,
    // reloaded from local memory every iteration)
    double *s_x = s_win + (WIN ? win_n : 0);
    double *s_di = s_x + (WIN ? (long)R * PTB : 0);
    double *s_bv = s_di + (WIN ? (long)R * PTB : 0);                          // [PERSIST_BS][bcap] boundary entries: values
    int32_t *s_bc = reinterpret_cast<int32_t *>(s_bv + (WIN ? (long)PERSIST_BS * a.bcap : 0));   // ... window-local columns
    int32_t *s_bn = s_bc + (WIN ? (long)PERSIST_BS * a.bcap : 0);             // [bcap] entries per staged row
    int32_t *s_bj = s_bn + (WIN ? a.bcap : 0);                                // [R][PTB]: >= 0 global boundary row, <= -2 staged slot, -1 none
    __shared__ int s_bcount;
    if (threadIdx.x == 0) s_bcount = 0;
    __syncthreads();
    double xr_[WIN ? 1 : R], dir_[WIN ? 1 : R];
    double r[R], p[R], s[R];
    int cw[R];                                          // class word of each row: read once, not once per iteration
#pragma unroll
    for (int k = 0; k < R; ++k) {
        const long row = t0 + k * T;
        const bool in = row < n;
        cw[k] = in ? (int)sd.rcls[row] : 0;
        p[k] = 0.0;
        s[k] = 0.0;
        r[k] = in ? a.b[row] : 0.0;
        const double dk = in ? a.dinv[row] : 0.0;
        if constexpr (WIN) {
            s_x[k * PTB + threadIdx.x] = 0.0;
            s_di[k * PTB + threadIdx.x] = dk;
            int bj = (in && (cw[k] & ROW_BND) && sd.bmat) ? sd.brow_of[row] : -1;
            if (BST && bj >= 0) {
                // the row's assembled exterior-facet entries are constant during the solve: keep them on chip
                const int nb = sd.bncol[bj];
                const int slot = nb <= PERSIST_BS ? atomicAdd(&s_bcount, 1) : a.bcap;
                if (slot < a.bcap) {
                    for (int i = 0; i < nb; ++i) {
                        s_bc[i * a.bcap + slot] = (int32_t)(sd.bcols[(long)i * sd.n_brows + bj] - win_lo);
                        s_bv[i * a.bcap + slot] = sd.bvals[(long)i * sd.n_brows + bj];
                    }
                    s_bn[slot] = nb;
                    bj = -2 - slot;
                }
            }
            s_bj[k * PTB + threadIdx.x] = bj;
        } else {
            xr_[k] = 0.0;
            dir_[k] = dk;
        }
        if (in) a.u[row] = dk * r[k];
    }
    double gamma_old = 1.0, alpha = 1.0, tol2 = 0.0, rr = 0.0;
    int it = 0, done = 0;
    const bool timing = a.dbg && blockIdx.x == 0 && threadIdx.x == 0;
    long long tph[7] = {0, 0, 0, 0, 0, 0, 0}, tc = timing ? clock64() : 0, t_row_acc = 0;
#define SG_PHASE(i)                                 \
    if (timing) {                                   \
        const long long tn = clock64();             \
        tph[i] += tn - tc;                          \
        tc = tn;                                    \
    }
    for (;; ++it) {
        grid_barrier(a.bar, gen);                       // u complete everywhere
        SG_PHASE(0)
        if constexpr (WIN) {                            // one L2 round trip fills the window; the gathers then hit shared memory
            for (long j = threadIdx.x; j < win_n; j += PTB) {
                const long g = win_lo + j;
                s_win[j] = (g >= 0 && g < n) ? __ldcg(a.u + g) : 0.0;
            }
            __syncthreads();
        }
        SG_PHASE(1)
        double acc[3] = {0.0, 0.0, 0.0}, tot[3];
        double w[R];
        const long long t_rows = a.dbg ? clock64() : 0;
        if constexpr (WIN) {
            // The R rows of a thread advance TOGETHER through their class lists: R independent FMA chains hide the
            // LDS -> LDS -> DFMA latency that a row-after-row loop exposes with only four warps per scheduler (each row's sum
            // still runs in entry order).  Rows past the end of the vector have an empty list.
            int k0[R], len[R], loc[R];
            int maxlen = 0;
#pragma unroll
            for (int k = 0; k < R; ++k) {
                const long row = t0 + k * T;
                const int c = cw[k] & (ROW_BND - 1);
                k0[k] = s_ptr[c];
                len[k] = row < n ? s_ptr[c + 1] - k0[k] : 0;
                loc[k] = (int)(row - win_lo);
                maxlen = len[k] > maxlen ? len[k] : maxlen;
                w[k] = 0.0;
            }
            for (int i = 0; i < maxlen; ++i) {
#pragma unroll
                for (int k = 0; k < R; ++k)
                    if (i < len[k]) {
                        const Entry e = s_ent[k0[k] + i];
                        w[k] = fma(e.coef, s_win[loc[k] + (int)e.off], w[k]);
                    }
            }
#pragma unroll
            for (int k = 0; k < R; ++k) {
                const long row = t0 + k * T;
                if (row < n) {
                    const double uk = s_di[k * PTB + threadIdx.x] * r[k];
                    const int bj = (cw[k] & ROW_BND) ? s_bj[k * PTB + threadIdx.x] : -1;
                    if (BST && bj <= -2) {
                        const int slot = -2 - bj, nb = s_bn[slot];
                        double bacc = 0.0;
                        for (int i = 0; i < nb; ++i) bacc = fma(s_bv[i * a.bcap + slot], s_win[s_bc[i * a.bcap + slot]], bacc);
                        w[k] += bacc;
                    } else if (bj >= 0) {
                        w[k] += stencil_boundary_row_window(sd, bj, s_win, (int)win_lo);
                    }
                    acc[0] += r[k] * uk;
                    acc[1] += w[k] * uk;
                    acc[2] += r[k] * r[k];
                }
            }
        } else {
#pragma unroll
            for (int k = 0; k < R; ++k) {
                const long row = t0 + k * T;
                w[k] = 0.0;
                if (row < n) {
                    const double uk = dir_[k] * r[k];
                    w[k] = stencil_row<false>(sd, s_ent, s_ptr, row, cw[k], a.u, a.u);
                    acc[0] += r[k] * uk;
                    acc[1] += w[k] * uk;
                    acc[2] += r[k] * r[k];
                }
            }
        }
        SG_PHASE(2)
        if (a.dbg) t_row_acc += clock64() - t_rows;
        block_sum3(acc, scratch, tot);
        double *part = a.partials + ((size_t)(it & 1) * gridDim.x + blockIdx.x) * 4;
        if (threadIdx.x < 3) part[threadIdx.x] = tot[threadIdx.x == 0 ? 0 : (threadIdx.x == 1 ? 1 : 2)];
        SG_PHASE(3)
        grid_barrier(a.bar, gen);                       // partials complete everywhere
        SG_PHASE(4)
        {
            const double *all = a.partials + (size_t)(it & 1) * gridDim.x * 4;
            double v[3] = {0.0, 0.0, 0.0};
            for (unsigned bq = threadIdx.x; bq < gridDim.x; bq += PTB) {   // gridDim.x <= 148 < PTB: one block's partials per thread
                v[0] += __ldcg(all + (size_t)bq * 4);
                v[1] += __ldcg(all + (size_t)bq * 4 + 1);
                v[2] += __ldcg(all + (size_t)bq * 4 + 2);
            }
            block_sum3(v, scratch, tot);                // identical order in every block: identical totals
        }
        SG_PHASE(5)
        const double gamma = tot[0], delta = tot[1];
        rr = tot[2];
        if (it == 0) {
            tol2 = a.pol.tol2(rr);
            if (blockIdx.x == 0 && threadIdx.x == 0) *a.rr0_out = rr;
        }
        if (!(rr > tol2) || !isfinite(rr)) {            // it == 0: x = 0 is the answer (or the right-hand side is not finite)
            done = isfinite(rr) ? 1 : 2;
            break;
        }
        if (it >= a.max_it) break;
        const double beta = it == 0 ? 0.0 : gamma / gamma_old;
        alpha = it == 0 ? gamma / delta : gamma / (delta - beta * gamma / alpha);
        gamma_old = gamma;
#pragma unroll
        for (int k = 0; k < R; ++k) {
            const long row = t0 + k * T;
            const double dk = WIN ? s_di[k * PTB + threadIdx.x] : dir_[k];
            p[k] = dk * r[k] + beta * p[k];             // u = D^-1 r of this iteration (r is updated two lines below)
            s[k] = w[k] + beta * s[k];
            if constexpr (WIN) s_x[k * PTB + threadIdx.x] += alpha * p[k];
            else xr_[k] += alpha * p[k];
            r[k] -= alpha * s[k];
            if (row < n) a.u[row] = dk * r[k];
        }
        SG_PHASE(6)
    }
#undef SG_PHASE
    if (timing) {
        for (int i = 0; i < 7; ++i) a.dbg[i] += tph[i];
        a.dbg[7] += it;
    }
    if (a.dbg && (threadIdx.x == 0 || threadIdx.x == PTB - 32))       // row phase of the first and last warp of every block
        a.dbg[8 + 2 * blockIdx.x + (threadIdx.x ? 1 : 0)] += t_row_acc;
#pragma unroll
    for (int k = 0; k < R; ++k) {
        const long row = t0 + k * T;
        if (row < n) a.x[row] = WIN ? s_x[k * PTB + threadIdx.x] : xr_[k];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        *a.ctrl_done = done;
        *a.ctrl_iters = it;
        *a.ctrl_rr = rr;
    }
}

using StencilKernel = void (*)(const StDev, const double *, double *, SgRed, double *, const int *, const SgHaloWait);
struct Variant {
    StencilKernel smem, global;
    const char *what;
};
// SG_STENCIL_VARIANT selects one at operator creation (measurement only); the default is variant 0.  Measured on one
// GPU's share of config 4 (3.86 M rows, tools/cg_apply_probe.py, profiles/r1s_cg_apply_probe.jsonl): 76.2 / 77.9 / 82.0 /
// 76.3 / - / 76.9 us - neither more loads in flight nor more resident warps help: ncu has the L1 at 65 % of its peak
// with 8.5 sectors per request and a 33 % hit rate (600 MB of L2->L1 traffic per apply), i.e. the gather is bound by
// L1 wavefronts (an unaligned 256-B warp load touches 3 lines), not by latency and not by HBM (42 MB of DRAM traffic).
const Variant VARIANTS[] = {
    {k_stencil_apply<true, 4, 4>, k_stencil_apply<false, 4, 4>, "4 loads in flight, 4 blocks/SM"},
    {k_stencil_apply<true, 8, 4>, k_stencil_apply<false, 8, 4>, "8 loads in flight, 4 blocks/SM"},
    {k_stencil_apply<true, 4, 8>, k_stencil_apply<false, 4, 8>, "4 loads in flight, 8 blocks/SM"},
    {k_stencil_apply<true, 8, 6>, k_stencil_apply<false, 8, 6>, "8 loads in flight, 6 blocks/SM"},
    {k_stencil_apply<true, 12, 3>, k_stencil_apply<false, 12, 3>, "12 loads in flight, 3 blocks/SM"},
    {k_stencil_apply<true, 6, 5>, k_stencil_apply<false, 6, 5>, "6 loads in flight, 5 blocks/SM"},
};
constexpr int N_VARIANTS = (int)(sizeof(VARIANTS) / sizeof(VARIANTS[0]));

struct DevBuf {
    void *p = nullptr;
    ~DevBuf() {
        if (p) cudaFree(p);
    }
    template <class T>
    T *as() { return static_cast<T *>(p); }
};

}  // namespace

struct SgStencil {
    StDev dev;
    uint16_t *rcls;
    int32_t *ptr;
    Entry *ent;
    int grid;
    size_t smem;   // bytes of the shared-memory copy of the class lists; 0: read through L1
    int max_nnz;
    long max_off;   // largest |column offset| of any class row
    StencilKernel kernel;
    int32_t *brow_of, *blist, *bcnt, *fd, *bcols, *bncol;   // exterior facets in gather form (sg_stencil_attach_boundary)
    double *bvals;
    long n_brows;
    double *pcg_partials;                   // persistent PCG scratch (allocated at first use)
    unsigned *pcg_bar;
};

void sg_stencil_destroy(SgStencil *s) {
    if (!s) return;
    if (s->rcls) cudaFree(s->rcls);
    if (s->ptr) cudaFree(s->ptr);
    if (s->ent) cudaFree(s->ent);
    if (s->brow_of) cudaFree(s->brow_of);
    if (s->blist) cudaFree(s->blist);
    if (s->bcnt) cudaFree(s->bcnt);
    if (s->fd) cudaFree(s->fd);
    if (s->bcols) cudaFree(s->bcols);
    if (s->bncol) cudaFree(s->bncol);
    if (s->bvals) cudaFree(s->bvals);
    if (s->pcg_partials) cudaFree(s->pcg_partials);
    if (s->pcg_bar) cudaFree(s->pcg_bar);
    delete s;
}

int sg_stencil_build(sg_ctx *ctx, const int32_t *dofmap, int64_t n_cells, int n_ld, int64_t cell_lo, int64_t cell_hi,
                     const uint16_t *cls16, const double *tab, int S, int64_t n_rows, SgStencil **out) {
    *out = nullptr;
    if (n_ld > MAX_NLD || n_rows <= 0 || n_rows >= (int64_t)0x7fffffff || cell_hi <= cell_lo) return SG_OK;
    const long ncell = cell_hi - cell_lo;
    const unsigned gcell = (unsigned)((ncell + 255) / 256);
    DevBuf sig, rcls32, rep, cnt, list, scratch, nnz_dev, ovf;
    SG_CHECK_CUDA(cudaMalloc(&sig.p, sizeof(uint64_t) * (size_t)n_rows));
    SG_CHECK_CUDA(cudaMemset(sig.p, 0, sizeof(uint64_t) * (size_t)n_rows));
    k_row_sig<<<gcell, 256>>>(dofmap, n_cells, n_ld, cell_lo, cell_hi, cls16, sig.as<unsigned long long>());
    SG_CHECK_CUDA(cudaGetLastError());
    int32_t R = 0;
    int rc = sg_classify_u64(sig.as<uint64_t>(), n_rows, (int32_t **)&rcls32.p, &R, (int32_t **)&rep.p);
    if (rc) return rc;
    if (R > MAX_CLASSES) return SG_OK;   // no repeating row pattern (unstructured numbering): keep the cell-centric kernel
    SG_CHECK_CUDA(cudaMalloc(&cnt.p, sizeof(int) * (size_t)R));
    SG_CHECK_CUDA(cudaMemset(cnt.p, 0, sizeof(int) * (size_t)R));
    SG_CHECK_CUDA(cudaMalloc(&list.p, sizeof(long long) * (size_t)R * MAX_CONTRIB));
    SG_CHECK_CUDA(cudaMalloc(&scratch.p, sizeof(Entry) * (size_t)R * MAX_NNZ));
    SG_CHECK_CUDA(cudaMalloc(&nnz_dev.p, sizeof(int) * (size_t)R));
    SG_CHECK_CUDA(cudaMalloc(&ovf.p, sizeof(unsigned)));
    SG_CHECK_CUDA(cudaMemset(ovf.p, 0, sizeof(unsigned)));
    k_rep_collect<<<gcell, 256>>>(dofmap, n_cells, n_ld, cell_lo, cell_hi, rcls32.as<int32_t>(), rep.as<int32_t>(), cnt.as<int>(),
                                  list.as<long long>(), ovf.as<unsigned>());
    SG_CHECK_CUDA(cudaGetLastError());
    k_rep_build<<<(R + 63) / 64, 64>>>(R, dofmap, n_cells, n_ld, cls16, tab, S, rep.as<int32_t>(), cnt.as<int>(), list.as<long long>(),
                                       scratch.as<Entry>(), nnz_dev.as<int>(), ovf.as<unsigned>());
    SG_CHECK_CUDA(cudaGetLastError());
    unsigned overflow = 0;
    SG_CHECK_CUDA(cudaMemcpy(&overflow, ovf.p, sizeof(unsigned), cudaMemcpyDeviceToHost));
    if (overflow) return SG_OK;
    std::vector<int> nnz(R);
    std::vector<Entry> wide((size_t)R * MAX_NNZ);
    SG_CHECK_CUDA(cudaMemcpy(nnz.data(), nnz_dev.p, sizeof(int) * (size_t)R, cudaMemcpyDeviceToHost));
    SG_CHECK_CUDA(cudaMemcpy(wide.data(), scratch.p, sizeof(Entry) * wide.size(), cudaMemcpyDeviceToHost));
    std::vector<int32_t> ptr(R + 1, 0);
    int max_nnz = 0;
    for (int c = 0; c < R; ++c) {
        ptr[c + 1] = ptr[c] + nnz[c];
        max_nnz = std::max(max_nnz, nnz[c]);
    }
    std::vector<Entry> ent((size_t)std::max(ptr[R], 1));
    for (int c = 0; c < R; ++c) std::copy(wide.begin() + (size_t)c * MAX_NNZ, wide.begin() + (size_t)c * MAX_NNZ + nnz[c], ent.begin() + ptr[c]);
    long long max_off = 0;
    for (int i = 0; i < ptr[R]; ++i) max_off = std::max(max_off, ent[i].off < 0 ? -ent[i].off : ent[i].off);

    SgStencil *s = new SgStencil();
    memset(s, 0, sizeof(*s));
    cudaError_t e = cudaMalloc(&s->rcls, sizeof(uint16_t) * (size_t)n_rows);
    if (e == cudaSuccess) e = cudaMalloc(&s->ptr, sizeof(int32_t) * (size_t)(R + 1));
    if (e == cudaSuccess) e = cudaMalloc(&s->ent, sizeof(Entry) * ent.size());
    if (e == cudaSuccess) e = cudaMemcpy(s->ptr, ptr.data(), sizeof(int32_t) * (size_t)(R + 1), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(s->ent, ent.data(), sizeof(Entry) * ent.size(), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        sg_set_error("sg_stencil_build: %s", cudaGetErrorString(e));
        sg_stencil_destroy(s);
        return SG_E_CUDA;
    }
    k_narrow16<<<(unsigned)((n_rows + 255) / 256), 256>>>(n_rows, rcls32.as<int32_t>(), s->rcls);
    SG_CHECK_CUDA(cudaGetLastError());
    SG_CHECK_CUDA(cudaDeviceSynchronize());
    const size_t table = sizeof(Entry) * (size_t)ptr[R] + sizeof(int32_t) * (size_t)(R + 1);
    s->smem = table <= MAX_SMEM_TABLE ? table : 0;
    int variant = 0;
    if (const char *v = getenv("SG_STENCIL_VARIANT")) variant = atoi(v);
    if (variant < 0 || variant >= N_VARIANTS) variant = 0;
    s->kernel = s->smem ? VARIANTS[variant].smem : VARIANTS[variant].global;
    int per_sm = 0;
    SG_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, s->kernel, STB, s->smem));
    long grid = (long)(per_sm > 0 ? per_sm : 1) * ctx->sm_count;
    const long need = (n_rows + STB - 1) / STB;
    if (grid > need) grid = need;
    if (grid > SG_MAX_BLOCKS) grid = SG_MAX_BLOCKS;
    s->grid = (int)grid;
    s->max_nnz = max_nnz;
    s->max_off = (long)max_off;
    s->dev.n_rows = n_rows;
    s->dev.own_lo = 0;
    s->dev.own_hi = n_rows;
    s->dev.rcls = s->rcls;
    s->dev.ptr = s->ptr;
    s->dev.ent = s->ent;
    s->dev.n_classes = R;
    s->dev.n_entries = ptr[R];
    *out = s;
    return SG_OK;
}

int sg_stencil_apply(const SgStencil *s, const double *x, double *y, int64_t own_lo, int64_t own_hi, SgRed red, double *dot2,
                     const int *skip, cudaStream_t st, const SgHaloWait *wait) {
    StDev sd = s->dev;
    sd.own_lo = own_lo;
    sd.own_hi = own_hi;
    SgHaloWait hw{};
    sd.safe_lo = 0;
    sd.safe_hi = sd.n_rows;
    if (wait && wait->n) {
        hw = *wait;
        if (own_lo > 0) sd.safe_lo = std::min<long>(sd.n_rows, own_lo + s->max_off);
        if (own_hi < sd.n_rows) sd.safe_hi = std::max<long>(sd.safe_lo, own_hi - s->max_off);
    }
    s->kernel<<<s->grid, STB, s->smem, st>>>(sd, x, y, red, dot2, skip, hw);
    SG_CHECK_CUDA(cudaGetLastError());
    sg_count_launch();
    return SG_OK;
}

// Cell part only (the attached exterior facets, if any, are left out), y = S x or - subtract - y -= S x: the two halves of
// the residual in gather form,  F_cells = S_J T - S_M T_prev  (thermal.cu).
int sg_stencil_apply_cells(const SgStencil *s, const double *x, double *y, int subtract, SgRed red, double *dot2, cudaStream_t st) {
    StDev sd = s->dev;
    sd.own_lo = 0;
    sd.own_hi = 0;                 // no dot product wanted
    sd.safe_lo = 0;
    sd.safe_hi = sd.n_rows;
    sd.bmat = nullptr;
    sd.sub = subtract ? 1 : 0;
    SgHaloWait hw{};
    s->kernel<<<s->grid, STB, s->smem, st>>>(sd, x, y, red, dot2, nullptr, hw);
    SG_CHECK_CUDA(cudaGetLastError());
    sg_count_launch();
    return SG_OK;
}

// Exterior facets in gather form (see StDev).  Returns SG_OK and leaves the stencil without a boundary part when a row lies on
// more than BND_MAXF facets (the caller then keeps its separate exterior-facet kernel).
int sg_stencil_attach_boundary(SgStencil *s, const int32_t *dofmap, int64_t n_cells, int64_t cell_lo, int64_t cell_hi, int64_t n_bf,
                               const int32_t *bf_cell, const int32_t *bf_facet, int nfd, const int *facet_dofs /* [4][6] */,
                               const double *bmat, int *attached) {
    *attached = 0;
    if (n_bf <= 0 || !bmat || nfd > 6) return SG_OK;
    StDev sd = s->dev;
    SG_CHECK_CUDA(cudaMalloc(&s->fd, sizeof(int32_t) * 24));
    SG_CHECK_CUDA(cudaMemcpy(s->fd, facet_dofs, sizeof(int32_t) * 24, cudaMemcpyHostToDevice));
    sd.fd = s->fd;
    const long n = sd.n_rows;
    DevBuf flag, cnt, ovf;
    SG_CHECK_CUDA(cudaMalloc(&flag.p, sizeof(int32_t) * (size_t)n));
    SG_CHECK_CUDA(cudaMemset(flag.p, 0, sizeof(int32_t) * (size_t)n));
    SG_CHECK_CUDA(cudaMalloc(&cnt.p, sizeof(int32_t)));
    SG_CHECK_CUDA(cudaMemset(cnt.p, 0, sizeof(int32_t)));
    SG_CHECK_CUDA(cudaMalloc(&ovf.p, sizeof(unsigned)));
    SG_CHECK_CUDA(cudaMemset(ovf.p, 0, sizeof(unsigned)));
    const unsigned gb = (unsigned)((n_bf + 255) / 256), gr = (unsigned)((n + 255) / 256);
    k_bnd_mark<<<gb, 256>>>(n_bf, bf_cell, bf_facet, dofmap, n_cells, cell_lo, cell_hi, nfd, sd, flag.as<int32_t>());
    SG_CHECK_CUDA(cudaGetLastError());
    SG_CHECK_CUDA(cudaMalloc(&s->brow_of, sizeof(int32_t) * (size_t)n));
    DevBuf pos;
    SG_CHECK_CUDA(cudaMalloc(&pos.p, sizeof(int32_t) * (size_t)n));
    SG_CHECK_CUDA(cudaMemcpy(pos.p, flag.p, sizeof(int32_t) * (size_t)n, cudaMemcpyDeviceToDevice));
    int64_t n_brows64 = 0;
    {
        const int rcs = sg_exclusive_scan_i32(pos.as<int32_t>(), n, &n_brows64);
        if (rcs) return rcs;
    }
    const int32_t n_brows = (int32_t)n_brows64;
    if (n_brows <= 0) return SG_OK;
    k_bnd_number<<<gr, 256>>>(n, flag.as<int32_t>(), pos.as<int32_t>(), s->brow_of);
    SG_CHECK_CUDA(cudaGetLastError());
    SG_CHECK_CUDA(cudaMalloc(&s->bcnt, sizeof(int32_t) * (size_t)n_brows));
    SG_CHECK_CUDA(cudaMemset(s->bcnt, 0, sizeof(int32_t) * (size_t)n_brows));
    SG_CHECK_CUDA(cudaMalloc(&s->blist, sizeof(int32_t) * (size_t)n_brows * BND_MAXF));
    k_bnd_collect<<<gb, 256>>>(n_bf, bf_cell, bf_facet, dofmap, n_cells, cell_lo, cell_hi, nfd, sd, s->brow_of, s->bcnt, s->blist,
                               ovf.as<unsigned>());
    SG_CHECK_CUDA(cudaGetLastError());
    unsigned overflow = 0;
    SG_CHECK_CUDA(cudaMemcpy(&overflow, ovf.p, sizeof(overflow), cudaMemcpyDeviceToHost));
    if (overflow) return SG_OK;
    k_bnd_sort_flag<<<(unsigned)((n_brows + 255) / 256), 256>>>(n_brows, s->bcnt, s->blist);
    SG_CHECK_CUDA(cudaGetLastError());
    SG_CHECK_CUDA(cudaMalloc(&s->bcols, sizeof(int32_t) * (size_t)n_brows * BND_MAXC));
    SG_CHECK_CUDA(cudaMalloc(&s->bncol, sizeof(int32_t) * (size_t)n_brows));
    SG_CHECK_CUDA(cudaMalloc(&s->bvals, sizeof(double) * (size_t)n_brows * BND_MAXC));
    SG_CHECK_CUDA(cudaMemset(s->bvals, 0, sizeof(double) * (size_t)n_brows * BND_MAXC));
    sd.blist = s->blist;
    sd.bcnt = s->bcnt;
    sd.bf_cell = bf_cell;
    sd.bf_facet = bf_facet;
    sd.dofmap = dofmap;
    sd.nc = (long)n_cells;
    sd.nfd = nfd;
    k_bnd_columns<<<(unsigned)((n_brows + 255) / 256), 256>>>(n_brows, sd, s->bcols, s->bncol, ovf.as<unsigned>());
    SG_CHECK_CUDA(cudaGetLastError());
    SG_CHECK_CUDA(cudaMemcpy(&overflow, ovf.p, sizeof(overflow), cudaMemcpyDeviceToHost));
    if (overflow) return SG_OK;
    k_bnd_flag_rows<<<gr, 256>>>(n, s->brow_of, s->rcls);
    SG_CHECK_CUDA(cudaGetLastError());
    SG_CHECK_CUDA(cudaDeviceSynchronize());
    s->n_brows = n_brows;
    s->dev.bcols = s->bcols;
    s->dev.bncol = s->bncol;
    s->dev.bvals = s->bvals;
    s->dev.n_brows = n_brows;
    s->dev.fd = s->fd;
    s->dev.brow_of = s->brow_of;
    s->dev.blist = s->blist;
    s->dev.bcnt = s->bcnt;
    s->dev.bf_cell = bf_cell;
    s->dev.bf_facet = bf_facet;
    s->dev.dofmap = dofmap;
    s->dev.bmat = bmat;
    s->dev.nc = (long)n_cells;
    s->dev.nfd = nfd;
    s->dev.maxf = BND_MAXF;
    *attached = 1;
    return SG_OK;
}

template <int R>
static int launch_persistent(SgStencil *s, int sm_count, PersistArgs &pa, cudaStream_t st) {
    if ((long)sm_count * PTB * R < pa.sd.n_rows) return 0;            // more rows than R per thread with one block per SM
    // window variant when the block's rows + the stencil's reach on both sides fit next to the class lists in shared memory
    const size_t tab = (s->smem + 15) & ~(size_t)15;
    // window + the per-row state kept in shared memory (x, 1/diag, boundary-row index)
    size_t win_bytes = sizeof(double) * ((size_t)R * PTB + 2 * (size_t)s->max_off) + (size_t)R * PTB * (8 + 8 + 4);
    const bool use_win = tab + win_bytes <= 200 * 1024;
    // what is left (up to 512 rows) holds the assembled boundary entries of the block's rows on exterior facets
    const size_t per_brow = (size_t)PERSIST_BS * 12 + 4;
    long bcap = use_win ? (long)((200 * 1024 - tab - win_bytes) / per_brow) : 0;
    bcap = bcap > 512 ? 512 : (bcap & ~1L);
    if (use_win) win_bytes += (size_t)bcap * per_brow;
    pa.bcap = (int)bcap;
    auto k = use_win ? (bcap > 0 ? k_cg_persistent<R, true, true> : k_cg_persistent<R, true, false>) : k_cg_persistent<R, false, false>;
    const size_t smem = use_win ? tab + win_bytes : s->smem;
    pa.win = use_win ? s->max_off : 0;
    SG_CHECK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    SG_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, PTB, smem));
    if (per_sm < 1) return 0;
    long grid = (pa.sd.n_rows + (long)R * PTB - 1) / ((long)R * PTB);
    if (grid > sm_count) grid = sm_count;
    if (!s->pcg_partials) SG_CHECK_CUDA(cudaMalloc(&s->pcg_partials, sizeof(double) * 2 * 4 * (size_t)sm_count));
    if (!s->pcg_bar) SG_CHECK_CUDA(cudaMalloc(&s->pcg_bar, sizeof(unsigned)));
    SG_CHECK_CUDA(cudaMemsetAsync(s->pcg_bar, 0, sizeof(unsigned), st));
    pa.partials = s->pcg_partials;
    pa.bar = s->pcg_bar;
    static const bool timing = [] {
        const char *e = getenv("SG_PERSIST_TIMING");
        return e && e[0] == '1';
    }();
    static long long *dbg = nullptr;
    if (timing && !dbg) {
        cudaMallocManaged(&dbg, (8 + 2 * 160) * sizeof(long long));
        memset(dbg, 0, (8 + 2 * 160) * sizeof(long long));
    }
    pa.dbg = timing ? dbg : nullptr;
    if (timing) {
        cudaStreamSynchronize(st);
        fprintf(stderr, "iterations so far %lld; ", dbg[7]);
        fprintf(stderr, "persistent PCG phases so far (cycles of block 0): barrierA %lld window %lld rows %lld blocksum %lld barrierB %lld totals %lld update %lld\n",
                dbg[0], dbg[1], dbg[2], dbg[3], dbg[4], dbg[5], dbg[6]);
        long long lo = dbg[8], hi = dbg[8], sum = 0;
        for (int i = 0; i < 2 * sm_count; ++i) {
            lo = dbg[8 + i] < lo ? dbg[8 + i] : lo;
            hi = dbg[8 + i] > hi ? dbg[8 + i] : hi;
            sum += dbg[8 + i];
        }
        fprintf(stderr, "  row phase over blocks x {first, last warp}: min %lld mean %lld max %lld; block 0: %lld %lld, block 1: %lld %lld, block 74: %lld %lld\n",
                lo, sum / (2 * sm_count), hi, dbg[8], dbg[9], dbg[10], dbg[11], dbg[8 + 148], dbg[8 + 149]);
    }
    void *args[] = {&pa};
    SG_CHECK_CUDA(cudaLaunchCooperativeKernel((const void *)k, dim3((unsigned)grid), dim3(PTB), args, smem, st));
    sg_count_launch();
    return 1;
}

int sg_stencil_pcg(SgStencil *s, int sm_count, const double *b, const double *dinv, double *x, double *work, const SgPcgPolicy &pol, int max_it,
                   double *rr0_out, int *ctrl_done, int *ctrl_iters, double *ctrl_rr, cudaStream_t st) {
    if (!s->smem) return 0;                               // class lists too large for shared memory: multi-kernel path
    PersistArgs pa;
    memset(&pa, 0, sizeof(pa));
    pa.sd = s->dev;
    pa.sd.own_lo = 0;
    pa.sd.own_hi = pa.sd.n_rows;
    pa.b = b;
    pa.dinv = dinv;
    pa.x = x;
    pa.u = work;
    pa.pol = pol;
    pa.max_it = max_it;
    pa.rr0_out = rr0_out;
    pa.ctrl_done = ctrl_done;
    pa.ctrl_iters = ctrl_iters;
    pa.ctrl_rr = ctrl_rr;
    // smallest rows-per-thread count that covers n with one 512-thread block per SM (B200: 75 776 threads)
    const long n = pa.sd.n_rows, per = (long)sm_count * PTB;
    if (n <= per * 1) return launch_persistent<1>(s, sm_count, pa, st);
    if (n <= per * 2) return launch_persistent<2>(s, sm_count, pa, st);
    if (n <= per * 3) return launch_persistent<3>(s, sm_count, pa, st);
    if (n <= per * 4) return launch_persistent<4>(s, sm_count, pa, st);
    if (n <= per * 5) return launch_persistent<5>(s, sm_count, pa, st);
    if (n <= per * 6) return launch_persistent<6>(s, sm_count, pa, st);
    if (n <= per * 8) return launch_persistent<8>(s, sm_count, pa, st);
    return 0;
}

// after sg_thermal_linearize has refreshed bmat: re-assemble the boundary rows' entries
int sg_stencil_refresh_boundary(const SgStencil *s, cudaStream_t st) {
    if (!s->dev.bmat || s->n_brows <= 0) return SG_OK;
    k_bnd_refresh<<<(unsigned)((s->n_brows + 127) / 128), 128, 0, st>>>(s->n_brows, s->dev, s->bvals);
    SG_CHECK_CUDA(cudaGetLastError());
    sg_count_launch();
    return SG_OK;
}

void sg_stencil_info(const SgStencil *s, int32_t *n_classes, int32_t *n_entries, int32_t *max_nnz) {
    if (n_classes) *n_classes = s->dev.n_classes;
    if (n_entries) *n_entries = s->dev.n_entries;
    if (max_nnz) *max_nnz = s->max_nnz;
}
