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
constexpr int MAX_CLASSES = 8192;
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
};

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
            sg_halo_wait_block(hw);
            waited = true;
        }
        const long idx = base + threadIdx.x;
        if (idx >= sd.n_rows) continue;
        const long j = idx - n_safe;
        const long row = idx < n_safe ? sd.safe_lo + idx : (j < sd.safe_lo ? j : sd.safe_hi + (j - sd.safe_lo));
        const int c = sd.rcls[row];
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
        y[row] = acc;
        if (row >= sd.own_lo && row < sd.own_hi) dsum[0] += __ldg(xr) * acc;
    }
    sg_grid_reduce<2>(dsum, red, dot_out);
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
};

void sg_stencil_destroy(SgStencil *s) {
    if (!s) return;
    if (s->rcls) cudaFree(s->rcls);
    if (s->ptr) cudaFree(s->ptr);
    if (s->ent) cudaFree(s->ent);
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

void sg_stencil_info(const SgStencil *s, int32_t *n_classes, int32_t *n_entries, int32_t *max_nnz) {
    if (n_classes) *n_classes = s->dev.n_classes;
    if (n_entries) *n_entries = s->dev.n_entries;
    if (max_nnz) *max_nnz = s->max_nnz;
}
