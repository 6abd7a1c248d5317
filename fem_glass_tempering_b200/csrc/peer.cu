// Peer-memory (NVLink / NVSwitch) halo exchange and small all-reduce of the multi-GPU solver.
//
// One process per GPU.  Every rank cudaMalloc's one communication block, exports it with
// cudaIpcGetMemHandle, and maps every other rank's block (cudaIpcOpenMemHandle); the handles travel through
// the host (torch.distributed).  After that the data path never leaves the GPUs and needs no NCCL launch:
//   halo forward   k_halo_push  stores this rank's boundary rows straight into the neighbour's mailbox over
//                  NVLink, fences, and bumps a sequence flag in the neighbour's memory;
//                  k_halo_pull  waits for the flags addressed to this rank and copies its mailbox into the
//                  ghost rows.  Mailboxes and flags are double-buffered by sequence parity.
//   all-reduce     k_allreduce_small  writes <= 4 doubles + a sequence tag into slot [my rank] of EVERY rank's
//                  block, waits until all tags of this round have arrived, and sums the slots in rank order:
//                  deterministic and bit-identical on all ranks.
// A PCG iteration needs 4 halos and 3 all-reduces of 1-2 doubles; each costs a few microseconds here instead of
// a 30-45 us NCCL operation.  Every wait is bounded (about 20 s): on expiry the kernel raises a flag in mapped host
// memory and returns, and the host reports SG_E_NCCL instead of hanging the GPU.
#include "sg_common.cuh"

namespace {

constexpr int PEER_MAX_RANKS = 16;
constexpr int RED_MAX_VALS = 4;
constexpr long long SPIN_LIMIT = 40000000000ll;   // clock64 ticks, about 20 s: ranks may arrive seconds apart after set-up

struct PeerLayout {
    // byte offsets inside a rank's communication block
    size_t flags;     // [2 sides][2 parities] unsigned long long
    size_t red_tags;  // [2 parities][PEER_MAX_RANKS] unsigned long long
    size_t red_vals;  // [2 parities][PEER_MAX_RANKS][RED_MAX_VALS] double
    size_t mailbox;   // [2 sides][2 parities][mailbox_doubles] double
    size_t total;
};

PeerLayout make_layout(size_t mailbox_doubles) {
    PeerLayout L;
    L.flags = 0;
    L.red_tags = 256;
    L.red_vals = L.red_tags + sizeof(unsigned long long) * 2 * PEER_MAX_RANKS;
    L.mailbox = (L.red_vals + sizeof(double) * 2 * PEER_MAX_RANKS * RED_MAX_VALS + 255) & ~(size_t)255;
    L.total = L.mailbox + sizeof(double) * 4 * mailbox_doubles;
    return L;
}

struct PushSeg {
    const double *src;            // first row to send (local)
    double *dst;                  // neighbour's mailbox [side seen by the neighbour][parity]
    unsigned long long *flag;     // neighbour's flag of that mailbox
    long count;                   // doubles
};
struct PushArgs {
    PushSeg seg[2];
    unsigned *counters;           // [2] local, zero between launches
    unsigned long long seq;
};

__global__ void __launch_bounds__(256) k_halo_push(const PushArgs a) {
    const PushSeg s = a.seg[blockIdx.y];
    if (s.count <= 0) return;
    if ((((uintptr_t)s.src | (uintptr_t)s.dst) & 15) == 0) {
        const long n2 = s.count >> 1;
        const double2 *src2 = reinterpret_cast<const double2 *>(s.src);
        double2 *dst2 = reinterpret_cast<double2 *>(s.dst);
        for (long i = (long)blockIdx.x * 256 + threadIdx.x; i < n2; i += (long)gridDim.x * 256) dst2[i] = src2[i];
        if ((s.count & 1) && blockIdx.x == 0 && threadIdx.x == 0) s.dst[s.count - 1] = s.src[s.count - 1];
    } else {   // ranges that start at an odd dof (small CG planes)
        for (long i = (long)blockIdx.x * 256 + threadIdx.x; i < s.count; i += (long)gridDim.x * 256) s.dst[i] = s.src[i];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned done = atomicAdd(&a.counters[blockIdx.y], 1u);
        if (done == gridDim.x - 1) {            // every block's stores are fenced: publish
            a.counters[blockIdx.y] = 0u;
            __threadfence_system();
            *reinterpret_cast<volatile unsigned long long *>(s.flag) = a.seq;
        }
    }
}

struct PullSeg {
    double *dst;                          // ghost rows (local)
    const double *src;                    // my mailbox [side][parity]
    const unsigned long long *flag;       // my flag of that mailbox
    long count;
};
struct PullArgs {
    PullSeg seg[2];
    unsigned long long seq;
    int *err;                             // mapped host memory
};

__device__ __forceinline__ bool spin_until(const unsigned long long *flag, unsigned long long seq) {
    const volatile unsigned long long *f = flag;
    const long long t0 = clock64();
    while (*f < seq) {
        if (clock64() - t0 > SPIN_LIMIT) return false;
        __nanosleep(64);
    }
    return true;
}

__global__ void __launch_bounds__(256) k_halo_pull(const PullArgs a) {
    const PullSeg s = a.seg[blockIdx.y];
    if (s.count <= 0) return;
    __shared__ int ok;
    if (threadIdx.x == 0) {
        ok = spin_until(s.flag, a.seq) ? 1 : 0;
        if (!ok) *a.err = 1;
        __threadfence_system();
    }
    __syncthreads();
    if (!ok) return;
    if ((((uintptr_t)s.src | (uintptr_t)s.dst) & 15) == 0) {
        const long n2 = s.count >> 1;
        const double2 *src2 = reinterpret_cast<const double2 *>(s.src);
        double2 *dst2 = reinterpret_cast<double2 *>(s.dst);
        for (long i = (long)blockIdx.x * 256 + threadIdx.x; i < n2; i += (long)gridDim.x * 256) {
            double2 v;
            asm volatile("ld.volatile.global.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(src2 + i));   // written by a peer
            dst2[i] = v;
        }
        if ((s.count & 1) && blockIdx.x == 0 && threadIdx.x == 0)
            s.dst[s.count - 1] = *reinterpret_cast<const volatile double *>(s.src + s.count - 1);
    } else {
        for (long i = (long)blockIdx.x * 256 + threadIdx.x; i < s.count; i += (long)gridDim.x * 256)
            s.dst[i] = *reinterpret_cast<const volatile double *>(s.src + i);
    }
}

struct RedArgs {
    char *base[PEER_MAX_RANKS];   // every rank's block (base[rank] is local)
    size_t tags_off, vals_off;
    int rank, nranks, count;
    unsigned long long seq;
    double *vals;                 // in/out, local device memory
    int *err;
};

__global__ void __launch_bounds__(32) k_allreduce_small(const RedArgs a) {
    const int t = threadIdx.x, par = (int)(a.seq & 1ull);
    if (t < a.nranks) {
        double *v = reinterpret_cast<double *>(a.base[t] + a.vals_off) + ((size_t)par * PEER_MAX_RANKS + a.rank) * RED_MAX_VALS;
        for (int k = 0; k < a.count; ++k) reinterpret_cast<volatile double *>(v)[k] = a.vals[k];
        __threadfence_system();
        unsigned long long *tag = reinterpret_cast<unsigned long long *>(a.base[t] + a.tags_off) + (size_t)par * PEER_MAX_RANKS + a.rank;
        *reinterpret_cast<volatile unsigned long long *>(tag) = a.seq;
    }
    bool ok = true;
    if (t < a.nranks) {
        const unsigned long long *tag = reinterpret_cast<const unsigned long long *>(a.base[a.rank] + a.tags_off) + (size_t)par * PEER_MAX_RANKS + t;
        ok = spin_until(tag, a.seq);
    }
    ok = __all_sync(0xffffffffu, ok);
    __threadfence_system();
    if (t == 0) {
        if (!ok) {
            *a.err = 1;
            __threadfence_system();
            return;
        }
        const volatile double *v = reinterpret_cast<const volatile double *>(a.base[a.rank] + a.vals_off) + (size_t)par * PEER_MAX_RANKS * RED_MAX_VALS;
        for (int k = 0; k < a.count; ++k) {
            double s = 0.0;
            for (int r = 0; r < a.nranks; ++r) s += v[(size_t)r * RED_MAX_VALS + k];   // fixed rank order on every rank
            a.vals[k] = s;
        }
    }
}

}  // namespace

struct SgPeer {
    sg_ctx *ctx;
    PeerLayout lay;
    size_t mailbox_doubles;
    char *local;
    char *remote[PEER_MAX_RANKS];
    bool opened;
    unsigned *counters;         // device [2]
    int *err_host;              // mapped pinned
    unsigned long long halo_seq, red_seq;
};

int sg_peer_create(sg_ctx *ctx, size_t mailbox_doubles, SgPeer **out, void *handle64) {
    SG_REQUIRE(ctx && out && handle64, "sg_peer_create: NULL argument");
    SG_REQUIRE(ctx->nranks <= PEER_MAX_RANKS, "sg_peer_create: at most %d ranks", PEER_MAX_RANKS);
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    SgPeer *p = new SgPeer();
    memset(p, 0, sizeof(*p));
    p->ctx = ctx;
    p->mailbox_doubles = (mailbox_doubles + 1) & ~(size_t)1;
    p->lay = make_layout(p->mailbox_doubles);
    cudaError_t e = cudaMalloc(&p->local, p->lay.total);
    if (e == cudaSuccess) e = cudaMemset(p->local, 0, p->lay.total);
    if (e == cudaSuccess) e = cudaMalloc(&p->counters, 2 * sizeof(unsigned));
    if (e == cudaSuccess) e = cudaMemset(p->counters, 0, 2 * sizeof(unsigned));
    if (e == cudaSuccess) e = cudaHostAlloc(&p->err_host, sizeof(int), cudaHostAllocMapped);
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) {
        *p->err_host = 0;
        e = cudaIpcGetMemHandle(&h, p->local);
    }
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        sg_set_error("sg_peer_create: %s", cudaGetErrorString(e));
        sg_peer_destroy(p);
        return SG_E_CUDA;
    }
    memcpy(handle64, &h, 64);
    p->remote[ctx->rank] = p->local;
    *out = p;
    return SG_OK;
}

int sg_peer_open(SgPeer *p, const void *handles) {
    SG_REQUIRE(p && handles, "sg_peer_open: NULL argument");
    for (int r = 0; r < p->ctx->nranks; ++r) {
        if (r == p->ctx->rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, (const char *)handles + 64 * (size_t)r, 64);
        void *ptr = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            sg_set_error("sg_peer_open: cudaIpcOpenMemHandle(rank %d) -> %s", r, cudaGetErrorString(e));
            cudaGetLastError();
            return SG_E_CUDA;
        }
        p->remote[r] = (char *)ptr;
    }
    p->opened = true;
    return SG_OK;
}

int sg_peer_destroy(SgPeer *p) {
    if (!p) return SG_OK;
    for (int r = 0; r < PEER_MAX_RANKS; ++r)
        if (p->remote[r] && p->remote[r] != p->local) cudaIpcCloseMemHandle(p->remote[r]);
    if (p->local) cudaFree(p->local);
    if (p->counters) cudaFree(p->counters);
    if (p->err_host) cudaFreeHost(p->err_host);
    delete p;
    return SG_OK;
}

bool sg_peer_ready(const SgPeer *p) { return p && p->opened; }
size_t sg_peer_mailbox_doubles(const SgPeer *p) { return p ? p->mailbox_doubles : 0; }

int sg_peer_check(SgPeer *p) {
    if (p && *p->err_host) {
        sg_set_error("peer-memory exchange timed out waiting for a neighbour (rank %d)", p->ctx->rank);
        return SG_E_NCCL;
    }
    return SG_OK;
}

// segments: at most one neighbour below and one above this rank (x-slab partition)
static dim3 halo_grid(int n_seg, const sg_halo_segment *seg) {
    long maxcount = 0;
    for (int i = 0; i < n_seg && i < 2; ++i) {
        if (seg[i].send_count > maxcount) maxcount = seg[i].send_count;
        if (seg[i].recv_count > maxcount) maxcount = seg[i].recv_count;
    }
    long gx = (maxcount / 2 + 255) / 256;
    if (gx < 1) gx = 1;
    if (gx > 64) gx = 64;
    return dim3((unsigned)gx, (unsigned)(n_seg < 2 ? (n_seg < 1 ? 1 : n_seg) : 2));
}

int sg_peer_halo_push(SgPeer *p, int n_seg, const sg_halo_segment *seg, const double *vec, cudaStream_t st) {
    const int rank = p->ctx->rank;
    const unsigned long long seq = ++p->halo_seq;
    const int par = (int)(seq & 1ull);
    PushArgs pa;
    memset(&pa, 0, sizeof(pa));
    for (int i = 0; i < n_seg && i < 2; ++i) {
        const sg_halo_segment &g = seg[i];
        const int side_there = rank < g.peer ? 0 : 1;   // how the neighbour sees me: I am below it -> its side 0
        SG_REQUIRE((size_t)g.send_count <= p->mailbox_doubles && (size_t)g.recv_count <= p->mailbox_doubles,
                   "sg_peer_halo: segment larger than the mailbox");
        char *rb = p->remote[g.peer];
        pa.seg[i].src = vec + g.send_offset;
        pa.seg[i].dst = reinterpret_cast<double *>(rb + p->lay.mailbox) + ((size_t)side_there * 2 + par) * p->mailbox_doubles;
        pa.seg[i].flag = reinterpret_cast<unsigned long long *>(rb + p->lay.flags) + side_there * 2 + par;
        pa.seg[i].count = g.send_count;
    }
    pa.counters = p->counters;
    pa.seq = seq;
    k_halo_push<<<halo_grid(n_seg, seg), 256, 0, st>>>(pa);
    SG_CHECK_CUDA(cudaGetLastError());
    sg_count_launch();
    return SG_OK;
}

// completes the exchange started by the LAST sg_peer_halo_push
int sg_peer_halo_pull(SgPeer *p, int n_seg, const sg_halo_segment *seg, double *vec, cudaStream_t st) {
    const int rank = p->ctx->rank;
    const unsigned long long seq = p->halo_seq;
    const int par = (int)(seq & 1ull);
    PullArgs pl;
    memset(&pl, 0, sizeof(pl));
    for (int i = 0; i < n_seg && i < 2; ++i) {
        const sg_halo_segment &g = seg[i];
        const int side_here = g.peer < rank ? 0 : 1;
        pl.seg[i].dst = vec + g.recv_offset;
        pl.seg[i].src = reinterpret_cast<const double *>(p->local + p->lay.mailbox) + ((size_t)side_here * 2 + par) * p->mailbox_doubles;
        pl.seg[i].flag = reinterpret_cast<const unsigned long long *>(p->local + p->lay.flags) + side_here * 2 + par;
        pl.seg[i].count = g.recv_count;
    }
    pl.seq = seq;
    pl.err = p->err_host;
    k_halo_pull<<<halo_grid(n_seg, seg), 256, 0, st>>>(pl);
    SG_CHECK_CUDA(cudaGetLastError());
    sg_count_launch();
    return SG_OK;
}

int sg_peer_halo_forward(SgPeer *p, int n_seg, const sg_halo_segment *seg, double *vec, cudaStream_t st) {
    const int rc = sg_peer_halo_push(p, n_seg, seg, vec, st);
    return rc ? rc : sg_peer_halo_pull(p, n_seg, seg, vec, st);
}

int sg_peer_allreduce(SgPeer *p, double *vals, int count, cudaStream_t st) {
    SG_REQUIRE(count >= 1 && count <= RED_MAX_VALS, "sg_peer_allreduce: 1..%d values", RED_MAX_VALS);
    RedArgs a;
    memset(&a, 0, sizeof(a));
    for (int r = 0; r < p->ctx->nranks; ++r) a.base[r] = p->remote[r];
    a.tags_off = p->lay.red_tags;
    a.vals_off = p->lay.red_vals;
    a.rank = p->ctx->rank;
    a.nranks = p->ctx->nranks;
    a.count = count;
    a.seq = ++p->red_seq;
    a.vals = vals;
    a.err = p->err_host;
    k_allreduce_small<<<1, 32, 0, st>>>(a);
    SG_CHECK_CUDA(cudaGetLastError());
    sg_count_launch();
    return SG_OK;
}
