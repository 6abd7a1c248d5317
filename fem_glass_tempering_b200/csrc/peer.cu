// Peer-memory (NVLink / NVSwitch) halo exchange and small all-reduce of the multi-GPU solver.
//
// One process per GPU.  Every rank cudaMalloc's one communication block — flags, all-reduce slots, mailboxes AND the
// solver's vector workspace — exports it with cudaIpcGetMemHandle, and maps every other rank's block
// (cudaIpcOpenMemHandle); the handles travel through the host (torch.distributed).  After that the data path never
// leaves the GPUs and needs no NCCL launch:
//   halo, direct   k_halo_put stores this rank's boundary rows of a workspace vector STRAIGHT INTO THE NEIGHBOUR'S GHOST
//                  ROWS of the same vector over NVLink, fences, and raises a sequence flag in the neighbour's memory.
//                  There is no receive-side copy: the consuming kernel (operator apply / Chebyshev step) processes the
//                  cells that read no ghost value first and waits for the flag only before its two boundary strips
//                  (SgHaloWait, sg_halo_wait_block), so the transfer and the rank-to-rank skew hide behind the interior.
//                  No double buffering is needed: between two puts into the same ghost rows the sender has consumed a
//                  put or an all-reduce that the receiver issued AFTER its last read of those rows (the solver alternates
//                  p, zA, zB and reduces in between), and every vector kernel of a partitioned solve writes owned rows only.
//   halo, mailbox  (vectors outside the workspace, e.g. the caller's temperature, a few times per time step)
//                  k_halo_push -> neighbour's mailbox + flag, k_halo_pull waits and fills the ghost rows; mailboxes and
//                  flags double-buffered by sequence parity.
//   all-reduce     runs INSIDE the reducing kernel: the last block of sg_grid_reduce (sg_common.cuh) writes <= 4 doubles +
//                  a tag into slot [rank] of EVERY rank's block, waits for all tags and sums the slots in rank order:
//                  deterministic, bit-identical on all ranks, and no separate launch.  k_allreduce_small remains for callers
//                  outside a reducing kernel.
// Every wait is bounded (about 20 s): on expiry the kernel raises a flag in mapped host memory and returns, and the host
// reports SG_E_NCCL instead of hanging the GPU.
#include "sg_common.cuh"

namespace {

constexpr int PEER_MAX_RANKS = SG_PEER_MAX_RANKS;
constexpr int RED_MAX_VALS = SG_PEER_RED_VALS;

struct PeerLayout {
    // byte offsets inside a rank's communication block
    size_t flags;     // [2 sides][2 parities] unsigned long long   (mailbox path)
    size_t dflags;    // [2 sides] unsigned long long               (direct path: monotonic sequence per side)
    size_t red_tags;  // [2 parities][PEER_MAX_RANKS] unsigned long long
    size_t red_vals;  // [2 parities][PEER_MAX_RANKS][RED_MAX_VALS] double
    size_t mailbox;   // [2 sides][2 parities][mailbox_doubles] double
    size_t workspace; // [workspace_doubles] double
    size_t total;
};

PeerLayout make_layout(size_t mailbox_doubles, size_t workspace_doubles) {
    PeerLayout L;
    L.flags = 0;
    L.dflags = 64;
    L.red_tags = 256;
    L.red_vals = L.red_tags + sizeof(unsigned long long) * 2 * PEER_MAX_RANKS;
    L.mailbox = (L.red_vals + sizeof(double) * 2 * PEER_MAX_RANKS * RED_MAX_VALS + 255) & ~(size_t)255;
    L.workspace = (L.mailbox + sizeof(double) * 4 * mailbox_doubles + 255) & ~(size_t)255;
    L.total = L.workspace + sizeof(double) * workspace_doubles;
    return L;
}

struct PushSeg {
    const double *src;            // first row to send (local)
    double *dst;                  // neighbour's mailbox [side seen by the neighbour][parity] / ghost rows of its vector
    unsigned long long *flag;     // neighbour's flag
    long count;                   // doubles
};
struct PushArgs {
    PushSeg seg[2];
    unsigned *counters;           // [2] local, zero between launches
    unsigned long long seq;
};

// also the direct put: only dst/flag differ.  128-thread blocks of <= 32 registers: 4096 registers is what the persistent
// operator kernels (3 blocks x 256 threads x 80 registers) leave free on an SM, so a put launched on the high-priority
// side stream runs NEXT TO the consumer kernel instead of waiting for its blocks to retire.
constexpr int PTB = 128;
__global__ void __launch_bounds__(PTB, 16) k_halo_push(const PushArgs a) {
    const PushSeg s = blockIdx.y == 0 ? a.seg[0] : a.seg[1];   // no dynamic indexing of the parameter struct (no local copy)
    if (s.count <= 0) return;
    if ((((uintptr_t)s.src | (uintptr_t)s.dst) & 15) == 0) {
        const long n2 = s.count >> 1;
        const double2 *src2 = reinterpret_cast<const double2 *>(s.src);
        double2 *dst2 = reinterpret_cast<double2 *>(s.dst);
        for (long i = (long)blockIdx.x * PTB + threadIdx.x; i < n2; i += (long)gridDim.x * PTB) dst2[i] = src2[i];
        if ((s.count & 1) && blockIdx.x == 0 && threadIdx.x == 0) s.dst[s.count - 1] = s.src[s.count - 1];
    } else {   // ranges that start at an odd dof (small CG planes)
        for (long i = (long)blockIdx.x * PTB + threadIdx.x; i < s.count; i += (long)gridDim.x * PTB) s.dst[i] = s.src[i];
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

__global__ void __launch_bounds__(256) k_halo_pull(const PullArgs a) {
    const PullSeg s = a.seg[blockIdx.y];
    if (s.count <= 0) return;
    __shared__ int ok;
    if (threadIdx.x == 0) {
        ok = sg_spin_until(s.flag, a.seq) ? 1 : 0;
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

__global__ void __launch_bounds__(32) k_halo_wait(const SgHaloWait w) { sg_halo_wait_block(w); }

__global__ void __launch_bounds__(32) k_allreduce_small(const SgPeerRedDev *a, double *vals, int count) {
    sg_peer_allreduce_warp(*a, vals, count);
}

}  // namespace

struct SgPeer {
    sg_ctx *ctx;
    PeerLayout lay;
    size_t mailbox_doubles, workspace_doubles;
    char *local;
    char *remote[PEER_MAX_RANKS];
    bool opened;
    unsigned *counters;         // device [2]
    int *err_host;              // mapped pinned
    unsigned long long halo_seq, direct_seq;
    unsigned long long *red_seq_dev;   // device counter of the in-kernel all-reduces
    SgPeerRedDev *red_dev;      // device copy of the all-reduce descriptor
    int64_t stride[PEER_MAX_RANKS], ghost_below[PEER_MAX_RANKS], ghost_above[PEER_MAX_RANKS];
};

int sg_peer_create(sg_ctx *ctx, size_t mailbox_doubles, size_t workspace_doubles, SgPeer **out, void *handle64) {
    SG_REQUIRE(ctx && out && handle64, "sg_peer_create: NULL argument");
    SG_REQUIRE(ctx->nranks <= PEER_MAX_RANKS, "sg_peer_create: at most %d ranks", PEER_MAX_RANKS);
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    SgPeer *p = new SgPeer();
    memset(p, 0, sizeof(*p));
    p->ctx = ctx;
    p->mailbox_doubles = (mailbox_doubles + 1) & ~(size_t)1;
    p->workspace_doubles = workspace_doubles;
    p->lay = make_layout(p->mailbox_doubles, workspace_doubles);
    cudaError_t e = cudaMalloc(&p->local, p->lay.total);
    if (e == cudaSuccess) e = cudaMemset(p->local, 0, p->lay.total);
    if (e == cudaSuccess) e = cudaMalloc(&p->counters, 2 * sizeof(unsigned));
    if (e == cudaSuccess) e = cudaMemset(p->counters, 0, 2 * sizeof(unsigned));
    if (e == cudaSuccess) e = cudaMalloc(&p->red_seq_dev, sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMemset(p->red_seq_dev, 0, sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMalloc(&p->red_dev, sizeof(SgPeerRedDev));
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

int sg_peer_open(SgPeer *p, const void *handles, const int64_t *layout3) {
    SG_REQUIRE(p && handles && layout3, "sg_peer_open: NULL argument");
    for (int r = 0; r < p->ctx->nranks; ++r) {
        p->stride[r] = layout3[3 * r];
        p->ghost_below[r] = layout3[3 * r + 1];
        p->ghost_above[r] = layout3[3 * r + 2];
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
    SgPeerRedDev rd;
    memset(&rd, 0, sizeof(rd));
    for (int r = 0; r < p->ctx->nranks; ++r) rd.base[r] = p->remote[r];
    rd.tags_off = p->lay.red_tags;
    rd.vals_off = p->lay.red_vals;
    rd.rank = p->ctx->rank;
    rd.nranks = p->ctx->nranks;
    rd.seq = p->red_seq_dev;
    rd.err = p->err_host;
    SG_CHECK_CUDA(cudaMemcpy(p->red_dev, &rd, sizeof(rd), cudaMemcpyHostToDevice));
    p->opened = true;
    return SG_OK;
}

int sg_peer_destroy(SgPeer *p) {
    if (!p) return SG_OK;
    for (int r = 0; r < PEER_MAX_RANKS; ++r)
        if (p->remote[r] && p->remote[r] != p->local) cudaIpcCloseMemHandle(p->remote[r]);
    if (p->local) cudaFree(p->local);
    if (p->counters) cudaFree(p->counters);
    if (p->red_seq_dev) cudaFree(p->red_seq_dev);
    if (p->red_dev) cudaFree(p->red_dev);
    if (p->err_host) cudaFreeHost(p->err_host);
    delete p;
    return SG_OK;
}

bool sg_peer_ready(const SgPeer *p) { return p && p->opened; }
size_t sg_peer_mailbox_doubles(const SgPeer *p) { return p ? p->mailbox_doubles : 0; }
double *sg_peer_workspace(const SgPeer *p) {
    return p && p->workspace_doubles ? reinterpret_cast<double *>(p->local + p->lay.workspace) : nullptr;
}
const SgPeerRedDev *sg_peer_red_dev(const SgPeer *p) { return sg_peer_ready(p) ? p->red_dev : nullptr; }

int sg_peer_check(SgPeer *p) {
    if (p && *p->err_host) {
        sg_set_error("peer-memory exchange timed out waiting for a neighbour (rank %d)", p->ctx->rank);
        return SG_E_NCCL;
    }
    return SG_OK;
}

// segments: at most one neighbour below and one above this rank (x-slab partition)
static dim3 halo_grid(int n_seg, const sg_halo_segment *seg, int push_scale = 1) {
    long maxcount = 0;
    for (int i = 0; i < n_seg && i < 2; ++i) {
        if (seg[i].send_count > maxcount) maxcount = seg[i].send_count;
        if (seg[i].recv_count > maxcount) maxcount = seg[i].recv_count;
    }
    long gx = (maxcount / 2 + 255) / 256;
    if (gx < 1) gx = 1;
    if (gx > 64) gx = 64;
    gx *= push_scale;
    return dim3((unsigned)gx, (unsigned)(n_seg < 2 ? (n_seg < 1 ? 1 : n_seg) : 2));
}

static int sg_peer_halo_push(SgPeer *p, int n_seg, const sg_halo_segment *seg, const double *vec, cudaStream_t st) {
    const int rank = p->ctx->rank;
    for (int i = 0; i < n_seg && i < 2; ++i)   // validate BEFORE the sequence number moves: a rejected call must not desynchronise
        SG_REQUIRE((size_t)seg[i].send_count <= p->mailbox_doubles && (size_t)seg[i].recv_count <= p->mailbox_doubles,
                   "sg_peer_halo: segment larger than the mailbox");
    const unsigned long long seq = ++p->halo_seq;
    const int par = (int)(seq & 1ull);
    PushArgs pa;
    memset(&pa, 0, sizeof(pa));
    for (int i = 0; i < n_seg && i < 2; ++i) {
        const sg_halo_segment &g = seg[i];
        const int side_there = rank < g.peer ? 0 : 1;   // how the neighbour sees me: I am below it -> its side 0
        char *rb = p->remote[g.peer];
        pa.seg[i].src = vec + g.send_offset;
        pa.seg[i].dst = reinterpret_cast<double *>(rb + p->lay.mailbox) + ((size_t)side_there * 2 + par) * p->mailbox_doubles;
        pa.seg[i].flag = reinterpret_cast<unsigned long long *>(rb + p->lay.flags) + side_there * 2 + par;
        pa.seg[i].count = g.send_count;
    }
    pa.counters = p->counters;
    pa.seq = seq;
    k_halo_push<<<halo_grid(n_seg, seg, 256 / PTB), PTB, 0, st>>>(pa);
    SG_CHECK_CUDA(cudaGetLastError());
    sg_count_launch();
    return SG_OK;
}

// completes the exchange started by the LAST sg_peer_halo_push
static int sg_peer_halo_pull(SgPeer *p, int n_seg, const sg_halo_segment *seg, double *vec, cudaStream_t st) {
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

bool sg_peer_in_workspace(const SgPeer *p, const double *vec) {
    const double *ws = sg_peer_workspace(p);
    const int64_t stride = p ? p->stride[p->ctx->rank] : 0;
    return ws && stride > 0 && vec >= ws && vec < ws + p->workspace_doubles && ((vec - ws) % stride) == 0;
}

int sg_peer_put(SgPeer *p, int n_seg, const sg_halo_segment *seg, double *vec, SgHaloWait *wait, cudaStream_t st) {
    memset(wait, 0, sizeof(*wait));
    const int rank = p->ctx->rank;
    double *ws = sg_peer_workspace(p);
    const int64_t stride = p->stride[rank];
    const bool inside = sg_peer_in_workspace(p, vec);
    if (!inside) return sg_peer_halo_forward(p, n_seg, seg, vec, st);     // mailbox path: complete on return (stream order)
    const int64_t slot = (vec - ws) / stride;
    const unsigned long long seq = ++p->direct_seq;
    PushArgs pa;
    memset(&pa, 0, sizeof(pa));
    for (int i = 0; i < n_seg && i < 2; ++i) {
        const sg_halo_segment &g = seg[i];
        const bool peer_above = g.peer > rank;
        const int side_there = peer_above ? 0 : 1;      // the neighbour above me receives from below (its side 0)
        const int side_here = peer_above ? 1 : 0;
        char *rb = p->remote[g.peer];
        const int64_t ghost = peer_above ? p->ghost_below[g.peer] : p->ghost_above[g.peer];
        pa.seg[i].src = vec + g.send_offset;
        pa.seg[i].dst = reinterpret_cast<double *>(rb + p->lay.workspace) + slot * p->stride[g.peer] + ghost;
        pa.seg[i].flag = reinterpret_cast<unsigned long long *>(rb + p->lay.dflags) + side_there;
        pa.seg[i].count = g.send_count;
        if (g.recv_count > 0) {
            wait->flag[wait->n] = reinterpret_cast<const unsigned long long *>(p->local + p->lay.dflags) + side_here;
            wait->n++;
        }
    }
    wait->seq = seq;
    wait->err = p->err_host;
    pa.counters = p->counters;
    pa.seq = seq;
    k_halo_push<<<halo_grid(n_seg, seg, 256 / PTB), PTB, 0, st>>>(pa);
    SG_CHECK_CUDA(cudaGetLastError());
    sg_count_launch();
    return SG_OK;
}

int sg_peer_wait(SgPeer *p, const SgHaloWait &wait, cudaStream_t st) {
    (void)p;
    if (wait.n == 0) return SG_OK;
    k_halo_wait<<<1, 32, 0, st>>>(wait);
    SG_CHECK_CUDA(cudaGetLastError());
    sg_count_launch();
    return SG_OK;
}

int sg_peer_allreduce(SgPeer *p, double *vals, int count, cudaStream_t st) {
    SG_REQUIRE(count >= 1 && count <= RED_MAX_VALS, "sg_peer_allreduce: 1..%d values", RED_MAX_VALS);
    k_allreduce_small<<<1, 32, 0, st>>>(p->red_dev, vals, count);
    SG_CHECK_CUDA(cudaGetLastError());
    sg_count_launch();
    return SG_OK;
}
