// Context, error reporting and NCCL binding of libsurroglas_b200.
#include <dlfcn.h>
#include <stdarg.h>
#include <stdlib.h>

#include "sg_common.cuh"
#include "sg_nccl.h"

#include <atomic>

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void sg_count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int sg_next_sweep_dir() {
    static const bool on = [] {
        const char *e = getenv("SG_SERPENTINE");
        return !(e && e[0] == '0');
    }();
    static std::atomic<unsigned> n{0};
    return on ? (int)(n.fetch_add(1u, std::memory_order_relaxed) & 1u) : 0;
}

void sg_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

const SgNccl *sg_nccl() {
    static SgNccl tab = {};
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) {
            sg_set_error("dlopen(libnccl.so.2) failed: %s", dlerror());
            return nullptr;
        }
#define SG_SYM(field, name)                                    \
    tab.field = (decltype(tab.field))dlsym(h, name);           \
    if (!tab.field) {                                          \
        sg_set_error("libnccl.so.2 lacks symbol %s", name);    \
        return nullptr;                                        \
    }
        SG_SYM(GetUniqueId, "ncclGetUniqueId")
        SG_SYM(CommInitRank, "ncclCommInitRank")
        SG_SYM(CommDestroy, "ncclCommDestroy")
        SG_SYM(GetErrorString, "ncclGetErrorString")
        SG_SYM(AllReduce, "ncclAllReduce")
        SG_SYM(Send, "ncclSend")
        SG_SYM(Recv, "ncclRecv")
        SG_SYM(GroupStart, "ncclGroupStart")
        SG_SYM(GroupEnd, "ncclGroupEnd")
#undef SG_SYM
        tab.ok = true;
    }
    return tab.ok ? &tab : nullptr;
}

extern "C" {

int sg_version(void) { return 100; }

int64_t sg_launch_count(void) { return (int64_t)g_launches.load(); }

const char *sg_last_error(void) { return g_err; }

int sg_nccl_unique_id(void *out128) {
    SG_REQUIRE(out128, "sg_nccl_unique_id: NULL output");
    const SgNccl *n = sg_nccl();
    if (!n) return SG_E_NCCL;
    ncclUniqueId id;
    SG_CHECK_NCCL(n->GetUniqueId(&id));
    memcpy(out128, &id, sizeof(id));
    return SG_OK;
}

int sg_ctx_create(int device, int rank, int nranks, const void *nccl_unique_id, sg_ctx **out) {
    SG_REQUIRE(out, "sg_ctx_create: NULL output");
    SG_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, "sg_ctx_create: bad rank %d of %d", rank, nranks);
    int ndev = 0;
    SG_CHECK_CUDA(cudaGetDeviceCount(&ndev));
    SG_REQUIRE(device >= 0 && device < ndev, "sg_ctx_create: device %d out of range (%d visible)", device, ndev);
    SG_CHECK_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    SG_CHECK_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        sg_set_error("sg_ctx_create: device %d is sm_%d%d; this library is built for sm_100a (B200) only", device,
                     prop.major, prop.minor);
        return SG_E_UNSUPPORTED;
    }
    sg_ctx *c = new sg_ctx();
    c->device = device;
    c->rank = rank;
    c->nranks = nranks;
    c->sm_count = prop.multiProcessorCount;
    c->comm = nullptr;
    if (nranks > 1) {
        if (!nccl_unique_id) {
            delete c;
            sg_set_error("sg_ctx_create: nranks > 1 needs an NCCL unique id");
            return SG_E_INVALID;
        }
        const SgNccl *n = sg_nccl();
        if (!n) {
            delete c;
            return SG_E_NCCL;
        }
        ncclUniqueId id;
        memcpy(&id, nccl_unique_id, sizeof(id));
        ncclComm_t comm;
        ncclResult_t r = n->CommInitRank(&comm, nranks, id, rank);
        if (r != ncclSuccess) {
            sg_set_error("ncclCommInitRank failed: %s", n->GetErrorString(r));
            delete c;
            return SG_E_NCCL;
        }
        c->comm = comm;
    }
    *out = c;
    return SG_OK;
}

int sg_ctx_destroy(sg_ctx *ctx) {
    if (!ctx) return SG_OK;
    if (ctx->comm) sg_nccl()->CommDestroy(ctx->comm);
    delete ctx;
    return SG_OK;
}

int sg_ctx_sm_count(const sg_ctx *ctx) { return ctx ? ctx->sm_count : -1; }

}  // extern "C"
