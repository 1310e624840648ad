// sbx_comm.cu -- NCCL plumbing for the latitude-band sharding (see sbx_comm.h).
#include "sbx_comm.h"

#include <dlfcn.h>

#include <cstdio>
#include <cstring>
#include <string>

namespace {

typedef void *ncclComm_t;
typedef int ncclResult_t;
struct ncclUniqueId_ { char internal[128]; };
enum { kNcclUint8 = 1, kNcclFloat64 = 8 };
enum { kNcclMax = 2 };

struct Nccl {
    void *lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId_ *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId_, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

thread_local std::string t_err;
Nccl g_nccl;

const char *set_err(const std::string &s)
{
    t_err = s;
    return t_err.c_str();
}

bool load_nccl(const char **err)
{
    if (g_nccl.lib) return true;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    void *lib = nullptr;
    for (const char *n : names) {
        lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (lib) break;
    }
    if (!lib) {
        *err = set_err(std::string("cannot dlopen libnccl.so.2: ") + dlerror());
        return false;
    }
#define SYM(field, name)                                                   \
    *(void **)(&g_nccl.field) = dlsym(lib, name);                          \
    if (!g_nccl.field) {                                                   \
        *err = set_err(std::string("libnccl is missing symbol ") + name);  \
        return false;                                                      \
    }
    SYM(GetUniqueId, "ncclGetUniqueId")
    SYM(CommInitRank, "ncclCommInitRank")
    SYM(CommDestroy, "ncclCommDestroy")
    SYM(Send, "ncclSend")
    SYM(Recv, "ncclRecv")
    SYM(GroupStart, "ncclGroupStart")
    SYM(GroupEnd, "ncclGroupEnd")
    SYM(AllReduce, "ncclAllReduce")
    SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
    g_nccl.lib = lib;
    return true;
}

}  // namespace

struct sbx_comm {
    ncclComm_t comm = nullptr;
    int nranks = 1, rank = 0, device = 0;
};

#define NC(expr)                                                                                   \
    do {                                                                                           \
        ncclResult_t r__ = (expr);                                                                 \
        if (r__ != 0) {                                                                            \
            *err = set_err(std::string(#expr) + ": " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r__) : "?")); \
            return -1;                                                                             \
        }                                                                                          \
    } while (0)

int sbx_comm_unique_id_impl(uint8_t out128[128], const char **err)
{
    if (!load_nccl(err)) return -1;
    ncclUniqueId_ id;
    NC(g_nccl.GetUniqueId(&id));
    memcpy(out128, id.internal, 128);
    return 0;
}

int sbx_comm_init_impl(const uint8_t id128[128], int nranks, int rank, int device, sbx_comm **out, const char **err)
{
    if (!load_nccl(err)) return -1;
    if (nranks < 1 || rank < 0 || rank >= nranks) {
        *err = set_err("bad nranks/rank");
        return -1;
    }
    ncclUniqueId_ id;
    memcpy(id.internal, id128, 128);
    sbx_comm *c = new sbx_comm();
    c->nranks = nranks;
    c->rank = rank;
    c->device = device;
    ncclResult_t r = g_nccl.CommInitRank(&c->comm, nranks, id, rank);
    if (r != 0) {
        *err = set_err(std::string("ncclCommInitRank: ") + g_nccl.GetErrorString(r));
        delete c;
        return -1;
    }
    *out = c;
    return 0;
}

void sbx_comm_destroy_impl(sbx_comm *c)
{
    if (!c) return;
    if (c->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
    delete c;
}

int sbx_comm_rank(const sbx_comm *c) { return c->rank; }
int sbx_comm_size(const sbx_comm *c) { return c->nranks; }

int sbx_comm_halo(sbx_comm *c, const void *first_row, void *top_ghost, const void *last_row, void *bottom_ghost,
                  size_t bytes, cudaStream_t stream, const char **err)
{
    if (c->nranks == 1) return 0;
    const int up = c->rank - 1, down = c->rank + 1;
    NC(g_nccl.GroupStart());
    if (up >= 0) {
        NC(g_nccl.Send(first_row, bytes, kNcclUint8, up, c->comm, stream));
        NC(g_nccl.Recv(top_ghost, bytes, kNcclUint8, up, c->comm, stream));
    }
    if (down < c->nranks) {
        NC(g_nccl.Send(last_row, bytes, kNcclUint8, down, c->comm, stream));
        NC(g_nccl.Recv(bottom_ghost, bytes, kNcclUint8, down, c->comm, stream));
    }
    NC(g_nccl.GroupEnd());
    return 0;
}

int sbx_comm_allreduce_max2(sbx_comm *c, double v[2], double *scratch_dev, cudaStream_t stream, const char **err)
{
    if (c->nranks == 1) return 0;
    if (cudaMemcpyAsync(scratch_dev, v, 16, cudaMemcpyHostToDevice, stream) != cudaSuccess) {
        *err = set_err("cudaMemcpyAsync failed");
        return -1;
    }
    NC(g_nccl.AllReduce(scratch_dev, scratch_dev, 2, kNcclFloat64, kNcclMax, c->comm, stream));
    if (cudaMemcpyAsync(v, scratch_dev, 16, cudaMemcpyDeviceToHost, stream) != cudaSuccess ||
        cudaStreamSynchronize(stream) != cudaSuccess) {
        *err = set_err("allreduce readback failed");
        return -1;
    }
    return 0;
}
