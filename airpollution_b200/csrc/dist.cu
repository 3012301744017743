// Communicator for the row-block partitioned solve: one process per GPU, NCCL
// over NVLink/NVSwitch.  Only two exchange patterns exist on the path
// (SURVEY.md section 8e): the halo-DOF exchange before each SpMV (grouped
// ncclSend/ncclRecv with the one or two strip neighbours) and the allreduce of
// the 1-3 BiCGStab dot products.  NCCL is confined to this translation unit.
#include <dlfcn.h>
#include <nccl.h>   // types and prototypes only: the library is bound at run time (see nccl_api below)
#include <string.h>

#include "crbe_common.cuh"

// libnccl is dlopen'ed on first use instead of being a link-time dependency: a process that has already
// loaded a (newer) libnccl.so.2 -- PyTorch bundles its own -- must keep using that one, and hosts that never
// go multi-GPU do not need NCCL at all.
struct nccl_api {
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclAllReduce) AllReduce = nullptr;
    decltype(&ncclSend) Send = nullptr;
    decltype(&ncclRecv) Recv = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    bool ok = false;
};

static nccl_api* nccl() {
    static nccl_api api;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (h) {
#define CRBE_BIND(name) api.name = (decltype(api.name))dlsym(h, "nccl" #name)
            CRBE_BIND(GetUniqueId);
            CRBE_BIND(CommInitRank);
            CRBE_BIND(CommDestroy);
            CRBE_BIND(AllReduce);
            CRBE_BIND(Send);
            CRBE_BIND(Recv);
            CRBE_BIND(GroupStart);
            CRBE_BIND(GroupEnd);
            CRBE_BIND(GetErrorString);
#undef CRBE_BIND
            api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllReduce && api.Send && api.Recv &&
                     api.GroupStart && api.GroupEnd && api.GetErrorString;
        }
    }
    return &api;
}

#define CRBE_NEED_NCCL()                                                                  \
    do {                                                                                  \
        if (!nccl()->ok) {                                                                \
            crbe_set_error("libnccl.so.2 could not be loaded: %s", dlerror() ? dlerror() : "missing symbols"); \
            return CRBE_ERR_COMM;                                                         \
        }                                                                                 \
    } while (0)

struct crbe_comm {
    ncclComm_t nccl = nullptr;
    int rank = 0, world = 1;
    crbe_ctx* ctx = nullptr;
};

#define CRBE_NCCL(call)                                                                          \
    do {                                                                                         \
        ncclResult_t r_ = (call);                                                                \
        if (r_ != ncclSuccess) {                                                                 \
            crbe_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, nccl()->GetErrorString(r_)); \
            return CRBE_ERR_COMM;                                                                \
        }                                                                                        \
    } while (0)

extern "C" int crbe_comm_unique_id_bytes(void) { return (int)sizeof(ncclUniqueId); }

extern "C" int crbe_comm_unique_id(void* id_out) {
    CRBE_REQUIRE(id_out != nullptr, "null argument");
    ncclUniqueId id;
    CRBE_NEED_NCCL();
    CRBE_NCCL(nccl()->GetUniqueId(&id));
    memcpy(id_out, &id, sizeof(id));
    return CRBE_OK;
}

extern "C" int crbe_comm_create(crbe_ctx* ctx, int rank, int world, const void* unique_id, crbe_comm** out) {
    CRBE_REQUIRE(ctx && out && unique_id && world >= 1 && rank >= 0 && rank < world, "bad argument");
    CRBE_NEED_NCCL();
    CRBE_CUDA(cudaSetDevice(ctx->device));
    crbe_comm* c = new crbe_comm();
    c->rank = rank;
    c->world = world;
    c->ctx = ctx;
    ncclUniqueId id;
    memcpy(&id, unique_id, sizeof(id));
    ncclResult_t r = nccl()->CommInitRank(&c->nccl, world, id, rank);
    if (r != ncclSuccess) {
        crbe_set_error("ncclCommInitRank failed: %s", nccl()->GetErrorString(r));
        delete c;
        return CRBE_ERR_COMM;
    }
    *out = c;
    return CRBE_OK;
}

extern "C" int crbe_comm_destroy(crbe_comm* c) {
    if (!c) return CRBE_OK;
    if (c->nccl) nccl()->CommDestroy(c->nccl);
    delete c;
    return CRBE_OK;
}

int crbe_comm_rank(const crbe_comm* c) { return c ? c->rank : 0; }
int crbe_comm_world(const crbe_comm* c) { return c ? c->world : 1; }

// sum of `count` doubles over all ranks, send_d -> recv_d (may be the same buffer), enqueued on `stream`
int crbe_comm_allreduce_sum(crbe_comm* c, const double* send_d, double* recv_d, int count, cudaStream_t stream) {
    CRBE_NCCL(nccl()->AllReduce(send_d, recv_d, (size_t)count, ncclDouble, ncclSum, c->nccl, stream));
    return CRBE_OK;
}

// One grouped exchange: send sendbuf[send_off[q] .. send_off[q+1]) to neighbour q and receive
// recv_off[q+1]-recv_off[q] doubles from it into recvbuf + recv_off[q].
int crbe_comm_exchange(crbe_comm* c, int n_neigh, const int* neigh, const double* sendbuf_d, const int64_t* send_off,
                       double* recvbuf_d, const int64_t* recv_off, cudaStream_t stream) {
    CRBE_NCCL(nccl()->GroupStart());
    for (int q = 0; q < n_neigh; ++q) {
        const int64_t ns = send_off[q + 1] - send_off[q], nr = recv_off[q + 1] - recv_off[q];
        if (ns > 0) CRBE_NCCL(nccl()->Send(sendbuf_d + send_off[q], (size_t)ns, ncclDouble, neigh[q], c->nccl, stream));
        if (nr > 0) CRBE_NCCL(nccl()->Recv(recvbuf_d + recv_off[q], (size_t)nr, ncclDouble, neigh[q], c->nccl, stream));
    }
    CRBE_NCCL(nccl()->GroupEnd());
    return CRBE_OK;
}

// test hooks through the ABI
extern "C" int crbe_comm_test_allreduce(crbe_comm* c, double* buf_d, int count) {
    CRBE_REQUIRE(c && buf_d && count > 0, "bad argument");
    return crbe_comm_allreduce_sum(c, buf_d, buf_d, count, c->ctx->stream);
}
