// slab_comm.cpp — NCCL transport of the ghost-row exchange (one process per GPU, NVLink 5 / NVSwitch).
//
// The reference has no distributed code at all; this is the B200 analogue of its only
// scaling axis (rayon threads over agents, sfm.rs:93-95): row slabs over GPUs. Per tick each rank
// sends its first two owned cell rows to rank-1 and its last two to rank+1 as one fixed-size message
// each (grid_sort.cuh HaloMessage) inside a single ncclGroup, on the handle's edge stream, so the
// exchange overlaps the interior force kernel running on the main stream.
//
// NCCL is also the bootstrap of the faster peer-memory transport (pedoni_cuda.cu, setup_peer_transport):
// the ranks swap CUDA IPC handles of their receive arenas through this communicator; after that the pack
// kernel stores the strips straight into the neighbour's memory and NCCL leaves the per-tick path.
//
// libnccl is resolved at run time (dlopen): a whole-domain user never needs it, and inside a torch
// process the already-loaded bundled libnccl.so.2 is the one that gets used.
#include "slab_comm.hpp"

#include <dlfcn.h>
#include <nccl.h>  // types and prototypes only; nothing links against libnccl

#include <cstdlib>
#include <cstring>

#include "../../include/pedoni_cuda.h"

namespace pedoni {

namespace {

struct NcclApi {
    void* lib = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclSend) Send = nullptr;
    decltype(&ncclRecv) Recv = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclAllReduce) AllReduce = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
};

NcclApi* load_api(std::string* err) {
    static NcclApi api;
    static bool tried = false;
    static std::string load_error;
    if (!tried) {
        tried = true;
        const char* names[] = {std::getenv("PEDONI_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            if (!n || !*n) continue;
            api.lib = dlopen(n, RTLD_NOW | RTLD_LOCAL);
            if (api.lib) break;
            load_error = dlerror();
        }
        if (api.lib) {
#define RESOLVE(name)                                                       \
    api.name = reinterpret_cast<decltype(api.name)>(dlsym(api.lib, "nccl" #name)); \
    if (!api.name) {                                                        \
        load_error = "libnccl lacks nccl" #name;                            \
        api.lib = nullptr;                                                  \
    }
            RESOLVE(GetUniqueId)
            RESOLVE(CommInitRank)
            RESOLVE(CommDestroy)
            RESOLVE(Send)
            RESOLVE(Recv)
            RESOLVE(GroupStart)
            RESOLVE(GroupEnd)
            RESOLVE(AllReduce)
            RESOLVE(GetErrorString)
#undef RESOLVE
        }
    }
    if (!api.lib) {
        if (err) *err = "cannot load NCCL (" + load_error + "); multi-GPU slabs need libnccl.so.2";
        return nullptr;
    }
    return &api;
}

}  // namespace

struct SlabComm {
    NcclApi* api = nullptr;
    ncclComm_t comm = nullptr;
    int rank = 0, count = 1;
};

static_assert(PEDONI_COMM_ID_BYTES == NCCL_UNIQUE_ID_BYTES, "unique id size");

int slab_comm_unique_id(void* out_id128, std::string* err) {
    NcclApi* api = load_api(err);
    if (!api) return PEDONI_ERR_COMM;
    ncclUniqueId id;
    ncclResult_t r = api->GetUniqueId(&id);
    if (r != ncclSuccess) {
        if (err) *err = std::string("ncclGetUniqueId: ") + api->GetErrorString(r);
        return PEDONI_ERR_COMM;
    }
    std::memcpy(out_id128, &id, sizeof id);
    return PEDONI_OK;
}

SlabComm* slab_comm_create(const void* id128, int rank, int count, std::string* err) {
    NcclApi* api = load_api(err);
    if (!api) return nullptr;
    ncclUniqueId id;
    std::memcpy(&id, id128, sizeof id);
    SlabComm* c = new SlabComm();
    c->api = api;
    c->rank = rank;
    c->count = count;
    ncclResult_t r = api->CommInitRank(&c->comm, count, id, rank);  // collective: every slab rank calls it
    if (r != ncclSuccess) {
        if (err) *err = std::string("ncclCommInitRank: ") + api->GetErrorString(r);
        delete c;
        return nullptr;
    }
    return c;
}

void slab_comm_destroy(SlabComm* c) {
    if (!c) return;
    if (c->comm) c->api->CommDestroy(c->comm);
    delete c;
}

int slab_comm_exchange(SlabComm* c, cudaStream_t stream, const void* send_dn, void* recv_below, const void* send_up,
                       void* recv_above, size_t bytes, bool has_below, bool has_above, std::string* err) {
    NcclApi* a = c->api;
    ncclResult_t r = a->GroupStart();
    if (r == ncclSuccess && has_below) r = a->Send(send_dn, bytes, ncclUint8, c->rank - 1, c->comm, stream);
    if (r == ncclSuccess && has_below) r = a->Recv(recv_below, bytes, ncclUint8, c->rank - 1, c->comm, stream);
    if (r == ncclSuccess && has_above) r = a->Send(send_up, bytes, ncclUint8, c->rank + 1, c->comm, stream);
    if (r == ncclSuccess && has_above) r = a->Recv(recv_above, bytes, ncclUint8, c->rank + 1, c->comm, stream);
    ncclResult_t e = a->GroupEnd();
    if (r == ncclSuccess) r = e;
    if (r != ncclSuccess) {
        if (err) *err = std::string("NCCL halo exchange: ") + a->GetErrorString(r);
        return PEDONI_ERR_COMM;
    }
    return PEDONI_OK;
}

int slab_comm_all_min(SlabComm* c, cudaStream_t stream, int* d_value, std::string* err) {
    ncclResult_t r = c->api->AllReduce(d_value, d_value, 1, ncclInt32, ncclMin, c->comm, stream);
    if (r != ncclSuccess) {
        if (err) *err = std::string("ncclAllReduce: ") + c->api->GetErrorString(r);
        return PEDONI_ERR_COMM;
    }
    return PEDONI_OK;
}

}  // namespace pedoni
