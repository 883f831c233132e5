// slab_comm.hpp — NCCL send/recv plumbing for the spatial slab decomposition (host side).
// NCCL is loaded lazily with dlopen so that a single-GPU user never needs libnccl.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <string>

namespace pedoni {

struct SlabComm;

int slab_comm_unique_id(void* out_id128, std::string* err);
SlabComm* slab_comm_create(const void* id128, int rank, int count, std::string* err);
void slab_comm_destroy(SlabComm* c);

// One grouped exchange with the slab neighbours (rank - 1 and rank + 1), enqueued on `stream`:
// send_dn -> rank-1's recv_above, send_up -> rank+1's recv_below, and the two matching receives.
// Every message has the same fixed size, so no size handshake (and no host sync) is needed.
int slab_comm_exchange(SlabComm* c, cudaStream_t stream, const void* send_dn, void* recv_below, const void* send_up,
                       void* recv_above, size_t bytes, bool has_below, bool has_above, std::string* err);

// Minimum over all slab ranks of a device-resident int32 (agreement on the transport). Enqueued on `stream`.
int slab_comm_all_min(SlabComm* c, cudaStream_t stream, int* d_value, std::string* err);

}  // namespace pedoni
