// slab_comm.hpp — NCCL send/recv plumbing for the spatial slab decomposition (host side).
// NCCL is loaded lazily with dlopen so that a single-GPU user never needs libnccl.
#pragma once
#include <string>

namespace pedoni {

struct SlabComm;

int slab_comm_unique_id(void* out_id128, std::string* err);
SlabComm* slab_comm_create(const void* id128, int rank, int count, std::string* err);
void slab_comm_destroy(SlabComm* c);

}  // namespace pedoni
