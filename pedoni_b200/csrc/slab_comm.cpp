// slab_comm.cpp — placeholder until the slab exchange lands (see DESIGN.md "Multi-GPU").
#include "slab_comm.hpp"

#include "../../include/pedoni_cuda.h"

namespace pedoni {

struct SlabComm {};

int slab_comm_unique_id(void*, std::string* err) {
    if (err) *err = "slab communication is not built yet";
    return PEDONI_ERR_UNSUPPORTED;
}
SlabComm* slab_comm_create(const void*, int, int, std::string* err) {
    if (err) *err = "slab communication is not built yet";
    return nullptr;
}
void slab_comm_destroy(SlabComm* c) { delete c; }

}  // namespace pedoni
