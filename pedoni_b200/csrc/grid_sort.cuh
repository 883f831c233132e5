// grid_sort.cuh — neighbor-grid rebuild as a deterministic, stable cell-key counting sort (sm_100a).
//
// Replaces NeighborGrid::update (neighbor_grid.rs:22-36) and the serial walk/gather of
// SocialForceModel::spawn_pedestrians (sfm.rs:58-77):
//
//   key        cell key per agent (fused into the force kernel's epilogue for agents that were just
//              integrated; this kernel only handles freshly spawned / uploaded agents)
//   histogram  per-cell population (atomicAdd on the cell counter; the returned ticket is a unique
//              but order-arbitrary slot inside the cell)
//   scan       exclusive prefix over cells, built from block-wide scans -> `neighbor_grid_indices`
//              (sfm.rs:61-75), length cells + 1
//   scatter    perm[start[cell] + ticket] = logical input index
//   gather     rank-by-counting inside the cell: an agent's final slot is start[cell] + #{members of
//              the cell with a smaller input index}. That is exactly "within a cell, ascending
//              previous index" (sfm.rs:66-68) and makes the output independent of the atomic
//              arrival order, i.e. run-to-run deterministic. Then one coalesced read / near-coalesced
//              write of the 24-byte state into the other buffer.
//
// The sort input is a virtual concatenation of up to kMaxSegments segments (inbound migrants from
// the slab below, the resident agents, inbound migrants from above, appended spawns): the order of
// the concatenation is the order of "previous index", which for slabs reproduces the single-GPU
// order exactly. Segment populations live in device memory so no host synchronisation is needed
// between ticks; grids are sized from host-known upper bounds.
#pragma once
#include "sfm_device.cuh"

namespace pedoni {

struct AgentArrays {
    float2* pos;     // sfm.rs:28 position
    float2* vel;     // sfm.rs:30 velocity
    float* v0;       // sfm.rs:31 desired_speed
    uint32_t* dest;  // sfm.rs:29 destination
};

constexpr int kMaxSegments = 4;

struct Segment {
    AgentArrays a;
    const uint32_t* d_range;  // device: [begin, end) of live entries inside the arrays; nullptr = [0, upper)
    uint32_t upper;           // host-known upper bound of (end - begin)
};

struct SortInput {
    Segment seg[kMaxSegments];
    uint32_t prefix[kMaxSegments + 1];  // exclusive prefix of `upper`
    int nseg;
};

// Logical input index t -> (segment, element index) or false if t is beyond the segment's live range.
__device__ __forceinline__ bool locate(const SortInput& in, uint32_t t, int& s, uint32_t& idx) {
    s = 0;
#pragma unroll
    for (int k = 1; k < kMaxSegments; ++k)
        if (k < in.nseg && t >= in.prefix[k]) s = k;
    // d_range == nullptr: the population is host-known, [0, upper) (appended spawns).
    uint32_t begin = 0, end = in.seg[s].upper;
    if (in.seg[s].d_range != nullptr) {
        begin = in.seg[s].d_range[0];
        end = in.seg[s].d_range[1];
    }
    idx = begin + (t - in.prefix[s]);
    return idx < end;
}

// ---- key: only for logical indices in [t_begin, t_end) whose keys are not fresh -------------------
__global__ void __launch_bounds__(256) key_kernel(SortInput in, uint32_t t_begin, uint32_t t_end, GridView g,
                                                  FieldView f, uint32_t* __restrict__ keys,
                                                  uint32_t* __restrict__ error_flag, bool foreign_rows_drop) {
    uint32_t t = t_begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= t_end) return;
    int s;
    uint32_t idx;
    uint32_t key = kKeyDrop;
    if (locate(in, t, s, idx)) {
        key = sort_key(g, f, in.seg[s].a.pos[idx], in.seg[s].a.dest[idx], error_flag);
        // Spawn lists are replicated to every slab; the owner keeps the agent, everybody else drops it.
        if (foreign_rows_drop && (key == kKeyMigrateDown || key == kKeyMigrateUp)) key = kKeyDrop;
    }
    keys[t] = key;
}

// ---- histogram -----------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) histogram_kernel(uint32_t total_upper, const uint32_t* __restrict__ keys,
                                                        uint32_t* __restrict__ cell_count,
                                                        uint32_t* __restrict__ ticket) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total_upper) return;
    uint32_t key = keys[t];
    if (key < kKeyFirstSpecial) ticket[t] = atomicAdd(cell_count + key, 1u);
}

// ---- scan: exclusive prefix over n_cells counters, three launches built from block-wide scans -----
constexpr int kScanThreads = 512;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;  // cells per block

__device__ __forceinline__ uint32_t warp_inclusive_scan(uint32_t v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t n = __shfl_up_sync(0xFFFFFFFFu, v, d);
        if (lane >= d) v += n;
    }
    return v;
}

// Block-wide exclusive scan of one value per thread; returns the exclusive prefix, *block_total = sum.
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* block_total) {
    __shared__ uint32_t warp_sums[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    uint32_t inc = warp_inclusive_scan(v, lane);
    if (lane == 31) warp_sums[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = lane < nwarps ? warp_sums[lane] : 0u;
        uint32_t winc = warp_inclusive_scan(w, lane);
        warp_sums[lane] = winc - w;  // exclusive prefix of warp sums
    }
    __syncthreads();
    uint32_t excl = inc - v + warp_sums[warp];
    // total = exclusive prefix of the last warp + its sum
    if (threadIdx.x == blockDim.x - 1) *block_total = excl + v;
    __syncthreads();
    return excl;
}

__global__ void __launch_bounds__(kScanThreads) scan_reduce_kernel(const uint32_t* __restrict__ cell_count,
                                                                   uint32_t n_cells, uint32_t* __restrict__ tile_sums) {
    __shared__ uint32_t total;
    const uint32_t base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
    uint32_t sum = 0;
    if (base + kScanItems <= n_cells) {
        const uint4* p = reinterpret_cast<const uint4*>(cell_count + base);
        uint4 a = p[0], b = p[1];
        sum = a.x + a.y + a.z + a.w + b.x + b.y + b.z + b.w;
    } else {
        for (int k = 0; k < kScanItems; ++k)
            if (base + k < n_cells) sum += cell_count[base + k];
    }
    block_exclusive_scan(sum, &total);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

// Single block: exclusive scan of the tile sums in place; writes the grand total.
__global__ void __launch_bounds__(kScanThreads) scan_tiles_kernel(uint32_t* __restrict__ tile_sums, uint32_t n_tiles,
                                                                  uint32_t* __restrict__ d_total) {
    __shared__ uint32_t total;
    uint32_t carry = 0;
    for (uint32_t base = 0; base < n_tiles; base += kScanThreads) {
        uint32_t i = base + threadIdx.x;
        uint32_t v = i < n_tiles ? tile_sums[i] : 0u;
        uint32_t excl = block_exclusive_scan(v, &total);
        if (i < n_tiles) tile_sums[i] = carry + excl;
        carry += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) *d_total = carry;
}

// cell_start[c] = exclusive prefix; cell_start[n_cells] = total.
__global__ void __launch_bounds__(kScanThreads) scan_apply_kernel(const uint32_t* __restrict__ cell_count,
                                                                  uint32_t n_cells,
                                                                  const uint32_t* __restrict__ tile_sums,
                                                                  uint32_t* __restrict__ cell_start) {
    __shared__ uint32_t total;
    const uint32_t base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
    uint32_t v[kScanItems];
    uint32_t sum = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        v[k] = (base + k < n_cells) ? cell_count[base + k] : 0u;
        sum += v[k];
    }
    uint32_t run = tile_sums[blockIdx.x] + block_exclusive_scan(sum, &total);
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        if (base + k < n_cells) cell_start[base + k] = run;
        run += v[k];
    }
    if (base < n_cells && base + kScanItems >= n_cells) cell_start[n_cells] = run;
}

// ---- scatter: perm[start[cell] + ticket] = t -----------------------------------------------------
__global__ void __launch_bounds__(256) scatter_kernel(uint32_t total_upper, const uint32_t* __restrict__ keys,
                                                      const uint32_t* __restrict__ ticket,
                                                      const uint32_t* __restrict__ cell_start,
                                                      uint32_t* __restrict__ perm) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total_upper) return;
    uint32_t key = keys[t];
    if (key < kKeyFirstSpecial) perm[__ldg(cell_start + key) + ticket[t]] = t;
}

// ---- gather: stable rank inside the cell, then move the 24-byte state ----------------------------
__global__ void __launch_bounds__(256) gather_kernel(SortInput in, uint32_t total_upper,
                                                     const uint32_t* __restrict__ keys,
                                                     const uint32_t* __restrict__ cell_start,
                                                     const uint32_t* __restrict__ perm, AgentArrays out) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total_upper) return;
    uint32_t key = keys[t];
    if (key >= kKeyFirstSpecial) return;
    int s;
    uint32_t idx;
    locate(in, t, s, idx);  // live by construction: dead logical indices carry kKeyDrop
    const uint32_t begin = __ldg(cell_start + key), end = __ldg(cell_start + key + 1);
    uint32_t rank = 0;
    for (uint32_t j = begin; j < end; ++j) rank += (__ldg(perm + j) < t) ? 1u : 0u;
    const uint32_t dst = begin + rank;
    const AgentArrays& a = in.seg[s].a;
    out.pos[dst] = a.pos[idx];
    out.vel[dst] = a.vel[idx];
    out.v0[dst] = a.v0[idx];
    out.dest[dst] = a.dest[idx];
}

// Publishes [begin, end) of the agents this handle owns after a rebuild, to the device-side range the
// next kernels read and to a pinned host slot (read by pedoni_count / pedoni_download after a sync).
__global__ void publish_range_kernel(const uint32_t* __restrict__ cell_start, uint32_t own_begin_cell,
                                     uint32_t own_end_cell, uint32_t* __restrict__ d_range,
                                     uint32_t* __restrict__ host_slot) {
    uint32_t b = cell_start[own_begin_cell], e = cell_start[own_end_cell];
    d_range[0] = b;
    d_range[1] = e;
    host_slot[0] = b;
    host_slot[1] = e;
}

}  // namespace pedoni
