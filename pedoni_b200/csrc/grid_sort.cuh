// grid_sort.cuh — neighbor-grid rebuild as a deterministic, stable cell-key counting sort (sm_100a),
// plus the ghost-row pack/unpack of the slab decomposition.
//
// Replaces NeighborGrid::update (neighbor_grid.rs:22-36) and the serial walk/gather of
// SocialForceModel::spawn_pedestrians (sfm.rs:58-77):
//
//   key        cell key per agent + per-cell population (atomicAdd on the cell counter; the returned
//              ticket is a unique but order-arbitrary slot inside the cell). Fused into the force kernel's
//              epilogue for agents that were just integrated; key_kernel only handles freshly spawned /
//              uploaded agents. The counters are zeroed at the end of every rebuild.
//   scan       exclusive prefix over cells, built from block-wide scans -> `neighbor_grid_indices`
//              (sfm.rs:61-75), length cells + 1
//   scatter    perm[start[cell] + ticket] = logical input index
//   gather     rank-by-counting inside the cell: an agent's final slot is start[cell] + #{members of
//              the cell with a smaller input index}. That is exactly "within a cell, ascending
//              previous index" (sfm.rs:66-68) and makes the output independent of the atomic
//              arrival order, i.e. run-to-run deterministic. Then one coalesced read / near-coalesced
//              write of the 24-byte state into the other buffer.
//
// The sort input is a virtual concatenation of segments (the resident agents this handle just
// integrated — its owned rows plus, on a slab handle, one ghost row each side — then the appended
// spawns): the order of the concatenation is the order of "previous index", which for slabs
// reproduces the single-GPU order exactly because ghost rows sit in the array where their rows sit in
// the global order. Segment populations live in device memory so no host synchronisation is needed
// between ticks; grids are sized from host-known upper bounds.
//
// Array layout of a slab handle (H = halo capacity), all four SoA columns alike:
//   [H - n_below, H)                      ghost rows r0-2, r0-1 (copied from the slab below)
//   [H, H + n_owned)                      owned rows r0 .. r1-1, cell-sorted
//   [H + n_owned, H + n_owned + n_above)  ghost rows r1, r1+1 (copied from the slab above)
// so the local cell table indexes one contiguous run and keeps the reference's (cells + 1) layout.
// A whole-domain handle is the same with H = 0 and no ghosts.
#pragma once
#include "sfm_device.cuh"

namespace pedoni {

struct AgentArrays {
    float2* pos;     // sfm.rs:28 position
    float2* vel;     // sfm.rs:30 velocity
    float* v0;       // sfm.rs:31 desired_speed
    uint32_t* dest;  // sfm.rs:29 destination
};

// Device-resident [begin, end) pairs describing the current layout (PedoniModel::d_ranges).
enum RangeId : int {
    kRangeOwned = 0,     // agents this handle owns (what count / download report)
    kRangeCompute = 1,   // agents this handle integrates: owned + one ghost row each side
    kRangeInterior = 2,  // owned agents whose 3x3 block touches no ghost row
    kRangeEdgeLo = 3,    // ghost row r0-1 + owned row r0 (needs the halo from below)
    kRangeEdgeHi = 4,    // owned row r1-1 + ghost row r1 (needs the halo from above)
    kNumRanges = 5
};

constexpr int kMaxSegments = 2;

struct Segment {
    AgentArrays a;
    uint32_t* keys;           // sort keys, indexed like the arrays
    uint32_t* ticket;         // slot inside the cell (kKeyDrop: not kept), indexed like the arrays
    const uint32_t* d_range;  // device: [begin, end) of live entries inside the arrays; nullptr = [0, upper)
    uint32_t upper;           // host-known upper bound of (end - begin)
};

struct SortInput {
    Segment seg[kMaxSegments];
    uint32_t prefix[kMaxSegments + 1];  // exclusive prefix of `upper`
    int nseg;
};

// Logical input index t -> the segment's arrays and the element index; live = false if t is beyond the
// segment's live range. Two segments, resolved with selects (no dynamic indexing of the kernel
// parameter, which would force a local-memory copy).
struct Located {
    AgentArrays a;
    uint32_t* keys;
    uint32_t* ticket;
    uint32_t idx;
    bool live;
};
static_assert(kMaxSegments == 2, "locate() selects between exactly two segments");

__device__ __forceinline__ Located locate(const SortInput& in, uint32_t t) {
    const bool second = in.nseg > 1 && t >= in.prefix[1];
    Located r;
    r.a.pos = second ? in.seg[1].a.pos : in.seg[0].a.pos;
    r.a.vel = second ? in.seg[1].a.vel : in.seg[0].a.vel;
    r.a.v0 = second ? in.seg[1].a.v0 : in.seg[0].a.v0;
    r.a.dest = second ? in.seg[1].a.dest : in.seg[0].a.dest;
    r.keys = second ? in.seg[1].keys : in.seg[0].keys;
    r.ticket = second ? in.seg[1].ticket : in.seg[0].ticket;
    const uint32_t* d_range = second ? in.seg[1].d_range : in.seg[0].d_range;
    // d_range == nullptr: the population is host-known, [0, upper) (appended spawns).
    uint32_t begin = 0, end = second ? in.seg[1].upper : in.seg[0].upper;
    if (d_range != nullptr) {
        begin = d_range[0];
        end = d_range[1];
    }
    r.idx = begin + (t - (second ? in.prefix[1] : in.prefix[0]));
    r.live = r.idx < end;
    return r;
}

// ---- device-side spawning (lib.rs:37-52,67-86; sfm.rs:49-56) with the harness's counter-based stream -----
__host__ __device__ inline unsigned long long splitmix64(unsigned long long x) {
    unsigned long long z = x + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__device__ __forceinline__ unsigned long long spawn_stream_u64(unsigned long long seed, unsigned long long k) {
    return splitmix64(seed ^ (k * 0x2545F4914F6CDD1Dull));
}

struct SpawnGroupDev {
    float p1_x, p1_y, p2_x, p2_y;
    uint32_t destination;
    uint32_t first;  // exclusive prefix of the counts
};

// One thread per spawned pedestrian j of the call (n in total): written at out[at + j].
__global__ void __launch_bounds__(256) spawn_groups_kernel(AgentArrays out, uint32_t at, uint32_t n,
                                                           const SpawnGroupDev* __restrict__ groups, uint32_t n_groups,
                                                           unsigned long long seed, unsigned long long counter) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    uint32_t g = 0;  // groups are few (one per [[pedestrians]] entry): linear search
    while (g + 1 < n_groups && groups[g + 1].first <= j) ++g;
    const SpawnGroupDev G = groups[g];
    // fastrand::f32(): 24 random mantissa bits
    const float u = static_cast<float>(spawn_stream_u64(seed, counter + j) >> 40) * (1.0f / 16777216.0f);
    // glam lerp: self + (rhs - self) * s
    const float px = S::add(G.p1_x, S::mul(S::sub(G.p2_x, G.p1_x), u));
    const float py = S::add(G.p1_y, S::mul(S::sub(G.p2_y, G.p1_y), u));
    // f32_normal_approx: popcount of 64 bits centred + difference of two 32-bit uniforms, unit variance
    const unsigned long long a = spawn_stream_u64(seed, counter + n + j), b = spawn_stream_u64(seed, counter + 2ull * n + j);
    const double pop = static_cast<double>(__popcll(a)) - 32.0;
    const double tri = (static_cast<double>(b & 0xFFFFFFFFull) - static_cast<double>(b >> 32)) * (1.0 / 4294967296.0);
    const double z = (pop + tri) * 0x1.fd5a9eebfd779p-3;  // 1 / sqrt(16 + 1/6) = 0.24870800168690346
    const float v0 = S::add(1.34f, S::mul(0.26f, static_cast<float>(z)));
    out.pos[at + j] = make_float2(px, py);
    out.vel[at + j] = make_float2(0.0f, 0.0f);  // sfm.rs:53
    out.v0[at + j] = v0;
    out.dest[at + j] = G.destination;
}

// Key + population count of one agent: shared by key_kernel and the force kernel's epilogue.
__device__ __forceinline__ void count_key(uint32_t key, uint32_t* __restrict__ cell_count, uint32_t* key_slot,
                                          uint32_t* ticket_slot) {
    *key_slot = key;
    *ticket_slot = key < kKeyFirstSpecial ? atomicAdd(cell_count + key, 1u) : kKeyDrop;
}

// ---- key: only for logical indices in [t_begin, t_end) whose keys are not fresh -------------------
__global__ void __launch_bounds__(256) key_kernel(SortInput in, uint32_t t_begin, uint32_t t_end, GridView g,
                                                  FieldView f, uint32_t* __restrict__ cell_count,
                                                  uint32_t* __restrict__ error_flag,
                                                  unsigned long long* __restrict__ arrived) {
    uint32_t t = t_begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= t_end) return;
    const Located l = locate(in, t);
    if (!l.live) return;
    // Spawn lists are replicated to every slab and ghost rows are copies: sort_key keeps an agent only
    // on the handle that owns its row.
    // a replicated spawn that stands on its destination is counted as arrived by the slab owning its row
    const float2 pos = l.a.pos[l.idx];
    const int row = __float2int_rz(S::div(pos.y, g.unit));
    count_key(sort_key(g, f, pos, l.a.dest[l.idx], error_flag, arrived, row >= g.own_row0 && row < g.own_row1),
              cell_count, l.keys + l.idx, l.ticket + l.idx);
}

// ---- scan: exclusive prefix over n_cells counters in ONE pass (chained scan over tile aggregates) -------
// A tile = 1024 threads x 16 cells (10 M pedestrians = 5.1 M cells = 312 tiles). Tiles are handed out by an atomic ticket (so a tile only ever waits for
// tiles whose CTAs are already running), publish their aggregate in a 64-bit status word tagged with the
// tick (no reset between launches), and sum their predecessors' aggregates for the exclusive prefix. The
// same pass
//   - zeroes the counters for the next tick's fused histogram (force epilogue / key_kernel),
//   - writes the layout ranges that depend on the owned rows (the thread that produces the cell-start
//     they are read from writes them), and
//   - publishes the owned population to the host: one aligned 64-bit store to pinned memory,
//     tick << 32 | n_owned, so a lagging reader never sees a torn pair.
constexpr int kScanThreads = 1024;
#ifndef PEDONI_SCAN_ITEMS
#define PEDONI_SCAN_ITEMS 16
#endif
constexpr int kScanItems = PEDONI_SCAN_ITEMS;
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

struct ScanLayout {  // what the scan publishes besides the table (see RangeId)
    uint32_t own_begin_cell, own_end_cell, nx;
    int has_below, has_above;
    uint32_t* ranges;
    unsigned long long* host_slot;
    uint32_t tick;
};

constexpr unsigned long long kScanAggregate = 1ull << 62, kScanPrefix = 2ull << 62;
__device__ __forceinline__ unsigned long long scan_status(unsigned long long flag, uint32_t tick, uint32_t value) {
    return flag | (static_cast<unsigned long long>(tick & 0x3FFFFFFFu) << 32) | value;
}

__device__ __forceinline__ void publish_cell_start(const ScanLayout& L, uint32_t cell, uint32_t start) {
    if (cell == L.own_end_cell) {  // one past the last owned agent
        L.ranges[2 * kRangeOwned + 1] = start;
        if (!L.has_above) {
            L.ranges[2 * kRangeInterior + 1] = start;
            L.ranges[2 * kRangeCompute + 1] = start;
            L.ranges[2 * kRangeEdgeHi] = start;
            L.ranges[2 * kRangeEdgeHi + 1] = start;
        }
        const uint32_t first = L.ranges[2 * kRangeOwned];  // constant: the array offset (reset_layout_kernel)
        *L.host_slot = (static_cast<unsigned long long>(L.tick) << 32) | static_cast<unsigned long long>(start - first);
    }
    if (L.has_below && cell == L.own_begin_cell + L.nx) L.ranges[2 * kRangeInterior] = start;
    if (L.has_above && cell == L.own_end_cell - L.nx) L.ranges[2 * kRangeInterior + 1] = start;
}

// cell_start[c] = offset + exclusive prefix; cell_start[n_cells] = offset + total. `offset` is the halo
// capacity H on a slab handle with a neighbour below (owned agents start at H), 0 otherwise.
__global__ void __launch_bounds__(kScanThreads) scan_cells_kernel(uint32_t* __restrict__ cell_count, uint32_t n_cells,
                                                                  uint32_t offset, uint32_t* __restrict__ cell_start,
                                                                  unsigned long long* __restrict__ tile_status,
                                                                  uint32_t* __restrict__ tile_ticket, uint32_t n_tiles,
                                                                  ScanLayout L) {
    __shared__ uint32_t s_tile, s_total, s_prefix;
    if (threadIdx.x == 0) {
        s_tile = atomicAdd(tile_ticket, 1u);
        if (s_tile == n_tiles - 1) *tile_ticket = 0;  // every ticket of this launch has been taken
    }
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint32_t base = tile * kScanTile + threadIdx.x * kScanItems;
    uint32_t v[kScanItems];
    uint32_t sum = 0;
    if (base + kScanItems <= n_cells) {
        uint4* p = reinterpret_cast<uint4*>(cell_count + base);
#pragma unroll
        for (int q = 0; q < kScanItems / 4; ++q) {
            const uint4 a = p[q];
            v[4 * q] = a.x, v[4 * q + 1] = a.y, v[4 * q + 2] = a.z, v[4 * q + 3] = a.w;
            p[q] = make_uint4(0u, 0u, 0u, 0u);
        }
    } else {
#pragma unroll
        for (int k = 0; k < kScanItems; ++k) {
            v[k] = (base + k < n_cells) ? cell_count[base + k] : 0u;
            if (base + k < n_cells) cell_count[base + k] = 0u;
        }
    }
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) sum += v[k];
    const uint32_t excl = block_exclusive_scan(sum, &s_total);

    // Exclusive prefix of the tile: every predecessor's aggregate, fetched in ONE round of loads spread over
    // the block (no chain of dependent round trips through earlier tiles' prefixes).
    // Predecessors hold smaller tickets, so their CTAs are running or done: the spin cannot deadlock.
    {
        volatile unsigned long long* status = tile_status;
        const unsigned long long tag = scan_status(kScanAggregate, L.tick, 0u) >> 32;
        if (threadIdx.x == 0) {
            s_prefix = 0;
            status[tile] = scan_status(kScanAggregate, L.tick, s_total);
        }
        __syncthreads();
        uint32_t part = 0;
        for (uint32_t j = threadIdx.x; j < tile; j += kScanThreads) {
            unsigned long long st;
            do {
                st = status[j];
            } while ((st >> 32) != tag);
            part += static_cast<uint32_t>(st);
        }
        part = __reduce_add_sync(0xFFFFFFFFu, part);
        if ((threadIdx.x & 31) == 0 && part != 0) atomicAdd(&s_prefix, part);
    }
    __syncthreads();

    uint32_t run = offset + s_prefix + excl;
    if (base + kScanItems <= n_cells) {
        // 16-byte stores: scalar ones reach L2 as one partial-sector write per cell
        uint32_t start[kScanItems];
#pragma unroll
        for (int k = 0; k < kScanItems; ++k) {
            start[k] = run;
            publish_cell_start(L, base + k, run);
            run += v[k];
        }
        uint4* p = reinterpret_cast<uint4*>(cell_start + base);
#pragma unroll
        for (int q = 0; q < kScanItems / 4; ++q)
            p[q] = make_uint4(start[4 * q], start[4 * q + 1], start[4 * q + 2], start[4 * q + 3]);
    } else {
#pragma unroll
        for (int k = 0; k < kScanItems; ++k) {
            if (base + k < n_cells) {
                cell_start[base + k] = run;
                publish_cell_start(L, base + k, run);
            }
            run += v[k];
        }
    }
    if (base < n_cells && base + kScanItems >= n_cells) {
        cell_start[n_cells] = run;
        publish_cell_start(L, n_cells, run);
    }
}

// ---- scatter: perm[start[cell] + ticket] = t -----------------------------------------------------
__device__ __forceinline__ void scatter_one(const SortInput& in, uint32_t t, const uint32_t* __restrict__ cell_start,
                                            uint32_t* __restrict__ perm) {
    const Located l = locate(in, t);
    if (!l.live) return;
    const uint32_t tk = l.ticket[l.idx];
    if (tk == kKeyDrop) return;
    perm[cell_start[l.keys[l.idx]] + tk] = t;
}

// Elements per thread, strided by the block so every access stays coalesced. The scatter is a chain of
// dependent loads (key -> cell start -> slot): two chains per thread keep the memory system busier (B200,
// 10 M: 0.057 -> 0.044 ms). The gather wants the opposite: one element per thread and <= 32 registers for full
// occupancy, with the 24-byte state loaded BEFORE the key -> cell start -> rank chain (0.187 -> 0.127 ms;
// 2 / 4 / 8 elements: 0.140 / 0.162 / 0.255 ms).
#ifndef PEDONI_SCATTER_ITEMS
#define PEDONI_SCATTER_ITEMS 2
#endif
#ifndef PEDONI_GATHER_ITEMS
#define PEDONI_GATHER_ITEMS 1
#endif
constexpr int kSortThreads = 256;
constexpr int kScatterItems = PEDONI_SCATTER_ITEMS, kScatterTile = kSortThreads * kScatterItems;
constexpr int kGatherItems = PEDONI_GATHER_ITEMS, kGatherTile = kSortThreads * kGatherItems;

__global__ void __launch_bounds__(kSortThreads) scatter_kernel(SortInput in, uint32_t total_upper,
                                                               const uint32_t* __restrict__ cell_start,
                                                               uint32_t* __restrict__ perm) {
    const uint32_t t0 = blockIdx.x * kScatterTile + threadIdx.x;
    uint32_t key[kScatterItems], tk[kScatterItems], start[kScatterItems];
#pragma unroll
    for (int k = 0; k < kScatterItems; ++k) {
        const uint32_t t = t0 + k * kSortThreads;
        tk[k] = kKeyDrop;
        key[k] = 0;
        if (t < total_upper) {
            const Located l = locate(in, t);
            if (l.live) tk[k] = l.ticket[l.idx], key[k] = l.keys[l.idx];
        }
    }
#pragma unroll
    for (int k = 0; k < kScatterItems; ++k) start[k] = tk[k] != kKeyDrop ? cell_start[key[k]] : 0u;
#pragma unroll
    for (int k = 0; k < kScatterItems; ++k)
        if (tk[k] != kKeyDrop) perm[start[k] + tk[k]] = t0 + k * kSortThreads;
}

// ---- gather: stable rank inside the cell, then move the 24-byte state ----------------------------
__device__ __forceinline__ void gather_one(const SortInput& in, uint32_t t, const uint32_t* __restrict__ cell_start,
                                           const uint32_t* __restrict__ perm, const AgentArrays& out) {
    const Located l = locate(in, t);
    if (!l.live) return;
    const uint32_t idx = l.idx;
    const uint32_t key = l.keys[idx];
    if (key >= kKeyFirstSpecial) return;  // count_key: exactly the entries whose ticket is kKeyDrop
    const uint32_t begin = cell_start[key], end = cell_start[key + 1];
    uint32_t rank = 0;
    for (uint32_t j = begin; j < end; ++j) rank += (perm[j] < t) ? 1u : 0u;
    const uint32_t dst = begin + rank;
    const AgentArrays& a = l.a;
    out.pos[dst] = a.pos[idx];
    out.vel[dst] = a.vel[idx];
    out.v0[dst] = a.v0[idx];
    out.dest[dst] = a.dest[idx];
}

__global__ void __launch_bounds__(kSortThreads) gather_kernel(SortInput in, uint32_t total_upper,
                                                              const uint32_t* __restrict__ cell_start,
                                                              const uint32_t* __restrict__ perm, AgentArrays out) {
    const uint32_t t0 = blockIdx.x * kGatherTile + threadIdx.x;
    uint32_t key[kGatherItems], begin[kGatherItems], end[kGatherItems], dest[kGatherItems];
    bool keep[kGatherItems];
    float2 pos[kGatherItems], vel[kGatherItems];
    float v0[kGatherItems];
    // every load that does not depend on another one first: key, ticket and the 24-byte state
#pragma unroll
    for (int k = 0; k < kGatherItems; ++k) {
        const uint32_t t = t0 + k * kSortThreads;
        keep[k] = false;
        key[k] = 0;
        if (t < total_upper) {
            const Located l = locate(in, t);
            if (l.live) {
                key[k] = l.keys[l.idx];
                keep[k] = key[k] < kKeyFirstSpecial;  // count_key: exactly the entries that hold a ticket
                pos[k] = l.a.pos[l.idx];
                vel[k] = l.a.vel[l.idx];
                v0[k] = l.a.v0[l.idx];
                dest[k] = l.a.dest[l.idx];
            }
        }
    }
#pragma unroll
    for (int k = 0; k < kGatherItems; ++k) {
        begin[k] = end[k] = 0;
        if (keep[k]) begin[k] = cell_start[key[k]], end[k] = cell_start[key[k] + 1];
    }
#pragma unroll
    for (int k = 0; k < kGatherItems; ++k) {
        if (!keep[k]) continue;
        const uint32_t t = t0 + k * kSortThreads;
        uint32_t rank = 0;  // stable: the number of cell mates that come earlier in the input
        for (uint32_t j = begin[k]; j < end[k]; ++j) rank += (perm[j] < t) ? 1u : 0u;
        const uint32_t dst = begin[k] + rank;
        out.pos[dst] = pos[k];
        out.vel[dst] = vel[k];
        out.v0[dst] = v0[k];
        out.dest[dst] = dest[k];
    }
}

// ---- the whole rebuild in ONE CTA, for small crowds -------------------------------------------------------
// A shipped scenario holds tens to a few thousand pedestrians; at that size a tick is nothing but launch
// latency (~6 us per dependent kernel). One 1024-thread CTA runs the scan (each thread owns a contiguous
// chunk of cells), the scatter and the gather back to back with block barriers in between.
constexpr uint32_t kSmallRebuildMaxAgents = 2048;  // beyond these one CTA is slower than three launches (measured:
constexpr uint32_t kSmallRebuildMaxCells = 8192;    // bottleneck.toml, 3 k agents on 20 k cells: 44 vs 34 us per tick)

__global__ void __launch_bounds__(1024) rebuild_small_kernel(SortInput in, uint32_t total_upper,
                                                             uint32_t* __restrict__ cell_count, uint32_t n_cells,
                                                             uint32_t offset, uint32_t* __restrict__ cell_start,
                                                             uint32_t* __restrict__ perm, AgentArrays out, ScanLayout L) {
    __shared__ uint32_t s_total;
    const uint32_t chunk = (n_cells + blockDim.x - 1) / blockDim.x;
    const uint32_t c0 = min(threadIdx.x * chunk, n_cells), c1 = min(c0 + chunk, n_cells);
    uint32_t sum = 0;
    for (uint32_t c = c0; c < c1; ++c) sum += cell_count[c];
    uint32_t run = offset + block_exclusive_scan(sum, &s_total);
    for (uint32_t c = c0; c < c1; ++c) {
        const uint32_t v = cell_count[c];
        cell_count[c] = 0;
        cell_start[c] = run;
        publish_cell_start(L, c, run);
        run += v;
    }
    if (threadIdx.x == blockDim.x - 1) {  // c1 == n_cells for the last thread (and for any thread past the end)
        cell_start[n_cells] = offset + s_total;
        publish_cell_start(L, n_cells, offset + s_total);
    }
    __syncthreads();  // block-wide visibility of the table in global memory
    for (uint32_t t = threadIdx.x; t < total_upper; t += blockDim.x) scatter_one(in, t, cell_start, perm);
    __syncthreads();
    for (uint32_t t = threadIdx.x; t < total_upper; t += blockDim.x) gather_one(in, t, cell_start, perm, out);
}

// ---- ghost rows -----------------------------------------------------------------------------------
// One message per direction, a flat array of 32-bit words:
//   [0] n = agents in the two rows   [1] tick   [2 .. 2 + 2nx] starts relative to the first agent
//   then pos (2H words), vel (2H), desired_speed (H), destination (H), H = halo capacity.
__host__ __device__ inline uint32_t halo_header_words(uint32_t nx) { return (2u * nx + 3u + 3u) & ~3u; }
__host__ __device__ inline size_t halo_message_words(uint32_t nx, uint32_t halo_cap) {
    return static_cast<size_t>(halo_header_words(nx)) + 6ull * halo_cap;
}

struct HaloMessage {
    uint32_t* words;
    __host__ __device__ uint32_t* starts() const { return words + 2; }
    __host__ __device__ float2* pos(uint32_t nx, uint32_t) const {
        return reinterpret_cast<float2*>(words + halo_header_words(nx));
    }
    __host__ __device__ float2* vel(uint32_t nx, uint32_t h) const {
        return reinterpret_cast<float2*>(words + halo_header_words(nx) + 2ull * h);
    }
    __host__ __device__ float* v0(uint32_t nx, uint32_t h) const {
        return reinterpret_cast<float*>(words + halo_header_words(nx) + 4ull * h);
    }
    __host__ __device__ uint32_t* dest(uint32_t nx, uint32_t h) const { return words + halo_header_words(nx) + 5ull * h; }
};

// How a packed strip announces itself to a receiver that polls (peer-memory transport): see PeerSignal.
struct PeerSignal {
    uint32_t* done_count;  // local: CTAs of this pack launch that have finished a side (2 counters)
    uint32_t* flag_down;   // in the slab BELOW's memory: its "strip from above has arrived" sequence number
    uint32_t* flag_up;     // in the slab ABOVE's memory: its "strip from below has arrived" sequence number
    uint32_t seq;          // ordinal of this exchange (every slab counts its rebuilds alike)
};

__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// blockIdx.y = 0: the first two owned rows -> message for the slab below;
// blockIdx.y = 1: the last two owned rows  -> message for the slab above.
// `down` / `up` may point into local staging buffers (NCCL and in-process transports) or straight into
// the neighbour's receive slot over NVLink (peer-memory transport: pack and send are this one kernel;
// the last CTA to finish a side publishes the sequence number behind a system-wide fence).
__global__ void __launch_bounds__(256) halo_pack_kernel(AgentArrays a, const uint32_t* __restrict__ cell_start,
                                                        uint32_t own_begin_cell, uint32_t own_end_cell, uint32_t nx,
                                                        uint32_t halo_cap, HaloMessage down, HaloMessage up,
                                                        int has_below, int has_above, uint32_t tick,
                                                        uint32_t* __restrict__ error_flag, PeerSignal sig) {
    const int side = blockIdx.y;
    if ((side == 0 && !has_below) || (side == 1 && !has_above)) return;
    const HaloMessage msg = side == 0 ? down : up;
    const uint32_t c0 = side == 0 ? own_begin_cell : own_end_cell - 2 * nx;
    const uint32_t first = cell_start[c0];
    uint32_t n = cell_start[c0 + 2 * nx] - first;
    const bool overflow = n > halo_cap;
    if (overflow) n = 0;  // ship nothing rather than a torn strip; the flag surfaces at the next blocking call
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) {
        msg.words[0] = n;
        msg.words[1] = tick;
        if (overflow) atomicOr(error_flag, kErrHaloOverflow);
    }
    if (i <= 2 * nx) msg.starts()[i] = overflow ? 0u : cell_start[c0 + i] - first;
    if (i < n) {
        msg.pos(nx, halo_cap)[i] = a.pos[first + i];
        msg.vel(nx, halo_cap)[i] = a.vel[first + i];
        msg.v0(nx, halo_cap)[i] = a.v0[first + i];
        msg.dest(nx, halo_cap)[i] = a.dest[first + i];
    }
    uint32_t* flag = side == 0 ? sig.flag_down : sig.flag_up;
    if (flag != nullptr) {
        __threadfence_system();  // this thread's peer stores are ordered before the signal
        __syncthreads();
        if (threadIdx.x == 0) {
            const uint32_t finished = atomicAdd(sig.done_count + side, 1u);
            if (finished == gridDim.x - 1) {
                sig.done_count[side] = 0;
                __threadfence_system();
                *reinterpret_cast<volatile uint32_t*>(flag) = sig.seq;
            }
        }
    }
}

// blockIdx.y = 0: ghost rows r0-2, r0-1 from `below` (right-aligned to end at H);
// blockIdx.y = 1: ghost rows r1, r1+1 from `above` (placed right after the owned agents).
// Also completes the cell table for the ghost rows and the compute / edge ranges.
// wait_seq != 0 (peer-memory transport): the strip is written by the neighbour's pack kernel; poll the
// local arrival flag until it reaches this exchange's ordinal (the neighbour may already be one ahead —
// it writes alternate slots), bounded by a 20 s time-out so that a dead peer cannot hang the GPU for good.
__global__ void __launch_bounds__(256) halo_unpack_kernel(AgentArrays a, uint32_t* __restrict__ cell_start,
                                                          uint32_t own_begin_cell, uint32_t own_end_cell, uint32_t nx,
                                                          uint32_t halo_cap, uint32_t array_cap, HaloMessage below,
                                                          HaloMessage above, int has_below, int has_above,
                                                          uint32_t* __restrict__ ranges,
                                                          uint32_t* __restrict__ error_flag,
                                                          const uint32_t* flag_below, const uint32_t* flag_above,
                                                          uint32_t wait_seq) {
    const int side = blockIdx.y;
    if ((side == 0 && !has_below) || (side == 1 && !has_above)) return;
    const HaloMessage msg = side == 0 ? below : above;
    __shared__ int s_timed_out;
    if (wait_seq != 0) {
        if (threadIdx.x == 0) {
            const volatile uint32_t* flag = side == 0 ? flag_below : flag_above;
            const unsigned long long t0 = global_timer_ns();
            int timed_out = 0;
            while (static_cast<int32_t>(*flag - wait_seq) < 0) {
                if (global_timer_ns() - t0 > 20000000000ull) {  // 20 s
                    timed_out = 1;
                    break;
                }
                __nanosleep(200);
            }
            s_timed_out = timed_out;
            __threadfence_system();
        }
        __syncthreads();
    } else if (threadIdx.x == 0) {
        s_timed_out = 0;
    }
    __syncthreads();
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    // Read the message around L1: with the peer-memory transport it was written by another GPU.
    uint32_t n = __ldcg(msg.words);
    if (s_timed_out) {
        if (i == 0) atomicOr(error_flag, kErrHaloTimeout);
        n = 0;
    }
    if (n > halo_cap) n = 0;  // cannot happen with a matching sender; never index out of bounds
    const uint32_t* rel = msg.starts();
    uint32_t base;
    if (side == 0) {
        base = halo_cap - n;
        if (i < 2 * nx) cell_start[i] = base + min(__ldcg(rel + i), n);  // entry 2nx == own_begin_cell == H already
        if (i == 0) {
            const uint32_t row_m1 = base + min(__ldcg(rel + nx), n);  // first agent of ghost row r0-1
            ranges[2 * kRangeCompute] = row_m1;
            ranges[2 * kRangeEdgeLo] = row_m1;
            ranges[2 * kRangeEdgeLo + 1] = cell_start[own_begin_cell + nx];
        }
    } else {
        base = cell_start[own_end_cell];  // H + n_owned, written by the scan
        if (base + n > array_cap) {       // host sizing bug or a burst of immigrants; drop the strip, flag it
            if (i == 0) atomicOr(error_flag, kErrHaloOverflow);
            n = 0;
        }
        if (i >= 1 && i <= 2 * nx) cell_start[own_end_cell + i] = base + min(__ldcg(rel + i), n);
        if (i == 0) {
            const uint32_t row_p1 = base + min(__ldcg(rel + nx), n);  // one past the last agent of ghost row r1
            ranges[2 * kRangeCompute + 1] = row_p1;
            ranges[2 * kRangeEdgeHi] = cell_start[own_end_cell - nx];
            ranges[2 * kRangeEdgeHi + 1] = row_p1;
        }
    }
    if (i < n) {
        a.pos[base + i] = __ldcg(msg.pos(nx, halo_cap) + i);
        a.vel[base + i] = __ldcg(msg.vel(nx, halo_cap) + i);
        a.v0[base + i] = __ldcg(msg.v0(nx, halo_cap) + i);
        a.dest[base + i] = __ldcg(msg.dest(nx, halo_cap) + i);
    }
}

// ---- observables: one pass over the owned pedestrians -----------------------
struct ObserveOut {  // device mirror of PedoniObservables' reduced fields
    unsigned int count;
    float speed_sum;
    unsigned int per_destination[16];
    float bin_vx_sum[64];
    unsigned int bin_count[64];
};

__global__ void __launch_bounds__(256) observe_kernel(AgentArrays a, const uint32_t* __restrict__ owned, uint32_t upper,
                                                      float y0, float inv_bin, uint32_t n_bins,
                                                      ObserveOut* __restrict__ out) {
    __shared__ unsigned int s_dest[16];
    __shared__ float s_vx[64];
    __shared__ unsigned int s_cnt[64];
    __shared__ float s_speed[8];
    for (int k = threadIdx.x; k < 64; k += blockDim.x) {
        s_vx[k] = 0.0f;
        s_cnt[k] = 0u;
        if (k < 16) s_dest[k] = 0u;
    }
    __syncthreads();
    const uint32_t begin = owned[0], end = owned[1];
    float speed = 0.0f;
    unsigned int mine = 0;
    for (uint32_t i = begin + blockIdx.x * blockDim.x + threadIdx.x; i < end && i - begin < upper;
         i += gridDim.x * blockDim.x) {
        const float2 v = a.vel[i], p = a.pos[i];
        speed += sqrtf(v.x * v.x + v.y * v.y);
        mine += 1;
        atomicAdd(s_dest + min(a.dest[i], 15u), 1u);
        const float b = (p.y - y0) * inv_bin;
        if (n_bins > 0 && b >= 0.0f && b < static_cast<float>(n_bins)) {
            atomicAdd(s_vx + static_cast<int>(b), v.x);
            atomicAdd(s_cnt + static_cast<int>(b), 1u);
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        speed += __shfl_down_sync(0xFFFFFFFFu, speed, d);
        mine += __shfl_down_sync(0xFFFFFFFFu, mine, d);
    }
    if ((threadIdx.x & 31) == 0) {
        s_speed[threadIdx.x >> 5] = speed;
        atomicAdd(&out->count, mine);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.0f;
        for (int w = 0; w < static_cast<int>(blockDim.x >> 5); ++w) t += s_speed[w];
        atomicAdd(&out->speed_sum, t);
    }
    for (int k = threadIdx.x; k < 64; k += blockDim.x) {
        if (k < 16 && s_dest[k]) atomicAdd(out->per_destination + k, s_dest[k]);
        if (s_cnt[k]) {
            atomicAdd(out->bin_vx_sum + k, s_vx[k]);
            atomicAdd(out->bin_count + k, s_cnt[k]);
        }
    }
}

}  // namespace pedoni
