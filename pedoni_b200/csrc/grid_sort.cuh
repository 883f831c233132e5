// grid_sort.cuh — neighbor-grid rebuild as a deterministic, stable cell-key counting sort (sm_100a),
// plus the ghost-row pack/unpack of the slab decomposition.
//
// Replaces NeighborGrid::update (neighbor_grid.rs:22-36) and the serial walk/gather of
// SocialForceModel::spawn_pedestrians (sfm.rs:58-77):
//
//   enrol      every pedestrian takes an atomic ticket on the counter of the cell it will belong to and leaves its
//              logical input index in the cell's slot row. Fused into the force kernel's epilogue for pedestrians
//              that were just integrated; key_kernel only handles freshly spawned / uploaded ones.
//   sort       ONE kernel (sort_cells_kernel): chained prefix scan over the cell populations ->
//              `neighbor_grid_indices` (sfm.rs:61-75, length cells + 1); then, per output slot, the member of the
//              cell with exactly `rank` smaller input indices — "within a cell, ascending previous index"
//              (sfm.rs:66-68), independent of the atomic arrival order, i.e. run-to-run deterministic — and one
//              gathered read / coalesced write of the 24-byte state into the other buffer.
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
    const uint32_t* d_range;  // device: [begin, end) of live entries inside the arrays; nullptr = [0, upper)
    uint32_t upper;           // host-known upper bound of (end - begin)
};

struct SortInput {
    Segment seg[kMaxSegments];
    uint32_t prefix[kMaxSegments + 1];  // exclusive prefix of `upper`
    int nseg;
    uint32_t cap[kMaxSegments];  // elements allocated per segment's arrays (bounds checks of debug builds)
    uint32_t out_cap;            // elements allocated in the rebuild's output arrays
    uint32_t* error_flag;
};

// Logical input index t -> the segment's arrays and the element index; live = false if t is outside the
// segment's live range. Segment 0 (the resident pedestrians) is indexed by the ABSOLUTE array index, t = idx:
// the force kernel's epilogue knows it without reading the layout ranges (on a slab handle the lower end of the
// compute range is only written by the ghost unpack, which runs concurrently with the interior force launch).
// Segment 1 (appended spawns) follows at prefix[1], a host bound above every resident index. Two segments,
// resolved with selects (no dynamic indexing of the kernel parameter, which would force a local-memory copy).
struct Located {
    AgentArrays a;
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
    if (second) {  // appended spawns: [0, upper) host-known, or [0, d_range[1]) when the device drew the counts
        r.idx = t - in.prefix[1];
        r.live = r.idx < (in.seg[1].d_range != nullptr ? in.seg[1].d_range[1] : in.seg[1].upper);
    } else {
        r.idx = t;
        r.live = t >= in.seg[0].d_range[0] && t < in.seg[0].d_range[1];
    }
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

// ---- Poisson arrivals drawn on the device (lib.rs:70-84, util.rs:78-89) -----------------------------------------
// The reference draws, every tick and for every periodic spawn group, count = poisson(frequency / 10) with Knuth's
// multiplication loop over fastrand::f64() (util.rs:78-89), then a position per pedestrian. With the counts drawn
// here too, a periodic-spawn scenario needs no host-side random numbers at all: the handle owns the position in the
// counter stream (SpawnStreamState), because how many numbers a Poisson draw consumes depends on its outcome.
// Stream order of one call = the order of pedoni_b200/simulator.py `Simulator._spawn`: every group's count first,
// then one uniform per pedestrian, then the desired speeds' two blocks.
struct SpawnRateDev {
    float p1_x, p1_y, p2_x, p2_y;
    uint32_t destination;
    uint32_t max_count;     // host bound used for grid sizes; a larger draw raises kErrSpawnBound (and is clamped)
    double exp_neg_lambda;  // exp(-frequency / 10), computed by the caller (same bits as the host-side loop)
};
struct SpawnStreamState {
    unsigned long long counter;       // next unused stream number
    unsigned long long counter_draw;  // first number of the current call's per-pedestrian draws
    unsigned long long spawned;       // pedestrians drawn so far
};

__global__ void poisson_counts_kernel(const SpawnRateDev* __restrict__ rates, uint32_t n_groups, unsigned long long seed,
                                      SpawnStreamState* __restrict__ st, SpawnGroupDev* __restrict__ groups,
                                      uint32_t* __restrict__ app_range, uint32_t at, uint32_t* __restrict__ error_flag) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    unsigned long long ctr = st->counter;
    uint32_t n = 0;
    for (uint32_t g = 0; g < n_groups; ++g) {
        const SpawnRateDev R = rates[g];
        // util.rs:78-89: y = 0; x = f64(); while x >= exp(-lambda) { x *= f64(); y += 1 }   (f64 = 53 random bits)
        uint32_t y = 0;
        double x = static_cast<double>(spawn_stream_u64(seed, ctr++) >> 11) * (1.0 / 9007199254740992.0);
        // (bounded: the host refuses rates whose exp(-lambda) underflows to 0, for which the reference's loop never ends)
        while (x >= R.exp_neg_lambda && y <= R.max_count) {
            x = __dmul_rn(x, static_cast<double>(spawn_stream_u64(seed, ctr++) >> 11) * (1.0 / 9007199254740992.0));
            ++y;
        }
        if (y > R.max_count) {
            atomicOr(error_flag, kErrSpawnBound);
            y = R.max_count;
        }
        groups[g] = SpawnGroupDev{R.p1_x, R.p1_y, R.p2_x, R.p2_y, R.destination, n};
        n += y;
    }
    app_range[0] = 0;
    app_range[1] = at + n;
    st->counter_draw = ctr;
    st->counter = ctr + 3ull * n;
    st->spawned += n;
}

// spawn_groups_kernel with the number of pedestrians and the stream position read from the device
__global__ void __launch_bounds__(256) spawn_groups_dev_kernel(AgentArrays out, uint32_t at, const uint32_t* __restrict__ app_range,
                                                               const SpawnGroupDev* __restrict__ groups, uint32_t n_groups,
                                                               unsigned long long seed, const SpawnStreamState* __restrict__ st) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t n = app_range[1] - at;
    if (j >= n) return;
    const unsigned long long counter = st->counter_draw;
    uint32_t g = 0;
    while (g + 1 < n_groups && groups[g + 1].first <= j) ++g;
    const SpawnGroupDev G = groups[g];
    const float u = static_cast<float>(spawn_stream_u64(seed, counter + j) >> 40) * (1.0f / 16777216.0f);
    const float px = S::add(G.p1_x, S::mul(S::sub(G.p2_x, G.p1_x), u));
    const float py = S::add(G.p1_y, S::mul(S::sub(G.p2_y, G.p1_y), u));
    const unsigned long long a = spawn_stream_u64(seed, counter + n + j), b = spawn_stream_u64(seed, counter + 2ull * n + j);
    const double pop = static_cast<double>(__popcll(a)) - 32.0;
    const double tri = (static_cast<double>(b & 0xFFFFFFFFull) - static_cast<double>(b >> 32)) * (1.0 / 4294967296.0);
    const double z = (pop + tri) * 0x1.fd5a9eebfd779p-3;
    out.pos[at + j] = make_float2(px, py);
    out.vel[at + j] = make_float2(0.0f, 0.0f);
    out.v0[at + j] = S::add(1.34f, S::mul(0.26f, static_cast<float>(z)));
    out.dest[at + j] = G.destination;
}

// ---- cell membership, written by the producers of the next rebuild -----------------------------------------
// A pedestrian ENROLS in the cell it will belong to: an atomic ticket on the cell's counter and its logical
// sort-input index t stored in the cell's slot row (kSlotsPerCell entries = one 32-byte sector per cell). Members
// beyond the row (a jam: more than 8 pedestrians on 1.96 m^2) go to a per-cell chain in an overflow list. The
// producers are the force kernel's epilogue (resident pedestrians, keyed on their just-integrated position) and
// key_kernel (freshly spawned / uploaded ones); the consumer is sort_cells_kernel, which therefore needs no
// per-pedestrian key, ticket or permutation arrays at all.
constexpr int kSlotsPerCell = 8;

struct OverflowEntry {
    uint32_t t;     // logical sort-input index of the member
    uint32_t next;  // 1-based index of the next entry of the same cell, 0 = end of chain
};

struct CellSort {
    uint32_t* cell_count;  // [cells] tickets handed out = population of the cell in the NEXT table (zeroed by the sort)
    uint32_t* slots;       // [cells][kSlotsPerCell] t of the member holding ticket j < kSlotsPerCell
    uint32_t* ovf_head;    // [cells] chain of the members with ticket >= kSlotsPerCell (reset by the sort)
    OverflowEntry* ovf;    // [ovf_cap]
    uint32_t* ovf_count;   // entries in use (reset by the sort)
    uint32_t ovf_cap;      // = capacity of the agent arrays: every pedestrian could overflow
    uint32_t n_cells;      // cells of the local table (bounds checks of debug builds)
};

__device__ __forceinline__ void enroll(const CellSort& cs, uint32_t key, uint32_t t, uint32_t* error_flag) {
    if (key >= kKeyFirstSpecial) return;  // dropped: outside the grid, despawned, or another slab's row
    if (!PEDONI_IN_BOUNDS(key < cs.n_cells && t < kKeyFirstSpecial, error_flag)) return;
    const uint32_t ticket = atomicAdd(cs.cell_count + key, 1u);
    if (ticket < static_cast<uint32_t>(kSlotsPerCell)) {
        cs.slots[static_cast<size_t>(key) * kSlotsPerCell + ticket] = t;
    } else {
        const uint32_t e = atomicAdd(cs.ovf_count, 1u);
        if (e < cs.ovf_cap) {
            cs.ovf[e].t = t;
            cs.ovf[e].next = atomicExch(cs.ovf_head + key, e + 1u);
        } else {
            atomicOr(error_flag, kErrSortOverflow);  // cannot happen with ovf_cap = array capacity
        }
    }
}

// ---- key: only for logical indices in [t_begin, t_end) that the force kernel has not enrolled ------------
__global__ void __launch_bounds__(256) key_kernel(SortInput in, uint32_t t_begin, uint32_t t_end, GridView g,
                                                  FieldView f, CellSort cs, uint32_t* __restrict__ error_flag,
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
    enroll(cs, sort_key(g, f, pos, l.a.dest[l.idx], error_flag, arrived, row >= g.own_row0 && row < g.own_row1), t,
           error_flag);
}

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

struct ScanLayout {  // what the sort publishes besides the table (see RangeId)
    uint32_t own_begin_cell, own_end_cell, nx;
    int has_below, has_above;
    uint32_t* ranges;
    unsigned long long* host_slot;
    uint32_t tick;
};

constexpr unsigned long long kScanAggregate = 1ull << 62;
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

// ---- the rebuild: ONE kernel, cell-centric ----------------------------------------------------------------------
// Replaces NeighborGrid::update + the serial walk of spawn_pedestrians (neighbor_grid.rs:22-36, sfm.rs:58-77).
// Persistent CTAs take tiles of kSortTile consecutive cells by ticket; a thread owns kSortItems consecutive cells.
// Per tile:
//   1. load (and zero, for the next tick) the cell populations and request the cells' slot rows; block scan; publish
//      the tile aggregate in a tick-tagged status word and sum ALL predecessors' aggregates in one round of loads
//      (tickets are handed out in order, so every predecessor is held by a running CTA or done: the spin cannot
//      deadlock) -> `neighbor_grid_indices` (sfm.rs:61-75) for the tile's cells;
//   2. warp by warp, 32 cells at a time: every cell's slot row is sorted ("within a cell, ascending previous index",
//      sfm.rs:66-68; independent of the order in which the atomics handed out the tickets, hence run-to-run
//      deterministic and identical across slab decompositions) and its members are staged by output slot;
//   3. one pass over the staged indices moves the 24-byte state: gathered reads (near-sequential: a pedestrian moves
//      less than a cell per tick), fully coalesced writes — the tile's pedestrians are one contiguous output range.
//   The same pass writes the layout ranges that depend on the owned rows and publishes the owned population to the
//   host (one aligned 64-bit store to pinned memory, tick << 32 | n_owned).
// Cells holding more than kSlotsPerCell pedestrians find their extra members by walking the cell's overflow chain:
// O(n^2) per cell, bounded, exact — a dense jam costs a few hundred instructions per pedestrian, nothing more.
#ifndef PEDONI_SORT_THREADS
#define PEDONI_SORT_THREADS 512
#endif
constexpr int kSortCtaThreads = PEDONI_SORT_THREADS;
constexpr int kSortItems = 8;
constexpr int kSortTile = kSortCtaThreads * kSortItems;  // cells per tile
#ifndef PEDONI_SORT_UNROLL
#define PEDONI_SORT_UNROLL 2  // pedestrians in flight per thread in the move pass (B200, 10 M: 2 -> 0.196 ms, 4 -> 0.207)
#endif
#ifndef PEDONI_SORT_STAGE
#define PEDONI_SORT_STAGE (PEDONI_SORT_THREADS * 20)  // output slots staged per round: a tile of 4096 cells holds ~8 000 pedestrians at 1 /m^2
#endif
constexpr int kSortStage = PEDONI_SORT_STAGE;
constexpr size_t kSortSmemBytes = sizeof(uint32_t) * (kSortCtaThreads * kSlotsPerCell + kSortTile + 4 + kSortStage);
#ifndef PEDONI_SORT_MIN_BLOCKS
#define PEDONI_SORT_MIN_BLOCKS 2
#endif

struct SortScratch {
    unsigned long long* tile_status;  // [n_tiles] tick-tagged aggregates (never reset)
    uint32_t* tile_ticket;            // [2] tile counters; launch k uses [k & 1] and zeroes the other one
    uint32_t* done_count;             // CTAs of this launch that have finished (the last one resets the overflow list)
    uint32_t n_tiles;
    uint32_t parity;
    uint32_t items;  // cells per thread of this launch (1 .. kSortItems): a tile is kSortCtaThreads * items cells
};

// The member of `cell` with exactly `rank` smaller logical indices among its n members.
__device__ __forceinline__ uint32_t select_member(const CellSort& cs, uint32_t cell, uint32_t n, uint32_t rank, uint4 lo,
                                                  uint4 hi) {
    uint32_t m[kSlotsPerCell] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
#pragma unroll
    for (int a = 0; a < kSlotsPerCell; ++a)
        if (static_cast<uint32_t>(a) >= n) m[a] = 0xFFFFFFFFu;  // unused slots hold stale values
    uint32_t t = m[0];
    if (n <= static_cast<uint32_t>(kSlotsPerCell)) {
#pragma unroll
        for (int a = 1; a < kSlotsPerCell; ++a) {
            if (static_cast<uint32_t>(a) < n) {  // uniform over the threads that share the cell
                uint32_t smaller = 0;
#pragma unroll
                for (int b = 0; b < kSlotsPerCell; ++b) smaller += (m[b] < m[a]) ? 1u : 0u;
                if (smaller == rank) t = m[a];
            }
        }
        return t;  // m[0] unless another member has the rank
    }
    // jam: the members beyond the slot row hang on the cell's chain
    const uint32_t head = cs.ovf_head[cell];
    auto smaller_than = [&](uint32_t x) {
        uint32_t c = 0;
#pragma unroll
        for (int b = 0; b < kSlotsPerCell; ++b) c += (m[b] < x) ? 1u : 0u;
        uint32_t guard = n;
        for (uint32_t e = head; e != 0 && guard-- != 0; e = cs.ovf[e - 1].next) c += (cs.ovf[e - 1].t < x) ? 1u : 0u;
        return c;
    };
#pragma unroll
    for (int a = 0; a < kSlotsPerCell; ++a)  // unrolled: m[] stays in registers
        if (smaller_than(m[a]) == rank) return m[a];
    uint32_t guard = n;
    for (uint32_t e = head; e != 0 && guard-- != 0; e = cs.ovf[e - 1].next) {
        const uint32_t x = cs.ovf[e - 1].t;
        if (smaller_than(x) == rank) return x;
    }
    return t;  // unreachable with a consistent chain
}

// kMinBlocks / kSortUnroll: 2 CTAs per SM with 2 pedestrians in flight per thread is fastest on a 10 M crowd (0.188 ms
// against 0.230 ms); a small problem (a slab of an 8-GPU run: one tile per CTA, latency only) gains from a third
// CTA per SM, i.e. fewer cells per thread, at one pedestrian in flight (1.25 M: 39 us against 44 us). The host picks.
template <int kMinBlocks, int kSortUnroll>
__global__ void __launch_bounds__(kSortCtaThreads, kMinBlocks)
    sort_cells_kernel(SortInput in, CellSort cs, uint32_t n_cells, uint32_t offset, uint32_t* __restrict__ cell_start,
                      SortScratch sc, ScanLayout L, AgentArrays out) {
    // dynamic shared memory, see kSortSmemBytes
    extern __shared__ __align__(16) uint32_t sort_smem[];
    uint32_t* const s_sorted = sort_smem;                                 // per warp: 32 sorted slot rows
    uint32_t* const s_start = s_sorted + kSortCtaThreads * kSlotsPerCell;  // [kSortTile + 1] starts of the tile's cells,
                                                                           // relative to the tile's first pedestrian
    uint32_t* const s_member = s_start + kSortTile + 4;  // [kSortStage] logical input index going to each staged slot
    __shared__ uint32_t s_tile, s_total, s_prefix;
    const uint32_t tid = threadIdx.x;
    uint32_t* const ticket = sc.tile_ticket + sc.parity;
    if (blockIdx.x == 0 && tid == 0) sc.tile_ticket[sc.parity ^ 1u] = 0;  // the next launch's counter, idle now

    for (;;) {
        __syncthreads();  // the previous tile's shared state is no longer read
        if (tid == 0) s_tile = atomicAdd(ticket, 1u);
        __syncthreads();
        const uint32_t tile = s_tile;
        if (tile >= sc.n_tiles) break;

        // ---- 1. populations -> starts
        const uint32_t items = sc.items, tile_cells = kSortCtaThreads * items;
        const uint32_t base = tile * tile_cells + tid * items;
        uint32_t v[kSortItems];
        if (items == kSortItems && base + kSortItems <= n_cells) {
            uint4* p = reinterpret_cast<uint4*>(cs.cell_count + base);
#pragma unroll
            for (int q = 0; q < kSortItems / 4; ++q) {
                const uint4 a = p[q];
                v[4 * q] = a.x, v[4 * q + 1] = a.y, v[4 * q + 2] = a.z, v[4 * q + 3] = a.w;
                p[q] = make_uint4(0u, 0u, 0u, 0u);
            }
        } else {
#pragma unroll
            for (int k = 0; k < kSortItems; ++k) {
                const bool mine = static_cast<uint32_t>(k) < items && base + k < n_cells;
                v[k] = mine ? cs.cell_count[base + k] : 0u;
                if (mine) cs.cell_count[base + k] = 0u;
            }
        }
        uint32_t sum = 0;
#pragma unroll
        for (int k = 0; k < kSortItems; ++k) sum += v[k];
        const uint32_t excl = block_exclusive_scan(sum, &s_total);
        {
            volatile unsigned long long* status = sc.tile_status;
            const unsigned long long tag = scan_status(kScanAggregate, L.tick, 0u) >> 32;
            if (tid == 0) {
                s_prefix = 0;
                status[tile] = scan_status(kScanAggregate, L.tick, s_total);
            }
            __syncthreads();
            uint32_t part = 0;
            for (uint32_t j = tid; j < tile; j += kSortCtaThreads) {
                unsigned long long st;
                do {
                    st = status[j];
                } while ((st >> 32) != tag);
                part += static_cast<uint32_t>(st);
            }
            part = __reduce_add_sync(0xFFFFFFFFu, part);
            if ((tid & 31) == 0 && part != 0) atomicAdd(&s_prefix, part);
        }
        __syncthreads();
        const uint32_t tile_base = offset + s_prefix;  // index of the tile's first pedestrian in the new arrays
        {
            uint32_t run = excl;
            uint32_t start[kSortItems];
#pragma unroll
            for (int k = 0; k < kSortItems; ++k) {
                start[k] = tile_base + run;
                if (static_cast<uint32_t>(k) < items) {
                    s_start[tid * items + k] = run;
                    if (base + k < n_cells) publish_cell_start(L, base + k, tile_base + run);
                }
                run += v[k];
            }
            if (items == kSortItems && base + kSortItems <= n_cells) {
                // 16-byte stores: scalar ones reach L2 as one partial-sector write per cell
                uint4* p = reinterpret_cast<uint4*>(cell_start + base);
#pragma unroll
                for (int q = 0; q < kSortItems / 4; ++q)
                    p[q] = make_uint4(start[4 * q], start[4 * q + 1], start[4 * q + 2], start[4 * q + 3]);
            } else {
#pragma unroll
                for (int k = 0; k < kSortItems; ++k)
                    if (static_cast<uint32_t>(k) < items && base + k < n_cells) cell_start[base + k] = start[k];
            }
            if (base < n_cells && base + items >= n_cells) {  // the table's closing entry
                cell_start[n_cells] = tile_base + run;
                publish_cell_start(L, n_cells, tile_base + run);
            }
            if (tid == 0) s_start[tile_cells] = s_total;
        }
        __syncthreads();

        // ---- 2. who goes where: for every output slot of the tile, the logical input index of its pedestrian.
        // A warp takes 32 consecutive cells at a time (lane = cell). Each lane loads its cell's slot row (a warp
        // reads 1 KB, contiguous; the next chunk's rows are requested before this chunk is finished), sorts it
        // (19 compare-exchanges: "within a cell, ascending previous index", sfm.rs:66-68; independent of the order
        // in which the atomics handed out the tickets) and parks it in the warp's slice of shared memory. The cells'
        // output range is contiguous; each round of 32 outputs finds its cell with a 5-step search over the lanes'
        // starts (shuffles) and stages member number `rank` of the sorted row.
        // ---- 3. the move, as a separate pass over the staged indices: nothing but gathered loads and coalesced
        // stores, kSortUnroll pedestrians per thread in flight.
        // A tile holding more than kSortStage pedestrians (a jam) takes several rounds of 2 + 3.
        const uint32_t total = s_total;
        for (uint32_t seg_lo = 0; seg_lo < total; seg_lo += kSortStage) {
            const uint32_t seg_hi = min(total, seg_lo + static_cast<uint32_t>(kSortStage));
            {
                constexpr uint32_t kFull = 0xFFFFFFFFu;
                const uint32_t kCellsPerWarp = 32u * items;  // the cells of this warp's threads
                const int kChunks = static_cast<int>(items);
                const uint32_t lane = tid & 31u, warp = tid >> 5;
                uint32_t* const sorted = s_sorted + warp * 32 * kSlotsPerCell;
                auto load_row = [&](int chunk, uint32_t& s, uint32_t& e, uint4& lo, uint4& hi) {
                    const uint32_t c = warp * kCellsPerWarp + chunk * 32 + lane;
                    s = s_start[c];
                    e = s_start[c + 1];
                    const uint32_t n = e - s;
                    const uint4* row = reinterpret_cast<const uint4*>(
                        cs.slots + (static_cast<size_t>(tile) * tile_cells + c) * kSlotsPerCell);
                    lo = n > 0 ? row[0] : make_uint4(0u, 0u, 0u, 0u);
                    hi = n > 4 ? row[1] : make_uint4(0u, 0u, 0u, 0u);
                };
                uint32_t s, e;
                uint4 lo, hi;
                load_row(0, s, e, lo, hi);
#pragma unroll 1
                for (int chunk = 0; chunk < kChunks; ++chunk) {
                    const uint32_t c0 = warp * kCellsPerWarp + chunk * 32;
                    const uint32_t n_mine = e - s;
                    if (__ballot_sync(kFull, n_mine != 0) == 0) {  // 32 empty cells (sparse crowds): nothing to stage
                        if (chunk + 1 < kChunks) load_row(chunk + 1, s, e, lo, hi);
                        continue;
                    }
                    {
                        uint32_t m[kSlotsPerCell] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
#pragma unroll
                        for (int a = 0; a < kSlotsPerCell; ++a)
                            if (static_cast<uint32_t>(a) >= n_mine) m[a] = 0xFFFFFFFFu;  // unused slots hold stale values
#define PEDONI_CE(i, j)                                              \
    {                                                                \
        const uint32_t lo_ = min(m[i], m[j]), hi_ = max(m[i], m[j]); \
        m[i] = lo_;                                                  \
        m[j] = hi_;                                                  \
    }
                        PEDONI_CE(0, 1) PEDONI_CE(2, 3) PEDONI_CE(4, 5) PEDONI_CE(6, 7) PEDONI_CE(0, 2) PEDONI_CE(1, 3)
                        PEDONI_CE(4, 6) PEDONI_CE(5, 7) PEDONI_CE(1, 2) PEDONI_CE(5, 6) PEDONI_CE(0, 4) PEDONI_CE(3, 7)
                        PEDONI_CE(1, 5) PEDONI_CE(2, 6) PEDONI_CE(1, 4) PEDONI_CE(3, 6) PEDONI_CE(2, 4) PEDONI_CE(3, 5)
                        PEDONI_CE(3, 4)
#undef PEDONI_CE
                        __syncwarp();  // the previous chunk's rows are no longer read
                        uint4* dst = reinterpret_cast<uint4*>(sorted + lane * kSlotsPerCell);
                        dst[0] = make_uint4(m[0], m[1], m[2], m[3]);
                        dst[1] = make_uint4(m[4], m[5], m[6], m[7]);
                        __syncwarp();
                    }
                    const uint32_t s_cur = s, e_cur = e;
                    if (chunk + 1 < kChunks) load_row(chunk + 1, s, e, lo, hi);  // in flight during the rest
                    const uint32_t first = __shfl_sync(kFull, s_cur, 0), last = __shfl_sync(kFull, e_cur, 31);
                    // the part of this chunk's output range [first, last) that lies in the staged segment
                    const uint32_t from = max(first, seg_lo), to = min(last, seg_hi);
#pragma unroll 1
                    for (uint32_t r0 = from; r0 < to; r0 += 32) {
                        const uint32_t r = r0 + lane;  // tile-relative output slot of this lane
                        uint32_t owner = 0;            // the last lane whose cell starts at or before r
#pragma unroll
                        for (uint32_t step = 16; step >= 1; step >>= 1) {
                            const uint32_t v_ = __shfl_sync(kFull, s_cur, owner + step);
                            if (v_ <= r) owner += step;
                        }
                        const uint32_t cs0 = __shfl_sync(kFull, s_cur, owner), cs1 = __shfl_sync(kFull, e_cur, owner);
                        if (r < to) {
                            const uint32_t rank = r - cs0, n = cs1 - cs0;
                            uint32_t t;
                            if (n <= static_cast<uint32_t>(kSlotsPerCell)) {
                                t = sorted[owner * kSlotsPerCell + rank];
                            } else {  // jam: members beyond the row hang on the cell's chain
                                const uint32_t cell = tile * tile_cells + c0 + owner;
                                const uint4* row =
                                    reinterpret_cast<const uint4*>(cs.slots + static_cast<size_t>(cell) * kSlotsPerCell);
                                t = select_member(cs, cell, n, rank, row[0], row[1]);
                            }
                            s_member[r - seg_lo] = t;
                        }
                    }
                }
            }
            __syncthreads();
            for (uint32_t r0 = seg_lo + tid; r0 < seg_hi; r0 += kSortCtaThreads * kSortUnroll) {
                float2 pos[kSortUnroll], vel[kSortUnroll];
                float v0[kSortUnroll];
                uint32_t dest[kSortUnroll];
#pragma unroll
                for (int u = 0; u < kSortUnroll; ++u) {
                    const uint32_t r = r0 + u * kSortCtaThreads;
                    if (r < seg_hi) {
                        const Located l = locate(in, s_member[r - seg_lo]);
                        if (!PEDONI_IN_BOUNDS(l.live && l.idx < in.cap[s_member[r - seg_lo] >= in.prefix[1] ? 1 : 0] &&
                                                  tile_base + r < in.out_cap, in.error_flag)) {
                            pos[u] = vel[u] = make_float2(0.f, 0.f), v0[u] = 0.f, dest[u] = 0u;
                            continue;
                        }
                        pos[u] = l.a.pos[l.idx];
                        vel[u] = l.a.vel[l.idx];
                        v0[u] = l.a.v0[l.idx];
                        dest[u] = l.a.dest[l.idx];
                    }
                }
#pragma unroll
                for (int u = 0; u < kSortUnroll; ++u) {
                    const uint32_t r = r0 + u * kSortCtaThreads;
                    if (r < seg_hi && PEDONI_IN_BOUNDS(tile_base + r < in.out_cap, in.error_flag)) {
                        const uint32_t dst = tile_base + r;
                        out.pos[dst] = pos[u];
                        out.vel[dst] = vel[u];
                        out.v0[dst] = v0[u];
                        out.dest[dst] = dest[u];
                    }
                }
            }
            if (seg_hi < total) __syncthreads();  // the stage is reused by the next segment
        }
        // ---- 4. chains of the cells that overflowed have been walked by every thread that needed them
        __syncthreads();
#pragma unroll
        for (int k = 0; k < kSortItems; ++k)
            if (v[k] > static_cast<uint32_t>(kSlotsPerCell) && static_cast<uint32_t>(k) < items && base + k < n_cells)
                cs.ovf_head[base + k] = 0u;
    }
    // the last CTA to leave empties the overflow list for the next tick's producers
    if (tid == 0) {
        __threadfence();
        if (atomicAdd(sc.done_count, 1u) == gridDim.x - 1) {
            *cs.ovf_count = 0u;
            *sc.done_count = 0u;
        }
    }
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
        // The CTA's peer stores are ordered before the signal by ONE system-scope fence: the barrier makes every
        // thread's stores visible to thread 0 (CTA scope), whose fence is cumulative over what it has observed —
        // the pattern of a grid-wide barrier. (A fence per thread cost most of this kernel's 16 us.)
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence_system();
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
