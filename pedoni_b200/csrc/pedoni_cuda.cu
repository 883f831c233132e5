// pedoni_cuda.cu — C ABI (include/pedoni_cuda.h) over the sm_100a kernels in grid_sort.cuh and
// force.cuh. One PedoniModel = the reference's `SocialForceModel` (sfm.rs:18-24) living on one GPU,
// or one row slab of it.
//
// Data layout in HBM (all SoA, coalesced):
//   two agent buffers buf[0/1], each {pos float2[cap], vel float2[cap], v0 float[cap], dest u32[cap]}
//   = 24 B/agent/buffer, the reference's PedestrianVec (sfm.rs:26-33). `cur` holds the live state; rebuild
//   sorts cur (+ appended spawns) -> other, step integrates cur -> other; both swap. Per cell: the population
//   counter and the slot row the next rebuild's members enrol in (written by the force epilogue, 8 x u32 = one
//   sector per cell, + an overflow list for jams), and cell_start u32[cells + 1] (neighbor_grid_indices).
//   Field maps f32 row-major (field.rs:194-205), uploaded once.
// Streams: `stream` (rebuild, interior force), `edge_stream` (slab handles, highest priority: ghost
//   exchange / unpack / force on the rows next to a slab boundary), `dl_stream` (pipelined download).
// Populations stay on the device (d_ranges, double-buffered with the layout); the host only tracks
//   upper bounds for grid sizes, refreshed from a pinned word the scan publishes, so spawn / rebuild /
//   step never synchronise. count / download / cell_table block.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/pedoni_cuda.h"
#include "force.cuh"
#include "slab_comm.hpp"

using namespace pedoni;

namespace {

thread_local std::string g_create_error;

// Indices of PedoniKernelTimes / PedoniLaunchRecord.kind. kGather is the sort kernel (prefix scan + reorder in one
// launch); kHistogram / kScan / kScatter were separate launches in round 1, are never timed now and keep their slots
// so that the struct layout of ABI 2 callers stays valid.
enum KernelKind { kKey = 0, kHistogram, kScan, kScatter, kGather, kForce, kComm, kForceEdge, kPack, kNumKinds };

struct TimedLaunch {
    int kind;
    int stream_id;  // 0: main stream, 1: edge stream
    cudaEvent_t start, stop;
    uint64_t agents;
};

}  // namespace

struct PedoniModel {
    int device = 0;
    cudaStream_t stream = nullptr;       // main: rebuild, interior force
    cudaStream_t edge_stream = nullptr;  // slab handles: halo exchange, ghost unpack, edge force
    bool own_stream = false;
    int math_mode = 0;
    bool use_distance_map = true;

    GridView grid{};
    FieldView field{};
    float* d_distance = nullptr;
    float* d_potential = nullptr;
    // fast math: the maps once more, tiled into one 2D CUDA array behind a point-sampled texture (FieldView::atlas)
    cudaArray_t field_atlas = nullptr;
    bool field_tex = false;
    uint32_t* d_far_mask = nullptr;  // FieldView::far_mask
    unsigned long long far_cells = 0, far_blocks_total = 0;  // marked / all blocks of the mask
    float* d_edges = nullptr;
    int n_obstacles = 0;
    uint32_t n_cells = 0;  // local table cells
    uint32_t own_begin_cell = 0, own_end_cell = 0;

    // slab decomposition
    int slab_rank = 0, slab_count = 1;
    bool has_below = false, has_above = false;
    uint32_t halo_cap = 0;     // H: capacity (agents) of one two-row ghost strip / message
    uint32_t array_offset = 0; // first owned agent sits at this index (H if has_below)
    size_t msg_bytes = 0;
    uint32_t* d_send_dn = nullptr;     // staging: first two owned rows, for the slab below (NCCL / in-process)
    uint32_t* d_send_up = nullptr;     // staging: last two owned rows, for the slab above
    // Receive arena, one allocation (one CUDA IPC handle): [control words | below slot 0 | below slot 1 |
    // above slot 0 | above slot 1]. Exchange k uses slot k % 2, so a neighbour that is already one tick
    // ahead never overwrites a strip that has not been unpacked yet.
    unsigned char* d_arena = nullptr;
    size_t arena_bytes = 0, slot_bytes = 0;
    uint32_t exchange_seq = 0;         // rebuilds that exchanged ghosts so far (identical on every slab)
    enum Transport { kTransportNone = 0, kTransportNccl, kTransportPeer } transport = kTransportNone;
    unsigned char* peer_arena_below = nullptr;  // the arenas of slab rank-1 / rank+1, mapped through CUDA IPC
    unsigned char* peer_arena_above = nullptr;
    pedoni::SlabComm* comm = nullptr;
    bool halo_pending = false;   // rebuilt, ghosts not exchanged yet (in-process transport)
    bool halo_inflight = false;  // exchange enqueued on edge_stream, main has not waited on ev_halo
    cudaEvent_t ev_packed = nullptr, ev_halo = nullptr, ev_edge = nullptr, ev_peer = nullptr, ev_sorted = nullptr;

    AgentArrays buf[2]{};
    uint32_t cap = 0;          // elements per array
    int cur = 0;
    uint32_t owned_upper = 0;  // host upper bound of owned agents in buf[cur]
    AgentArrays app{};         // appended spawns, not yet rebuilt
    void* d_spawn_groups = nullptr;  // SpawnGroupDev table of pedoni_spawn_groups
    uint32_t spawn_groups_cap = 0;
    // device-side Poisson arrivals (pedoni_spawn_poisson)
    SpawnRateDev* d_spawn_rates = nullptr;
    uint32_t spawn_rates_cap = 0;
    SpawnStreamState* d_spawn_stream = nullptr;  // the handle's position in the counter stream
    unsigned long long spawn_seed = 0;
    uint32_t* d_app_range = nullptr;  // device [0, n): live entries of the appended spawns when the DEVICE drew the count
    bool app_on_device = false;       // app_n is an upper bound, the count lives in d_app_range
    uint32_t app_cap = 0, app_n = 0;
    // Spawn staging: the caller's arrays (pageable or pinned) are copied into a pinned ring slot on the host and
    // travel from there, so pedoni_spawn never waits for the stream and never reads a borrowed buffer after it
    // returns. A slot is reused once the copies that read it have completed (event per slot).
    static constexpr int kStageSlots = 4;
    static constexpr size_t kStageBytes = 1u << 20;
    unsigned char* h_stage[kStageSlots] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev_stage[kStageSlots] = {nullptr, nullptr, nullptr, nullptr};
    bool stage_busy[kStageSlots] = {false, false, false, false};
    int stage_next = 0;

    uint32_t* d_slots = nullptr;         // [cells][kSlotsPerCell] members of the next table (CellSort)
    uint32_t* d_ovf_head = nullptr;      // [cells]
    OverflowEntry* d_ovf = nullptr;      // [ovf_cap]
    uint32_t* d_ovf_count = nullptr;
    uint32_t ovf_cap = 0;
    uint32_t* d_sort_done = nullptr;
    uint32_t sort_launches = 0;
    int sm_count = 148;
    uint32_t* d_cell_count = nullptr;
    uint32_t* d_cell_start = nullptr;
    unsigned long long* d_tile_status = nullptr;  // chained-scan status words (tick-tagged, never reset)
    uint32_t* d_tile_ticket = nullptr;            // [2], alternating per sort launch
    uint32_t n_tiles = 0;
    // [2][kNumRanges][2], see RangeId. Double-buffered: the sort kernel's scan publishes the NEXT layout while its
    // index and move passes still locate the sort input through the current one.
    uint32_t* d_ranges = nullptr;
    int rcur = 0;
    uint32_t* d_error = nullptr;
    unsigned long long* d_updates = nullptr;
    unsigned long long* d_arrived = nullptr;  // [16] cumulative arrivals by destination
    ObserveOut* d_observe = nullptr;
    uint64_t launches = 0;  // kernels launched by this handle

    // pinned, mapped: [0] = tick << 32 | n_owned (published after every rebuild)
    unsigned long long* h_pub = nullptr;
    unsigned long long* h_pub_dev = nullptr;
    uint32_t tick = 0;                  // rebuilds (and state resets) so far
    uint64_t inflow_cum[64] = {0};      // inflow_cum[t % 64]: upper bound of agents added up to tick t

    bool keys_fresh = false;   // the compute range of buf[cur] has enrolled in the next table (force epilogue)
    bool table_valid = false;  // d_cell_start indexes buf[cur]
    bool ever_rebuilt = false;

    // pipelined list_pedestrians (pedoni_download_begin / _end)
    cudaStream_t dl_stream = nullptr;
    // Two downloads may be in flight (begin k, begin k+1, end k, ...): while the host widens the byte-sized
    // destinations of tick k, tick k+1's copy keeps PCIe busy. pedoni_download_end completes the oldest.
    struct DownloadSlot {
        cudaEvent_t ev_snap = nullptr, ev_done = nullptr;
        float2* d_pos = nullptr;
        uint32_t* d_dest = nullptr;   // destinations as stored, or
        uint8_t* d_dest8 = nullptr;   // one byte each when there are at most 256 potential maps: packed on the
        uint8_t* h_dest8 = nullptr;   // device, staged in pinned memory, widened into the caller's array by _end
        uint32_t* d_range = nullptr;  // owned [begin, end) copied on the main stream with the snapshot
        uint32_t* h_range = nullptr;  // pinned
        uint32_t cap = 0;             // elements of the snapshot buffers
        uint32_t* user_dest = nullptr;
        bool widen = false;           // _end widens h_dest8 into user_dest
        uint32_t user_cap = 0;
        bool inflight = false;
    } dl[2];
    uint32_t dl_head = 0, dl_count = 0;  // oldest slot in flight, number in flight

    bool profiling = false;
    std::vector<TimedLaunch> timed;
    std::vector<PedoniLaunchRecord> timeline;  // per-launch offsets since the last pedoni_timer_begin (profiling on)
    bool timer_armed = false;
    std::vector<cudaEvent_t> event_pool;
    double acc_ms[kNumKinds] = {0};
    uint64_t acc_launches[kNumKinds] = {0};
    uint64_t acc_force_agents = 0;
    cudaEvent_t timer_start = nullptr, timer_stop = nullptr;

    std::string last_error;

    // arena layout
    static constexpr size_t kArenaControlBytes = 256;  // [0] flag_below, [1] flag_above, [4..5] pack counters
    uint32_t* flag_below(unsigned char* arena) const { return reinterpret_cast<uint32_t*>(arena); }
    uint32_t* flag_above(unsigned char* arena) const { return reinterpret_cast<uint32_t*>(arena) + 1; }
    uint32_t* pack_counters() const { return reinterpret_cast<uint32_t*>(d_arena) + 4; }
    uint32_t* recv_below(unsigned char* arena, uint32_t slot) const {
        return reinterpret_cast<uint32_t*>(arena + kArenaControlBytes + slot * slot_bytes);
    }
    uint32_t* recv_above(unsigned char* arena, uint32_t slot) const {
        return reinterpret_cast<uint32_t*>(arena + kArenaControlBytes + (2 + slot) * slot_bytes);
    }
    uint32_t n_sides() const { return (has_below ? 1u : 0u) + (has_above ? 1u : 0u); }
    uint32_t compute_upper() const { return owned_upper + n_sides() * halo_cap; }
    // host bound above the array index of every pedestrian this handle integrates (owned + the ghost row above)
    uint32_t resident_hi() const { return array_offset + owned_upper + (has_above ? halo_cap : 0u); }
    CellSort cell_sort() const { return CellSort{d_cell_count, d_slots, d_ovf_head, d_ovf, d_ovf_count, ovf_cap, n_cells}; }
    uint32_t* ranges(int which) const { return d_ranges + which * 2 * kNumRanges; }
    const uint32_t* range(int id) const { return ranges(rcur) + 2 * id; }
};

namespace {

int fail(PedoniModel* m, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (m)
        m->last_error = buf;
    else
        g_create_error = buf;
    return code;
}

#define CUDA_TRY(m, expr)                                                                                  \
    do {                                                                                                   \
        cudaError_t err__ = (expr);                                                                        \
        if (err__ != cudaSuccess)                                                                          \
            return fail(m, PEDONI_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(err__), __FILE__, \
                        __LINE__);                                                                         \
    } while (0)

inline uint32_t div_up(uint32_t a, uint32_t b) { return (a + b - 1) / b; }

// f32 helpers with the reference's (glam) formulations, evaluated on the host in plain IEEE fp32
// (this TU is compiled with -ffp-contract=off for the host pass).
struct HVec2 {
    float x, y;
};

cudaError_t alloc_agents(AgentArrays& a, uint32_t cap) {
    cudaError_t e;
    // + 2: the force kernel's bulk copies round their windows outward to 16-byte boundaries
    if ((e = cudaMalloc(&a.pos, sizeof(float2) * ((size_t)cap + 2))) != cudaSuccess) return e;
    if ((e = cudaMalloc(&a.vel, sizeof(float2) * ((size_t)cap + 2))) != cudaSuccess) return e;
    if ((e = cudaMalloc(&a.v0, sizeof(float) * (size_t)cap)) != cudaSuccess) return e;
    if ((e = cudaMalloc(&a.dest, sizeof(uint32_t) * (size_t)cap)) != cudaSuccess) return e;
    return cudaSuccess;
}
void free_agents(AgentArrays& a) {
    cudaFree(a.pos);
    cudaFree(a.vel);
    cudaFree(a.v0);
    cudaFree(a.dest);
    a = AgentArrays{};
}
cudaError_t copy_agents(AgentArrays& dst, const AgentArrays& src, uint32_t n, cudaStream_t s) {
    cudaError_t e;
    if (n == 0) return cudaSuccess;
    if ((e = cudaMemcpyAsync(dst.pos, src.pos, sizeof(float2) * (size_t)n, cudaMemcpyDeviceToDevice, s))) return e;
    if ((e = cudaMemcpyAsync(dst.vel, src.vel, sizeof(float2) * (size_t)n, cudaMemcpyDeviceToDevice, s))) return e;
    if ((e = cudaMemcpyAsync(dst.v0, src.v0, sizeof(float) * (size_t)n, cudaMemcpyDeviceToDevice, s))) return e;
    if ((e = cudaMemcpyAsync(dst.dest, src.dest, sizeof(uint32_t) * (size_t)n, cudaMemcpyDeviceToDevice, s))) return e;
    return cudaSuccess;
}

int sync_all(PedoniModel* m) {
    CUDA_TRY(m, cudaStreamSynchronize(m->stream));
    if (m->edge_stream) CUDA_TRY(m, cudaStreamSynchronize(m->edge_stream));
    return PEDONI_OK;
}

// Grow the state buffers (keeping buf[cur]: the layout uses absolute indices) and the overflow list of the cell
// slots (keeping its entries: a step may already have enrolled pedestrians in the next table). Blocking; only runs
// when a host upper bound outgrows the allocation.
int ensure_capacity(PedoniModel* m, uint32_t need_arrays) {
    if (need_arrays <= m->cap) return PEDONI_OK;
    int rc = sync_all(m);
    if (rc != PEDONI_OK) return rc;
    uint32_t ncap = std::max<uint32_t>(need_arrays, m->cap + m->cap / 2);
    ncap = std::max<uint32_t>(ncap, 1024);
    for (int b = 0; b < 2; ++b) {
        AgentArrays fresh{};
        CUDA_TRY(m, alloc_agents(fresh, ncap));
        if (b == m->cur && m->cap > 0) CUDA_TRY(m, copy_agents(fresh, m->buf[b], m->cap, m->stream));
        CUDA_TRY(m, cudaStreamSynchronize(m->stream));
        free_agents(m->buf[b]);
        m->buf[b] = fresh;
    }
    OverflowEntry* fresh_ovf = nullptr;
    CUDA_TRY(m, cudaMalloc(&fresh_ovf, sizeof(OverflowEntry) * (size_t)ncap));
    if (m->ovf_cap > 0)
        CUDA_TRY(m, cudaMemcpyAsync(fresh_ovf, m->d_ovf, sizeof(OverflowEntry) * (size_t)m->ovf_cap,
                                    cudaMemcpyDeviceToDevice, m->stream));
    CUDA_TRY(m, cudaStreamSynchronize(m->stream));
    cudaFree(m->d_ovf);
    m->d_ovf = fresh_ovf;
    m->ovf_cap = ncap;
    m->cap = ncap;
    return PEDONI_OK;
}

int ensure_app_capacity(PedoniModel* m, uint32_t need) {
    if (need <= m->app_cap) return PEDONI_OK;
    uint32_t ncap = std::max<uint32_t>(need, m->app_cap * 2);
    ncap = std::max<uint32_t>(ncap, 1024);
    AgentArrays fresh{};
    CUDA_TRY(m, alloc_agents(fresh, ncap));
    CUDA_TRY(m, copy_agents(fresh, m->app, m->app_n, m->stream));
    CUDA_TRY(m, cudaStreamSynchronize(m->stream));
    free_agents(m->app);
    m->app = fresh;
    m->app_cap = ncap;
    return PEDONI_OK;
}

// ---- optional per-kernel event timing -----------------------------------------------------------
cudaEvent_t take_event(PedoniModel* m) {
    if (!m->event_pool.empty()) {
        cudaEvent_t e = m->event_pool.back();
        m->event_pool.pop_back();
        return e;
    }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}

struct ScopedTimer {
    PedoniModel* m;
    TimedLaunch t{};
    cudaStream_t s;
    bool on;
    ScopedTimer(PedoniModel* m_, int kind, cudaStream_t stream, uint64_t agents = 0)
        : m(m_), s(stream), on(m_->profiling) {
        if (!on) return;
        t.kind = kind;
        t.stream_id = (m->edge_stream && stream == m->edge_stream) ? 1 : 0;
        t.agents = agents;
        t.start = take_event(m);
        t.stop = take_event(m);
        cudaEventRecord(t.start, s);
    }
    ~ScopedTimer() {
        if (!on) return;
        cudaEventRecord(t.stop, s);
        m->timed.push_back(t);
    }
};

int drain_timed(PedoniModel* m) {
    if (m->timed.empty()) return PEDONI_OK;
    int rc = sync_all(m);
    if (rc != PEDONI_OK) return rc;
    for (auto& t : m->timed) {
        float ms = 0.f;
        CUDA_TRY(m, cudaEventElapsedTime(&ms, t.start, t.stop));
        m->acc_ms[t.kind] += ms;
        m->acc_launches[t.kind] += 1;
        if (t.kind == kForce) m->acc_force_agents += t.agents;
        if (m->timer_armed && m->timeline.size() < 4096) {  // offsets from the last pedoni_timer_begin
            float t0 = 0.f, t1 = 0.f;
            if (cudaEventElapsedTime(&t0, m->timer_start, t.start) == cudaSuccess &&
                cudaEventElapsedTime(&t1, m->timer_start, t.stop) == cudaSuccess)
                m->timeline.push_back(PedoniLaunchRecord{t.kind, t.stream_id, t0, t1});
            else
                (void)cudaGetLastError();
        }
        m->event_pool.push_back(t.start);
        m->event_pool.push_back(t.stop);
    }
    m->timed.clear();
    return PEDONI_OK;
}

// Obstacle -> 4 edges, sfm.rs:194-205, evaluated once on the host with the reference's f32 ops.
void build_edges(const float* obstacles, int n, std::vector<float>& out) {
    out.assign((size_t)n * kEdgeFloats, 0.0f);
    for (int k = 0; k < n; ++k) {
        const float* o = obstacles + 5 * k;
        HVec2 v0{o[0], o[1]}, v1{o[2], o[3]};
        float w = o[4];
        HVec2 d{v1.x - v0.x, v1.y - v0.y};
        float h = std::sqrt((d.x * d.x) + (d.y * d.y));
        // vec2(d.y, -d.x).normalize_or_zero() * w * 0.5
        HVec2 nn{d.y, -d.x};
        float rcp = 1.0f / std::sqrt((nn.x * nn.x) + (nn.y * nn.y));
        if (std::isfinite(rcp) && rcp > 0.0f) {
            nn.x = nn.x * rcp;
            nn.y = nn.y * rcp;
        } else {
            nn.x = 0.0f;
            nn.y = 0.0f;
        }
        nn.x = nn.x * w * 0.5f;
        nn.y = nn.y * w * 0.5f;
        HVec2 lines[4][2] = {
            {{v0.x + nn.x, v0.y + nn.y}, {v0.x - nn.x, v0.y - nn.y}},
            {{v1.x + nn.x, v1.y + nn.y}, {v1.x - nn.x, v1.y - nn.y}},
            {{v0.x + nn.x, v0.y + nn.y}, {v1.x + nn.x, v1.y + nn.y}},
            {{v0.x - nn.x, v0.y - nn.y}, {v1.x - nn.x, v1.y - nn.y}},
        };
        float* e = out.data() + (size_t)k * kEdgeFloats;
        for (int j = 0; j < 4; ++j) {
            HVec2 b{lines[j][1].x - lines[j][0].x, lines[j][1].y - lines[j][0].y};
            e[5 * j + 0] = lines[j][0].x;
            e[5 * j + 1] = lines[j][0].y;
            e[5 * j + 2] = b.x;
            e[5 * j + 3] = b.y;
            e[5 * j + 4] = (b.x * b.x) + (b.y * b.y);
        }
        e[20] = w;
        e[21] = h;
    }
}

// Sort input = [compute range of buf[cur]] ++ [appended spawns]: "previous index" order.
SortInput make_sort_input(PedoniModel* m) {
    SortInput in{};
    in.nseg = 2;
    // segment 0 is addressed by absolute array index: everything below resident_hi() (see locate())
    in.seg[0] = Segment{m->buf[m->cur], m->range(kRangeCompute), m->resident_hi()};
    in.seg[1] = Segment{m->app, m->app_on_device ? m->d_app_range : nullptr, m->app_n};
    in.prefix[0] = 0;
    in.prefix[1] = m->resident_hi();
    in.prefix[2] = m->resident_hi() + m->app_n;
    in.cap[0] = m->cap;
    in.cap[1] = m->app_cap;
    in.out_cap = m->cap;
    in.error_flag = m->d_error;
    return in;
}

template <Math M, bool D, bool T>
void launch_force_t(PedoniModel* m, const ForceParams& p, dim3 blocks, size_t smem, cudaStream_t s) {
    force_integrate_kernel<M, D, T><<<blocks, kForceThreads, smem, s>>>(p);
    m->launches += 1;
}

// range_id2 >= 0: a second range of the same size bound, integrated by the CTAs with blockIdx.y == 1.
void launch_force(PedoniModel* m, int range_id, uint32_t count_upper, cudaStream_t s, int range_id2 = -1) {
    const dim3 blocks(div_up(count_upper, kForceThreads), range_id2 >= 0 ? 2 : 1);
    if (blocks.x == 0) return;
    ForceParams p{};
    p.in = m->buf[m->cur];
    p.out = m->buf[m->cur ^ 1];
    p.d_range = m->range(range_id);
    p.d_range_hi = m->range(range_id2 >= 0 ? range_id2 : range_id);
    p.d_owned = m->range(kRangeOwned);
    p.count_upper = count_upper;
    p.cap = m->cap;
    p.table_cells = m->n_cells;
    p.cell_start = m->d_cell_start;
    p.grid = m->grid;
    p.field = m->field;
    p.cs = m->cell_sort();
    p.error_flag = m->d_error;
    p.updates_total = m->d_updates;
    p.arrived = m->d_arrived;
    p.obstacle_edges = m->d_edges;
    p.n_obstacles = m->use_distance_map ? 0 : m->n_obstacles;
    const size_t smem = kForceSmemBytes;  // tile + neighbour lists; the segment-wall variant reuses the tile
    ScopedTimer t(m, (m->edge_stream && s == m->edge_stream) ? kForceEdge : kForce, s, count_upper);
    if (m->math_mode == PEDONI_MATH_STRICT) {
        if (m->use_distance_map)
            launch_force_t<Math::Strict, true, false>(m, p, blocks, smem, s);
        else
            launch_force_t<Math::Strict, false, false>(m, p, blocks, smem, s);
    } else if (m->field_tex) {
        if (m->use_distance_map)
            launch_force_t<Math::Fast, true, true>(m, p, blocks, smem, s);
        else
            launch_force_t<Math::Fast, false, true>(m, p, blocks, smem, s);
    } else {
        if (m->use_distance_map)
            launch_force_t<Math::Fast, true, false>(m, p, blocks, smem, s);
        else
            launch_force_t<Math::Fast, false, false>(m, p, blocks, smem, s);
    }
}

// Blocking: callers have synchronised (or are about to block anyway).
int check_device_error(PedoniModel* m) {
    uint32_t bits = 0;
    CUDA_TRY(m, cudaMemcpyAsync(&bits, m->d_error, sizeof bits, cudaMemcpyDeviceToHost, m->stream));
    CUDA_TRY(m, cudaStreamSynchronize(m->stream));
    if (bits == 0) return PEDONI_OK;
    CUDA_TRY(m, cudaMemsetAsync(m->d_error, 0, sizeof(uint32_t), m->stream));
    // Every raised bit is reported; the return code is that of the most severe one.
    std::string msg;
    int code = PEDONI_OK;
    auto add = [&](int c, const char* fmt, ...) {
        char buf[400];
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(buf, sizeof buf, fmt, ap);
        va_end(ap);
        if (!msg.empty()) msg += "; ";
        msg += buf;
        if (code == PEDONI_OK) code = c;
    };
    if (bits & kErrStageTimeout)
        add(PEDONI_ERR_CUDA, "force kernel: the bulk copies staging a warp's neighbour tile never completed");
    if (bits & kErrDebugBounds)
        add(PEDONI_ERR_CUDA, "debug build: an index derived on the device left its array (the access was skipped)");
    if (bits & kErrSpawnBound)
        add(PEDONI_ERR_CAPACITY, "a device-side Poisson draw exceeded mean + 10 sigma + 10 and was clamped");
    if (bits & kErrSortOverflow)
        add(PEDONI_ERR_CUDA, "rebuild: the overflow list of the cell slot rows ran out of entries");
    if (bits & kErrHaloTimeout)
        add(PEDONI_ERR_COMM,
            "slab %d of %d waited 20 s for a neighbour's ghost strip (peer-memory transport): a rank died or the "
            "ranks do not call pedoni_rebuild in lockstep", m->slab_rank, m->slab_count);
    if (bits & kErrHaloOverflow)
        add(PEDONI_ERR_CAPACITY,
            "two boundary rows of a slab hold more than halo_capacity = %u agents (or the ghost strip overran the "
            "arrays); raise PedoniConfig.halo_capacity", m->halo_cap);
    if (bits & kErrBadDestination)
        add(PEDONI_ERR_INVALID,
            "device flagged an invalid agent (destination >= n_potential_maps); the reference would panic with an "
            "index-out-of-bounds at field.rs:237");
    if (bits & kErrRowJump)
        add(PEDONI_ERR_STATE,
            "a pedestrian crossed two or more neighbor-grid rows in one step; the slab decomposition exchanges two "
            "ghost rows per tick and cannot follow it (speed > %.1f m/s)", m->grid.unit / 0.1f);
    if (code == PEDONI_OK) add(PEDONI_ERR_STATE, "unknown device error bits 0x%x", bits);
    return fail(m, code, "%s", msg.c_str());
}

__global__ void reset_layout_kernel(uint32_t* ranges, uint32_t offset, unsigned long long* host_slot, uint32_t tick) {
    for (int k = 0; k < 2 * 2 * kNumRanges; ++k) ranges[k] = offset;  // both layout buffers
    *host_slot = static_cast<unsigned long long>(tick) << 32;
}

// Upper bound of the owned population after the rebuild that just became tick m->tick: the last count
// the device published (it may lag the host by a few ticks) plus everything that can have been added
// since. Never synchronises.
uint32_t owned_bound(PedoniModel* m, uint32_t sort_input_upper) {
    const unsigned long long pub = *reinterpret_cast<volatile unsigned long long*>(m->h_pub);
    const uint32_t pub_tick = static_cast<uint32_t>(pub >> 32), pub_n = static_cast<uint32_t>(pub);
    uint64_t bound = sort_input_upper;
    if (pub_tick <= m->tick && m->tick - pub_tick < 64) {
        const uint64_t since = m->inflow_cum[m->tick % 64] - m->inflow_cum[pub_tick % 64];
        bound = std::min<uint64_t>(bound, pub_n + since);
    }
    return static_cast<uint32_t>(bound);
}

void advance_tick(PedoniModel* m, uint64_t inflow) {
    const uint64_t prev = m->inflow_cum[m->tick % 64];
    m->tick += 1;
    m->inflow_cum[m->tick % 64] = prev + inflow;
}

HaloMessage msg_of(uint32_t* words) { return HaloMessage{words}; }

// Enqueue the ghost unpack (edge stream). NCCL / in-process transports: the strips have landed in this
// exchange's receive slots (stream order). Peer-memory transport: the kernel polls the arrival flags.
void enqueue_unpack(PedoniModel* m) {
    const uint32_t threads = std::max<uint32_t>(m->halo_cap, 2 * m->grid.nx + 1);
    const uint32_t slot = m->exchange_seq & 1u;
    dim3 grid(div_up(threads, 256), 2);
    halo_unpack_kernel<<<grid, 256, 0, m->edge_stream>>>(
        m->buf[m->cur], m->d_cell_start, m->own_begin_cell, m->own_end_cell, m->grid.nx, m->halo_cap, m->cap,
        msg_of(m->recv_below(m->d_arena, slot)), msg_of(m->recv_above(m->d_arena, slot)), m->has_below, m->has_above,
        m->ranges(m->rcur), m->d_error, m->flag_below(m->d_arena), m->flag_above(m->d_arena),
        m->transport == PedoniModel::kTransportPeer ? m->exchange_seq : 0u);
    m->launches += 1;
    cudaEventRecord(m->ev_halo, m->edge_stream);
    m->halo_pending = false;
    m->halo_inflight = true;
}

int exchange_nccl(PedoniModel* m) {
    CUDA_TRY(m, cudaStreamWaitEvent(m->edge_stream, m->ev_packed, 0));
    ScopedTimer t(m, kComm, m->edge_stream);
    std::string err;
    const uint32_t slot = m->exchange_seq & 1u;
    int rc = pedoni::slab_comm_exchange(m->comm, m->edge_stream, m->d_send_dn, m->recv_below(m->d_arena, slot),
                                        m->d_send_up, m->recv_above(m->d_arena, slot), m->msg_bytes, m->has_below,
                                        m->has_above, &err);
    if (rc != PEDONI_OK) return fail(m, PEDONI_ERR_COMM, "%s", err.c_str());
    enqueue_unpack(m);
    return PEDONI_OK;
}

// Peer-memory transport: nothing to enqueue but the unpack, which polls for the neighbours' strips.
int exchange_peer(PedoniModel* m) {
    CUDA_TRY(m, cudaStreamWaitEvent(m->edge_stream, m->ev_packed, 0));  // the ghost rows go next to the new table
    ScopedTimer t(m, kComm, m->edge_stream);
    enqueue_unpack(m);
    return PEDONI_OK;
}

// Map the neighbours' receive arenas (CUDA IPC) so that the pack kernel can store into them over NVLink.
// The 64-byte handles travel through the NCCL communicator; every rank must succeed or all stay on NCCL.
int setup_peer_transport(PedoniModel* m) {
    const char* forced = std::getenv("PEDONI_SLAB_TRANSPORT");
    int ok = !(forced && std::string(forced) == "nccl");
    unsigned char* d_handles = nullptr;  // [mine | from below | from above], 64 bytes each
    int* d_ok = nullptr;
    cudaIpcMemHandle_t mine{}, from_below{}, from_above{};
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    // every exit path frees the scratch buffers; a rank that ends up on NCCL closes what it had mapped
    auto finish = [&](int rc) {
        cudaFree(d_handles);
        cudaFree(d_ok);
        if (rc != PEDONI_OK || m->transport != PedoniModel::kTransportPeer) {
            if (m->peer_arena_below) cudaIpcCloseMemHandle(m->peer_arena_below);
            if (m->peer_arena_above) cudaIpcCloseMemHandle(m->peer_arena_above);
            m->peer_arena_below = m->peer_arena_above = nullptr;
            (void)cudaGetLastError();
        }
        return rc;
    };
#define PEER_TRY(expr)                                                                                           \
    do {                                                                                                         \
        cudaError_t err__ = (expr);                                                                              \
        if (err__ != cudaSuccess)                                                                                \
            return finish(fail(m, PEDONI_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(err__),    \
                               __FILE__, __LINE__));                                                             \
    } while (0)
    PEER_TRY(cudaMalloc(&d_handles, 3 * 64));
    PEER_TRY(cudaMalloc(&d_ok, sizeof(int)));
    if (ok && cudaIpcGetMemHandle(&mine, m->d_arena) != cudaSuccess) {
        ok = 0;
        (void)cudaGetLastError();
    }
    PEER_TRY(cudaMemcpyAsync(d_handles, &mine, 64, cudaMemcpyHostToDevice, m->edge_stream));
    std::string err;
    // my handle goes to both neighbours; theirs come back (same pattern as a ghost exchange)
    int rc = pedoni::slab_comm_exchange(m->comm, m->edge_stream, d_handles, d_handles + 64, d_handles, d_handles + 128, 64,
                                        m->has_below, m->has_above, &err);
    if (rc != PEDONI_OK) return finish(fail(m, PEDONI_ERR_COMM, "%s", err.c_str()));
    PEER_TRY(cudaMemcpyAsync(&from_below, d_handles + 64, 64, cudaMemcpyDeviceToHost, m->edge_stream));
    PEER_TRY(cudaMemcpyAsync(&from_above, d_handles + 128, 64, cudaMemcpyDeviceToHost, m->edge_stream));
    PEER_TRY(cudaStreamSynchronize(m->edge_stream));
    void* p = nullptr;
    if (ok && m->has_below) {
        if (cudaIpcOpenMemHandle(&p, from_below, cudaIpcMemLazyEnablePeerAccess) == cudaSuccess)
            m->peer_arena_below = static_cast<unsigned char*>(p);
        else
            ok = 0;
    }
    if (ok && m->has_above) {
        if (cudaIpcOpenMemHandle(&p, from_above, cudaIpcMemLazyEnablePeerAccess) == cudaSuccess)
            m->peer_arena_above = static_cast<unsigned char*>(p);
        else
            ok = 0;
    }
    (void)cudaGetLastError();
    PEER_TRY(cudaMemcpyAsync(d_ok, &ok, sizeof ok, cudaMemcpyHostToDevice, m->edge_stream));
    rc = pedoni::slab_comm_all_min(m->comm, m->edge_stream, d_ok, &err);
    if (rc != PEDONI_OK) return finish(fail(m, PEDONI_ERR_COMM, "%s", err.c_str()));
    PEER_TRY(cudaMemcpyAsync(&ok, d_ok, sizeof ok, cudaMemcpyDeviceToHost, m->edge_stream));
    PEER_TRY(cudaStreamSynchronize(m->edge_stream));
#undef PEER_TRY
    m->transport = ok ? PedoniModel::kTransportPeer : PedoniModel::kTransportNccl;
    return finish(PEDONI_OK);
}

// ---- field maps as one texture atlas (fast math) -------------------------------------------------------
// Footprints at pseudo-random places of every map, fetched with footprint_gather and with plain loads: any
// difference (a gather component order or a coordinate convention other than the one force.cuh assumes)
// raises the flag and the handle keeps the load path.
__global__ void footprint_check_kernel(FieldView f, uint32_t* mismatch) {
    const int map = blockIdx.y;  // 0: distance map, 1 + k: potential map k
    const unsigned long long r = splitmix64((static_cast<unsigned long long>(map) << 32) | (blockIdx.x * blockDim.x + threadIdx.x));
    const int x0 = static_cast<int>((r & 0xFFFFFFFFull) % static_cast<unsigned>(f.fx - 3));
    const int y0 = static_cast<int>((r >> 32) % static_cast<unsigned>(f.fy - 3));
    const float* g = map == 0 ? f.distance_map : f.potential_maps + static_cast<size_t>(map - 1) * f.fy * f.fx;
    float t[4][4];
    footprint_gather(f.atlas, x0 + (map % f.atlas_tiles_x) * f.fx, y0 + (map / f.atlas_tiles_x) * f.fy, t);
    bool same = true;
    for (int r4 = 0; r4 < 4; ++r4)
        for (int c4 = 0; c4 < 4; ++c4)
            same = same && __float_as_uint(t[r4][c4]) == __float_as_uint(g[static_cast<size_t>(y0 + r4) * f.fx + x0 + c4]);
    if (!same) atomicOr(mismatch, 1u);
}

// ---- pipelined download: destinations as bytes --------------------------------------------------------
bool download_packs_destinations(const PedoniModel* m) {
    // Whole-domain handles only: there ONE PCIe link carries every pedestrian and the bytes on it bound the
    // tick. With one process per slab each GPU has its own link and the shared host (memory bandwidth,
    // cores) is the limit; widening on the host then costs more than the bytes save (measured at 2 GPUs:
    // 7.5e9 -> 3.6e9 updates/s end to end). PEDONI_DOWNLOAD_PACK=0 / 1 overrides.
    const char* env = std::getenv("PEDONI_DOWNLOAD_PACK");
    if (m->field.n_maps > 256) return false;  // live pedestrians have destination < n_maps (sort_key)
    if (env && (env[0] == '0' || env[0] == '1')) return env[0] == '1';
    return m->slab_count <= 1;
}

// four destinations per thread -> one 32-bit store of four bytes
__global__ void __launch_bounds__(256) pack_dest_kernel(const uint32_t* __restrict__ dest, uint32_t n, uint32_t* __restrict__ out) {
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x, i = 4 * q;
    if (i >= n) return;
    uint32_t w = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k)
        if (i + k < n) w |= (dest[i + k] & 0xFFu) << (8 * k);
    out[q] = w;
}

void release_field_textures(PedoniModel* m) {
    if (m->field.atlas) cudaDestroyTextureObject(m->field.atlas);
    if (m->field_atlas) cudaFreeArray(m->field_atlas);
    m->field_atlas = nullptr;
    m->field.atlas = 0;
    m->field_tex = false;
}

// Builds the atlas + texture object from the device copies of the maps and verifies it. Any failure (no
// memory for a second copy, maps that do not tile into a 32768 x 32768 gather array, unexpected gather
// layout) leaves the handle on the load path: slower, same results. PEDONI_FIELD_TEXTURES=0 skips the attempt.
void build_field_textures(PedoniModel* m) {
    const char* env = std::getenv("PEDONI_FIELD_TEXTURES");
    if (env && env[0] == '0') return;
    constexpr int kMaxGather = 32768;  // cudaDeviceProp::maxTexture2DGather
    const int fx = m->field.fx, fy = m->field.fy, n = 1 + m->field.n_maps;
    if (fx < 4 || fy < 4 || fx > kMaxGather || fy > kMaxGather) return;
    int tiles_x = 1, shift = 0;  // a power of two, so that the kernels find a map's tile with a mask and a shift
    while (2 * tiles_x <= std::min(n, kMaxGather / fx)) tiles_x *= 2, shift += 1;
    const int tiles_y = (n + tiles_x - 1) / tiles_x;
    if (static_cast<long long>(tiles_y) * fy > kMaxGather) return;
    const size_t map_elems = static_cast<size_t>(fx) * fy;
    const cudaChannelFormatDesc desc = cudaCreateChannelDesc<float>();
    bool ok = cudaMallocArray(&m->field_atlas, &desc, static_cast<size_t>(tiles_x) * fx, static_cast<size_t>(tiles_y) * fy,
                              cudaArrayTextureGather) == cudaSuccess;
    for (int k = 0; k < n && ok; ++k) {
        const float* src = k == 0 ? m->d_distance : m->d_potential + static_cast<size_t>(k - 1) * map_elems;
        ok = cudaMemcpy2DToArrayAsync(m->field_atlas, sizeof(float) * (k % tiles_x) * fx, static_cast<size_t>(k / tiles_x) * fy, src,
                                      sizeof(float) * fx, sizeof(float) * fx, fy, cudaMemcpyDeviceToDevice,
                                      m->stream) == cudaSuccess;
    }
    if (ok) {
        cudaResourceDesc res{};
        res.resType = cudaResourceTypeArray;
        res.res.array.array = m->field_atlas;
        cudaTextureDesc td{};
        td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp;  // never exercised: footprints are in bounds
        td.filterMode = cudaFilterModePoint;
        td.readMode = cudaReadModeElementType;
        td.normalizedCoords = 0;
        ok = cudaCreateTextureObject(&m->field.atlas, &res, &td, nullptr) == cudaSuccess;
        m->field.atlas_tiles_x = tiles_x;
        m->field.atlas_shift = shift;
    }
    uint32_t* d_flag = nullptr;
    uint32_t flag = 1;
    if (ok) ok = cudaMalloc(&d_flag, sizeof(uint32_t)) == cudaSuccess;
    if (ok) {
        cudaMemsetAsync(d_flag, 0, sizeof(uint32_t), m->stream);
        footprint_check_kernel<<<dim3(8, n), 128, 0, m->stream>>>(m->field, d_flag);
        ok = cudaMemcpyAsync(&flag, d_flag, sizeof flag, cudaMemcpyDeviceToHost, m->stream) == cudaSuccess &&
             cudaStreamSynchronize(m->stream) == cudaSuccess && flag == 0;
    }
    if (d_flag) cudaFree(d_flag);
    if (!ok) {
        cudaGetLastError();  // an optimisation that did not fit is not an error of the handle
        release_field_textures(m);
        return;
    }
    m->field_tex = true;
}

// ---- far-from-walls mask (fast math, distance-map walls) -----------------------------------------------
// The wall term is 10 * 0.2 * exp(-distance / 0.2) * direction (sfm.rs:188-192): at 8 m it is 2 e^-40 = 8.5e-18 m/s^2,
// ten orders of magnitude below one ulp of any acceleration a pedestrian has. One bit per block of 8 x 8 texels of
// the distance map says: for EVERY position whose field coordinate (pos / unit - 0.5, field.rs:243) floors into the
// block, all texels the wall term's 4x4 footprints can touch (floor - 1 .. floor + 2, plus one texel of margin) lie
// inside the map and hold a distance >= kWallCutoff, and the map is strictly monotone along x or along y over that
// region (steps of one sign, each beyond 1e-3 * unit) — then the Sobel gradient cannot vanish, so the reference's
// `normalize()` cannot produce the NaN that removes a pedestrian (edge semantics), and
// skipping the term changes the acceleration by less than 1e-17. Ridges of the distance map (corridor mid-lines, the
// diagonals of an open square), map borders and anything NaN keep the full evaluation.
constexpr float kWallCutoff = 8.0f;  // metres

__global__ void __launch_bounds__(256) far_mask_kernel(FieldView f, int blocks_y, uint32_t* __restrict__ mask,
                                                       unsigned long long* __restrict__ n_far) {
    const uint32_t block = blockIdx.x * blockDim.x + threadIdx.x;
    if (block >= static_cast<uint32_t>(f.far_bw) * static_cast<uint32_t>(blocks_y)) return;
    const int bx = static_cast<int>(block % static_cast<uint32_t>(f.far_bw)), by = static_cast<int>(block / static_cast<uint32_t>(f.far_bw));
    constexpr int kB = 1 << kFarShift;
    const int x0 = bx * kB - 2, x1 = bx * kB + kB - 1 + 3;
    const int y0 = by * kB - 2, y1 = by * kB + kB - 1 + 3;
    if (x0 < 0 || y0 < 0 || x1 >= f.fx || y1 >= f.fy) return;  // map border: out-of-bounds taps read 1e12 (util.rs:45)
    const float eps = 1.0e-3f * f.unit;
    bool far = true, up_x = true, down_x = true, up_y = true, down_y = true;
    for (int y = y0; y <= y1; ++y) {
        const float* row = f.distance_map + static_cast<size_t>(y) * f.fx;
        for (int x = x0; x <= x1; ++x) {
            const float v = row[x];
            far = far && v >= kWallCutoff;  // false for NaN
            if (x < x1) {
                const float d = row[x + 1] - v;
                up_x = up_x && d > eps;
                down_x = down_x && d < -eps;
            }
            if (y < y1) {
                const float d = row[x + f.fx] - v;
                up_y = up_y && d > eps;
                down_y = down_y && d < -eps;
            }
        }
    }
    if (far && (up_x || down_x || up_y || down_y)) {
        atomicOr(mask + (block >> 5), 1u << (block & 31u));
        atomicAdd(n_far, 1ull);
    }
}

// Any failure leaves the handle without a mask (every wall term evaluated): an optimisation, not a requirement.
void build_far_mask(PedoniModel* m) {
    const char* env = std::getenv("PEDONI_WALL_CUTOFF");
    if (env && env[0] == '0') return;
    constexpr int kB = 1 << kFarShift;
    const int bw = (m->field.fx + kB - 1) / kB, bh = (m->field.fy + kB - 1) / kB;
    const uint32_t blocks = static_cast<uint32_t>(bw) * static_cast<uint32_t>(bh);
    const size_t words = (static_cast<size_t>(blocks) + 31) / 32;
    m->field.far_bw = bw;
    m->far_blocks_total = blocks;
    unsigned long long* d_n = nullptr;
    bool ok = cudaMalloc(&m->d_far_mask, words * sizeof(uint32_t)) == cudaSuccess &&
              cudaMalloc(&d_n, sizeof(unsigned long long)) == cudaSuccess;
    if (ok) {
        cudaMemsetAsync(m->d_far_mask, 0, words * sizeof(uint32_t), m->stream);
        cudaMemsetAsync(d_n, 0, sizeof(unsigned long long), m->stream);
        far_mask_kernel<<<div_up(blocks, 256), 256, 0, m->stream>>>(m->field, bh, m->d_far_mask, d_n);
        ok = cudaMemcpyAsync(&m->far_cells, d_n, sizeof(unsigned long long), cudaMemcpyDeviceToHost, m->stream) == cudaSuccess &&
             cudaStreamSynchronize(m->stream) == cudaSuccess;
    }
    cudaFree(d_n);
    if (!ok || m->far_cells == 0) {  // nothing to skip: keep the kernel off the extra load
        (void)cudaGetLastError();
        cudaFree(m->d_far_mask);
        m->d_far_mask = nullptr;
        m->far_cells = 0;
        return;
    }
    m->field.far_mask = m->d_far_mask;
}

}  // namespace

// ===================================================================================================
extern "C" {

int pedoni_abi_version(void) { return PEDONI_ABI_VERSION; }

const char* pedoni_last_error(const PedoniModel* model) {
    return model ? model->last_error.c_str() : g_create_error.c_str();
}

int pedoni_slab_rows(int32_t ny, int32_t count, int32_t rank, int32_t* row0, int32_t* row1) {
    if (ny <= 0 || count <= 0 || rank < 0 || rank >= count || !row0 || !row1) return PEDONI_ERR_INVALID;
    // Balanced contiguous row ranges; the first (ny % count) slabs get one extra row.
    int32_t base = ny / count, extra = ny % count;
    *row0 = rank * base + std::min(rank, extra);
    *row1 = *row0 + base + (rank < extra ? 1 : 0);
    return PEDONI_OK;
}

int pedoni_create(const PedoniConfig* c, PedoniModel** out) {
    if (!c || !out) return fail(nullptr, PEDONI_ERR_INVALID, "null config or out pointer");
    *out = nullptr;
    if (c->struct_size != sizeof(PedoniConfig))
        return fail(nullptr, PEDONI_ERR_INVALID, "PedoniConfig.struct_size %u != %zu (ABI mismatch)", c->struct_size,
                    sizeof(PedoniConfig));
    if (!c->use_neighbor_grid)
        return fail(nullptr, PEDONI_ERR_UNSUPPORTED,
                    "use_neighbor_grid = false (the O(N^2) path, sfm.rs:157-185) is not implemented on CUDA");
    if (!(c->neighbor_grid_unit > 0.f) || !(c->field_grid_unit > 0.f))
        return fail(nullptr, PEDONI_ERR_INVALID, "grid units must be positive");
    if (c->field_ny <= 0 || c->field_nx <= 0 || c->n_potential_maps <= 0 || !c->distance_map || !c->potential_maps)
        return fail(nullptr, PEDONI_ERR_INVALID, "field maps missing or empty");
    if (c->n_obstacles < 0 || (c->n_obstacles > 0 && !c->obstacles))
        return fail(nullptr, PEDONI_ERR_INVALID, "obstacles pointer missing");
    if (c->math_mode != PEDONI_MATH_STRICT && c->math_mode != PEDONI_MATH_FAST)
        return fail(nullptr, PEDONI_ERR_INVALID, "unknown math_mode %d", c->math_mode);

    int n_dev = 0;
    cudaError_t err = cudaGetDeviceCount(&n_dev);
    if (err != cudaSuccess || n_dev == 0)
        return fail(nullptr, PEDONI_ERR_CUDA, "no CUDA device available (%s); there is no CPU fallback",
                    err != cudaSuccess ? cudaGetErrorString(err) : "device count 0");
    if (c->device < 0 || c->device >= n_dev)
        return fail(nullptr, PEDONI_ERR_INVALID, "device %d out of range [0, %d)", c->device, n_dev);

    PedoniModel* m = new PedoniModel();
    auto bail = [&](int code) {
        g_create_error = m->last_error;
        pedoni_destroy(m);
        return code;
    };
#define CREATE_TRY(expr)                                                                                \
    do {                                                                                                \
        cudaError_t e__ = (expr);                                                                       \
        if (e__ != cudaSuccess) {                                                                       \
            fail(m, PEDONI_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(e__));                  \
            return bail(PEDONI_ERR_CUDA);                                                               \
        }                                                                                               \
    } while (0)

    m->device = c->device;
    CREATE_TRY(cudaSetDevice(m->device));
    if (c->stream) {
        m->stream = static_cast<cudaStream_t>(c->stream);
    } else {
        CREATE_TRY(cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking));
        m->own_stream = true;
    }
    m->math_mode = c->math_mode;
    m->use_distance_map = c->use_distance_map != 0;

    // neighbor_grid.rs:14-17: shape = ceil(size / unit) as (ny, nx)
    const float gx = std::ceil(c->field_size_x / c->neighbor_grid_unit);
    const float gy = std::ceil(c->field_size_y / c->neighbor_grid_unit);
    if (!(gx >= 1.f) || !(gy >= 1.f) || gx * gy > 2.0e9f) {
        fail(m, PEDONI_ERR_INVALID, "neighbor grid %g x %g is empty or too large", gx, gy);
        return bail(PEDONI_ERR_INVALID);
    }
    m->grid.unit = c->neighbor_grid_unit;
    m->grid.nx = static_cast<int>(gx);
    m->grid.ny = static_cast<int>(gy);

    m->slab_count = c->slab_count > 1 ? c->slab_count : 1;
    m->slab_rank = m->slab_count > 1 ? c->slab_rank : 0;
    int32_t r0 = 0, r1 = m->grid.ny;
    if (m->slab_count > 1) {
        // every slab ships two whole rows to each neighbour, so it must own at least two
        if (pedoni_slab_rows(m->grid.ny, m->slab_count, m->slab_rank, &r0, &r1) != PEDONI_OK ||
            m->grid.ny / m->slab_count < 2) {
            fail(m, PEDONI_ERR_INVALID, "slab %d of %d: a %d-row grid gives a slab fewer than 2 rows", m->slab_rank,
                 m->slab_count, m->grid.ny);
            return bail(PEDONI_ERR_INVALID);
        }
        m->has_below = m->slab_rank > 0;
        m->has_above = m->slab_rank < m->slab_count - 1;
    }
    m->grid.slab = m->slab_count > 1 ? 1 : 0;
    m->grid.own_row0 = r0;
    m->grid.own_row1 = r1;
    m->grid.row_base = r0 - (m->has_below ? 2 : 0);
    const int table_end = r1 + (m->has_above ? 2 : 0);
    m->grid.table_rows = table_end - m->grid.row_base;
    m->n_cells = static_cast<uint32_t>(m->grid.table_rows) * static_cast<uint32_t>(m->grid.nx);
    m->own_begin_cell = static_cast<uint32_t>(r0 - m->grid.row_base) * m->grid.nx;
    m->own_end_cell = static_cast<uint32_t>(r1 - m->grid.row_base) * m->grid.nx;

    const uint32_t capacity = c->capacity ? c->capacity : 4096;
    if (m->slab_count > 1) {
        if (c->halo_capacity) {
            m->halo_cap = c->halo_capacity;
        } else {  // 3x the two-row population of a uniformly filled slab
            // (same on every rank: the message size must agree, so use the common floor of rows per slab)
            const uint64_t per_row = capacity / static_cast<uint32_t>(m->grid.ny / m->slab_count) + 1;
            m->halo_cap = static_cast<uint32_t>(std::max<uint64_t>(4096, 6 * per_row));
        }
        m->halo_cap = (m->halo_cap + 255u) & ~255u;
        m->array_offset = m->has_below ? m->halo_cap : 0;
        m->msg_bytes = halo_message_words(m->grid.nx, m->halo_cap) * sizeof(uint32_t);
        // Highest priority: the exchange, unpack and edge kernels are tiny but sit behind tens of thousands
        // of queued CTAs of the interior force kernel; without priority they are dispatched only when that
        // kernel drains (measured at 8 GPUs: exchange 167 us ~ the whole interior kernel).
        int prio_least = 0, prio_greatest = 0;
        CREATE_TRY(cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest));
        CREATE_TRY(cudaStreamCreateWithPriority(&m->edge_stream, cudaStreamNonBlocking, prio_greatest));
        for (uint32_t** b : {&m->d_send_dn, &m->d_send_up}) {
            CREATE_TRY(cudaMalloc(b, m->msg_bytes));
            CREATE_TRY(cudaMemsetAsync(*b, 0, m->msg_bytes, m->stream));
        }
        m->slot_bytes = (m->msg_bytes + 255) & ~static_cast<size_t>(255);
        m->arena_bytes = PedoniModel::kArenaControlBytes + 4 * m->slot_bytes;
        CREATE_TRY(cudaMalloc(&m->d_arena, m->arena_bytes));
        CREATE_TRY(cudaMemsetAsync(m->d_arena, 0, m->arena_bytes, m->stream));
        for (cudaEvent_t* e : {&m->ev_packed, &m->ev_halo, &m->ev_edge, &m->ev_peer, &m->ev_sorted})
            CREATE_TRY(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
    }

    // Field (field.rs:194-205)
    m->field.unit = c->field_grid_unit;
    {
        int exp2 = 0;  // unit = 0.5 * 2^exp2 exactly <=> power of two; then 1/unit is exact as well
        m->field.inv_unit = (std::frexp(c->field_grid_unit, &exp2) == 0.5f && exp2 > -120 && exp2 < 120)
                                ? 1.0f / c->field_grid_unit : 0.0f;
    }
    m->field.fy = c->field_ny;
    m->field.fx = c->field_nx;
    m->field.n_maps = c->n_potential_maps;
    const size_t map_elems = static_cast<size_t>(c->field_ny) * c->field_nx;
    CREATE_TRY(cudaMalloc(&m->d_distance, map_elems * sizeof(float)));
    CREATE_TRY(cudaMalloc(&m->d_potential, map_elems * sizeof(float) * c->n_potential_maps));
    CREATE_TRY(cudaMemcpyAsync(m->d_distance, c->distance_map, map_elems * sizeof(float), cudaMemcpyHostToDevice,
                               m->stream));
    CREATE_TRY(cudaMemcpyAsync(m->d_potential, c->potential_maps, map_elems * sizeof(float) * c->n_potential_maps,
                               cudaMemcpyHostToDevice, m->stream));
    m->field.distance_map = m->d_distance;
    m->field.potential_maps = m->d_potential;
    if (m->math_mode == PEDONI_MATH_FAST) build_field_textures(m);
    if (m->math_mode == PEDONI_MATH_FAST && m->use_distance_map) build_far_mask(m);

    m->n_obstacles = c->n_obstacles;
    if (!m->use_distance_map && m->n_obstacles > 0) {
        std::vector<float> edges;
        build_edges(c->obstacles, c->n_obstacles, edges);
        CREATE_TRY(cudaMalloc(&m->d_edges, edges.size() * sizeof(float)));
        CREATE_TRY(cudaMemcpyAsync(m->d_edges, edges.data(), edges.size() * sizeof(float), cudaMemcpyHostToDevice,
                                   m->stream));
        CREATE_TRY(cudaStreamSynchronize(m->stream));  // `edges` dies at scope end
    }

    // cell table + scan scratch
    m->n_tiles = div_up(m->n_cells, kSortCtaThreads);  // status words for the smallest tile (one cell per thread)
    {
        cudaDeviceProp prop{};
        CREATE_TRY(cudaGetDeviceProperties(&prop, m->device));
        m->sm_count = prop.multiProcessorCount;
    }
    CREATE_TRY(cudaFuncSetAttribute(sort_cells_kernel<3, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    static_cast<int>(kSortSmemBytes)));
    CREATE_TRY(cudaFuncSetAttribute(sort_cells_kernel<PEDONI_SORT_MIN_BLOCKS, PEDONI_SORT_UNROLL>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kSortSmemBytes)));
    CREATE_TRY(cudaMalloc(&m->d_slots, sizeof(uint32_t) * kSlotsPerCell * ((size_t)m->n_cells + 1)));
    CREATE_TRY(cudaMalloc(&m->d_ovf_head, sizeof(uint32_t) * ((size_t)m->n_cells + 8)));
    CREATE_TRY(cudaMemsetAsync(m->d_ovf_head, 0, sizeof(uint32_t) * ((size_t)m->n_cells + 8), m->stream));
    CREATE_TRY(cudaMalloc(&m->d_ovf_count, sizeof(uint32_t)));
    CREATE_TRY(cudaMemsetAsync(m->d_ovf_count, 0, sizeof(uint32_t), m->stream));
    CREATE_TRY(cudaMalloc(&m->d_sort_done, sizeof(uint32_t)));
    CREATE_TRY(cudaMemsetAsync(m->d_sort_done, 0, sizeof(uint32_t), m->stream));
    CREATE_TRY(cudaMalloc(&m->d_cell_count, sizeof(uint32_t) * ((size_t)m->n_cells + 8)));
    CREATE_TRY(cudaMalloc(&m->d_cell_start, sizeof(uint32_t) * ((size_t)m->n_cells + 8)));
    CREATE_TRY(cudaMalloc(&m->d_tile_status, sizeof(unsigned long long) * std::max<uint32_t>(m->n_tiles, 1)));
    CREATE_TRY(cudaMemsetAsync(m->d_tile_status, 0, sizeof(unsigned long long) * std::max<uint32_t>(m->n_tiles, 1),
                               m->stream));
    CREATE_TRY(cudaMalloc(&m->d_tile_ticket, 2 * sizeof(uint32_t)));
    CREATE_TRY(cudaMemsetAsync(m->d_tile_ticket, 0, 2 * sizeof(uint32_t), m->stream));
    CREATE_TRY(cudaMalloc(&m->d_ranges, sizeof(uint32_t) * 2 * 2 * kNumRanges));
    CREATE_TRY(cudaMalloc(&m->d_error, sizeof(uint32_t)));
    CREATE_TRY(cudaMalloc(&m->d_updates, sizeof(unsigned long long)));
    CREATE_TRY(cudaMemsetAsync(m->d_updates, 0, sizeof(unsigned long long), m->stream));
    CREATE_TRY(cudaMalloc(&m->d_arrived, 16 * sizeof(unsigned long long)));
    CREATE_TRY(cudaMemsetAsync(m->d_arrived, 0, 16 * sizeof(unsigned long long), m->stream));
    CREATE_TRY(cudaMalloc(&m->d_observe, sizeof(ObserveOut)));
    CREATE_TRY(cudaMemsetAsync(m->d_cell_start, 0, sizeof(uint32_t) * ((size_t)m->n_cells + 8), m->stream));
    CREATE_TRY(cudaMemsetAsync(m->d_cell_count, 0, sizeof(uint32_t) * ((size_t)m->n_cells + 8), m->stream));
    CREATE_TRY(cudaMemsetAsync(m->d_error, 0, sizeof(uint32_t), m->stream));
    CREATE_TRY(cudaHostAlloc(&m->h_pub, sizeof(unsigned long long) * 2, cudaHostAllocMapped));
    m->h_pub[0] = m->h_pub[1] = 0;
    CREATE_TRY(cudaHostGetDevicePointer(&m->h_pub_dev, m->h_pub, 0));
    reset_layout_kernel<<<1, 1, 0, m->stream>>>(m->d_ranges, m->array_offset, m->h_pub_dev, m->tick);
    CREATE_TRY(cudaEventCreate(&m->timer_start));
    CREATE_TRY(cudaEventCreate(&m->timer_stop));

    if (ensure_capacity(m, m->array_offset + capacity + (m->has_above ? m->halo_cap : 0)) != PEDONI_OK)
        return bail(PEDONI_ERR_CUDA);
    CREATE_TRY(cudaStreamSynchronize(m->stream));  // borrowed map pointers may die after return
#undef CREATE_TRY
    *out = m;
    return PEDONI_OK;
}

void pedoni_destroy(PedoniModel* m) {
    if (!m) return;
    cudaSetDevice(m->device);
    if (m->stream) cudaStreamSynchronize(m->stream);
    if (m->edge_stream) cudaStreamSynchronize(m->edge_stream);
    if (m->peer_arena_below) cudaIpcCloseMemHandle(m->peer_arena_below);
    if (m->peer_arena_above) cudaIpcCloseMemHandle(m->peer_arena_above);
    if (m->comm) pedoni::slab_comm_destroy(m->comm);
    for (auto& t : m->timed) {
        cudaEventDestroy(t.start);
        cudaEventDestroy(t.stop);
    }
    for (auto e : m->event_pool) cudaEventDestroy(e);
    for (cudaEvent_t e : {m->timer_start, m->timer_stop, m->ev_packed, m->ev_halo, m->ev_edge, m->ev_peer, m->ev_sorted})
        if (e) cudaEventDestroy(e);
    release_field_textures(m);
    free_agents(m->buf[0]);
    free_agents(m->buf[1]);
    free_agents(m->app);
    for (void* p : {(void*)m->d_slots, (void*)m->d_ovf_head, (void*)m->d_ovf, (void*)m->d_ovf_count, (void*)m->d_sort_done,
                    (void*)m->d_cell_count, (void*)m->d_cell_start, (void*)m->d_tile_status, (void*)m->d_tile_ticket,
                    (void*)m->d_ranges,
                    (void*)m->d_error, (void*)m->d_updates, (void*)m->d_arrived, (void*)m->d_observe, (void*)m->d_distance, (void*)m->d_potential,
                    (void*)m->d_edges, (void*)m->d_send_dn, (void*)m->d_send_up, (void*)m->d_arena, m->d_spawn_groups,
                    (void*)m->d_spawn_rates, (void*)m->d_spawn_stream, (void*)m->d_app_range, (void*)m->d_far_mask})
        cudaFree(p);
    if (m->h_pub) cudaFreeHost(m->h_pub);
    for (int k = 0; k < PedoniModel::kStageSlots; ++k) {
        if (m->ev_stage[k]) cudaEventDestroy(m->ev_stage[k]);
        if (m->h_stage[k]) cudaFreeHost(m->h_stage[k]);
    }
    if (m->dl_stream) {
        cudaStreamSynchronize(m->dl_stream);
        cudaStreamDestroy(m->dl_stream);
    }
    for (auto& d : m->dl) {
        for (cudaEvent_t e : {d.ev_snap, d.ev_done})
            if (e) cudaEventDestroy(e);
        cudaFree(d.d_pos);
        cudaFree(d.d_dest);
        cudaFree(d.d_dest8);
        cudaFree(d.d_range);
        if (d.h_dest8) cudaFreeHost(d.h_dest8);
        if (d.h_range) cudaFreeHost(d.h_range);
    }
    if (m->edge_stream) cudaStreamDestroy(m->edge_stream);
    if (m->own_stream && m->stream) cudaStreamDestroy(m->stream);
    delete m;
}

int pedoni_synchronize(PedoniModel* m) {
    if (!m) return PEDONI_ERR_INVALID;
    CUDA_TRY(m, cudaSetDevice(m->device));
    int rc = sync_all(m);
    if (rc != PEDONI_OK) return rc;
    return check_device_error(m);
}

static int append_agents(PedoniModel* m, uint32_t n, const float* pos_xy, const uint32_t* dest, const float* vel_xy,
                         const float* v0) {
    if (n == 0) return PEDONI_OK;
    if (!pos_xy || !dest || !v0) return fail(m, PEDONI_ERR_INVALID, "null agent array with n = %u", n);
    if (m->app_on_device)
        return fail(m, PEDONI_ERR_STATE, "pedoni_spawn_poisson must be the last spawn before pedoni_rebuild");
    if ((uint64_t)m->array_offset + m->compute_upper() + m->app_n + n + m->halo_cap > 0xFFFFFFF0ull)
        return fail(m, PEDONI_ERR_CAPACITY, "too many agents");
    int rc = ensure_app_capacity(m, m->app_n + n);
    if (rc != PEDONI_OK) return rc;
    // Through the pinned staging ring, chunk by chunk: one slot holds [pos | vel | v0 | dest] of up to
    // `per_slot` pedestrians. The host copy into the slot ends before this call returns (the caller's arrays
    // are borrowed for the call only, whatever kind of memory they are); the H2D copies read the slot.
    const size_t per_agent = sizeof(float2) * (vel_xy ? 2 : 1) + sizeof(float) + sizeof(uint32_t);
    const uint32_t per_slot = static_cast<uint32_t>((PedoniModel::kStageBytes - 64) / per_agent) & ~3u;  // 64: padding between the arrays
    for (uint32_t done = 0; done < n;) {
        const uint32_t k = std::min(per_slot, n - done), at = m->app_n + done;
        const int slot = m->stage_next;
        m->stage_next = (slot + 1) % PedoniModel::kStageSlots;
        if (!m->h_stage[slot]) {
            CUDA_TRY(m, cudaHostAlloc(&m->h_stage[slot], PedoniModel::kStageBytes, cudaHostAllocDefault));
            CUDA_TRY(m, cudaEventCreateWithFlags(&m->ev_stage[slot], cudaEventDisableTiming));
        }
        if (m->stage_busy[slot]) CUDA_TRY(m, cudaEventSynchronize(m->ev_stage[slot]));  // copies issued 4 chunks ago
        unsigned char* h = m->h_stage[slot];
        size_t off = 0;
        auto stage = [&](void* dst, const void* src, size_t bytes) -> cudaError_t {
            std::memcpy(h + off, src, bytes);
            cudaError_t e = cudaMemcpyAsync(dst, h + off, bytes, cudaMemcpyHostToDevice, m->stream);
            off += (bytes + 15) & ~static_cast<size_t>(15);
            return e;
        };
        CUDA_TRY(m, stage(m->app.pos + at, pos_xy + 2 * (size_t)done, sizeof(float2) * (size_t)k));
        if (vel_xy)
            CUDA_TRY(m, stage(m->app.vel + at, vel_xy + 2 * (size_t)done, sizeof(float2) * (size_t)k));
        else  // sfm.rs:53 velocity: Vec2::ZERO
            CUDA_TRY(m, cudaMemsetAsync(m->app.vel + at, 0, sizeof(float2) * (size_t)k, m->stream));
        CUDA_TRY(m, stage(m->app.v0 + at, v0 + done, sizeof(float) * (size_t)k));
        CUDA_TRY(m, stage(m->app.dest + at, dest + done, sizeof(uint32_t) * (size_t)k));
        CUDA_TRY(m, cudaEventRecord(m->ev_stage[slot], m->stream));
        m->stage_busy[slot] = true;
        done += k;
    }
    m->app_n += n;
    m->table_valid = false;
    return PEDONI_OK;
}

int pedoni_spawn(PedoniModel* m, uint32_t n, const float* pos_xy, const uint32_t* dest, const float* v0) {
    if (!m) return PEDONI_ERR_INVALID;
    CUDA_TRY(m, cudaSetDevice(m->device));
    return append_agents(m, n, pos_xy, dest, nullptr, v0);
}

int pedoni_spawn_groups(PedoniModel* m, uint32_t n_groups, const PedoniSpawnGroup* groups, uint64_t seed,
                        uint64_t counter) {
    if (!m || (n_groups > 0 && !groups)) return PEDONI_ERR_INVALID;
    CUDA_TRY(m, cudaSetDevice(m->device));
    std::vector<SpawnGroupDev> dev;
    uint64_t n = 0;
    for (uint32_t g = 0; g < n_groups; ++g) {
        if (groups[g].count == 0) continue;
        dev.push_back(SpawnGroupDev{groups[g].p1_x, groups[g].p1_y, groups[g].p2_x, groups[g].p2_y, groups[g].destination,
                                    static_cast<uint32_t>(n)});
        n += groups[g].count;
    }
    if (n == 0) return PEDONI_OK;
    if (m->app_on_device)
        return fail(m, PEDONI_ERR_STATE, "pedoni_spawn_poisson must be the last spawn before pedoni_rebuild");
    if ((uint64_t)m->array_offset + m->compute_upper() + m->app_n + n + m->halo_cap > 0xFFFFFFF0ull)
        return fail(m, PEDONI_ERR_CAPACITY, "too many agents");
    int rc = ensure_app_capacity(m, m->app_n + static_cast<uint32_t>(n));
    if (rc != PEDONI_OK) return rc;
    if (dev.size() > m->spawn_groups_cap) {
        cudaFree(m->d_spawn_groups);
        m->d_spawn_groups = nullptr;
        m->spawn_groups_cap = static_cast<uint32_t>(std::max<size_t>(dev.size(), 64));
        CUDA_TRY(m, cudaMalloc(&m->d_spawn_groups, sizeof(SpawnGroupDev) * m->spawn_groups_cap));
    }
    // the group table is a few dozen bytes: a pageable copy is staged before the call returns
    CUDA_TRY(m, cudaMemcpyAsync(m->d_spawn_groups, dev.data(), sizeof(SpawnGroupDev) * dev.size(), cudaMemcpyHostToDevice,
                                m->stream));
    spawn_groups_kernel<<<div_up(static_cast<uint32_t>(n), 256), 256, 0, m->stream>>>(
        m->app, m->app_n, static_cast<uint32_t>(n), static_cast<const SpawnGroupDev*>(m->d_spawn_groups),
        static_cast<uint32_t>(dev.size()), seed, counter);
    m->launches += 1;
    CUDA_TRY(m, cudaGetLastError());
    m->app_n += static_cast<uint32_t>(n);
    m->table_valid = false;
    return PEDONI_OK;
}

int pedoni_spawn_stream_seek(PedoniModel* m, uint64_t seed, uint64_t counter) {
    if (!m) return PEDONI_ERR_INVALID;
    CUDA_TRY(m, cudaSetDevice(m->device));
    if (!m->d_spawn_stream) {
        CUDA_TRY(m, cudaMalloc(&m->d_spawn_stream, sizeof(SpawnStreamState)));
        CUDA_TRY(m, cudaMalloc(&m->d_app_range, 2 * sizeof(uint32_t)));
    }
    const SpawnStreamState st{counter, counter, 0ull};
    CUDA_TRY(m, cudaMemcpyAsync(m->d_spawn_stream, &st, sizeof st, cudaMemcpyHostToDevice, m->stream));
    CUDA_TRY(m, cudaStreamSynchronize(m->stream));  // `st` is a local
    m->spawn_seed = seed;
    return PEDONI_OK;
}

int pedoni_spawn_stream_tell(PedoniModel* m, uint64_t* counter, uint64_t* spawned) {
    if (!m) return PEDONI_ERR_INVALID;
    CUDA_TRY(m, cudaSetDevice(m->device));
    if (!m->d_spawn_stream) return fail(m, PEDONI_ERR_STATE, "no spawn stream: call pedoni_spawn_stream_seek first");
    SpawnStreamState st{};
    CUDA_TRY(m, cudaMemcpyAsync(&st, m->d_spawn_stream, sizeof st, cudaMemcpyDeviceToHost, m->stream));
    CUDA_TRY(m, cudaStreamSynchronize(m->stream));
    if (counter) *counter = st.counter;
    if (spawned) *spawned = st.spawned;
    return check_device_error(m);
}

int pedoni_spawn_poisson(PedoniModel* m, uint32_t n_groups, const PedoniSpawnRate* rates) {
    if (!m || (n_groups > 0 && !rates)) return PEDONI_ERR_INVALID;
    CUDA_TRY(m, cudaSetDevice(m->device));
    if (!m->d_spawn_stream) return fail(m, PEDONI_ERR_STATE, "no spawn stream: call pedoni_spawn_stream_seek first");
    if (m->app_on_device)
        return fail(m, PEDONI_ERR_STATE, "pedoni_spawn_poisson must be the last spawn before pedoni_rebuild");
    if (n_groups == 0) return PEDONI_OK;
    std::vector<SpawnRateDev> dev(n_groups);
    uint64_t bound = 0;
    for (uint32_t g = 0; g < n_groups; ++g) {
        const double lambda = rates[g].frequency / 10.0;  // lib.rs:73
        // exp(-lambda) underflows to 0 beyond lambda ~ 745, and Knuth's loop (util.rs:82-85) then never terminates
        if (!(lambda >= 0.0) || lambda > 700.0)
            return fail(m, PEDONI_ERR_INVALID, "spawn frequency %g /s out of range [0, 7000]: the reference's Poisson loop "
                                               "(util.rs:78-89) does not terminate once exp(-frequency / 10) underflows", rates[g].frequency);
        // a Poisson draw above mean + 10 sigma + 10 has probability < 1e-20: the bound sizes grids and buffers
        const uint32_t max_count = static_cast<uint32_t>(std::ceil(lambda + 10.0 * std::sqrt(lambda) + 10.0));
        dev[g] = SpawnRateDev{rates[g].p1_x, rates[g].p1_y, rates[g].p2_x, rates[g].p2_y, rates[g].destination, max_count,
                              std::exp(-lambda)};
        bound += max_count;
    }
    if ((uint64_t)m->array_offset + m->compute_upper() + m->app_n + bound + m->halo_cap > 0xFFFFFFF0ull)
        return fail(m, PEDONI_ERR_CAPACITY, "too many agents");
    int rc = ensure_app_capacity(m, m->app_n + static_cast<uint32_t>(bound));
    if (rc != PEDONI_OK) return rc;
    if (n_groups > m->spawn_rates_cap || n_groups > m->spawn_groups_cap) {
        cudaFree(m->d_spawn_rates);
        cudaFree(m->d_spawn_groups);
        m->d_spawn_rates = nullptr, m->d_spawn_groups = nullptr;
        const uint32_t cap = std::max<uint32_t>(n_groups, 64);
        CUDA_TRY(m, cudaMalloc(&m->d_spawn_rates, sizeof(SpawnRateDev) * cap));
        CUDA_TRY(m, cudaMalloc(&m->d_spawn_groups, sizeof(SpawnGroupDev) * cap));
        m->spawn_rates_cap = m->spawn_groups_cap = cap;
    }
    // the rate table is a few hundred bytes of pageable memory: staged by the driver before the call returns
    CUDA_TRY(m, cudaMemcpyAsync(m->d_spawn_rates, dev.data(), sizeof(SpawnRateDev) * n_groups, cudaMemcpyHostToDevice,
                                m->stream));
    poisson_counts_kernel<<<1, 32, 0, m->stream>>>(m->d_spawn_rates, n_groups, m->spawn_seed, m->d_spawn_stream,
                                                   static_cast<SpawnGroupDev*>(m->d_spawn_groups), m->d_app_range, m->app_n,
                                                   m->d_error);
    spawn_groups_dev_kernel<<<div_up(static_cast<uint32_t>(bound), 256), 256, 0, m->stream>>>(
        m->app, m->app_n, m->d_app_range, static_cast<const SpawnGroupDev*>(m->d_spawn_groups), n_groups, m->spawn_seed,
        m->d_spawn_stream);
    m->launches += 2;
    CUDA_TRY(m, cudaGetLastError());
    m->app_n += static_cast<uint32_t>(bound);  // an upper bound from here on: the count lives in d_app_range
    m->app_on_device = true;
    m->table_valid = false;
    return PEDONI_OK;
}

int pedoni_upload_state(PedoniModel* m, uint32_t n, const float* pos_xy, const uint32_t* dest, const float* vel_xy,
                        const float* v0) {
    if (!m) return PEDONI_ERR_INVALID;
    CUDA_TRY(m, cudaSetDevice(m->device));
    if (n > 0 && !vel_xy) return fail(m, PEDONI_ERR_INVALID, "null velocity array");
    if (m->halo_inflight) {
        CUDA_TRY(m, cudaStreamWaitEvent(m->stream, m->ev_halo, 0));
        m->halo_inflight = false;
    }
    advance_tick(m, 0);
    // a preceding pedoni_step has already enrolled the (now discarded) residents in the next cell table
    CUDA_TRY(m, cudaMemsetAsync(m->d_cell_count, 0, sizeof(uint32_t) * (size_t)m->n_cells, m->stream));
    CUDA_TRY(m, cudaMemsetAsync(m->d_ovf_head, 0, sizeof(uint32_t) * (size_t)m->n_cells, m->stream));
    CUDA_TRY(m, cudaMemsetAsync(m->d_ovf_count, 0, sizeof(uint32_t), m->stream));
    reset_layout_kernel<<<1, 1, 0, m->stream>>>(m->d_ranges, m->array_offset, m->h_pub_dev, m->tick);
    m->launches += 1;
    m->owned_upper = 0;
    m->app_n = 0;
    m->app_on_device = false;
    m->keys_fresh = false;
    m->table_valid = false;
    m->halo_pending = false;
    // compute_upper() still counts ghost capacity, but every range is empty now: nothing is adopted.
    return append_agents(m, n, pos_xy, dest, vel_xy, v0);
}

static int rebuild_impl(PedoniModel* m) {
    cudaStream_t s = m->stream;
    const uint32_t resident = m->resident_hi();  // logical indices below this are array indices of buf[cur]
    const uint32_t total = resident + m->app_n;
    // worst case every input pedestrian is kept, and the ghosts from above land right behind them
    const uint64_t need = (uint64_t)m->array_offset + m->compute_upper() + m->app_n + (m->has_above ? m->halo_cap : 0);
    if (need > 0xFFFFFFF0ull) return fail(m, PEDONI_ERR_CAPACITY, "too many agents");
    int rc = ensure_capacity(m, static_cast<uint32_t>(need));
    if (rc != PEDONI_OK) return rc;
    SortInput in = make_sort_input(m);

    if (total > 0) {
        // The resident pedestrians enrolled in the next table in the force kernel's epilogue; only appended
        // spawns (or everybody, when no step preceded this rebuild) are keyed here.
        const uint32_t t_begin = m->keys_fresh ? resident : 0u;
        if (t_begin < total) {
            ScopedTimer t(m, kKey, s);
            key_kernel<<<div_up(total - t_begin, 256), 256, 0, s>>>(in, t_begin, total, m->grid, m->field, m->cell_sort(),
                                                                   m->d_error, m->d_arrived);
            m->launches += 1;
        }
    }
    advance_tick(m, (uint64_t)m->app_n + (uint64_t)m->n_sides() * m->halo_cap);
    ScanLayout layout{m->own_begin_cell, m->own_end_cell, static_cast<uint32_t>(m->grid.nx), m->has_below,
                      m->has_above,     m->ranges(m->rcur ^ 1), m->h_pub_dev,               m->tick};
    {
        // one persistent kernel: prefix scan over the cells + the stable reorder of the 24-byte state
        // Two instantiations: 2 CTAs per SM x 2 pedestrians in flight (large crowds) or 3 CTAs per SM x 1 (small ones,
        // where a CTA has one tile and only latency counts). Cells per thread: as few as keep every resident CTA busy
        // with one tile (a tile's latency grows with the cells a thread owns; a 10 M crowd on one GPU takes the full 8).
        // (small = fewer cells than one full-size tile per CTA at 2 CTAs per SM; 2.5 M pedestrians at 1 /m^2 are not: 65 vs 63 us)
        const bool small = m->n_cells <= 2u * static_cast<uint32_t>(m->sm_count) * kSortCtaThreads * kSortItems;
        const uint32_t slots = static_cast<uint32_t>((small ? 3 : PEDONI_SORT_MIN_BLOCKS) * m->sm_count);
        const uint32_t items = std::min<uint32_t>(kSortItems, std::max<uint32_t>(1u, div_up(m->n_cells, slots * kSortCtaThreads)));
        const uint32_t n_tiles = div_up(m->n_cells, items * kSortCtaThreads);
        SortScratch scratch{m->d_tile_status, m->d_tile_ticket, m->d_sort_done, n_tiles, m->sort_launches & 1u, items};
        m->sort_launches += 1;
        const uint32_t ctas = std::min<uint32_t>(n_tiles, slots);
        ScopedTimer t(m, kGather, s);
        if (small)
            sort_cells_kernel<3, 1><<<ctas, kSortCtaThreads, kSortSmemBytes, s>>>(
                in, m->cell_sort(), m->n_cells, m->array_offset, m->d_cell_start, scratch, layout, m->buf[m->cur ^ 1]);
        else
            sort_cells_kernel<PEDONI_SORT_MIN_BLOCKS, PEDONI_SORT_UNROLL><<<ctas, kSortCtaThreads, kSortSmemBytes, s>>>(
                in, m->cell_sort(), m->n_cells, m->array_offset, m->d_cell_start, scratch, layout, m->buf[m->cur ^ 1]);
        m->launches += 1;
    }
    m->cur ^= 1;
    m->rcur ^= 1;  // the layout the scan published describes buf[cur] from here on
    m->owned_upper = owned_bound(m, total);
    m->app_n = 0;
    m->app_on_device = false;
    m->keys_fresh = false;
    m->table_valid = true;
    m->ever_rebuilt = true;

    if (m->slab_count > 1) {
        const uint32_t threads = std::max<uint32_t>(m->halo_cap, 2 * m->grid.nx + 1);
        dim3 grid(div_up(threads, 256), 2);
        m->exchange_seq += 1;
        const uint32_t slot = m->exchange_seq & 1u;
        HaloMessage down = msg_of(m->d_send_dn), up = msg_of(m->d_send_up);
        PeerSignal sig{m->pack_counters(), nullptr, nullptr, m->exchange_seq};
        if (m->transport == PedoniModel::kTransportPeer) {
            // pack + send in one kernel: my first rows are the slab below's strip "from above", and vice versa
            if (m->has_below) {
                down = msg_of(m->recv_above(m->peer_arena_below, slot));
                sig.flag_down = m->flag_above(m->peer_arena_below);
            }
            if (m->has_above) {
                up = msg_of(m->recv_below(m->peer_arena_above, slot));
                sig.flag_up = m->flag_below(m->peer_arena_above);
            }
        }
        // Pack (and, with the peer-memory transport, send) on the EDGE stream, behind the sort: the main stream goes
        // straight on to the interior force launch, which only reads what the pack reads.
        CUDA_TRY(m, cudaEventRecord(m->ev_sorted, s));
        CUDA_TRY(m, cudaStreamWaitEvent(m->edge_stream, m->ev_sorted, 0));
        {
            ScopedTimer t(m, kPack, m->edge_stream);
            halo_pack_kernel<<<grid, 256, 0, m->edge_stream>>>(m->buf[m->cur], m->d_cell_start, m->own_begin_cell,
                                                              m->own_end_cell, m->grid.nx, m->halo_cap, down, up,
                                                              m->has_below, m->has_above, m->tick, m->d_error, sig);
            m->launches += 1;
        }
        CUDA_TRY(m, cudaEventRecord(m->ev_packed, m->edge_stream));
        m->halo_pending = true;
        if (m->transport == PedoniModel::kTransportPeer)
            rc = exchange_peer(m);
        else if (m->transport == PedoniModel::kTransportNccl)
            rc = exchange_nccl(m);
        if (rc != PEDONI_OK) return rc;
    }
    CUDA_TRY(m, cudaGetLastError());
    return PEDONI_OK;
}

int pedoni_rebuild(PedoniModel* m) {
    if (!m) return PEDONI_ERR_INVALID;
    CUDA_TRY(m, cudaSetDevice(m->device));
    if (m->halo_inflight) {  // rebuild without a step in between: the compute range is still being written
        CUDA_TRY(m, cudaStreamWaitEvent(m->stream, m->ev_halo, 0));
        m->halo_inflight = false;
    }
    // Worst case every input agent is kept; ghosts from above land right behind them.
    const uint64_t need = (uint64_t)m->array_offset + m->compute_upper() + m->app_n + (m->has_above ? m->halo_cap : 0);
    if (need > m->cap) {
        // The host bound may be stale: refresh it with the exact population before growing anything.
        int rc = sync_all(m);
        if (rc != PEDONI_OK) return rc;
        m->owned_upper = std::min<uint32_t>(m->owned_upper, static_cast<uint32_t>(m->h_pub[0]));
    }
    return rebuild_impl(m);
}

static int slab_exchange_local_impl(PedoniModel* const* models, int32_t n) {
    if (!models || n < 1) return PEDONI_ERR_INVALID;
    for (int i = 0; i < n; ++i) {
        PedoniModel* m = models[i];
        if (!m) return PEDONI_ERR_INVALID;
        if (m->slab_count != n || m->slab_rank != i)
            return fail(m, PEDONI_ERR_INVALID, "models[%d] is slab %d of %d, expected %d of %d", i, m->slab_rank,
                        m->slab_count, i, n);
        if (m->comm) return fail(m, PEDONI_ERR_STATE, "handle exchanges by itself; the in-process transport is for "
                                                       "handles without pedoni_comm_init");
        if (models[0]->exchange_seq != m->exchange_seq)
            return fail(m, PEDONI_ERR_STATE, "slabs of one group must be rebuilt in lockstep");
        if (n > 1 && !m->halo_pending) return fail(m, PEDONI_ERR_STATE, "no rebuild pending an exchange");
        if (i > 0 && (models[i - 1]->msg_bytes != m->msg_bytes))
            return fail(m, PEDONI_ERR_INVALID, "slabs disagree on the halo message size (halo_capacity)");
    }
    if (n == 1) return PEDONI_OK;
    // Strip copies run on the RECEIVER's edge stream, after the sender's pack.
    for (int i = 0; i < n; ++i) {
        PedoniModel* m = models[i];
        CUDA_TRY(m, cudaSetDevice(m->device));
        CUDA_TRY(m, cudaStreamWaitEvent(m->edge_stream, m->ev_packed, 0));
        if (m->has_below) {
            PedoniModel* o = models[i - 1];
            CUDA_TRY(m, cudaStreamWaitEvent(m->edge_stream, o->ev_packed, 0));
            CUDA_TRY(m, cudaMemcpyAsync(m->recv_below(m->d_arena, m->exchange_seq & 1u), o->d_send_up, m->msg_bytes,
                                        cudaMemcpyDefault, m->edge_stream));
        }
        if (m->has_above) {
            PedoniModel* o = models[i + 1];
            CUDA_TRY(m, cudaStreamWaitEvent(m->edge_stream, o->ev_packed, 0));
            CUDA_TRY(m, cudaMemcpyAsync(m->recv_above(m->d_arena, m->exchange_seq & 1u), o->d_send_dn, m->msg_bytes,
                                        cudaMemcpyDefault, m->edge_stream));
        }
        CUDA_TRY(m, cudaEventRecord(m->ev_peer, m->edge_stream));
    }
    // A sender may not repack (next tick, its main stream) before its neighbours have read the strips:
    // its edge stream waits for their copies, and its main stream waits for its edge stream every step.
    for (int i = 0; i < n; ++i) {
        PedoniModel* m = models[i];
        CUDA_TRY(m, cudaSetDevice(m->device));
        if (m->has_below) CUDA_TRY(m, cudaStreamWaitEvent(m->edge_stream, models[i - 1]->ev_peer, 0));
        if (m->has_above) CUDA_TRY(m, cudaStreamWaitEvent(m->edge_stream, models[i + 1]->ev_peer, 0));
        enqueue_unpack(m);
        CUDA_TRY(m, cudaGetLastError());
    }
    return PEDONI_OK;
}

int pedoni_slab_exchange_local(PedoniModel* const* models, int32_t n) {
    int rc = slab_exchange_local_impl(models, n);
    if (rc != PEDONI_OK) {  // a group call has no single handle: mirror the message into the handle-less slot
        g_create_error = "invalid slab group";
        for (int i = 0; models && i < n; ++i)
            if (models[i] && !models[i]->last_error.empty()) g_create_error = models[i]->last_error;
    }
    return rc;
}

int pedoni_step(PedoniModel* m) {
    if (!m) return PEDONI_ERR_INVALID;
    CUDA_TRY(m, cudaSetDevice(m->device));
    if (m->app_n > 0 || !m->table_valid)
        return fail(m, PEDONI_ERR_STATE,
                    "pedoni_step needs a freshly rebuilt neighbor grid: call pedoni_rebuild before every "
                    "pedoni_step and after pedoni_spawn / pedoni_upload_state (the reference rebuilds inside "
                    "spawn_pedestrians every tick, sfm.rs:58-77, lib.rs:85-90)");
    if (m->halo_pending)
        return fail(m, PEDONI_ERR_STATE,
                    "slab %d of %d has no ghost rows for this tick: join the ranks with pedoni_comm_init, or call "
                    "pedoni_slab_exchange_local after every handle's pedoni_rebuild", m->slab_rank, m->slab_count);
    // Interior rows need no ghost data: they run on the main stream while the halo is still in flight.
    launch_force(m, kRangeInterior, m->owned_upper, m->stream);
    if (m->slab_count > 1) {
        // the rows next to the slab boundaries: one launch for both edges
        if (m->has_below && m->has_above)
            launch_force(m, kRangeEdgeLo, 2 * m->halo_cap, m->edge_stream, kRangeEdgeHi);
        else
            launch_force(m, m->has_below ? kRangeEdgeLo : kRangeEdgeHi, 2 * m->halo_cap, m->edge_stream);
        CUDA_TRY(m, cudaEventRecord(m->ev_edge, m->edge_stream));
        CUDA_TRY(m, cudaStreamWaitEvent(m->stream, m->ev_edge, 0));
        m->halo_inflight = false;  // ev_edge is behind ev_halo on the edge stream
    }
    CUDA_TRY(m, cudaGetLastError());
    m->cur ^= 1;
    m->keys_fresh = true;
    m->table_valid = false;  // positions moved; the reference, too, rebuilds before every update (lib.rs:85-90)
    return PEDONI_OK;
}

// Blocks; returns the owned range of buf[cur].
static int sync_range(PedoniModel* m, uint32_t* begin, uint32_t* end) {
    if (m->app_on_device)
        return fail(m, PEDONI_ERR_STATE, "pedestrians drawn by pedoni_spawn_poisson are pending: call pedoni_rebuild first");
    int rc = sync_all(m);
    if (rc != PEDONI_OK) return rc;
    rc = check_device_error(m);
    if (rc != PEDONI_OK) return rc;
    *begin = m->array_offset;
    *end = m->array_offset + static_cast<uint32_t>(m->h_pub[0]);
    return PEDONI_OK;
}

int32_t pedoni_count(PedoniModel* m) {
    if (!m) return PEDONI_ERR_INVALID;
    CUDA_TRY(m, cudaSetDevice(m->device));
    uint32_t b, e;
    int rc = sync_range(m, &b, &e);
    if (rc != PEDONI_OK) return rc;
    // Spawn lists are replicated to every slab of a group and only the rebuild decides which slab keeps a
    // newcomer: a slab handle reports rebuilt pedestrians only (a whole-domain handle, like the reference's
    // PedestrianVec, also counts the ones appended since the last rebuild).
    return static_cast<int32_t>(e - b + (m->slab_count > 1 ? 0u : m->app_n));
}

// Non-blocking population: what the device last published (after the most recent COMPLETED rebuild).
int pedoni_count_published(PedoniModel* m, int32_t* count, uint32_t* tick) {
    if (!m || !count) return PEDONI_ERR_INVALID;
    const unsigned long long pub = *reinterpret_cast<volatile unsigned long long*>(m->h_pub);
    *count = static_cast<int32_t>(static_cast<uint32_t>(pub));
    if (tick) *tick = static_cast<uint32_t>(pub >> 32);
    return PEDONI_OK;
}

int pedoni_download(PedoniModel* m, float* pos_xy, uint32_t* dest, float* vel_xy, float* v0, uint32_t cap,
                    uint32_t* n_out) {
    if (!m) return PEDONI_ERR_INVALID;
    CUDA_TRY(m, cudaSetDevice(m->device));
    uint32_t b, e;
    int rc = sync_range(m, &b, &e);
    if (rc != PEDONI_OK) return rc;
    const uint32_t pending = m->slab_count > 1 ? 0u : m->app_n;  // see pedoni_count
    const uint32_t n_cur = e - b, n = n_cur + pending;
    if (n_out) *n_out = n;
    const uint32_t take_cur = std::min(n_cur, cap), take_app = std::min(pending, cap - take_cur);
    const AgentArrays& a = m->buf[m->cur];
    auto pull = [&](void* dst, const void* src_cur, const void* src_app, size_t elem) -> cudaError_t {
        if (!dst) return cudaSuccess;
        cudaError_t er = cudaSuccess;
        if (take_cur)
            er = cudaMemcpyAsync(dst, static_cast<const char*>(src_cur) + elem * b, elem * take_cur,
                                 cudaMemcpyDeviceToHost, m->stream);
        if (er == cudaSuccess && take_app)
            er = cudaMemcpyAsync(static_cast<char*>(dst) + elem * take_cur, src_app, elem * take_app,
                                 cudaMemcpyDeviceToHost, m->stream);
        return er;
    };
    CUDA_TRY(m, pull(pos_xy, a.pos, m->app.pos, sizeof(float2)));
    CUDA_TRY(m, pull(dest, a.dest, m->app.dest, sizeof(uint32_t)));
    CUDA_TRY(m, pull(vel_xy, a.vel, m->app.vel, sizeof(float2)));
    CUDA_TRY(m, pull(v0, a.v0, m->app.v0, sizeof(float)));
    CUDA_TRY(m, cudaStreamSynchronize(m->stream));
    if (cap < n) return fail(m, PEDONI_ERR_CAPACITY, "download capacity %u < %u agents", cap, n);
    return PEDONI_OK;
}

// Pipelined list_pedestrians: snapshot the owned (pos, destination) columns on the device (one D2D pass
// behind the work already enqueued), then copy the snapshot to the caller's buffers on a separate
// stream. The model may keep stepping meanwhile; pedoni_download_end waits for the oldest copy in flight.
static int download_begin_impl(PedoniModel* m, float* pos_xy, uint32_t* dest, uint8_t* dest8, uint32_t cap) {
    if (!m || !pos_xy || (!dest && !dest8)) return PEDONI_ERR_INVALID;
    CUDA_TRY(m, cudaSetDevice(m->device));
    if (dest8 && m->field.n_maps > 256)
        return fail(m, PEDONI_ERR_UNSUPPORTED, "byte-sized destinations need at most 256 potential maps (have %d)",
                    m->field.n_maps);
    // destinations travel as bytes if the caller wants bytes, or if this handle packs them and widens on the host
    const bool as_bytes = dest8 != nullptr || download_packs_destinations(m);
    if (m->dl_count == 2) return fail(m, PEDONI_ERR_STATE, "two pipelined downloads are already in flight");
    if (m->app_n > 0) return fail(m, PEDONI_ERR_STATE, "pedoni_download_begin with un-rebuilt spawns");
    if (!m->dl_stream) CUDA_TRY(m, cudaStreamCreateWithFlags(&m->dl_stream, cudaStreamNonBlocking));
    PedoniModel::DownloadSlot& d = m->dl[(m->dl_head + m->dl_count) & 1];
    if (!d.ev_snap) {
        CUDA_TRY(m, cudaEventCreateWithFlags(&d.ev_snap, cudaEventDisableTiming));
        CUDA_TRY(m, cudaEventCreateWithFlags(&d.ev_done, cudaEventDisableTiming));
        CUDA_TRY(m, cudaHostAlloc(&d.h_range, 2 * sizeof(uint32_t), cudaHostAllocDefault));
        CUDA_TRY(m, cudaMalloc(&d.d_range, 2 * sizeof(uint32_t)));
    }
    const uint32_t upper = m->owned_upper;  // host bound of the owned population; the exact count follows
    if (upper > d.cap) {  // the slot is free: its previous copy finished before pedoni_download_end returned
        cudaFree(d.d_pos);
        cudaFree(d.d_dest);
        cudaFree(d.d_dest8);
        if (d.h_dest8) cudaFreeHost(d.h_dest8);
        d.d_pos = nullptr, d.d_dest = nullptr, d.d_dest8 = nullptr, d.h_dest8 = nullptr, d.cap = 0;
        const uint32_t ncap = std::max<uint32_t>(upper + upper / 8, 1024);
        CUDA_TRY(m, cudaMalloc(&d.d_pos, sizeof(float2) * (size_t)ncap));
        d.cap = ncap;
    }
    // the destination staging this call needs (a handle may be asked for bytes and for words in turn)
    if (as_bytes && !d.d_dest8) CUDA_TRY(m, cudaMalloc(&d.d_dest8, (size_t)d.cap + 4));
    if (as_bytes && !dest8 && !d.h_dest8) CUDA_TRY(m, cudaHostAlloc(&d.h_dest8, (size_t)d.cap + 4, cudaHostAllocDefault));
    if (!as_bytes && !d.d_dest) CUDA_TRY(m, cudaMalloc(&d.d_dest, sizeof(uint32_t) * (size_t)d.cap));
    const AgentArrays& a = m->buf[m->cur];
    const uint32_t take = std::min(upper, cap);
    if (upper) {
        CUDA_TRY(m, cudaMemcpyAsync(d.d_pos, a.pos + m->array_offset, sizeof(float2) * (size_t)upper,
                                    cudaMemcpyDeviceToDevice, m->stream));
        if (as_bytes) {
            pack_dest_kernel<<<div_up(div_up(upper, 4), 256), 256, 0, m->stream>>>(a.dest + m->array_offset, upper,
                                                                                   reinterpret_cast<uint32_t*>(d.d_dest8));
            m->launches += 1;
        } else
            CUDA_TRY(m, cudaMemcpyAsync(d.d_dest, a.dest + m->array_offset, sizeof(uint32_t) * (size_t)upper,
                                        cudaMemcpyDeviceToDevice, m->stream));
    }
    CUDA_TRY(m, cudaMemcpyAsync(d.d_range, m->range(kRangeOwned), 2 * sizeof(uint32_t), cudaMemcpyDeviceToDevice,
                                m->stream));  // the next rebuild rewrites d_ranges: freeze the count with the data
    CUDA_TRY(m, cudaEventRecord(d.ev_snap, m->stream));
    CUDA_TRY(m, cudaStreamWaitEvent(m->dl_stream, d.ev_snap, 0));
    CUDA_TRY(m, cudaMemcpyAsync(d.h_range, d.d_range, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, m->dl_stream));
    if (take) {
        CUDA_TRY(m, cudaMemcpyAsync(pos_xy, d.d_pos, sizeof(float2) * (size_t)take, cudaMemcpyDeviceToHost, m->dl_stream));
        if (as_bytes)  // straight into the caller's byte array, or into the pinned staging _end widens from
            CUDA_TRY(m, cudaMemcpyAsync(dest8 ? dest8 : d.h_dest8, d.d_dest8, (size_t)take, cudaMemcpyDeviceToHost,
                                        m->dl_stream));
        else
            CUDA_TRY(m, cudaMemcpyAsync(dest, d.d_dest, sizeof(uint32_t) * (size_t)take, cudaMemcpyDeviceToHost,
                                        m->dl_stream));
    }
    CUDA_TRY(m, cudaEventRecord(d.ev_done, m->dl_stream));
    d.user_dest = dest;
    d.widen = as_bytes && !dest8;
    d.user_cap = cap;
    d.inflight = true;
    m->dl_count += 1;
    return PEDONI_OK;
}

int pedoni_download_begin(PedoniModel* m, float* pos_xy, uint32_t* dest, uint32_t cap) {
    if (!dest) return PEDONI_ERR_INVALID;
    return download_begin_impl(m, pos_xy, dest, nullptr, cap);
}
int pedoni_download_begin_u8(PedoniModel* m, float* pos_xy, uint8_t* dest8, uint32_t cap) {
    if (!dest8) return PEDONI_ERR_INVALID;
    return download_begin_impl(m, pos_xy, nullptr, dest8, cap);
}

int pedoni_download_end(PedoniModel* m, uint32_t* n_out) {
    if (!m) return PEDONI_ERR_INVALID;
    CUDA_TRY(m, cudaSetDevice(m->device));
    if (m->dl_count == 0) return fail(m, PEDONI_ERR_STATE, "no pipelined download in flight");
    PedoniModel::DownloadSlot& d = m->dl[m->dl_head];
    CUDA_TRY(m, cudaEventSynchronize(d.ev_done));
    d.inflight = false;
    m->dl_head ^= 1;
    m->dl_count -= 1;
    const uint32_t n = d.h_range[1] - d.h_range[0];
    if (n_out) *n_out = n;
    if (n > d.user_cap) return fail(m, PEDONI_ERR_CAPACITY, "download capacity %u < %u agents", d.user_cap, n);
    if (d.widen) {  // widen the byte-sized destinations into the caller's array (a later download may be copying meanwhile)
        const uint8_t* src = d.h_dest8;
        uint32_t* dst = d.user_dest;
        const long long count = n;
#pragma omp parallel for simd schedule(static) if (count > (1 << 16))
        for (long long i = 0; i < count; ++i) dst[i] = src[i];
    }
    return PEDONI_OK;
}

int pedoni_download_wire_bytes(const PedoniModel* m) {
    if (!m) return 0;
    return static_cast<int>(sizeof(float2)) + (download_packs_destinations(m) ? 1 : 4);
}

int pedoni_observe(PedoniModel* m, float y0, float y1, uint32_t n_bins, PedoniObservables* out) {
    if (!m || !out || n_bins > 64 || (n_bins > 0 && !(y1 > y0))) return PEDONI_ERR_INVALID;
    CUDA_TRY(m, cudaSetDevice(m->device));
    int rc = sync_all(m);
    if (rc != PEDONI_OK) return rc;
    rc = check_device_error(m);
    if (rc != PEDONI_OK) return rc;
    std::memset(out, 0, sizeof *out);
    out->n_bins = n_bins;
    CUDA_TRY(m, cudaMemsetAsync(m->d_observe, 0, sizeof(ObserveOut), m->stream));
    const uint32_t upper = std::max<uint32_t>(m->owned_upper, 1);
    const uint32_t blocks = std::min<uint32_t>(div_up(upper, 256), static_cast<uint32_t>(m->sm_count) * 8);
    observe_kernel<<<blocks, 256, 0, m->stream>>>(m->buf[m->cur], m->range(kRangeOwned), upper, y0,
                                                  n_bins ? static_cast<float>(n_bins) / (y1 - y0) : 0.0f, n_bins,
                                                  m->d_observe);
    m->launches += 1;
    ObserveOut h{};
    unsigned long long arrived[16];
    CUDA_TRY(m, cudaMemcpyAsync(&h, m->d_observe, sizeof h, cudaMemcpyDeviceToHost, m->stream));
    CUDA_TRY(m, cudaMemcpyAsync(arrived, m->d_arrived, sizeof arrived, cudaMemcpyDeviceToHost, m->stream));
    CUDA_TRY(m, cudaStreamSynchronize(m->stream));
    out->count = h.count;
    out->mean_speed = h.count ? h.speed_sum / static_cast<float>(h.count) : 0.0f;
    for (int k = 0; k < 16; ++k) {
        out->per_destination[k] = h.per_destination[k];
        out->arrived[k] = arrived[k];
    }
    for (uint32_t k = 0; k < n_bins; ++k) {
        out->bin_count[k] = h.bin_count[k];
        out->bin_mean_vx[k] = h.bin_count[k] ? h.bin_vx_sum[k] / static_cast<float>(h.bin_count[k]) : 0.0f;
    }
    return PEDONI_OK;
}

int pedoni_grid_shape(const PedoniModel* m, int32_t* ny, int32_t* nx) {
    if (!m || !ny || !nx) return PEDONI_ERR_INVALID;
    *ny = m->grid.ny;
    *nx = m->grid.nx;
    return PEDONI_OK;
}

int pedoni_cell_table(PedoniModel* m, uint32_t* indices, uint32_t cap, uint32_t* n_out) {
    if (!m) return PEDONI_ERR_INVALID;
    CUDA_TRY(m, cudaSetDevice(m->device));
    if (!m->ever_rebuilt) return fail(m, PEDONI_ERR_STATE, "no rebuild yet");
    const uint32_t n = m->own_end_cell - m->own_begin_cell + 1;
    if (n_out) *n_out = n;
    if (cap < n || !indices) return fail(m, PEDONI_ERR_CAPACITY, "cell table needs %u entries", n);
    int rc = sync_all(m);
    if (rc != PEDONI_OK) return rc;
    CUDA_TRY(m, cudaMemcpyAsync(indices, m->d_cell_start + m->own_begin_cell, sizeof(uint32_t) * (size_t)n,
                                cudaMemcpyDeviceToHost, m->stream));
    CUDA_TRY(m, cudaStreamSynchronize(m->stream));
    const uint32_t base = indices[0];  // local offsets: the first owned agent is entry 0
    if (base)
        for (uint32_t k = 0; k < n; ++k) indices[k] -= base;
    return PEDONI_OK;
}

// ---- measurement ----------------------------------------------------------------------------------
int pedoni_profile_enable(PedoniModel* m, int32_t enable) {
    if (!m) return PEDONI_ERR_INVALID;
    CUDA_TRY(m, cudaSetDevice(m->device));
    int rc = drain_timed(m);
    m->profiling = enable != 0;
    return rc;
}
int pedoni_profile_reset(PedoniModel* m) {
    if (!m) return PEDONI_ERR_INVALID;
    CUDA_TRY(m, cudaSetDevice(m->device));
    int rc = drain_timed(m);
    for (int k = 0; k < kNumKinds; ++k) {
        m->acc_ms[k] = 0;
        m->acc_launches[k] = 0;
    }
    m->acc_force_agents = 0;
    m->timeline.clear();
    return rc;
}
int pedoni_profile_read(PedoniModel* m, PedoniKernelTimes* out) {
    if (!m || !out) return PEDONI_ERR_INVALID;
    CUDA_TRY(m, cudaSetDevice(m->device));
    int rc = drain_timed(m);
    if (rc != PEDONI_OK) return rc;
    out->key_ms = m->acc_ms[kKey];
    out->histogram_ms = m->acc_ms[kHistogram];
    out->scan_ms = m->acc_ms[kScan];
    out->scatter_ms = m->acc_ms[kScatter];
    out->gather_ms = m->acc_ms[kGather];
    out->force_ms = m->acc_ms[kForce];
    out->comm_ms = m->acc_ms[kComm];
    out->key_launches = m->acc_launches[kKey];
    out->histogram_launches = m->acc_launches[kHistogram];
    out->scan_launches = m->acc_launches[kScan];
    out->scatter_launches = m->acc_launches[kScatter];
    out->gather_launches = m->acc_launches[kGather];
    out->force_launches = m->acc_launches[kForce];
    out->comm_launches = m->acc_launches[kComm];
    out->force_agents = m->acc_force_agents;
    out->force_edge_ms = m->acc_ms[kForceEdge];
    out->pack_ms = m->acc_ms[kPack];
    out->force_edge_launches = m->acc_launches[kForceEdge];
    out->pack_launches = m->acc_launches[kPack];
    return PEDONI_OK;
}
int pedoni_profile_timeline(PedoniModel* m, PedoniLaunchRecord* out, uint32_t cap, uint32_t* n_out) {
    if (!m || !n_out) return PEDONI_ERR_INVALID;
    CUDA_TRY(m, cudaSetDevice(m->device));
    int rc = drain_timed(m);
    if (rc != PEDONI_OK) return rc;
    *n_out = static_cast<uint32_t>(m->timeline.size());
    for (uint32_t k = 0; out && k < cap && k < m->timeline.size(); ++k) out[k] = m->timeline[k];
    return PEDONI_OK;
}
void* pedoni_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) {
        (void)cudaGetLastError();
        return nullptr;
    }
    return p;
}
void pedoni_host_free(void* p) {
    if (p) cudaFreeHost(p);
}
int pedoni_counters(PedoniModel* m, uint64_t* kernel_launches, uint64_t* pedestrian_updates) {
    if (!m) return PEDONI_ERR_INVALID;
    CUDA_TRY(m, cudaSetDevice(m->device));
    if (kernel_launches) *kernel_launches = m->launches;
    if (pedestrian_updates) {
        unsigned long long v = 0;
        int rc = sync_all(m);
        if (rc != PEDONI_OK) return rc;
        CUDA_TRY(m, cudaMemcpyAsync(&v, m->d_updates, sizeof v, cudaMemcpyDeviceToHost, m->stream));
        CUDA_TRY(m, cudaStreamSynchronize(m->stream));
        *pedestrian_updates = v;
    }
    return PEDONI_OK;
}
int pedoni_timer_begin(PedoniModel* m) {
    if (!m) return PEDONI_ERR_INVALID;
    CUDA_TRY(m, cudaSetDevice(m->device));
    CUDA_TRY(m, cudaEventRecord(m->timer_start, m->stream));
    m->timer_armed = true;
    m->timeline.clear();
    return PEDONI_OK;
}
int pedoni_timer_end(PedoniModel* m, float* ms) {
    if (!m || !ms) return PEDONI_ERR_INVALID;
    CUDA_TRY(m, cudaSetDevice(m->device));
    if (m->halo_inflight) {  // the last rebuild's exchange belongs to the timed region
        CUDA_TRY(m, cudaStreamWaitEvent(m->stream, m->ev_halo, 0));
        m->halo_inflight = false;
    }
    CUDA_TRY(m, cudaEventRecord(m->timer_stop, m->stream));
    CUDA_TRY(m, cudaEventSynchronize(m->timer_stop));
    CUDA_TRY(m, cudaEventElapsedTime(ms, m->timer_start, m->timer_stop));
    return PEDONI_OK;
}

// ---- multi-GPU slabs ------------------------------------------------------------------------------
int pedoni_comm_unique_id(void* out_id128) {
    std::string err;
    int rc = pedoni::slab_comm_unique_id(out_id128, &err);
    if (rc != PEDONI_OK) g_create_error = err;
    return rc;
}
int pedoni_comm_init(PedoniModel* m, const void* id128) {
    if (!m || !id128) return PEDONI_ERR_INVALID;
    CUDA_TRY(m, cudaSetDevice(m->device));
    if (m->slab_count <= 1) return fail(m, PEDONI_ERR_STATE, "pedoni_comm_init on a whole-domain handle");
    if (m->comm) return fail(m, PEDONI_ERR_STATE, "communicator already initialised");
    std::string err;
    if (m->exchange_seq != 0)
        return fail(m, PEDONI_ERR_STATE, "join the ranks with pedoni_comm_init before the first pedoni_rebuild");
    m->comm = pedoni::slab_comm_create(id128, m->slab_rank, m->slab_count, &err);
    if (!m->comm) return fail(m, PEDONI_ERR_COMM, "%s", err.c_str());
    return setup_peer_transport(m);
}
const char* pedoni_slab_transport(const PedoniModel* m) {
    if (!m || m->slab_count <= 1) return "none";
    switch (m->transport) {
        case PedoniModel::kTransportPeer: return "peer-memory (CUDA IPC over NVLink; pack kernel stores into the neighbour)";
        case PedoniModel::kTransportNccl: return "nccl send/recv";
        default: return "in-process (pedoni_slab_exchange_local)";
    }
}
int pedoni_field_textures(const PedoniModel* m) { return m && m->field_tex ? 1 : 0; }
int pedoni_wall_far_cells(const PedoniModel* m, uint64_t* far_cells, uint64_t* cells) {
    if (!m || !far_cells || !cells) return PEDONI_ERR_INVALID;
    *far_cells = m->far_cells;
    *cells = m->far_blocks_total;
    return PEDONI_OK;
}

int pedoni_halo_capacity(const PedoniModel* m, uint32_t* halo_capacity) {
    if (!m || !halo_capacity) return PEDONI_ERR_INVALID;
    *halo_capacity = m->halo_cap;
    return PEDONI_OK;
}

}  // extern "C"
