/*
 * pedoni_cuda.h — C ABI of the B200-native social-force backend (libpedoni_cuda.so).
 *
 * This is the drop-in boundary for ONE path of qt2/pedoni: the per-timestep pedestrian update that
 * sits behind the reference's plugin trait
 *
 *     pub trait PedestrianModel: Send + Sync            (pedoni-simulator/src/models/mod.rs:13-25)
 *         fn new(&SimulatorOptions, &Scenario, &Field) -> Self
 *         fn spawn_pedestrians(&mut self, &Field, Vec<Pedestrian>)
 *         fn update_states(&mut self, &Scenario, &Field)
 *         fn list_pedestrians(&self) -> Vec<Pedestrian>
 *         fn get_pedestrian_count(&self) -> i32
 *
 * A third implementation `SocialForceModelCuda` (ffi/) binds exactly these
 * entry points. Conventions:
 *   - plain C types only; every pointer argument is caller-owned and borrowed for the duration of
 *     the call; outputs go to caller-provided buffers with an explicit capacity;
 *   - every function returns PEDONI_OK (0) or a negative PedoniStatus; the message is available from
 *     pedoni_last_error(); nothing throws or aborts across the boundary (the reference panics via
 *     unwrap(), sfm_gpu.rs:51,69,79,127 — the Rust shim turns a non-zero status into panic!);
 *   - a handle may be used from a thread other than its creator (main.rs:79-97 moves the Simulator
 *     to a spawned thread) but is not re-entrant: one caller at a time;
 *   - pedoni_spawn / pedoni_rebuild / pedoni_step only ENQUEUE device work; pedoni_count,
 *     pedoni_download, pedoni_cell_table and pedoni_synchronize block;
 *   - there is no CPU fallback: without a CUDA device pedoni_create fails with PEDONI_ERR_CUDA.
 */
#ifndef PEDONI_CUDA_H
#define PEDONI_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PEDONI_ABI_VERSION 3

typedef struct PedoniModel PedoniModel;

typedef enum PedoniStatus {
    PEDONI_OK = 0,
    PEDONI_ERR_INVALID = -1,     /* bad argument / config */
    PEDONI_ERR_CUDA = -2,        /* CUDA runtime or driver error (message has the CUDA string) */
    PEDONI_ERR_STATE = -3,       /* call sequence error (e.g. step without a fresh rebuild) */
    PEDONI_ERR_CAPACITY = -4,    /* caller buffer too small / device buffer overflow */
    PEDONI_ERR_UNSUPPORTED = -5, /* option the CUDA path does not implement (use_neighbor_grid = 0) */
    PEDONI_ERR_COMM = -6         /* NCCL failure (multi-GPU slabs) */
} PedoniStatus;

/* Arithmetic of the force kernel. Cell keys, neighbor sets and the despawn predicate are IEEE-exact
 * in both modes. */
typedef enum PedoniMathMode {
    PEDONI_MATH_STRICT = 0, /* IEEE div/sqrt, no FMA contraction, reference summation order */
    PEDONI_MATH_FAST = 1    /* MUFU rcp/rsqrt/ex2 approximations + FMA; same summation order */
} PedoniMathMode;

/*
 * Creation parameters = the three arguments of PedestrianModel::new (models/mod.rs:14):
 *   SimulatorOptions  (lib.rs:108-135)  -> neighbor_grid_unit, field_grid_unit, use_neighbor_grid,
 *                                          use_distance_map        (gpu_work_size is unused, as in
 *                                          the reference: args.rs:40 vs :47-66)
 *   Scenario          (scenario.rs:10-27) -> field_size_*, obstacles
 *   Field             (field.rs:194-205)  -> field_ny/nx, distance_map, potential_maps
 * plus device placement. Set struct_size = sizeof(PedoniConfig).
 */
typedef struct PedoniConfig {
    uint32_t struct_size;
    int32_t device;              /* CUDA device ordinal */

    float field_size_x;          /* Scenario.field.size (metres) */
    float field_size_y;
    float neighbor_grid_unit;    /* SimulatorOptions.neighbor_grid_unit, default 1.4 */
    float field_grid_unit;       /* SimulatorOptions.field_grid_unit, default 0.25 (= Field.unit) */
    int32_t use_neighbor_grid;   /* must be 1; the O(N^2) debugging path (sfm.rs:157-185) is not built */
    int32_t use_distance_map;    /* 1: distance-map walls (sfm.rs:188-192); 0: segment walls (sfm.rs:193-237) */

    int32_t field_ny;            /* Field.shape.0 */
    int32_t field_nx;            /* Field.shape.1 */
    int32_t n_potential_maps;    /* Field.potential_maps.len() == Scenario.waypoints.len() */
    int32_t n_obstacles;         /* Scenario.obstacles.len() */
    const float* distance_map;   /* [field_ny * field_nx] row-major (y, x); copied to the device */
    const float* potential_maps; /* [n_potential_maps * field_ny * field_nx]; copied to the device */
    const float* obstacles;      /* [n_obstacles * 5]: x0, y0, x1, y1, width; may be NULL if 0 */

    uint32_t capacity;           /* initial agent capacity (0 = default); buffers grow on demand */
    int32_t math_mode;           /* PedoniMathMode */

    /* Spatial slab decomposition (one handle per GPU / process). slab_count <= 1: whole domain. */
    int32_t slab_rank;
    int32_t slab_count;

    void* stream;                /* optional cudaStream_t to enqueue on (NULL: the handle owns one) */

    /* Slab handles: capacity, in pedestrians, of one two-row ghost strip (= of one halo message).
     * 0 = auto (3x the two-row population of a uniformly filled slab of `capacity` agents, >= 4096).
     * All slabs of one decomposition must use the same value. */
    uint32_t halo_capacity;
} PedoniConfig;

int pedoni_abi_version(void);

/* PedestrianModel::new */
int pedoni_create(const PedoniConfig* config, PedoniModel** out_model);
void pedoni_destroy(PedoniModel* model);

/* Message of the last failure on this handle (or of the last failed pedoni_create if model == NULL). */
const char* pedoni_last_error(const PedoniModel* model);

/*
 * PedestrianModel::spawn_pedestrians, first half (sfm.rs:49-56): append n pedestrians with zero
 * velocity. desired_speed is an INPUT (the reference draws it from an unseeded global RNG,
 * sfm.rs:54; the caller draws it so device code stays RNG-free). n may be 0.
 * On a slab handle every rank is given the same list; a rank keeps the agents whose cell row it owns.
 * The arrays may be pageable or pinned, and all of them may be reused as soon as the call returns: they
 * are copied into a pinned staging ring of the handle on the host and travel from there, so the call
 * never waits for the device (it only blocks if more than 4 MiB of spawns are still in flight).
 */
int pedoni_spawn(PedoniModel* model, uint32_t n, const float* pos_xy, const uint32_t* destination,
                 const float* desired_speed);

/*
 * The same append, with the pedestrians DRAWN ON THE DEVICE: what
 * Simulator::new / Simulator::tick do before calling the plugin (lib.rs:37-52, 67-86) — for every spawn
 * group `count` positions p1.lerp(p2, u) on the origin waypoint's line and a desired speed
 * N(1.34, 0.26) per pedestrian (sfm.rs:54) — without a host-to-device copy of the arrays.
 * The draws come from the counter-based stream of pedoni_b200/simulator.py `SpawnStream`
 * (u64 number k of the stream = splitmix64(seed ^ k * 0x2545F4914F6CDD1D)): u for pedestrian j of the
 * call is u64 number counter + j (top 24 bits), the speed's two u64 are numbers counter + n + j and
 * counter + 2n + j, n = the sum of the counts — bit for bit what SpawnStream.f32 / .normal_approx return,
 * so a host-drawn and a device-drawn run are identical. The per-group counts come from the caller here;
 * pedoni_spawn_poisson draws them on the device as well. Consumes 3n stream numbers.
 */
typedef struct PedoniSpawnGroup {
    float p1_x, p1_y, p2_x, p2_y; /* Scenario.waypoints[origin].line */
    uint32_t destination;
    uint32_t count;
} PedoniSpawnGroup;
int pedoni_spawn_groups(PedoniModel* model, uint32_t n_groups, const PedoniSpawnGroup* groups, uint64_t seed,
                        uint64_t counter);

/*
 * The arrivals of a tick drawn entirely ON THE DEVICE: per periodic spawn group the
 * count = poisson(frequency / 10) of lib.rs:74 / util.rs:78-89 (Knuth's multiplication loop over fastrand::f64()),
 * then positions and desired speeds as in pedoni_spawn_groups. How many stream numbers a Poisson draw consumes
 * depends on its outcome, so the HANDLE owns the position in the counter stream: pedoni_spawn_stream_seek sets it
 * (and the seed), every pedoni_spawn_poisson advances it on the device in the order of pedoni_b200/simulator.py
 * `Simulator._spawn` (all counts, then one uniform per pedestrian, then the two speed blocks), and
 * pedoni_spawn_stream_tell reads it back together with the number of pedestrians drawn so far (blocks). A host
 * loop that draws from the same stream (tests: the oracle Simulator) sees bit-identical arrivals. No host-to-device
 * copy per tick beyond the few bytes of the rate table. Must be the last spawn before pedoni_rebuild; until that
 * rebuild pedoni_count / pedoni_download return PEDONI_ERR_STATE (only the device knows how many were drawn).
 * frequency must lie in [0, 7000] pedestrians / s: beyond ~7450 exp(-frequency / 10) underflows to 0 and the
 * reference's loop (util.rs:82-85) never terminates; such a rate is refused with PEDONI_ERR_INVALID.
 */
typedef struct PedoniSpawnRate {
    float p1_x, p1_y, p2_x, p2_y; /* Scenario.waypoints[origin].line */
    uint32_t destination;
    double frequency;             /* PedestrianSpawnConfig::Periodic { frequency: f64 } (scenario.rs:64), pedestrians / s */
} PedoniSpawnRate;
int pedoni_spawn_stream_seek(PedoniModel* model, uint64_t seed, uint64_t counter);
int pedoni_spawn_stream_tell(PedoniModel* model, uint64_t* counter, uint64_t* pedestrians_drawn);
int pedoni_spawn_poisson(PedoniModel* model, uint32_t n_groups, const PedoniSpawnRate* rates);

/*
 * PedestrianModel::spawn_pedestrians, second half (sfm.rs:58-77): neighbor-grid rebuild.
 * Cell key = trunc(pos / unit) (neighbor_grid.rs:27), out-of-grid agents are dropped
 * (neighbor_grid.rs:29), agents with potential <= 0.25 are despawned (sfm.rs:69), survivors are
 * reordered cell-major, stably within a cell (sfm.rs:66-68), and the cell-start table
 * `neighbor_grid_indices` (sfm.rs:61-75) is rebuilt.
 */
int pedoni_rebuild(PedoniModel* model);

/* PedestrianModel::update_states (sfm.rs:91-255): forces + integration, dt = 0.1 s. */
int pedoni_step(PedoniModel* model);

/* PedestrianModel::get_pedestrian_count (sfm.rs:267-269). Blocks. Negative = PedoniStatus.
 * A whole-domain handle counts the pedestrians appended since the last rebuild too (the reference's
 * PedestrianVec holds them). A slab handle does not, here and in pedoni_download: spawn lists are
 * replicated to every slab of a group and only the rebuild decides which slab keeps a newcomer, so
 * between pedoni_spawn and pedoni_rebuild a slab reports its rebuilt pedestrians only. */
int32_t pedoni_count(PedoniModel* model);

/* The same without blocking: the population of the owned rows as of the most
 * recent rebuild the DEVICE has completed, and that rebuild's ordinal. The reference reads the count after
 * every tick (lib.rs:95), which costs a host/device synchronisation per tick; a headless loop that only
 * logs the series can read this lagging value instead (it trails the host by the ticks still in flight). */
int pedoni_count_published(PedoniModel* model, int32_t* count, uint32_t* rebuild_ordinal);

/*
 * PedestrianModel::list_pedestrians (sfm.rs:257-265) plus the model's private columns.
 * Writes min(count, cap) agents in the model's current order; *n_out = count. pos_xy and
 * destination are what the trait returns; vel_xy / desired_speed may be NULL. Blocks.
 * Returns PEDONI_ERR_CAPACITY (after filling cap agents) if cap < count.
 */
int pedoni_download(PedoniModel* model, float* pos_xy, uint32_t* destination, float* vel_xy,
                    float* desired_speed, uint32_t cap, uint32_t* n_out);

/*
 * list_pedestrians without stalling the model (the reference copies every pedestrian to the host every
 * tick, main.rs:95). pedoni_download_begin snapshots (position, destination) of the owned pedestrians on
 * the device behind the work already enqueued and starts copying the snapshot to the caller's buffers
 * on a separate stream; it returns at once, and spawn / rebuild / step may be called meanwhile.
 * pedoni_download_end blocks until the OLDEST copy in flight has landed and reports how many pedestrians
 * were written. The buffers must stay valid (pinned memory makes the copy truly asynchronous) until the
 * matching _end returns; up to two pipelined downloads may be in flight per handle (begin k, begin k+1,
 * end k, ...: PCIe stays busy while _end finishes tick k on the host, see pedoni_download_wire_bytes).
 * Needs a rebuilt state (no pending pedoni_spawn).
 */
int pedoni_download_begin(PedoniModel* model, float* pos_xy, uint32_t* destination, uint32_t cap);
int pedoni_download_end(PedoniModel* model, uint32_t* n_out);
/* The same with the destinations delivered as BYTES (9 instead of 12 bytes per pedestrian over PCIe, nothing to
 * widen on the host; a caller that builds `Pedestrian { pos, destination: usize }` records touches every element
 * anyway). Needs at most 256 potential maps (PEDONI_ERR_UNSUPPORTED otherwise). Completed by pedoni_download_end. */
int pedoni_download_begin_u8(PedoniModel* model, float* pos_xy, uint8_t* destination8, uint32_t cap);

/*
 * Aggregate observables reduced ON THE DEVICE instead of downloading every
 * pedestrian: population (total and by destination), mean speed, the lane histogram of a counter-flow
 * corridor (mean x-velocity and population per y-bin over [y0, y1), at most 64 bins; n_bins = 0 skips
 * it) and the cumulative number of pedestrians that reached their destination, by destination (removed by
 * the predicate of sfm.rs:69 — not those that left the grid or turned NaN); its time derivative is the
 * flow rate. Destinations >= 15 are lumped into entry 15. Slab handles report their own rows. Blocks.
 */
typedef struct PedoniObservables {
    uint32_t count;
    float mean_speed;
    uint32_t per_destination[16];
    uint64_t arrived[16];
    uint32_t n_bins;
    uint32_t bin_count[64];
    float bin_mean_vx[64];
} PedoniObservables;
int pedoni_observe(PedoniModel* model, float y0, float y1, uint32_t n_bins, PedoniObservables* out);

/* Replace the whole agent state (parity tests, checkpoint restore). Unsorted: call pedoni_rebuild. */
int pedoni_upload_state(PedoniModel* model, uint32_t n, const float* pos_xy, const uint32_t* destination,
                        const float* vel_xy, const float* desired_speed);

/* Neighbor grid shape (neighbor_grid.rs:15-16): *ny = ceil(size_y / unit), *nx = ceil(size_x / unit). */
int pedoni_grid_shape(const PedoniModel* model, int32_t* ny, int32_t* nx);

/* `neighbor_grid_indices` (sfm.rs:22): ny*nx + 1 exclusive cell starts of the last rebuild. Blocks.
 * On a slab handle: the rows this rank owns, (row1 - row0) * nx + 1 entries, local offsets. */
int pedoni_cell_table(PedoniModel* model, uint32_t* indices, uint32_t cap, uint32_t* n_out);

/* Wait for all enqueued work of this handle. */
int pedoni_synchronize(PedoniModel* model);

/* ---- field precompute on the host (the step before the path) -----------
 *
 * `Field::from_scenario` (field.rs:220-232): outline rasterisation of obstacles and waypoints plus the
 * reference's fast-marching variant. Pure host code (OpenMP over the potential maps), no GPU needed.
 * obstacles / waypoints: 5 floats each (x0, y0, x1, y1, width). Outputs are caller-allocated:
 * obstacle_exist[fy*fx] (0/1), distance_map[fy*fx], potential_maps[n_waypoints*fy*fx], row-major (y, x),
 * with (fy, fx) from pedoni_field_shape = ceil(size / unit) (field.rs:25-26). The arrays are exactly what
 * PedoniConfig.distance_map / potential_maps expect.
 */
int pedoni_field_shape(float size_x, float size_y, float unit, int32_t* field_ny, int32_t* field_nx);
int pedoni_field_build(float size_x, float size_y, float unit, int32_t n_obstacles, const float* obstacles,
                       int32_t n_waypoints, const float* waypoints, uint8_t* obstacle_exist, float* distance_map,
                       float* potential_maps);

/* The same precompute ON THE GPU (second half), for domains where the serial heap
 * marching takes minutes (the 10 M synthetic crowd: three maps of 12 656^2 cells). Same rasterised inputs, same
 * per-cell upwind update (field.rs:177-187) and costs, solved as the fixed point of that update by a block-iterative
 * scheme (32 x 32 tiles relaxed in shared memory, a changed edge wakes the neighbouring tile) instead of by marching.
 * Equal to pedoni_field_build to rounding (<= 1e-3 field cells on every shipped scenario, most maps bit-identical:
 * tests/test_gpu_field_device.py), not bit for bit: sums are formed in a different order. Outputs as for
 * pedoni_field_build (host arrays); *passes_out = relaxation passes over the grid, summed over the maps. Needs a CUDA
 * device (PEDONI_ERR_CUDA otherwise: the host builder is pedoni_field_build). */
int pedoni_field_build_device(int32_t device, float size_x, float size_y, float unit, int32_t n_obstacles,
                              const float* obstacles, int32_t n_waypoints, const float* waypoints,
                              uint8_t* obstacle_exist, float* distance_map, float* potential_maps, int32_t* passes_out);

/* ---- measurement (bench.py): device-side timers on the handle's own stream ---------------------- */

typedef struct PedoniKernelTimes {
    /* accumulated since pedoni_profile_reset, milliseconds, measured with CUDA events.
     * key = key_kernel (spawned / uploaded agents only); gather = sort_cells_kernel, the whole rebuild (chained scan
     * over the cell populations + stable reorder of the state, one launch); histogram, scan and scatter = 0 (the
     * per-cell count and the cell membership are fused into the force epilogue and key_kernel, the scan and the
     * scatter into the sort; kept for ABI stability); force: see below; comm = exchange + unpack on the edge stream
     * (overlaps force). */
    double key_ms, histogram_ms, scan_ms, scatter_ms, gather_ms, force_ms, comm_ms;
    uint64_t key_launches, histogram_launches, scan_launches, scatter_launches, gather_launches,
        force_launches, comm_launches;
    uint64_t force_agents; /* agents processed by the timed force launches */
    /* slab handles (ABI 3): force = the interior launch on the main stream; force_edge = the launches on the rows
     * next to a slab boundary (edge stream, concurrent with the interior launch); pack = halo_pack_kernel (main
     * stream; with the peer-memory transport it also stores the strips into the neighbours). */
    double force_edge_ms, pack_ms;
    uint64_t force_edge_launches, pack_launches;
} PedoniKernelTimes;

/* One timed launch (profiling on): kind = 0 key, 4 sort (the rebuild), 5 force (interior), 6 exchange +
 * unpack, 7 force (edge rows), 8 pack; stream = 0 main, 1 edge; start / stop in milliseconds after the
 * last pedoni_timer_begin (CUDA events; both streams share the origin). */
typedef struct PedoniLaunchRecord {
    int32_t kind;
    int32_t stream;
    float start_ms, stop_ms;
} PedoniLaunchRecord;
/* The launches timed since the last pedoni_timer_begin / pedoni_profile_reset (at most 4096): a two-stream
 * timeline of the tick, for finding the critical path of a slab handle. *n_out = records available. Blocks. */
int pedoni_profile_timeline(PedoniModel* model, PedoniLaunchRecord* out, uint32_t cap, uint32_t* n_out);

int pedoni_profile_enable(PedoniModel* model, int32_t enable); /* per-kernel events; off by default */
int pedoni_profile_reset(PedoniModel* model);
int pedoni_profile_read(PedoniModel* model, PedoniKernelTimes* out); /* blocks */

/* Totals since creation: kernels this library launched, and pedestrian-updates integrated
 * (sum over pedoni_step calls of the live population; accumulated on the device). Blocks. */
int pedoni_counters(PedoniModel* model, uint64_t* kernel_launches, uint64_t* pedestrian_updates);

/* Bracket a region with two events on the handle's stream; end blocks and returns elapsed ms. */
int pedoni_timer_begin(PedoniModel* model);
int pedoni_timer_end(PedoniModel* model, float* elapsed_ms);

/* ---- multi-GPU slabs (one handle per GPU; NCCL send/recv between slab neighbours) -----------------
 *
 * The neighbor grid is cut into `slab_count` bands of whole cell rows (pedoni_slab_rows). A slab handle
 * owns the pedestrians whose cell row is in its band and additionally holds two GHOST rows each side,
 * copies of its neighbours' boundary rows. Per tick it integrates its own rows plus the nearest ghost
 * row each side (redundantly, bit-identically to the owner), so a pedestrian that walks across a band
 * boundary is adopted by the receiving slab's own rebuild: halo strips and migrating pedestrians
 * travel in the same message. After every pedoni_rebuild the handle sends its first/last two owned
 * rows (24 B per pedestrian + the rows' cell table) to rank-1 / rank+1 on a second stream while
 * pedoni_step's interior kernel runs; only the four rows next to a boundary wait for the exchange.
 * Concatenating the ranks' pedoni_download outputs in rank order reproduces the whole-domain order.
 */

/* Rows [*row0, *row1) of the ny-row neighbor grid owned by `rank` of `count` slabs. Pure host. */
int pedoni_slab_rows(int32_t ny, int32_t count, int32_t rank, int32_t* row0, int32_t* row1);

#define PEDONI_COMM_ID_BYTES 128
/* Rank 0 creates the NCCL unique id; the host program (torch.distributed, MPI, a socket) hands the
 * 128 bytes to every rank, which then joins with pedoni_comm_init (collective over the slab ranks; call it
 * before the first pedoni_rebuild). */
int pedoni_comm_unique_id(void* out_id128);
int pedoni_comm_init(PedoniModel* model, const void* id128);

/* In-process transport for slabs that live in ONE process (several handles on one GPU, or one per GPU
 * under one host thread): after pedoni_rebuild has been called on every handle, exchange their ghost
 * rows with device-to-device copies. models[i] must be slab i of n. Used where NCCL cannot be (tests
 * of the slab path on a single GPU); handles joined with pedoni_comm_init exchange by themselves. */
int pedoni_slab_exchange_local(PedoniModel* const* models, int32_t n);

/* How this handle's ghost rows travel: "peer-memory ..." (default after pedoni_comm_init when the ranks can
 * map each other's memory through CUDA IPC: the pack kernel stores the strips straight into the
 * neighbour over NVLink and raises a flag its unpack kernel polls; NCCL only bootstraps the handles),
 * "nccl send/recv" (fallback, or PEDONI_SLAB_TRANSPORT=nccl), "in-process ...", or "none". */
const char* pedoni_slab_transport(const PedoniModel* model);

/* 1 if the force kernel reads the field maps through texture gathers (PEDONI_MATH_FAST handles keep a
 * second copy of the maps tiled into one 2D CUDA array: four gathers instead of sixteen loads per 4x4
 * footprint, same texel values, verified against the linear copy at creation), 0 if it uses plain loads
 * (strict handles; no memory for the second copy; maps that do not tile into 32768 x 32768 texels;
 * PEDONI_FIELD_TEXTURES=0). */
int pedoni_field_textures(const PedoniModel* model);

/* Wall term far from every obstacle (PEDONI_MATH_FAST, distance-map walls; sfm.rs:188-192). The term is
 * 10 * 0.2 * exp(-distance / 0.2): 8.5e-18 m/s^2 at 8 m, ten orders of magnitude below one ulp of a pedestrian's
 * acceleration. At creation the handle marks the blocks of 8 x 8 distance-map texels in which, for every position
 * whose field coordinate falls into the block, all texels the term's 4x4 footprints can touch lie inside the map,
 * hold a distance >= 8 m, and rise or fall strictly along one axis (so that the gradient the reference normalises
 * cannot vanish: its NaN, which removes the pedestrian, cannot occur there). In marked blocks the force kernel does
 * not fetch the distance map at all; ridges of the map, map borders and everything closer than 8 m to an obstacle
 * keep the full evaluation. Results differ from the unmasked fast path by < 1e-17 m/s^2 per tick (tests:
 * bit-identical trajectories). Strict handles never skip. PEDONI_WALL_CUTOFF=0 disables the mask. Returns how many
 * of the mask's blocks are marked (0 of n: no mask). */
int pedoni_wall_far_cells(const PedoniModel* model, uint64_t* far_blocks, uint64_t* blocks);

/* Bytes per pedestrian that pedoni_download_begin / _end move over PCIe: 8 (position) + 4 (destination), or
 * + 1 on a whole-domain handle with at most 256 potential maps — destinations then travel as bytes and
 * pedoni_download_end widens them into the caller's uint32 array on the host. Slab handles keep 4-byte
 * destinations (each GPU has its own link; the shared host is the limit there). PEDONI_DOWNLOAD_PACK=0 / 1
 * overrides. */
int pedoni_download_wire_bytes(const PedoniModel* model);

/* Page-locked host memory for the caller's download buffers (pedoni_download / pedoni_download_begin copy at the
 * PCIe rate only into pinned memory; a Rust or C host need not link the CUDA runtime for it). NULL on failure. */
void* pedoni_host_alloc(size_t bytes);
void pedoni_host_free(void* ptr);

/* The halo capacity in effect (0 on a whole-domain handle). */
int pedoni_halo_capacity(const PedoniModel* model, uint32_t* halo_capacity);

#ifdef __cplusplus
}
#endif
#endif /* PEDONI_CUDA_H */
