//! UNVERIFIED SOURCE (no Rust toolchain in the build image): raw bindings of include/pedoni_cuda.h,
//! ABI version 3. Field order and types mirror the C header one to one; tests/test_capi_symbols.py
//! checks the same layout against the header for the ctypes mirror.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

pub const PEDONI_ABI_VERSION: c_int = 3;
pub const PEDONI_OK: c_int = 0;
pub const PEDONI_MATH_STRICT: i32 = 0;
pub const PEDONI_MATH_FAST: i32 = 1;
pub const PEDONI_COMM_ID_BYTES: usize = 128;

#[repr(C)]
pub struct PedoniModel {
    _private: [u8; 0],
}

#[repr(C)]
pub struct PedoniConfig {
    pub struct_size: u32,
    pub device: i32,
    pub field_size_x: f32,
    pub field_size_y: f32,
    pub neighbor_grid_unit: f32,
    pub field_grid_unit: f32,
    pub use_neighbor_grid: i32,
    pub use_distance_map: i32,
    pub field_ny: i32,
    pub field_nx: i32,
    pub n_potential_maps: i32,
    pub n_obstacles: i32,
    pub distance_map: *const f32,
    pub potential_maps: *const f32,
    pub obstacles: *const f32,
    pub capacity: u32,
    pub math_mode: i32,
    pub slab_rank: i32,
    pub slab_count: i32,
    pub stream: *mut c_void,
    pub halo_capacity: u32,
}

#[repr(C)]
pub struct PedoniSpawnGroup {
    pub p1_x: f32,
    pub p1_y: f32,
    pub p2_x: f32,
    pub p2_y: f32,
    pub destination: u32,
    pub count: u32,
}

#[repr(C)]
pub struct PedoniObservables {
    pub count: u32,
    pub mean_speed: f32,
    pub per_destination: [u32; 16],
    pub arrived: [u64; 16],
    pub n_bins: u32,
    pub bin_count: [u32; 64],
    pub bin_mean_vx: [f32; 64],
}

#[repr(C)]
pub struct PedoniSpawnRate {
    pub p1_x: f32,
    pub p1_y: f32,
    pub p2_x: f32,
    pub p2_y: f32,
    pub destination: u32,
    pub frequency: f64,
}

#[repr(C)]
pub struct PedoniKernelTimes {
    pub key_ms: f64,
    pub histogram_ms: f64,
    pub scan_ms: f64,
    pub scatter_ms: f64,
    pub gather_ms: f64,
    pub force_ms: f64,
    pub comm_ms: f64,
    pub key_launches: u64,
    pub histogram_launches: u64,
    pub scan_launches: u64,
    pub scatter_launches: u64,
    pub gather_launches: u64,
    pub force_launches: u64,
    pub comm_launches: u64,
    pub force_agents: u64,
    pub force_edge_ms: f64,
    pub pack_ms: f64,
    pub force_edge_launches: u64,
    pub pack_launches: u64,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct PedoniLaunchRecord {
    pub kind: i32,
    pub stream: i32,
    pub start_ms: f32,
    pub stop_ms: f32,
}

extern "C" {
    pub fn pedoni_abi_version() -> c_int;
    pub fn pedoni_create(config: *const PedoniConfig, out_model: *mut *mut PedoniModel) -> c_int;
    pub fn pedoni_destroy(model: *mut PedoniModel);
    pub fn pedoni_last_error(model: *const PedoniModel) -> *const c_char;
    pub fn pedoni_spawn(model: *mut PedoniModel, n: u32, pos_xy: *const f32, destination: *const u32,
                        desired_speed: *const f32) -> c_int;
    pub fn pedoni_rebuild(model: *mut PedoniModel) -> c_int;
    pub fn pedoni_step(model: *mut PedoniModel) -> c_int;
    pub fn pedoni_count(model: *mut PedoniModel) -> i32;
    pub fn pedoni_download(model: *mut PedoniModel, pos_xy: *mut f32, destination: *mut u32, vel_xy: *mut f32,
                           desired_speed: *mut f32, cap: u32, n_out: *mut u32) -> c_int;
    pub fn pedoni_synchronize(model: *mut PedoniModel) -> c_int;
    pub fn pedoni_slab_rows(ny: i32, count: i32, rank: i32, row0: *mut i32, row1: *mut i32) -> c_int;
    pub fn pedoni_comm_unique_id(out_id128: *mut c_void) -> c_int;
    pub fn pedoni_comm_init(model: *mut PedoniModel, id128: *const c_void) -> c_int;
    // everything below: not needed by the trait shim (ffi/sfm_cuda.rs), declared so that this block mirrors
    // the whole header (tests/test_capi_symbols.py checks the names)
    pub fn pedoni_spawn_groups(model: *mut PedoniModel, n_groups: u32, groups: *const PedoniSpawnGroup, seed: u64,
                               counter: u64) -> c_int;
    pub fn pedoni_spawn_stream_seek(model: *mut PedoniModel, seed: u64, counter: u64) -> c_int;
    pub fn pedoni_spawn_stream_tell(model: *mut PedoniModel, counter: *mut u64, pedestrians_drawn: *mut u64) -> c_int;
    pub fn pedoni_spawn_poisson(model: *mut PedoniModel, n_groups: u32, rates: *const PedoniSpawnRate) -> c_int;
    pub fn pedoni_count_published(model: *mut PedoniModel, count: *mut i32, rebuild_ordinal: *mut u32) -> c_int;
    pub fn pedoni_download_begin(model: *mut PedoniModel, pos_xy: *mut f32, destination: *mut u32, cap: u32) -> c_int;
    pub fn pedoni_download_end(model: *mut PedoniModel, n_out: *mut u32) -> c_int;
    pub fn pedoni_download_begin_u8(model: *mut PedoniModel, pos_xy: *mut f32, destination8: *mut u8, cap: u32) -> c_int;
    pub fn pedoni_download_wire_bytes(model: *const PedoniModel) -> c_int;
    pub fn pedoni_observe(model: *mut PedoniModel, y0: f32, y1: f32, n_bins: u32, out: *mut PedoniObservables) -> c_int;
    pub fn pedoni_upload_state(model: *mut PedoniModel, n: u32, pos_xy: *const f32, destination: *const u32,
                               vel_xy: *const f32, desired_speed: *const f32) -> c_int;
    pub fn pedoni_grid_shape(model: *const PedoniModel, ny: *mut i32, nx: *mut i32) -> c_int;
    pub fn pedoni_cell_table(model: *mut PedoniModel, indices: *mut u32, cap: u32, n_out: *mut u32) -> c_int;
    pub fn pedoni_field_shape(size_x: f32, size_y: f32, unit: f32, field_ny: *mut i32, field_nx: *mut i32) -> c_int;
    pub fn pedoni_field_build(size_x: f32, size_y: f32, unit: f32, n_obstacles: i32, obstacles: *const f32,
                              n_waypoints: i32, waypoints: *const f32, obstacle_exist: *mut u8,
                              distance_map: *mut f32, potential_maps: *mut f32) -> c_int;
    pub fn pedoni_field_build_device(device: i32, size_x: f32, size_y: f32, unit: f32, n_obstacles: i32,
                                     obstacles: *const f32, n_waypoints: i32, waypoints: *const f32,
                                     obstacle_exist: *mut u8, distance_map: *mut f32, potential_maps: *mut f32,
                                     passes_out: *mut i32) -> c_int;
    pub fn pedoni_field_textures(model: *const PedoniModel) -> c_int;
    pub fn pedoni_wall_far_cells(model: *const PedoniModel, far_cells: *mut u64, cells: *mut u64) -> c_int;
    pub fn pedoni_profile_enable(model: *mut PedoniModel, enable: i32) -> c_int;
    pub fn pedoni_profile_reset(model: *mut PedoniModel) -> c_int;
    pub fn pedoni_profile_read(model: *mut PedoniModel, out: *mut PedoniKernelTimes) -> c_int;
    pub fn pedoni_counters(model: *mut PedoniModel, kernel_launches: *mut u64, pedestrian_updates: *mut u64) -> c_int;
    pub fn pedoni_timer_begin(model: *mut PedoniModel) -> c_int;
    pub fn pedoni_timer_end(model: *mut PedoniModel, elapsed_ms: *mut f32) -> c_int;
    pub fn pedoni_slab_exchange_local(models: *const *mut PedoniModel, n: i32) -> c_int;
    pub fn pedoni_slab_transport(model: *const PedoniModel) -> *const c_char;
    pub fn pedoni_halo_capacity(model: *const PedoniModel, halo_capacity: *mut u32) -> c_int;
    pub fn pedoni_profile_timeline(model: *mut PedoniModel, out: *mut PedoniLaunchRecord, cap: u32, n_out: *mut u32) -> c_int;
    pub fn pedoni_host_alloc(bytes: usize) -> *mut c_void;
    pub fn pedoni_host_free(ptr: *mut c_void);
}
