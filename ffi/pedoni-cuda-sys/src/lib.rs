//! UNVERIFIED SOURCE (no Rust toolchain in the build image): raw bindings of include/pedoni_cuda.h,
//! ABI version 2. Field order and types mirror the C header one to one; tests/test_capi_symbols.py
//! checks the same layout against the header for the ctypes mirror.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

pub const PEDONI_ABI_VERSION: c_int = 2;
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
}
