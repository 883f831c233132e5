//! UNVERIFIED SOURCE (no Rust toolchain in the build image) — the third `PedestrianModel`
//! implementation, to live at pedoni-simulator/src/models/sfm_cuda.rs next to sfm.rs and sfm_gpu.rs.
//! See INTEGRATION.md for the three one-line edits that wire it in (models/mod.rs, lib.rs, args.rs).
use std::ffi::CStr;
use std::sync::Mutex;

use glam::Vec2;
use pedoni_cuda_sys as sys;

use super::PedestrianModel;
use crate::{field::Field, scenario::Scenario, SimulatorOptions};

/// Page-locked staging for `list_pedestrians`: the device-to-host copy runs at the PCIe rate only into pinned
/// memory (pageable `Vec`s are staged by the driver at about a third of it). Grown on demand, reused every tick.
struct Staging {
    pos: *mut f32,
    dest: *mut u32,
    cap: usize,
}

impl Staging {
    fn ensure(&mut self, n: usize) {
        if n <= self.cap {
            return;
        }
        unsafe {
            sys::pedoni_host_free(self.pos as *mut _);
            sys::pedoni_host_free(self.dest as *mut _);
            let cap = (n + n / 8).max(4096);
            self.pos = sys::pedoni_host_alloc(cap * 8) as *mut f32;
            self.dest = sys::pedoni_host_alloc(cap * 4) as *mut u32;
            assert!(!self.pos.is_null() && !self.dest.is_null(), "pedoni_host_alloc failed");
            self.cap = cap;
        }
    }
}

pub struct SocialForceModelCuda {
    handle: *mut sys::PedoniModel,
    staging: Mutex<Staging>, // list_pedestrians takes &self (models/mod.rs:23)
}

// The handle is moved to the simulation thread (main.rs:79-97) and used from one thread at a time.
unsafe impl Send for SocialForceModelCuda {}
unsafe impl Sync for SocialForceModelCuda {}

impl SocialForceModelCuda {
    /// The trait's methods are infallible and the reference's GPU backend unwrap()s (sfm_gpu.rs:51,69,79,127).
    fn check(&self, rc: i32) -> i32 {
        if rc < 0 {
            let msg = unsafe { CStr::from_ptr(sys::pedoni_last_error(self.handle)) };
            panic!("pedoni_cuda error {rc}: {}", msg.to_string_lossy());
        }
        rc
    }
}

impl PedestrianModel for SocialForceModelCuda {
    fn new(options: &SimulatorOptions, scenario: &Scenario, field: &Field) -> Self {
        let potentials: Vec<f32> = field.potential_maps.iter().flat_map(|m| m.iter().cloned()).collect();
        let distance: Vec<f32> = field.distance_map.iter().cloned().collect();
        let obstacles: Vec<f32> = scenario
            .obstacles
            .iter()
            .flat_map(|o| [o.line[0].x, o.line[0].y, o.line[1].x, o.line[1].y, o.width])
            .collect();
        let config = sys::PedoniConfig {
            struct_size: std::mem::size_of::<sys::PedoniConfig>() as u32,
            device: 0,
            field_size_x: scenario.field.size.x,
            field_size_y: scenario.field.size.y,
            neighbor_grid_unit: options.neighbor_grid_unit,
            field_grid_unit: field.unit,
            use_neighbor_grid: options.use_neighbor_grid as i32,
            use_distance_map: options.use_distance_map as i32,
            field_ny: field.shape.0 as i32,
            field_nx: field.shape.1 as i32,
            n_potential_maps: field.potential_maps.len() as i32,
            n_obstacles: scenario.obstacles.len() as i32,
            distance_map: distance.as_ptr(),
            potential_maps: potentials.as_ptr(),
            obstacles: obstacles.as_ptr(),
            capacity: 0,
            math_mode: sys::PEDONI_MATH_FAST,
            slab_rank: 0,
            slab_count: 1,
            stream: std::ptr::null_mut(),
            halo_capacity: 0,
        };
        let mut handle = std::ptr::null_mut();
        let rc = unsafe { sys::pedoni_create(&config, &mut handle) };
        if rc < 0 {
            let msg = unsafe { CStr::from_ptr(sys::pedoni_last_error(std::ptr::null())) };
            panic!("pedoni_create failed ({rc}): {}", msg.to_string_lossy());
        }
        SocialForceModelCuda { handle, staging: Mutex::new(Staging { pos: std::ptr::null_mut(), dest: std::ptr::null_mut(), cap: 0 }) }
    }

    fn spawn_pedestrians(&mut self, _field: &Field, spawned: Vec<super::Pedestrian>) {
        let pos: Vec<f32> = spawned.iter().flat_map(|p| [p.pos.x, p.pos.y]).collect();
        let dest: Vec<u32> = spawned.iter().map(|p| p.destination as u32).collect();
        // sfm.rs:54 — the draw stays on the host so the device path has no RNG
        let v0: Vec<f32> = spawned.iter().map(|_| fastrand_contrib::f32_normal_approx(1.34, 0.26)).collect();
        let n = dest.len() as u32;
        self.check(unsafe { sys::pedoni_spawn(self.handle, n, pos.as_ptr(), dest.as_ptr(), v0.as_ptr()) });
        self.check(unsafe { sys::pedoni_rebuild(self.handle) }); // sfm.rs:58-77
    }

    fn update_states(&mut self, _scenario: &Scenario, _field: &Field) {
        self.check(unsafe { sys::pedoni_step(self.handle) }); // sfm.rs:91-255
    }

    fn list_pedestrians(&self) -> Vec<super::Pedestrian> {
        let n = self.get_pedestrian_count() as usize;
        let mut st = self.staging.lock().unwrap();
        st.ensure(n);
        let mut n_out = 0u32;
        self.check(unsafe {
            sys::pedoni_download(self.handle, st.pos, st.dest, std::ptr::null_mut(), std::ptr::null_mut(), n as u32,
                                 &mut n_out)
        });
        let (pos, dest) = unsafe {
            (std::slice::from_raw_parts(st.pos, 2 * n_out as usize), std::slice::from_raw_parts(st.dest, n_out as usize))
        };
        (0..n_out as usize)
            .map(|i| super::Pedestrian { pos: Vec2::new(pos[2 * i], pos[2 * i + 1]), destination: dest[i] as usize })
            .collect()
    }

    fn get_pedestrian_count(&self) -> i32 {
        self.check(unsafe { sys::pedoni_count(self.handle) })
    }
}

impl Drop for SocialForceModelCuda {
    fn drop(&mut self) {
        let st = self.staging.lock().unwrap();
        unsafe {
            sys::pedoni_host_free(st.pos as *mut _);
            sys::pedoni_host_free(st.dest as *mut _);
            sys::pedoni_destroy(self.handle)
        }
    }
}
