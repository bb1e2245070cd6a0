//! Raw bindings to include/p3d.h — only what `Particles::update` needs.
pub const P3D_OPT_FAITHFUL: std::os::raw::c_int = 5;
use std::os::raw::{c_char, c_int};

#[repr(C)]
pub struct P3dEngine { _private: [u8; 0] }

/// `p3d_params` (include/p3d.h): the scalar fields of `Particles` that `update` reads.
#[repr(C)]
pub struct P3dParams {
    pub world_size: f32,
    pub coefficient: f32,
    pub interaction_force: f32,
    pub min_pull_ratio: f32,
    pub particle_effect_radius: f32,
    pub accel: [f32; 3],
    pub walls: u32,
    pub id_count: u32,
    pub attraction_matrix: *const f32,
}

unsafe extern "C" {
    pub fn p3d_create(device: c_int, out: *mut *mut P3dEngine) -> c_int;
    /// One handle over several devices of the node (include/p3d.h); `update` is unchanged above it.
    pub fn p3d_create_multi(devices: *const c_int, n_dev: c_int, out: *mut *mut P3dEngine) -> c_int;
    /// `P3D_OPT_FAITHFUL` = 5: reproduce the reference's bucket double-visit quirk (src/lib.rs:195-206).
    pub fn p3d_set_option(eng: *mut P3dEngine, option: c_int, value: c_int) -> c_int;
    pub fn p3d_destroy(eng: *mut P3dEngine);
    pub fn p3d_last_error() -> *const c_char;
    pub fn p3d_update(eng: *mut P3dEngine, prm: *const P3dParams, ts: f32,
                      input: *const crate::Particle, output: *mut crate::Particle, n: usize) -> c_int;
}
