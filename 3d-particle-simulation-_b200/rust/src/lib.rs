//! Drop-in `src/lib.rs` for the `particle_3d` crate: the public types and the signature of
//! `Particles::update` are those of the reference (src/lib.rs:12-33,130); the body of `update`
//! hands the step to the B200 engine through the C ABI in include/p3d.h.  `src/bin/main.rs` of the
//! reference compiles against this file unchanged (it constructs `Particles { .. }` with a struct
//! literal, so the engine handle cannot live in the struct: it is kept in a thread-local).
//!
//! NOT BUILT IN THIS IMAGE (no cargo/rustc); executable verification goes through the C ABI.
use std::{cell::RefCell, ffi::CStr};

use encase::ShaderType;

mod ffi;

/// Same fields as the reference; `#[repr(C)]` is the one addition, so that a `Vec<Particle>` can be
/// handed to the engine without a copy (28 bytes: two Vector3<f32> and a u32).
#[repr(C)]
#[derive(Clone, Copy, ShaderType, Debug)]
pub struct Particle {
    pub position: cgmath::Vector3<f32>,
    pub velocity: cgmath::Vector3<f32>,
    pub id: u32,
}

pub struct Particles {
    pub world_size: f32,
    pub active_particles: Vec<Particle>,
    pub past_particles: Vec<Particle>,
    pub id_count: u32,
    pub attraction_matrix: Vec<f32>,
    pub colors: Vec<cgmath::Vector3<f32>>,
    pub coefficient: f32,
    pub interaction_force: f32,
    pub min_pull_ratio: f32,
    pub particle_effect_radius: f32,
    pub walls: bool,
    pub acceleration: cgmath::Vector3<f32>,
}

struct Engine(*mut ffi::P3dEngine);
impl Drop for Engine {
    fn drop(&mut self) { unsafe { ffi::p3d_destroy(self.0) } }
}
thread_local! { static ENGINE: RefCell<Option<Engine>> = const { RefCell::new(None) }; }

/// Devices the engine runs on: `P3D_DEVICES="0,1,2,3"` (one handle drives them all), default device 0.
fn devices_from_env() -> Vec<i32> {
    std::env::var("P3D_DEVICES").ok()
        .map(|s| s.split(',').filter_map(|d| d.trim().parse().ok()).collect::<Vec<i32>>())
        .filter(|v| !v.is_empty()).unwrap_or_else(|| vec![0])
}

/// `P3D_FAITHFUL=1`: reproduce the reference's bucket double-visit quirk (two of the 27 hashed cells colliding
/// modulo N scan a bucket twice, src/lib.rs:195-206).  The default evaluates every in-range pair exactly once,
/// which deviates from the reference for the ~26*k/N of the particles per step that the quirk touches.
fn faithful_from_env() -> bool {
    std::env::var("P3D_FAITHFUL").map(|v| !v.is_empty() && v != "0").unwrap_or(false)
}

fn last_error() -> String {
    unsafe { CStr::from_ptr(ffi::p3d_last_error()).to_string_lossy().into_owned() }
}

impl Particles {
    /// One time step on the GPU.  Panics exactly where the reference panics: when
    /// `world_size < 2 * particle_effect_radius` and when a particle id is not below `id_count`;
    /// it also panics when no B200 is available (there is no CPU fallback).
    pub fn update(&mut self, ts: f32) -> Vec<Particle> {
        assert!(self.attraction_matrix.len() >= (self.id_count * self.id_count) as usize);
        let prm = ffi::P3dParams {
            world_size: self.world_size,
            coefficient: self.coefficient,
            interaction_force: self.interaction_force,
            min_pull_ratio: self.min_pull_ratio,
            particle_effect_radius: self.particle_effect_radius,
            accel: [self.acceleration.x, self.acceleration.y, self.acceleration.z],
            walls: self.walls as u32,
            id_count: self.id_count,
            attraction_matrix: self.attraction_matrix.as_ptr(),
        };
        let n = self.active_particles.len();
        // old `active` becomes `past`; the engine writes the new state into the recycled buffer
        std::mem::swap(&mut self.active_particles, &mut self.past_particles);
        self.active_particles.clear();
        self.active_particles.reserve(n);
        let rc = ENGINE.with(|cell| {
            let mut slot = cell.borrow_mut();
            if slot.is_none() {
                let mut raw = std::ptr::null_mut();
                let devs = devices_from_env();
                let rc = unsafe {
                    if devs.len() > 1 { ffi::p3d_create_multi(devs.as_ptr(), devs.len() as i32, &mut raw) }
                    else { ffi::p3d_create(devs[0], &mut raw) }
                };
                if rc != 0 { return rc; }
                let rc = unsafe { ffi::p3d_set_option(raw, ffi::P3D_OPT_FAITHFUL, faithful_from_env() as i32) };
                if rc != 0 { return rc; }
                *slot = Some(Engine(raw));
            }
            unsafe {
                ffi::p3d_update(slot.as_ref().unwrap().0, &prm, ts, self.past_particles.as_ptr(),
                                self.active_particles.as_mut_ptr(), n)
            }
        });
        if rc != 0 { panic!("particle_3d GPU step failed ({rc}): {}", last_error()); }
        unsafe { self.active_particles.set_len(n) };
        self.active_particles.clone()
    }
}
