//! Headless stepper: the default scene of the reference's `SimulationApp::new` stepped without a
//! window (BASELINE.json config 1).  NOT BUILT IN THIS IMAGE; `host/headless.cpp` is the built twin.
use particle_3d::{Particle, Particles};

fn main() {
    let n: usize = std::env::args().nth(1).and_then(|s| s.parse().ok()).unwrap_or(1000);
    let steps: usize = std::env::args().nth(2).and_then(|s| s.parse().ok()).unwrap_or(100);
    // splitmix64, the same stream as p3d_scene_uniform (csrc/p3d_scene.cpp)
    struct SplitMix(u64);
    impl SplitMix {
        fn next(&mut self) -> u64 {
            self.0 = self.0.wrapping_add(0x9E3779B97F4A7C15);
            let mut z = self.0;
            z = (z ^ (z >> 30)).wrapping_mul(0xBF58476D1CE4E5B9);
            z = (z ^ (z >> 27)).wrapping_mul(0x94D049BB133111EB);
            z ^ (z >> 31)
        }
        fn unit(&mut self) -> f32 { (self.next() >> 40) as f32 * (1.0 / 16_777_216.0) }
    }
    let mut rng = SplitMix(42);
    let world = 10.0f32;
    let mut parts = Vec::with_capacity(n);
    for _ in 0..n {
        let x = -5.0 + world * rng.unit();
        let y = -5.0 + world * rng.unit();
        let z = -5.0 + world * rng.unit();
        let id = (rng.next() % 5) as u32;
        parts.push(Particle { position: cgmath::vec3(x, y, z), velocity: cgmath::vec3(0.0, 0.0, 0.0), id });
    }
    let mut sim = Particles {
        world_size: world, id_count: 5, colors: vec![],
        attraction_matrix: vec![0.5, 1.0, -0.5, 0.0, -1.0, 1.0, 1.0, 1.0, 0.0, -1.0, 0.0, 0.0, 0.5, 1.5, -1.0,
                                0.0, 0.0, 0.0, 0.0, -1.0, 1.0, 1.0, 1.0, 1.0, 0.5],
        particle_effect_radius: 2.0, coefficient: 0.97, interaction_force: 1.0, min_pull_ratio: 0.3,
        active_particles: parts, past_particles: vec![], walls: false, acceleration: cgmath::vec3(0.0, 0.0, 0.0),
    };
    let t0 = std::time::Instant::now();
    for _ in 0..steps { sim.update(1.0 / 60.0); }
    let ke: f64 = sim.active_particles.iter().map(|p| 0.5 * (p.velocity.x as f64).powi(2)
        + 0.5 * (p.velocity.y as f64).powi(2) + 0.5 * (p.velocity.z as f64).powi(2)).sum();
    println!("n={n} steps={steps} ms_per_step={:.4} ke={ke:.9e}", t0.elapsed().as_secs_f64() * 1e3 / steps as f64);
}
