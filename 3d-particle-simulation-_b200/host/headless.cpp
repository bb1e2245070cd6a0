// headless.cpp — steps the default scene without a window (BASELINE.json config 1: the reference's
// built-in scene at its built-in particle count, 100 steps).  Replaces the fixed-timestep driver of
// src/bin/main.rs:183-203 minus the UI.  Usage: headless [n=1000] [steps=100] [seed=42] [per_step=1]
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "particle_3d.hpp"

int main(int argc, char **argv) {
    const size_t n = argc > 1 ? strtoull(argv[1], nullptr, 10) : 1000;
    const int steps = argc > 2 ? atoi(argv[2]) : 100;
    const uint64_t seed = argc > 3 ? strtoull(argv[3], nullptr, 10) : 42;
    const bool per_step = argc > 4 ? atoi(argv[4]) != 0 : true;
    const float ts = 1.0f / 60.0f;  // main.rs:164,194
    try {
        particle_3d::Particles sim;
        particle_3d::Particles::default_scene(sim, n, seed);
        if (n != 1000) {  // keep density 1 like the default scene
            sim.world_size = 10.0f * cbrtf((float)n / 1000.0f);
            p3d_scene_uniform(seed, n, sim.world_size, sim.id_count,
                              reinterpret_cast<p3d_particle *>(sim.active_particles.data()));
        }
        const auto t0 = std::chrono::steady_clock::now();
        if (per_step) {
            for (int s = 0; s < steps; ++s) sim.update(ts);  // what main.rs:199 does
        } else {
            sim.run(ts, steps);
        }
        const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        double ke = 0, px = 0, py = 0, pz = 0;
        for (const auto &p : sim.active_particles) {
            ke += 0.5 * ((double)p.velocity.x * p.velocity.x + (double)p.velocity.y * p.velocity.y +
                         (double)p.velocity.z * p.velocity.z);
            px += p.velocity.x; py += p.velocity.y; pz += p.velocity.z;
        }
        printf("{\"n\": %zu, \"steps\": %d, \"ms_per_step\": %.6f, \"ke\": %.9e, \"p\": [%.6e, %.6e, %.6e]}\n", n, steps,
               ms / steps, ke, px, py, pz);
    } catch (const particle_3d::Panic &e) {
        fprintf(stderr, "panic (%d): %s\n", e.code, e.what());
        return 101;  // Rust's panic exit code
    }
    return 0;
}
