// p3d_scene.cpp — seeded scene generation (host only).
//
// The reference's generator is private to its binary and unseeded (`rand::thread_rng` per rayon
// thread, src/bin/main.rs:60-87), so no seed can reproduce it.  These functions restate its
// DISTRIBUTION (uniform positions in [-W/2, W/2]^3, zero velocity, uniform id) with a
// counter-based splitmix64 stream, and expose the default scene constants of
// src/bin/main.rs:123-148.  Shared by the bench, the tests and the headless stepper.
#include <cmath>
#include <cstring>

#include "p3d.h"

namespace {
struct SplitMix64 {
    uint64_t s;
    explicit SplitMix64(uint64_t seed) : s(seed) {}
    uint64_t next() {
        uint64_t z = (s += 0x9E3779B97F4A7C15ULL);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
        return z ^ (z >> 31);
    }
    // 24 random mantissa bits -> [0,1)
    float unit_f32() { return (float)(next() >> 40) * (1.0f / 16777216.0f); }
    double unit_f64() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
};
}  // namespace

extern "C" void p3d_scene_default_params(p3d_params *prm, float matrix25[25]) {
    // src/bin/main.rs:133-139
    static const float kDefault[25] = {0.5f, 1.0f, -0.5f, 0.0f, -1.0f, 1.0f, 1.0f, 1.0f, 0.0f,
                                       -1.0f, 0.0f, 0.0f, 0.5f, 1.5f, -1.0f, 0.0f, 0.0f, 0.0f,
                                       0.0f, -1.0f, 1.0f, 1.0f, 1.0f, 1.0f, 0.5f};
    std::memcpy(matrix25, kDefault, sizeof(kDefault));
    prm->world_size = 10.0f;             // main.rs:124
    prm->id_count = 5;                   // main.rs:125 (MAX_PARTICLE_TYPES, main.rs:13)
    prm->particle_effect_radius = 2.0f;  // main.rs:140
    prm->coefficient = 0.97f;            // main.rs:141
    prm->interaction_force = 1.0f;       // main.rs:142
    prm->min_pull_ratio = 0.3f;          // main.rs:143
    prm->walls = 0;                      // main.rs:146
    prm->accel[0] = prm->accel[1] = prm->accel[2] = 0.0f;  // main.rs:147
    prm->attraction_matrix = matrix25;
}

extern "C" void p3d_scene_uniform(uint64_t seed, size_t n, float world_size, uint32_t id_count,
                                  p3d_particle *out) {
    SplitMix64 rng(seed);
    const float half = world_size * 0.5f;  // main.rs:66
    for (size_t i = 0; i < n; ++i) {
        p3d_particle p;
        p.px = -half + world_size * rng.unit_f32();  // main.rs:67-71 (distribution)
        p.py = -half + world_size * rng.unit_f32();
        p.pz = -half + world_size * rng.unit_f32();
        p.vx = p.vy = p.vz = 0.0f;                   // main.rs:73
        p.id = id_count ? (uint32_t)(rng.next() % id_count) : 0u;  // main.rs:75
        out[i] = p;
    }
}

extern "C" void p3d_scene_plummer(uint64_t seed, size_t n, float world_size, float scale_a,
                                  uint32_t id_count, p3d_particle *out) {
    SplitMix64 rng(seed);
    const double half = 0.5 * (double)world_size;
    const double two_pi = 6.283185307179586476925286766559;
    for (size_t i = 0; i < n; ++i) {
        double x, y, z;
        for (;;) {
            // Plummer cumulative mass M(r) = r^3 / (r^2 + a^2)^(3/2); invert for r.
            double u = rng.unit_f64();
            if (u < 1e-12) u = 1e-12;
            const double rad = (double)scale_a / std::sqrt(std::pow(u, -2.0 / 3.0) - 1.0);
            const double cz = 2.0 * rng.unit_f64() - 1.0;
            const double phi = two_pi * rng.unit_f64();
            const double sz = std::sqrt(1.0 - cz * cz);
            x = rad * sz * std::cos(phi);
            y = rad * sz * std::sin(phi);
            z = rad * cz;
            if (std::fabs(x) < half && std::fabs(y) < half && std::fabs(z) < half) break;
        }
        p3d_particle p;
        p.px = (float)x; p.py = (float)y; p.pz = (float)z;
        p.vx = p.vy = p.vz = 0.0f;
        p.id = id_count ? (uint32_t)(rng.next() % id_count) : 0u;
        out[i] = p;
    }
}
