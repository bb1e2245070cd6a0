// particle_3d.hpp — C++ mirror of the reference crate's public API (src/lib.rs) over the C ABI
// (include/p3d.h).  Header only.  The reference is Rust and no Rust toolchain exists in this image,
// so this is the compiled-language host side that can actually be built and run here; the Rust shim
// with the same shape is under ../rust/.
//
//   particle_3d::Particle   <- `pub struct Particle`   src/lib.rs:12-17
//   particle_3d::Particles  <- `pub struct Particles`  src/lib.rs:20-33 (every field public)
//   Particles::update(ts)   <- `pub fn update(&mut self, ts: f32) -> Vec<Particle>`  src/lib.rs:130
//
// Error behaviour: the reference panics (assert! at src/lib.rs:132, slice index at :225-228);
// here the same conditions throw particle_3d::Panic carrying the engine's message.
#pragma once
#include <cstdint>
#include <cstdlib>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "p3d.h"

namespace particle_3d {

struct Vector3 {  // stands in for cgmath::Vector3<f32>
    float x = 0.f, y = 0.f, z = 0.f;
};

struct Particle {  // src/lib.rs:12-17; layout-compatible with p3d_particle (28 bytes)
    Vector3 position;
    Vector3 velocity;
    uint32_t id = 0;
};
static_assert(sizeof(Particle) == sizeof(p3d_particle), "Particle must match the ABI struct");

struct Panic : std::runtime_error {
    int code;
    Panic(int c, const std::string &what) : std::runtime_error(what), code(c) {}
};

class Particles {
   public:
    // src/lib.rs:21-32, same names
    float world_size = 10.0f;
    std::vector<Particle> active_particles;
    std::vector<Particle> past_particles;
    uint32_t id_count = 0;
    std::vector<float> attraction_matrix;
    std::vector<Vector3> colors;  // render only
    float coefficient = 0.97f;
    float interaction_force = 1.0f;
    float min_pull_ratio = 0.3f;
    float particle_effect_radius = 2.0f;
    bool walls = false;
    Vector3 acceleration;
    // Not a field of the reference: true reproduces its bucket double-visit quirk (P3D_OPT_FAITHFUL; SURVEY.md
    // Appendix B.1).  The default evaluates every in-range pair exactly once, which DEVIATES from
    // src/lib.rs:195-206 for the ~26*k/N of the particles whose 27 hashed cells collide modulo N per step.
    // The environment variable P3D_FAITHFUL=1 flips the default.
    bool faithful = env_flag("P3D_FAITHFUL");

    // device >= 0: that CUDA device.  The environment variable P3D_DEVICES="0,1,2,3" overrides device 0 with a
    // list: one engine handle then drives all of them (p3d_create_multi) behind the unchanged update().
    explicit Particles(int device = 0) : device_(device) {}
    Particles(const Particles &) = delete;
    Particles &operator=(const Particles &) = delete;
    ~Particles() {
        if (engine_) p3d_destroy(engine_);
    }

    // The default scene of src/bin/main.rs:123-148 with a seeded generator (main.rs:60-87 is unseeded).
    static void default_scene(Particles &out, size_t n = 1000, uint64_t seed = 42) {
        p3d_params prm;
        float m[25];
        p3d_scene_default_params(&prm, m);
        out.world_size = prm.world_size;
        out.id_count = prm.id_count;
        out.attraction_matrix.assign(m, m + 25);
        out.colors = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {1, 1, 0}, {1, 0, 1}};  // main.rs:126-132
        out.coefficient = prm.coefficient;
        out.interaction_force = prm.interaction_force;
        out.min_pull_ratio = prm.min_pull_ratio;
        out.particle_effect_radius = prm.particle_effect_radius;
        out.walls = prm.walls != 0;
        out.acceleration = {prm.accel[0], prm.accel[1], prm.accel[2]};
        out.active_particles.resize(n);
        p3d_scene_uniform(seed, n, out.world_size, out.id_count,
                          reinterpret_cast<p3d_particle *>(out.active_particles.data()));
        out.past_particles.clear();
    }

    // src/lib.rs:130.  After the call past_particles holds the pre-step state (:167), active_particles
    // the post-step state in the same index order (:171-173,268); the return value is a copy (:271).
    std::vector<Particle> update(float ts) {
        if (attraction_matrix.size() < (size_t)id_count * id_count)
            throw Panic(P3D_ERR_BAD_ID, "attraction_matrix shorter than id_count^2 (src/lib.rs:225-228)");
        ensure_engine();
        check(p3d_set_option(engine_, P3D_OPT_FAITHFUL, faithful ? 1 : 0));
        const p3d_params prm = params();
        std::vector<Particle> next(active_particles.size());
        const int rc = p3d_update(engine_, &prm, ts, reinterpret_cast<const p3d_particle *>(active_particles.data()),
                                  reinterpret_cast<p3d_particle *>(next.data()), active_particles.size());
        if (rc != P3D_OK) throw Panic(rc, p3d_last_error());
        std::swap(active_particles, past_particles);  // :167
        active_particles = std::move(next);
        return active_particles;                      // :271 clone
    }

    // Device-resident multi-step run for headless use: n_steps x update(ts) with one upload and one
    // download (the per-step host round trip of update() is what main.rs:199 + :445 would pay).
    void run(float ts, int n_steps) {
        ensure_engine();
        check(p3d_set_option(engine_, P3D_OPT_FAITHFUL, faithful ? 1 : 0));
        const p3d_params prm = params();
        check(p3d_upload(engine_, reinterpret_cast<const p3d_particle *>(active_particles.data()),
                         active_particles.size(), id_count));
        check(p3d_step(engine_, &prm, ts, n_steps));
        past_particles = active_particles;
        check(p3d_download(engine_, reinterpret_cast<p3d_particle *>(active_particles.data()), active_particles.size()));
    }

    p3d_engine *engine() {
        ensure_engine();
        return engine_;
    }

    p3d_params params() const {
        p3d_params prm;
        prm.world_size = world_size;
        prm.coefficient = coefficient;
        prm.interaction_force = interaction_force;
        prm.min_pull_ratio = min_pull_ratio;
        prm.particle_effect_radius = particle_effect_radius;
        prm.accel[0] = acceleration.x;
        prm.accel[1] = acceleration.y;
        prm.accel[2] = acceleration.z;
        prm.walls = walls ? 1u : 0u;
        prm.id_count = id_count;
        prm.attraction_matrix = attraction_matrix.data();
        return prm;
    }

   private:
    static void check(int rc) {
        if (rc != P3D_OK) throw Panic(rc, p3d_last_error());
    }
    static bool env_flag(const char *name) {
        const char *v = std::getenv(name);
        return v && *v && !(v[0] == '0' && v[1] == 0);
    }
    void ensure_engine() {
        if (engine_) return;
        std::vector<int> devs;
        if (const char *env = (device_ == 0) ? std::getenv("P3D_DEVICES") : nullptr) {
            for (const char *p = env; *p;) {
                char *end = nullptr;
                const long d = std::strtol(p, &end, 10);
                if (end == p) break;
                devs.push_back((int)d);
                p = (*end == ',') ? end + 1 : end;
            }
        }
        if (devs.size() > 1) check(p3d_create_multi(devs.data(), (int)devs.size(), &engine_));
        else check(p3d_create(devs.empty() ? device_ : devs[0], &engine_));
    }
    int device_ = 0;
    p3d_engine *engine_ = nullptr;
};

}  // namespace particle_3d
