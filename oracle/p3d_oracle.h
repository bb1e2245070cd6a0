/* p3d_oracle.h — CPU ORACLE for the particle_3d hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This is a plain-C restatement of the reference's `Particles::update`
 * (/root/reference/src/lib.rs:130-272) and its private helpers (lib.rs:37-127).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load it.  The product (libp3d.so) never links or calls anything in oracle/.
 *
 * PARITY STATUS: "parity unpinned".  The reference ships no tests, golden vectors or
 * fixtures and no Rust toolchain exists in this image, so the
 * reference binary cannot be run.  The oracle is pinned instead by (i) SipHash
 * known-answer vectors from the SipHash paper, (ii) hand-derived known answers for
 * every helper, (iii) an independent brute-force f64 all-pairs
 * evaluation over the 27 periodic images which must agree with the cell-list walk,
 * (iv) a second restatement written separately in Python (oracle/lib_rs_twin.py) that must
 * agree with this file's faithful mode bit for bit (tests/test_oracle_twin.py).
 * (v) hash_cell against CPython's own SipHash-1-3 with a zero key (PYTHONHASHSEED=0), an implementation
 * written by neither of us (tests/test_oracle_kat.py).
 */
#ifndef P3D_ORACLE_H
#define P3D_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* lib.rs:12-17 — 28-byte particle (position, velocity, id). */
typedef struct {
    float px, py, pz;
    float vx, vy, vz;
    uint32_t id;
} ora_particle;

/* lib.rs:20-33 — the scalar fields of `Particles` that the step reads. */
typedef struct {
    float world_size;
    float coefficient;
    float interaction_force;
    float min_pull_ratio;
    float particle_effect_radius;
    float accel[3];
    uint32_t walls;
    uint32_t id_count;
    const float *attraction_matrix; /* id_count*id_count, row = self id, col = other id */
} ora_params;

enum { ORA_FAITHFUL = 0, ORA_IDEAL = 1 };

typedef struct {
    uint64_t candidates;      /* distance tests executed (lib.rs:211-213) */
    uint64_t in_radius;       /* passed 0 < d2 < r^2 (lib.rs:216-220), incl. double visits */
    uint64_t nonzero;         /* in_radius with f != 0 */
    uint64_t dup_bucket_queries; /* (particle,image) queries whose 27 cells hit a bucket twice */
    uint64_t affected;        /* particles with a double-counted non-zero force */
    double t_build_s;         /* wall time of the counting sort (lib.rs:135-164) */
    double t_force_s;         /* wall time of the per-particle pass (lib.rs:171-268) */
} ora_stats;

/* Generic SipHash-c-d, 64-bit output (used with c=1,d=3,k=0/0 by hash_cell; c=2,d=4 for KATs). */
uint64_t ora_siphash(int c_rounds, int d_rounds, uint64_t k0, uint64_t k1,
                     const uint8_t *msg, size_t len);
/* lib.rs:46-52 */
uint64_t ora_hash_cell(int64_t x, int64_t y, int64_t z);
/* lib.rs:37-43 (Rust `as isize`: truncate toward zero, saturate, NaN -> 0) */
void ora_cell_coord(float radius, const float v[3], int64_t out[3]);
/* lib.rs:55-67 */
float ora_calculate_force(float min_pull_ratio, float distance, float attraction);
/* lib.rs:70-127 */
void ora_handle_wall_collision(float world_size, uint32_t walls, ora_particle *p);

/* lib.rs:130-272.  Returns 0, or 1 if world_size < 2*radius (the reference panics, lib.rs:132),
 * or 2 if some id >= id_count (the reference would index out of bounds, lib.rs:225-228).
 * mode: ORA_FAITHFUL scans all 27 hashed cells even when two collide modulo N (the
 *       reference's behaviour); ORA_IDEAL visits each distinct bucket once per image query.
 * acc64: accumulate the force sum in double (error budgeting only; 0 = reference arithmetic).
 * force_out (nullable): n*3 floats, total_force of lib.rs:177-243 before integration.
 * affected (nullable): n bytes, 1 where FAITHFUL double-counted a non-zero force.
 * nthreads <= 0 means all cores. */
int ora_update(const ora_params *prm, float ts, const ora_particle *in, ora_particle *out,
               size_t n, int mode, int acc64, float *force_out, uint8_t *affected,
               ora_stats *stats, int nthreads);

/* Advance only particles [i_begin, i_end) (out holds i_end - i_begin entries); the hash table still
 * covers all n.  For timing a bounded sample of a large step. */
int ora_update_sample(const ora_params *prm, float ts, const ora_particle *in, ora_particle *out,
                      size_t n, size_t i_begin, size_t i_end, int mode, ora_stats *stats, int nthreads);

/* Advance only the particles idx[0..n_idx) (out[k] = updated particle idx[k]); the hash table covers all n.
 * Returns 4 for an index >= n. */
int ora_update_indices(const ora_params *prm, float ts, const ora_particle *in, ora_particle *out, size_t n,
                       const size_t *idx, size_t n_idx, int mode, ora_stats *stats, int nthreads);

/* Independent check: O(N^2 * 27) brute force over all particles and all 27 image offsets in
 * double precision, same cutoff and force law, each (j,image) counted once.  force_out n*3 doubles. */
int ora_bruteforce_forces(const ora_params *prm, const ora_particle *in, size_t n,
                          double *force_out, int nthreads);

/* Integration only (lib.rs:245-264) given a force array (n*3 floats). */
void ora_integrate(const ora_params *prm, float ts, const ora_particle *in, const float *force,
                   ora_particle *out, size_t n);

int ora_num_threads(void);

/* Seeded scenes for the bench's reference arm (same streams as the product's p3d_scene_*; byte equality is a
 * CPU test): default constants of src/bin/main.rs:123-148, uniform cloud (main.rs:60-87), Plummer cloud. */
void ora_scene_default_params(ora_params *prm, float matrix25[25]);
void ora_scene_uniform(uint64_t seed, size_t n, float world_size, uint32_t id_count, ora_particle *out);
void ora_scene_plummer(uint64_t seed, size_t n, float world_size, float scale_a, uint32_t id_count,
                       ora_particle *out);

#ifdef __cplusplus
}
#endif
#endif
