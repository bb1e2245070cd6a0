/* p3d_oracle.c — CPU ORACLE (test infrastructure only; see p3d_oracle.h).
 *
 * Restates /root/reference/src/lib.rs in plain C, operation by operation, f32 with no
 * fused multiply-add (build with -ffp-contract=off).  Each function cites the lines it follows.
 * Third-party arithmetic that is not vendored in the reference is restated from its published
 * definition:
 *   - std::collections::hash_map::DefaultHasher = SipHash-1-3, keys (0,0); `isize::hash`
 *     feeds 8 native-endian bytes (Rust std; toolchain unpinned, edition 2024).
 *   - cgmath 0.18.0 (Cargo.lock) Vector3: component-wise + - *scalar /scalar,
 *     magnitude2 = x*x + y*y + z*z summed left to right.
 *   - rayon 1.10.0 only decides the grouping of the f32 sums and the order inside a bucket;
 *     the canonical order used here is what a sequential execution produces.
 */
#include "p3d_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------ SipHash */
static inline uint64_t rotl64(uint64_t x, int b) { return (x << b) | (x >> (64 - b)); }

#define SIPROUND(v0, v1, v2, v3) \
    do {                         \
        v0 += v1;                \
        v1 = rotl64(v1, 13);     \
        v1 ^= v0;                \
        v0 = rotl64(v0, 32);     \
        v2 += v3;                \
        v3 = rotl64(v3, 16);     \
        v3 ^= v2;                \
        v0 += v3;                \
        v3 = rotl64(v3, 21);     \
        v3 ^= v0;                \
        v2 += v1;                \
        v1 = rotl64(v1, 17);     \
        v1 ^= v2;                \
        v2 = rotl64(v2, 32);     \
    } while (0)

uint64_t ora_siphash(int c_rounds, int d_rounds, uint64_t k0, uint64_t k1, const uint8_t *msg,
                     size_t len) {
    uint64_t v0 = k0 ^ 0x736f6d6570736575ULL;
    uint64_t v1 = k1 ^ 0x646f72616e646f6dULL;
    uint64_t v2 = k0 ^ 0x6c7967656e657261ULL;
    uint64_t v3 = k1 ^ 0x7465646279746573ULL;
    size_t nwords = len / 8;
    for (size_t w = 0; w < nwords; ++w) {
        uint64_t m = 0;
        for (int b = 0; b < 8; ++b) m |= (uint64_t)msg[8 * w + b] << (8 * b);
        v3 ^= m;
        for (int r = 0; r < c_rounds; ++r) SIPROUND(v0, v1, v2, v3);
        v0 ^= m;
    }
    uint64_t last = (uint64_t)(len & 0xff) << 56;
    for (size_t b = 0; b < (len & 7); ++b) last |= (uint64_t)msg[8 * nwords + b] << (8 * b);
    v3 ^= last;
    for (int r = 0; r < c_rounds; ++r) SIPROUND(v0, v1, v2, v3);
    v0 ^= last;
    v2 ^= 0xff;
    for (int r = 0; r < d_rounds; ++r) SIPROUND(v0, v1, v2, v3);
    return v0 ^ v1 ^ v2 ^ v3;
}

/* lib.rs:46-52: DefaultHasher::new(); x.hash(); y.hash(); z.hash(); finish().
 * Three 8-byte little-endian words, 24-byte message, SipHash-1-3 with zero keys.
 * Specialised (no byte loop) because the step evaluates it 729 times per particle. */
uint64_t ora_hash_cell(int64_t x, int64_t y, int64_t z) {
    uint64_t v0 = 0x736f6d6570736575ULL, v1 = 0x646f72616e646f6dULL;
    uint64_t v2 = 0x6c7967656e657261ULL, v3 = 0x7465646279746573ULL;
    uint64_t m;
    m = (uint64_t)x; v3 ^= m; SIPROUND(v0, v1, v2, v3); v0 ^= m;
    m = (uint64_t)y; v3 ^= m; SIPROUND(v0, v1, v2, v3); v0 ^= m;
    m = (uint64_t)z; v3 ^= m; SIPROUND(v0, v1, v2, v3); v0 ^= m;
    m = (uint64_t)24 << 56; v3 ^= m; SIPROUND(v0, v1, v2, v3); v0 ^= m;
    v2 ^= 0xff;
    SIPROUND(v0, v1, v2, v3);
    SIPROUND(v0, v1, v2, v3);
    SIPROUND(v0, v1, v2, v3);
    return v0 ^ v1 ^ v2 ^ v3;
}

/* Rust `f32 as isize`: truncate toward zero, saturate at the i64 range, NaN -> 0. */
static inline int64_t f32_as_isize(float f) {
    if (f != f) return 0;
    if (f >= 9223372036854775808.0f) return INT64_MAX;
    if (f <= -9223372036854775808.0f) return INT64_MIN;
    return (int64_t)f; /* C conversion truncates toward zero */
}

/* lib.rs:37-43 */
void ora_cell_coord(float radius, const float v[3], int64_t out[3]) {
    out[0] = f32_as_isize(v[0] / radius);
    out[1] = f32_as_isize(v[1] / radius);
    out[2] = f32_as_isize(v[2] / radius);
}

/* lib.rs:55-67 */
float ora_calculate_force(float m, float distance, float attraction) {
    if (distance < m) {
        return distance / m - 1.0f;
    } else if (m < distance && distance < 1.0f) {
        float t = 2.0f * distance;
        t = t - 1.0f;
        t = t - m;
        return attraction * (1.0f - fabsf(t) / (1.0f - m));
    } else {
        return 0.0f;
    }
}

/* Rust f32::min / f32::max ignore a NaN operand. */
static inline float rs_min(float a, float b) { return (a != a) ? b : (b != b) ? a : (a < b ? a : b); }
static inline float rs_max(float a, float b) { return (a != a) ? b : (b != b) ? a : (a > b ? a : b); }

static inline void wall_axis(float half, float world, uint32_t walls, float *pos, float *vel) {
    if (*pos > half) {
        if (walls) {
            *pos = half;
            *vel = rs_min(*vel, 0.0f);
        } else {
            *pos -= world;
        }
    } else if (*pos < -half) {
        if (walls) {
            *pos = -half;
            *vel = rs_max(*vel, 0.0f);
        } else {
            *pos += world;
        }
    }
}

/* lib.rs:70-127 */
void ora_handle_wall_collision(float world_size, uint32_t walls, ora_particle *p) {
    float half = world_size * 0.5f;
    wall_axis(half, world_size, walls, &p->px, &p->vx);
    wall_axis(half, world_size, walls, &p->py, &p->vy);
    wall_axis(half, world_size, walls, &p->pz, &p->vz);
}

/* lib.rs:245-264 for one particle, given total_force. */
static inline void integrate_one(const ora_params *prm, float ts, const float F[3],
                                 ora_particle *p) {
    const float k = prm->interaction_force, r = prm->particle_effect_radius;
    /* :246-247  velocity += ((F * k) * r) * ts */
    p->vx = p->vx + ((F[0] * k) * r) * ts;
    p->vy = p->vy + ((F[1] * k) * r) * ts;
    p->vz = p->vz + ((F[2] * k) * r) * ts;
    /* :249 */
    p->vx = p->vx + prm->accel[0] * ts;
    p->vy = p->vy + prm->accel[1] * ts;
    p->vz = p->vz + prm->accel[2] * ts;
    /* :252-259 */
    float cx = (p->vx * prm->coefficient) * ts;
    float cy = (p->vy * prm->coefficient) * ts;
    float cz = (p->vz * prm->coefficient) * ts;
    float c2 = cx * cx + cy * cy + cz * cz;
    float v2 = p->vx * p->vx + p->vy * p->vy + p->vz * p->vz;
    if (c2 > v2) {
        p->vx = 0.0f; p->vy = 0.0f; p->vz = 0.0f;
    } else {
        p->vx = p->vx - cx; p->vy = p->vy - cy; p->vz = p->vz - cz;
    }
    /* :262 */
    p->px = p->px + p->vx * ts;
    p->py = p->py + p->vy * ts;
    p->pz = p->pz + p->vz * ts;
    /* :264 */
    ora_handle_wall_collision(prm->world_size, prm->walls, p);
}

void ora_integrate(const ora_params *prm, float ts, const ora_particle *in, const float *force,
                   ora_particle *out, size_t n) {
    for (size_t i = 0; i < n; ++i) {
        ora_particle p = in[i];
        integrate_one(prm, ts, force + 3 * i, &p);
        out[i] = p;
    }
}

int ora_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* lib.rs:130-272 */
static int ora_update_impl(const ora_params *prm, float ts, const ora_particle *in, ora_particle *out,
                           size_t n, size_t i_begin, size_t i_end, const size_t *idx, int mode, int acc64,
                           float *force_out, uint8_t *affected, ora_stats *stats, int nthreads);

int ora_update(const ora_params *prm, float ts, const ora_particle *in, ora_particle *out,
               size_t n, int mode, int acc64, float *force_out, uint8_t *affected,
               ora_stats *stats, int nthreads) {
    return ora_update_impl(prm, ts, in, out, n, 0, n, NULL, mode, acc64, force_out, affected, stats, nthreads);
}

/* Same step, but only particles [i_begin, i_end) are advanced (out, force_out, affected hold
 * i_end - i_begin entries).  The hash table is still built over all n particles.  Used by the
 * bench to time a bounded sample of a large step. */
int ora_update_sample(const ora_params *prm, float ts, const ora_particle *in, ora_particle *out,
                      size_t n, size_t i_begin, size_t i_end, int mode, ora_stats *stats, int nthreads) {
    if (i_end > n) i_end = n;
    if (i_begin > i_end) i_begin = i_end;
    return ora_update_impl(prm, ts, in, out, n, i_begin, i_end, NULL, mode, 0, NULL, NULL, stats, nthreads);
}

/* Same step, but only the particles idx[0..n_idx) are advanced (out[k] = updated particle idx[k]); the hash
 * table covers all n.  Lets a parity check at N = 1M sample particles from every type segment / every rank's
 * slot range instead of one contiguous index range. */
int ora_update_indices(const ora_params *prm, float ts, const ora_particle *in, ora_particle *out, size_t n,
                       const size_t *idx, size_t n_idx, int mode, ora_stats *stats, int nthreads) {
    if (n_idx && !idx) return 4;
    for (size_t k = 0; k < n_idx; ++k)
        if (idx[k] >= n) return 4;
    return ora_update_impl(prm, ts, in, out, n, 0, n_idx, idx, mode, 0, NULL, NULL, stats, nthreads);
}

static double now_s(void) {
#ifdef _OPENMP
    return omp_get_wtime();
#else
    return 0.0;
#endif
}

static int ora_update_impl(const ora_params *prm, float ts, const ora_particle *in, ora_particle *out,
                           size_t n, size_t i_begin, size_t i_end, const size_t *idx, int mode, int acc64,
                           float *force_out, uint8_t *affected, ora_stats *stats, int nthreads) {
    const double t_start = now_s();
    const float W = prm->world_size, r = prm->particle_effect_radius, m = prm->min_pull_ratio;
    const uint32_t T = prm->id_count;
    if (stats) memset(stats, 0, sizeof(*stats));
    /* :132 assert!(world_size >= 2.0 * particle_effect_radius) */
    if (!(W >= 2.0f * r)) return 1;
    if (n == 0) return 0;
    for (size_t i = 0; i < n; ++i)
        if (in[i].id >= T) return 2; /* :225-228 would index past the matrix (or a wrong row) */
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
    nthreads = 1;
#endif

    /* :135-164 counting sort of particle indices by hash_cell(cell_coord(pos)) % N */
    size_t *table = (size_t *)calloc(n + 1, sizeof(size_t));
    size_t *indices = (size_t *)malloc(n * sizeof(size_t));
    size_t *bucket_of = (size_t *)malloc(n * sizeof(size_t));
    if (!table || !indices || !bucket_of) { free(table); free(indices); free(bucket_of); return 3; }
#pragma omp parallel for schedule(static) num_threads(nthreads)
    for (size_t i = 0; i < n; ++i) {
        float v[3] = {in[i].px, in[i].py, in[i].pz};
        int64_t c[3];
        ora_cell_coord(r, v, c);
        bucket_of[i] = (size_t)(ora_hash_cell(c[0], c[1], c[2]) % (uint64_t)n); /* :142 */
    }
    for (size_t i = 0; i < n; ++i) table[bucket_of[i]] += 1;          /* :141-144 */
    for (size_t i = 1; i < n + 1; ++i) table[i] += table[i - 1];      /* :147-149 */
    for (size_t i = 0; i < n; ++i) {                                  /* :157-164, sequential order */
        size_t slot = table[bucket_of[i]]--;
        indices[slot - 1] = i;
    }
    free(bucket_of);
    /* now table[b] = start of bucket b, table[b+1] = its end */
    const double t_built = now_s();

    const float r2 = r * r; /* :218-219 */
    uint64_t s_cand = 0, s_in = 0, s_nz = 0, s_dupq = 0, s_aff = 0;

#pragma omp parallel for schedule(dynamic, 64) num_threads(nthreads) \
    reduction(+ : s_cand, s_in, s_nz, s_dupq, s_aff)
    for (size_t k = i_begin; k < i_end; ++k) {
        const size_t i = idx ? idx[k] : k;
        const ora_particle p = in[i]; /* past_particles[i], :171-174 */
        float ax = 0.0f, ay = 0.0f, az = 0.0f;
        double dax = 0.0, day = 0.0, daz = 0.0;
        int hit_dup = 0;
        for (int ox = -1; ox <= 1; ++ox)
            for (int oy = -1; oy <= 1; ++oy)
                for (int oz = -1; oz <= 1; ++oz) { /* :177-185, canonical order x outermost */
                    /* :190-192 */
                    const float offx = (float)ox * W, offy = (float)oy * W, offz = (float)oz * W;
                    const float q0[3] = {p.px + offx, p.py + offy, p.pz + offz};
                    int64_t c0[3];
                    ora_cell_coord(r, q0, c0);
                    size_t seen[27];
                    int nseen = 0, query_has_dup = 0;
                    for (int cx = -1; cx <= 1; ++cx)
                        for (int cy = -1; cy <= 1; ++cy)
                            for (int cz = -1; cz <= 1; ++cz) { /* :195-199 */
                                /* isize add; wraps like release-mode Rust */
                                int64_t ccx = (int64_t)((uint64_t)c0[0] + (uint64_t)(int64_t)cx);
                                int64_t ccy = (int64_t)((uint64_t)c0[1] + (uint64_t)(int64_t)cy);
                                int64_t ccz = (int64_t)((uint64_t)c0[2] + (uint64_t)(int64_t)cz);
                                size_t b = (size_t)(ora_hash_cell(ccx, ccy, ccz) % (uint64_t)n); /* :202 */
                                int dup = 0;
                                for (int s = 0; s < nseen; ++s)
                                    if (seen[s] == b) { dup = 1; break; }
                                seen[nseen++] = b;
                                if (dup) {
                                    query_has_dup = 1;
                                    if (mode == ORA_IDEAL) continue;
                                }
                                for (size_t s = table[b]; s < table[b + 1]; ++s) { /* :203-206 */
                                    const ora_particle *q = &in[indices[s]];      /* :207-208 */
                                    /* :211-212 other.position - (position + offset) */
                                    const float rx = q->px - (p.px + offx);
                                    const float ry = q->py - (p.py + offy);
                                    const float rz = q->pz - (p.pz + offz);
                                    const float d2 = rx * rx + ry * ry + rz * rz; /* :213 */
                                    ++s_cand;
                                    if (d2 > 0.0f && d2 < r2) { /* :216-220 */
                                        ++s_in;
                                        const float d = sqrtf(d2); /* :221 */
                                        const float a =
                                            prm->attraction_matrix[p.id * T + q->id]; /* :225-228 */
                                        const float f = ora_calculate_force(m, d, a);
                                        if (f != 0.0f) {
                                            ++s_nz;
                                            if (dup) hit_dup = 1;
                                        }
                                        /* :231 acc += rel / d * f */
                                        const float fx = rx / d * f, fy = ry / d * f, fz = rz / d * f;
                                        if (acc64) {
                                            dax += (double)fx; day += (double)fy; daz += (double)fz;
                                        } else {
                                            ax = ax + fx; ay = ay + fy; az = az + fz;
                                        }
                                    }
                                }
                            }
                    s_dupq += (uint64_t)query_has_dup;
                }
        if (acc64) { ax = (float)dax; ay = (float)day; az = (float)daz; }
        const float F[3] = {ax, ay, az};
        const size_t o = k - i_begin;
        if (force_out) { force_out[3 * o] = ax; force_out[3 * o + 1] = ay; force_out[3 * o + 2] = az; }
        if (affected) affected[o] = (uint8_t)hit_dup;
        s_aff += (uint64_t)hit_dup;
        ora_particle u = p;
        integrate_one(prm, ts, F, &u); /* :245-264 */
        out[o] = u;                     /* :266-268, index order preserved */
    }
    if (stats) {
        stats->candidates = s_cand; stats->in_radius = s_in; stats->nonzero = s_nz;
        stats->dup_bucket_queries = s_dupq; stats->affected = s_aff;
        stats->t_build_s = t_built - t_start;
        stats->t_force_s = now_s() - t_built;
    }
    free(table);
    free(indices);
    return 0;
}

/* Independent brute force in double precision: every particle j, every one of the 27 image
 * offsets of lib.rs:177-185, counted once.  The image position is still rounded in f32 the way
 * the reference rounds it (p + k*W, lib.rs:211-212) so that cutoff decisions agree. */
int ora_bruteforce_forces(const ora_params *prm, const ora_particle *in, size_t n,
                          double *force_out, int nthreads) {
    const float W = prm->world_size, r = prm->particle_effect_radius;
    const double m = (double)prm->min_pull_ratio;
    const uint32_t T = prm->id_count;
    if (!(W >= 2.0f * r)) return 1;
    for (size_t i = 0; i < n; ++i)
        if (in[i].id >= T) return 2;
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
    nthreads = 1;
#endif
    const float r2f = r * r;
#pragma omp parallel for schedule(dynamic, 16) num_threads(nthreads)
    for (size_t i = 0; i < n; ++i) {
        const ora_particle p = in[i];
        double ax = 0, ay = 0, az = 0;
        for (int ox = -1; ox <= 1; ++ox)
            for (int oy = -1; oy <= 1; ++oy)
                for (int oz = -1; oz <= 1; ++oz) {
                    const float qx = p.px + (float)ox * W, qy = p.py + (float)oy * W,
                                qz = p.pz + (float)oz * W;
                    for (size_t j = 0; j < n; ++j) {
                        const float rxf = in[j].px - qx, ryf = in[j].py - qy, rzf = in[j].pz - qz;
                        const float d2f = rxf * rxf + ryf * ryf + rzf * rzf;
                        if (!(d2f > 0.0f && d2f < r2f)) continue; /* same decision as f32 */
                        const double rx = rxf, ry = ryf, rz = rzf;
                        const double d = sqrt(rx * rx + ry * ry + rz * rz);
                        const double a = (double)prm->attraction_matrix[p.id * T + in[j].id];
                        double f;
                        if (d < m) f = d / m - 1.0;
                        else if (m < d && d < 1.0) f = a * (1.0 - fabs(2.0 * d - 1.0 - m) / (1.0 - m));
                        else f = 0.0;
                        ax += rx / d * f; ay += ry / d * f; az += rz / d * f;
                    }
                }
        force_out[3 * i] = ax; force_out[3 * i + 1] = ay; force_out[3 * i + 2] = az;
    }
    return 0;
}


/* ---- seeded scenes for the bench's reference arm -------------------------------------------------
 * The reference's generator (src/bin/main.rs:60-87) is unseeded, so the bench uses its own
 * counter-based stream (splitmix64, seed 42).  The product has the same generator
 * (p3d_scene_uniform / p3d_scene_plummer); it is restated here so that the reference arm builds its
 * inputs without loading the product library.  tests/test_oracle_scene.py requires byte equality. */
static uint64_t sm64_next(uint64_t *state) {
    uint64_t z = (*state += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
static float sm64_unit_f32(uint64_t *state) { return (float)(sm64_next(state) >> 40) * (1.0f / 16777216.0f); }
static double sm64_unit_f64(uint64_t *state) {
    return (double)(sm64_next(state) >> 11) * (1.0 / 9007199254740992.0);
}

/* default scene constants, src/bin/main.rs:123-148; matrix25 receives the 5x5 matrix of :133-139 */
void ora_scene_default_params(ora_params *prm, float matrix25[25]) {
    static const float rows[5][5] = {{0.5f, 1.0f, -0.5f, 0.0f, -1.0f},
                                     {1.0f, 1.0f, 1.0f, 0.0f, -1.0f},
                                     {0.0f, 0.0f, 0.5f, 1.5f, -1.0f},
                                     {0.0f, 0.0f, 0.0f, 0.0f, -1.0f},
                                     {1.0f, 1.0f, 1.0f, 1.0f, 0.5f}};
    for (int i = 0; i < 5; ++i)
        for (int j = 0; j < 5; ++j) matrix25[5 * i + j] = rows[i][j];
    prm->world_size = 10.0f;
    prm->coefficient = 0.97f;
    prm->interaction_force = 1.0f;
    prm->min_pull_ratio = 0.3f;
    prm->particle_effect_radius = 2.0f;
    prm->accel[0] = prm->accel[1] = prm->accel[2] = 0.0f;
    prm->walls = 0;
    prm->id_count = 5;
    prm->attraction_matrix = matrix25;
}

/* uniform box [-W/2, W/2)^3, v = 0, id uniform in 0..id_count (distribution of main.rs:60-87) */
void ora_scene_uniform(uint64_t seed, size_t n, float world_size, uint32_t id_count, ora_particle *out) {
    uint64_t st = seed;
    const float half = world_size * 0.5f;
    for (size_t i = 0; i < n; ++i) {
        ora_particle *p = &out[i];
        p->px = -half + world_size * sm64_unit_f32(&st);
        p->py = -half + world_size * sm64_unit_f32(&st);
        p->pz = -half + world_size * sm64_unit_f32(&st);
        p->vx = 0.0f; p->vy = 0.0f; p->vz = 0.0f;
        p->id = id_count ? (uint32_t)(sm64_next(&st) % id_count) : 0u;
    }
}

/* Plummer-like cloud (BASELINE.json config 3): radius from the inverted Plummer mass profile with scale a,
 * isotropic direction, samples outside the box rejected. */
void ora_scene_plummer(uint64_t seed, size_t n, float world_size, float scale_a, uint32_t id_count,
                       ora_particle *out) {
    uint64_t st = seed;
    const double half = 0.5 * (double)world_size;
    for (size_t i = 0; i < n; ++i) {
        double x, y, z;
        do {
            double u = sm64_unit_f64(&st);
            if (u < 1e-12) u = 1e-12;
            const double rad = (double)scale_a / sqrt(pow(u, -2.0 / 3.0) - 1.0);
            const double cz = 2.0 * sm64_unit_f64(&st) - 1.0;
            const double phi = 6.283185307179586476925286766559 * sm64_unit_f64(&st);
            const double sz = sqrt(1.0 - cz * cz);
            x = rad * sz * cos(phi);
            y = rad * sz * sin(phi);
            z = rad * cz;
        } while (!(fabs(x) < half && fabs(y) < half && fabs(z) < half));
        ora_particle *p = &out[i];
        p->px = (float)x; p->py = (float)y; p->pz = (float)z;
        p->vx = 0.0f; p->vy = 0.0f; p->vz = 0.0f;
        p->id = id_count ? (uint32_t)(sm64_next(&st) % id_count) : 0u;
    }
}
