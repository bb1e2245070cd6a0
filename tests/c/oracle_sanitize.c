/* Drives every entry point of the CPU oracle under AddressSanitizer + UndefinedBehaviorSanitizer
 * ("Sanitizers": host ASan/UBSan on the oracle).  Built and run by tests/test_oracle_sanitize.py
 * with -fsanitize=address,undefined -fno-sanitize-recover=all: any out-of-bounds access, signed overflow,
 * invalid float->int conversion or misaligned access aborts the process.  Inputs include the awkward ones:
 * n = 0 and 1, coincident particles, positions far outside the box, NaN / infinite coordinates, huge velocities,
 * a 7-bucket hash table, r < 1 and m > 1. */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "p3d_oracle.h"

static uint64_t lcg_state = 0x9E3779B97F4A7C15ull;
static float unit(void) {
    lcg_state = lcg_state * 6364136223846793005ull + 1442695040888963407ull;
    return (float)(lcg_state >> 40) * (1.0f / 16777216.0f);
}

static void cloud(ora_particle *p, size_t n, float W, uint32_t T, float speed) {
    for (size_t i = 0; i < n; ++i) {
        p[i].px = (unit() - 0.5f) * W; p[i].py = (unit() - 0.5f) * W; p[i].pz = (unit() - 0.5f) * W;
        p[i].vx = (unit() - 0.5f) * speed; p[i].vy = (unit() - 0.5f) * speed; p[i].vz = (unit() - 0.5f) * speed;
        p[i].id = (uint32_t)(unit() * (float)T) % T;
    }
}

static int run(const ora_params *prm, ora_particle *in, size_t n, int steps) {
    ora_particle *out = malloc((n ? n : 1) * sizeof(*out));
    float *force = malloc((n ? n : 1) * 3 * sizeof(float));
    double *brute = malloc((n ? n : 1) * 3 * sizeof(double));
    uint8_t *aff = malloc(n ? n : 1);
    ora_stats st;
    int rc = 0;
    for (int s = 0; s < steps && !rc; ++s) {
        for (int mode = 0; mode < 2 && !rc; ++mode)
            for (int acc64 = 0; acc64 < 2 && !rc; ++acc64)
                rc = ora_update(prm, 1.0f / 60.0f, in, out, n, mode, acc64, force, aff, &st, 2);
        if (!rc) rc = ora_bruteforce_forces(prm, in, n, brute, 1);
        if (!rc && n > 2) rc = ora_update_sample(prm, 1.0f / 60.0f, in, out, n, 1, n - 1, ORA_FAITHFUL, &st, 1);
        if (!rc) {
            rc = ora_update(prm, 1.0f / 60.0f, in, out, n, ORA_FAITHFUL, 0, force, NULL, NULL, 1);
            if (!rc && n) {
                ora_integrate(prm, 1.0f / 60.0f, in, force, out, n);
                memcpy(in, out, n * sizeof(*in));
            }
        }
    }
    free(out); free(force); free(brute); free(aff);
    return rc;
}

int main(void) {
    static const float A5[25] = {0.5f, 1.0f, -0.5f, 0.0f, -1.0f, 1.0f, 1.0f, 1.0f, 0.0f, -1.0f, 0.0f, 0.0f, 0.5f,
                                 1.5f, -1.0f, 0.0f, 0.0f, 0.0f, 0.0f, -1.0f, 1.0f, 1.0f, 1.0f, 1.0f, 0.5f};
    ora_params prm = {10.0f, 0.97f, 1.0f, 0.3f, 2.0f, {0.0f, 0.0f, 0.0f}, 0, 5, A5};
    enum { NMAX = 400 };
    ora_particle *p = malloc(NMAX * sizeof(*p));
    int fails = 0;

    /* helpers on hostile scalars */
    const float hostile[] = {0.0f, -0.0f, 1e30f, -1e30f, INFINITY, -INFINITY, NAN, 1.9999999f, -1.9999999f, 9.3e18f};
    for (size_t k = 0; k < sizeof(hostile) / sizeof(hostile[0]); ++k) {
        float v[3] = {hostile[k], -hostile[k], 0.5f * hostile[k]};
        int64_t c[3];
        ora_cell_coord(2.0f, v, c);
        (void)ora_hash_cell(c[0], c[1], c[2]);
        (void)ora_calculate_force(0.3f, hostile[k], 1.0f);
        (void)ora_calculate_force(hostile[k], 0.5f, -1.0f);
        ora_particle q = {hostile[k], 0.0f, -hostile[k], hostile[k], 1.0f, -1.0f, 0};
        ora_handle_wall_collision(10.0f, (uint32_t)(k & 1), &q);
    }
    uint8_t msg[15];
    for (int k = 0; k < 15; ++k) msg[k] = (uint8_t)k;
    for (size_t len = 0; len <= 15; ++len) (void)ora_siphash(2, 4, 0x0706050403020100ull, 0x0f0e0d0c0b0a0908ull, msg, len);

    /* default scene density, two steps */
    cloud(p, 300, 10.0f, 5, 0.0f);
    fails += run(&prm, p, 300, 2) != 0;
    /* empty, single, tiny hash tables */
    fails += run(&prm, p, 0, 1) != 0;
    cloud(p, 1, 10.0f, 5, 1.0f);
    fails += run(&prm, p, 1, 1) != 0;
    cloud(p, 7, 4.0f, 5, 1.0f);
    prm.world_size = 4.0f;
    fails += run(&prm, p, 7, 2) != 0;
    /* walls + gravity + fast particles + r < 1 */
    prm.world_size = 6.0f; prm.walls = 1; prm.accel[1] = -9.8f; prm.particle_effect_radius = 0.7f;
    cloud(p, 250, 6.0f, 5, 900.0f);
    fails += run(&prm, p, 250, 2) != 0;
    /* m > 1, strong drag, coincident and far-outside particles, a NaN and an infinite coordinate */
    prm.walls = 0; prm.accel[1] = 0.0f; prm.particle_effect_radius = 2.5f; prm.min_pull_ratio = 1.5f; prm.coefficient = 70.0f;
    cloud(p, 200, 6.0f, 5, 5.0f);
    p[1] = p[0];
    p[2].px += 6.0f; p[3].py -= 60.0f; p[4].pz = 3.0e9f; p[5].px = NAN; p[6].py = INFINITY; p[7].vz = 1.0e30f;
    fails += run(&prm, p, 200, 2) != 0;
    /* the two error returns */
    prm.world_size = 4.9f;
    fails += ora_update(&prm, 1.0f / 60.0f, p, p, 0, 0, 0, NULL, NULL, NULL, 1) != 1;
    prm.world_size = 6.0f;
    p[9].id = 5;
    {
        ora_particle *out = malloc(200 * sizeof(*out));
        fails += ora_update(&prm, 1.0f / 60.0f, p, out, 200, 0, 0, NULL, NULL, NULL, 1) != 2;
        free(out);
    }
    free(p);
    printf("oracle_sanitize: %s (%d unexpected return codes)\n", fails ? "FAILED" : "clean", fails);
    return fails ? 1 : 0;
}
