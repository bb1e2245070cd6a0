/* Plain-C consumer of the multi-device handle: p3d_create_multi over the devices named in P3D_DEVICES (default
 * "0,0": two members sharing device 0), one p3d_update of a seeded cloud with the all-pairs kernel and with the
 * default (cell-list) kernel, checked against the CPU oracle (oracle/p3d_oracle.h — tests may link it, the product
 * never does) with the parity metric of tests/helpers.py, then a device-resident run against the single-device
 * engine.  Exit code 0 = all good. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "p3d.h"
#include "p3d_oracle.h"

#define CHECK(cond)                                                          \
    do {                                                                     \
        if (!(cond)) {                                                       \
            fprintf(stderr, "FAILED %s:%d: %s (last error: %s)\n", __FILE__, __LINE__, #cond, p3d_last_error()); \
            return 1;                                                        \
        }                                                                    \
    } while (0)

static int parity_ok(const p3d_particle *got, const ora_particle *ref, size_t n, float W, double tol, double *worst) {
    double v2 = 0.0;
    for (size_t i = 0; i < n; ++i) v2 += (double)ref[i].vx * ref[i].vx + (double)ref[i].vy * ref[i].vy + (double)ref[i].vz * ref[i].vz;
    const double vrms = sqrt(v2 / (double)(n ? n : 1));
    *worst = 0.0;
    for (size_t i = 0; i < n; ++i) {
        if (got[i].id != ref[i].id) return 0;
        const double dvx = got[i].vx - ref[i].vx, dvy = got[i].vy - ref[i].vy, dvz = got[i].vz - ref[i].vz;
        const double dpx = got[i].px - ref[i].px, dpy = got[i].py - ref[i].py, dpz = got[i].pz - ref[i].pz;
        const double vn = sqrt((double)ref[i].vx * ref[i].vx + (double)ref[i].vy * ref[i].vy + (double)ref[i].vz * ref[i].vz);
        const double pn = sqrt((double)ref[i].px * ref[i].px + (double)ref[i].py * ref[i].py + (double)ref[i].pz * ref[i].pz);
        const double ev = sqrt(dvx * dvx + dvy * dvy + dvz * dvz) / fmax(vn, fmax(vrms, 1e-30));
        const double ep = sqrt(dpx * dpx + dpy * dpy + dpz * dpz) / fmax(pn, 0.5 * W);
        if (ev > *worst) *worst = ev;
        if (ep > *worst) *worst = ep;
    }
    return *worst <= tol;
}

int main(void) {
    int devices[8], n_dev = 0;
    const char *env = getenv("P3D_DEVICES");
    char buf[128];
    snprintf(buf, sizeof(buf), "%s", (env && *env) ? env : "0,0");
    for (char *tok = strtok(buf, ","); tok && n_dev < 8; tok = strtok(NULL, ",")) devices[n_dev++] = atoi(tok);
    CHECK(n_dev >= 2);

    p3d_params prm;
    float matrix[25];
    p3d_scene_default_params(&prm, matrix);
    prm.world_size = 30.0f;
    const float ts = 1.0f / 60.0f;
    const size_t n = 27000;
    p3d_particle *cloud = (p3d_particle *)malloc(n * sizeof(p3d_particle));
    p3d_particle *out = (p3d_particle *)malloc(n * sizeof(p3d_particle));
    p3d_particle *single = (p3d_particle *)malloc(n * sizeof(p3d_particle));
    ora_particle *ref = (ora_particle *)malloc(n * sizeof(ora_particle));
    CHECK(cloud && out && single && ref);
    p3d_scene_uniform(42, n, prm.world_size, prm.id_count, cloud);

    ora_params op;
    memcpy(&op, &prm, sizeof(op) < sizeof(prm) ? sizeof(op) : sizeof(prm)); /* same field order by construction */
    op.attraction_matrix = matrix;
    CHECK(ora_update(&op, ts, (const ora_particle *)cloud, ref, n, ORA_IDEAL, 0, NULL, NULL, NULL, 0) == 0);

    p3d_engine *eng = NULL;
    CHECK(p3d_create_multi(devices, n_dev, &eng) == P3D_OK && eng != NULL);
    double worst = 0.0;
    const int kernels[3] = {P3D_FORCE_PAIR, P3D_FORCE_AUTO, P3D_FORCE_REFERENCE_ORDER};
    for (int k = 0; k < 3; ++k) {
        CHECK(p3d_set_option(eng, P3D_OPT_FORCE_KERNEL, kernels[k]) == P3D_OK);
        memset(out, 0, n * sizeof(p3d_particle));
        CHECK(p3d_update(eng, &prm, ts, cloud, out, n) == P3D_OK);
        CHECK(parity_ok(out, ref, n, prm.world_size, 1e-5, &worst));
        printf("multi_update: %d devices, kernel %d: worst parity error %.3g (tol 1e-5)\n", n_dev, kernels[k], worst);
    }
    /* the calls that belong to one-process-per-GPU drivers are refused on the handle */
    size_t a, b;
    CHECK(p3d_shard_range(eng, &a, &b) == P3D_ERR_INVALID);
    CHECK(p3d_set_shard(eng, 0, 2) == P3D_ERR_INVALID);
    /* the reference's panics, through the handle */
    prm.world_size = 3.9f;
    CHECK(p3d_update(eng, &prm, ts, cloud, out, n) == P3D_ERR_WORLD_TOO_SMALL);
    prm.world_size = 30.0f;
    cloud[n / 2].id = 5;
    CHECK(p3d_set_option(eng, P3D_OPT_FORCE_KERNEL, P3D_FORCE_PAIR) == P3D_OK);
    CHECK(p3d_update(eng, &prm, ts, cloud, out, n) == P3D_ERR_BAD_ID);
    CHECK(p3d_set_option(eng, P3D_OPT_FORCE_KERNEL, P3D_FORCE_CELLS) == P3D_OK);
    CHECK(p3d_update(eng, &prm, ts, cloud, out, n) == P3D_ERR_BAD_ID);
    cloud[n / 2].id = 1;
    p3d_scene_uniform(42, n, prm.world_size, prm.id_count, cloud);

    /* device-resident run: 20 steps on the handle vs 20 steps on one device (cell list: deterministic kernels, but the
     * sum over the devices' partial forces groups the additions differently, hence a tolerance) */
    p3d_engine *one = NULL;
    CHECK(p3d_create(devices[0], &one) == P3D_OK);
    CHECK(p3d_set_option(one, P3D_OPT_FORCE_KERNEL, P3D_FORCE_CELLS) == P3D_OK);
    CHECK(p3d_upload(eng, cloud, n, prm.id_count) == P3D_OK && p3d_upload(one, cloud, n, prm.id_count) == P3D_OK);
    CHECK(p3d_step(eng, &prm, ts, 20) == P3D_OK && p3d_step(one, &prm, ts, 20) == P3D_OK);
    CHECK(p3d_download(eng, out, n) == P3D_OK && p3d_download(one, single, n) == P3D_OK);
    CHECK(parity_ok(out, (const ora_particle *)single, n, prm.world_size, 1e-4, &worst));
    double d_multi[8], d_one[8];
    CHECK(p3d_diagnostics(eng, d_multi) == P3D_OK && p3d_diagnostics(one, d_one) == P3D_OK);
    CHECK(d_multi[5] == (double)n && fabs(d_multi[0] - d_one[0]) <= 1e-4 * d_one[0]);
    uint64_t cnt[4];
    /* (the cell-list run lives on the first device alone unless P3D_MULTI_CELLS_MIN says otherwise) */
    CHECK(p3d_get_counters(eng, cnt) == P3D_OK && cnt[1] >= 20u && cnt[2] >= 20u);
    p3d_destroy(one);
    p3d_destroy(eng);
    free(cloud); free(out); free(single); free(ref);
    printf("multi_update ok\n");
    return 0;
}
