/* Plain-C consumer of include/p3d.h: the two-body known answer of SURVEY.md Appendix C through p3d_update,
 * the reference's panics as error codes, the seeded scene, the render-buffer layout.  Exit code 0 = all good. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "p3d.h"

#define CHECK(cond)                                                          \
    do {                                                                     \
        if (!(cond)) {                                                       \
            fprintf(stderr, "FAILED %s:%d: %s (last error: %s)\n", __FILE__, __LINE__, #cond, p3d_last_error()); \
            return 1;                                                        \
        }                                                                    \
    } while (0)

int main(void) {
    p3d_params prm;
    float matrix[25];
    p3d_scene_default_params(&prm, matrix);
    CHECK(prm.world_size == 10.0f && prm.id_count == 5 && matrix[13] == 1.5f);

    p3d_engine *eng = NULL;
    CHECK(p3d_abi_version() == P3D_ABI_VERSION);
    CHECK(p3d_create(0, &eng) == P3D_OK && eng != NULL);

    /* two particles at rest, 0.65 apart, types 0 and 1: A[0][1] = A[1][0] = 1 at the triangle peak */
    p3d_particle in[2], out[2];
    memset(in, 0, sizeof(in));
    in[1].px = 0.65f;
    in[1].id = 1;
    const float ts = 1.0f / 60.0f;
    CHECK(p3d_update(eng, &prm, ts, in, out, 2) == P3D_OK);
    const double v = (2.0 / 60.0) * (1.0 - 0.97 / 60.0);
    CHECK(fabs(out[0].vx - v) < 2e-6 * v && fabs(out[1].vx + v) < 2e-6 * v);
    CHECK(fabs(out[0].px - v / 60.0) < 1e-8 && out[0].id == 0 && out[1].id == 1);

    /* the reference's panics */
    prm.world_size = 3.9f;
    CHECK(p3d_update(eng, &prm, ts, in, out, 2) == P3D_ERR_WORLD_TOO_SMALL);
    prm.world_size = 10.0f;
    in[1].id = 5;
    CHECK(p3d_update(eng, &prm, ts, in, out, 2) == P3D_ERR_BAD_ID);
    CHECK(strlen(p3d_last_error()) > 0);
    in[1].id = 1;
    CHECK(p3d_update(eng, &prm, ts, in, out, 0) == P3D_OK); /* empty system */

    /* device-resident run + render buffer */
    const size_t n = 3000;
    p3d_particle *cloud = (p3d_particle *)malloc(n * sizeof(p3d_particle));
    p3d_particle *back = (p3d_particle *)malloc(n * sizeof(p3d_particle));
    unsigned char *render = (unsigned char *)malloc(16 + 32 * n);
    p3d_scene_uniform(42, n, prm.world_size, prm.id_count, cloud);
    CHECK(p3d_upload(eng, cloud, n, prm.id_count) == P3D_OK);
    CHECK(p3d_step(eng, &prm, ts, 10) == P3D_OK);
    CHECK(p3d_download(eng, back, n) == P3D_OK);
    CHECK(p3d_download_render(eng, prm.world_size, render, 16 + 32 * n, n) == P3D_OK);
    for (size_t i = 0; i < n; ++i) {
        float p[3];
        unsigned id;
        memcpy(p, render + 16 + 32 * i, 12);
        memcpy(&id, render + 16 + 32 * i + 28, 4);
        CHECK(p[0] == back[i].px && p[1] == back[i].py && p[2] == back[i].pz && id == back[i].id);
        CHECK(fabsf(back[i].px) <= 5.0f);
    }
    double d[8];
    CHECK(p3d_diagnostics(eng, d) == P3D_OK && d[5] == (double)n && d[0] > 0.0);
    p3d_destroy(eng);
    free(cloud); free(back); free(render);
    printf("abi_smoke ok\n");
    return 0;
}
