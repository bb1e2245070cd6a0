/* The host-only part of the product (csrc/p3d_scene.cpp: seeded generators, default scene) under ASan + UBSan.
 * Built and run by tests/test_oracle_sanitize.py. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "p3d.h"

int main(void) {
    p3d_params prm;
    float m[25];
    p3d_scene_default_params(&prm, m);
    int bad = !(prm.world_size == 10.0f && prm.id_count == 5 && prm.attraction_matrix == m && m[24] == 0.5f);
    const size_t sizes[] = {0, 1, 7, 1000, 50000};
    for (size_t k = 0; k < sizeof(sizes) / sizeof(sizes[0]); ++k) {
        const size_t n = sizes[k];
        p3d_particle *p = malloc((n ? n : 1) * sizeof(*p));
        for (int plummer = 0; plummer < 2; ++plummer) {
            if (plummer) p3d_scene_plummer(42 + k, n, 64.0f, 64.0f / 6.0f, 5, p);
            else p3d_scene_uniform(42 + k, n, 10.0f, 5, p);
            const float half = plummer ? 32.0f : 5.0f;
            for (size_t i = 0; i < n; ++i)
                bad += !(fabsf(p[i].px) <= half && fabsf(p[i].py) <= half && fabsf(p[i].pz) <= half && p[i].id < 5 &&
                         p[i].vx == 0.0f && p[i].vy == 0.0f && p[i].vz == 0.0f);
        }
        free(p);
    }
    p3d_particle one;
    p3d_scene_uniform(1, 1, 1.0e30f, 1, &one);   /* extreme box, single type */
    p3d_scene_uniform(1, 1, 0.0f, 64, &one);
    printf("scene_sanitize: %s (%d)\n", bad ? "FAILED" : "clean", bad);
    return bad ? 1 : 0;
}
