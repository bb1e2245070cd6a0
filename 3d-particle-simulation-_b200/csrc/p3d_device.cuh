// p3d_device.cuh — device-side parameter block and small helpers shared by all kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define P3D_GHOST_ID 0xFFFFFFFFu
// Ghost slots pad each type segment to a whole number of blocks.  They sit far outside any
// legal world so every distance test fails, and all share one position so ghost-ghost
// separations are exactly zero (contributing exactly zero).
#define P3D_GHOST_COORD 1.0e15f

// Canonicalised copy of p3d_params, passed by value to every kernel.
struct DevParams {
    float W;      // world_size                       (src/lib.rs:21)
    float half;   // world_size * 0.5f                (src/lib.rs:71)
    float r;      // particle_effect_radius           (src/lib.rs:30)
    float r2;     // r * r                            (src/lib.rs:218-219)
    float m;      // min_pull_ratio                   (src/lib.rs:29)
    float kf;     // interaction_force                (src/lib.rs:28)
    float coef;   // coefficient                      (src/lib.rs:27)
    float ax, ay, az;  // acceleration                (src/lib.rs:32)
    int walls;
    int T;        // id_count
    // constants of the branch-free force law used by the fast kernel:
    //   s(d) = f(d)/d = min(1/m - 1/d, 0) + a * max(0, min(c2 - c2*m/d, c2/d - c2))
    float inv_m;  // 1/m, or +inf when m <= 0 (the repulsion branch is then unreachable)
    float c2;     // 2/(1-m), or 0 when m >= 1 (the attraction branch is then unreachable)
    int rcut;     // 1 when r < max(1, m): the d2 < r^2 test cuts inside the force range and must be explicit
    float reach;  // min(r, max(1, m)): beyond this distance the force is exactly zero
};

__device__ __forceinline__ uint32_t f2u(float f) { return __float_as_uint(f); }
__device__ __forceinline__ float u2f(uint32_t u) { return __uint_as_float(u); }
