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

// Optional self-checking build (-DP3D_BOUNDS_CHECK; compute-sanitizer is not available on the target pool):
// every data-dependent slot / cell index is compared with the extent the engine published before the
// launch; a violation is counted and the access skipped.  Compiles to nothing in the product build.
#ifdef P3D_BOUNDS_CHECK
static __device__ unsigned int g_p3d_slots, g_p3d_cells;
static __device__ unsigned long long g_p3d_oob;
__device__ __forceinline__ bool p3d_in_range(unsigned long long i, unsigned long long n) {
    if (i < n) return true;
    atomicAdd(&g_p3d_oob, 1ull);
    return false;
}
// extents are published by a kernel (by-value arguments survive CUDA-graph capture; a memcpy from a host
// variable would be replayed from a dangling pointer); 0xFFFFFFFF leaves a value as it is
__global__ void k_debug_publish(unsigned int slots, unsigned int cells) {
    if (slots != 0xFFFFFFFFu) g_p3d_slots = slots;
    if (cells != 0xFFFFFFFFu) g_p3d_cells = cells;
}
#define P3D_SLOT_OK(s) p3d_in_range((unsigned long long)(s), (unsigned long long)g_p3d_slots)
#define P3D_SLOT_END_OK(s) p3d_in_range((unsigned long long)(s), (unsigned long long)g_p3d_slots + 1ull)
#define P3D_CELL_OK(c) p3d_in_range((unsigned long long)(c), (unsigned long long)g_p3d_cells + 1ull)
#else
#define P3D_SLOT_OK(s) true
#define P3D_SLOT_END_OK(s) true
#define P3D_CELL_OK(c) true
#endif

__device__ __forceinline__ uint32_t f2u(float f) { return __float_as_uint(f); }
__device__ __forceinline__ float u2f(uint32_t u) { return __uint_as_float(u); }
