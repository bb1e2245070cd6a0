// p3d_kernels_cells.cuh — uniform-grid (cell list) force path: SURVEY.md §8f row 1, the GPU
// analogue of the reference's spatial hash (src/lib.rs:135-236).
//
// The reference hashes cells of edge r into N buckets and scans 27 images x 27 cells.  Here the box
// [-W/2, W/2]^3 is cut into nc^3 cells of edge W/nc >= reach = min(r, 1) — beyond `reach` the force
// law of src/lib.rs:55-67 (and the cutoff of :216-220) is exactly zero — particles are sorted by
// cell every step, and each particle scans the 27 neighbouring cells, periodic neighbours included.
// Every in-range (particle, image) pair is visited exactly once ("ideal" physics: the reference's
// bucket double visits are not reproduced).  Relative positions use the reference's image arithmetic:
// when a neighbour cell is reached through a face, the i-particle is shifted by the rounded
// `position + offset` (src/lib.rs:190-192,211-212), decided once per neighbour cell.
//
// This path does far fewer pair evaluations than N^2, so its throughput is reported as steps/s only,
// never as a fraction of the FP32 roofline.
#pragma once
#include "p3d_device.cuh"
#include "p3d_kernels_pair.cuh"

struct CellGrid {
    int nc;         // cells per axis (>= 3)
    float inv_cs;   // nc / W
    float half;     // W / 2
};

__device__ __forceinline__ int cell_axis(float x, const CellGrid g) {
    int c = (int)floorf((x + g.half) * g.inv_cs);
    return min(max(c, 0), g.nc - 1);  // x == +W/2 lands in the last cell
}

// keys[s] = linear cell index of slot s (ghosts: nc^3, sorted to the end); vals[s] = s.
__global__ void __launch_bounds__(256) k_cell_keys(const float4 *__restrict__ pos, int n_slots, CellGrid g,
                                                   uint32_t *__restrict__ keys, uint32_t *__restrict__ vals,
                                                   int *__restrict__ flag_to_clear) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s == 0) *flag_to_clear = 0;
    if (s >= n_slots) return;
    const float4 p = pos[s];
    uint32_t key = (uint32_t)(g.nc * g.nc * g.nc);
    if (f2u(p.w) != P3D_GHOST_ID) {
        const int cx = cell_axis(p.x, g), cy = cell_axis(p.y, g), cz = cell_axis(p.z, g);
        key = (uint32_t)((cz * g.nc + cy) * g.nc + cx);
    }
    keys[s] = key;
    vals[s] = (uint32_t)s;
}

// After the sort: gather positions into cell order and mark where each cell starts.
// cell_start has nc^3 + 1 entries, pre-filled with 0xFFFFFFFF ("empty") by a memset.
__global__ void __launch_bounds__(256) k_cell_gather(const float4 *__restrict__ pos, int n_slots,
                                                     const uint32_t *__restrict__ keys_sorted,
                                                     const uint32_t *__restrict__ vals_sorted,
                                                     float4 *__restrict__ cpos, uint32_t *__restrict__ cell_start,
                                                     uint32_t *__restrict__ cell_end) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_slots) return;
    const uint32_t key = keys_sorted[k];
    cpos[k] = pos[vals_sorted[k]];
    if (k == 0 || keys_sorted[k - 1] != key) cell_start[key] = (uint32_t)k;
    if (k == n_slots - 1 || keys_sorted[k + 1] != key) cell_end[key] = (uint32_t)(k + 1);
}

// One thread per particle in cell order.  i_begin/i_end shard the sorted range across GPUs.
template <bool RCUT>
__global__ void __launch_bounds__(128) k_force_cells(const float4 *__restrict__ cpos,
                                                     const uint32_t *__restrict__ keys_sorted,
                                                     const uint32_t *__restrict__ vals_sorted,
                                                     const uint32_t *__restrict__ cell_start,
                                                     const uint32_t *__restrict__ cell_end, int n_slots, int i_begin,
                                                     int i_end, CellGrid g, float4 *__restrict__ frc, DevParams P,
                                                     const float *__restrict__ matrix,
                                                     const int *__restrict__ flags) {
    if (flags[0] != 0) return;  // out-of-box input: the reference-order kernel takes the step
    extern __shared__ float smat_dyn[];
    for (int k = threadIdx.x; k < P.T * P.T; k += blockDim.x) smat_dyn[k] = matrix[k];
    __syncthreads();
    const int k = i_begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= i_end) return;
    const uint32_t key = keys_sorted[k];
    const int nc = g.nc;
    if (key >= (uint32_t)(nc * nc * nc)) return;  // ghost
    const float4 pi = cpos[k];
    const int cx = (int)(key % (uint32_t)nc), cy = (int)((key / (uint32_t)nc) % (uint32_t)nc),
              cz = (int)(key / (uint32_t)(nc * nc));
    const float *arow = smat_dyn + f2u(pi.w) * (uint32_t)P.T;
    const float c2 = P.c2, ncm = -P.c2 * P.m, nc2 = -P.c2, im = P.inv_m, r2 = P.r2;
    // `position + offset` for offset = -W and +W (src/lib.rs:190-192), rounded like the reference
    const float pxm = __fadd_rn(pi.x, -P.W), pxp = __fadd_rn(pi.x, P.W);
    const float pym = __fadd_rn(pi.y, -P.W), pyp = __fadd_rn(pi.y, P.W);
    const float pzm = __fadd_rn(pi.z, -P.W), pzp = __fadd_rn(pi.z, P.W);
    float ax = 0.f, ay = 0.f, az = 0.f;
#pragma unroll 1
    for (int dz = -1; dz <= 1; ++dz) {
        int nz = cz + dz;
        float pz = pi.z;
        if (nz < 0) { nz += nc; pz = pzp; }        // neighbour lies through the -z face: it sees us at z + W
        else if (nz >= nc) { nz -= nc; pz = pzm; }
#pragma unroll 1
        for (int dy = -1; dy <= 1; ++dy) {
            int ny = cy + dy;
            float py = pi.y;
            if (ny < 0) { ny += nc; py = pyp; }
            else if (ny >= nc) { ny -= nc; py = pym; }
#pragma unroll 1
            for (int dx = -1; dx <= 1; ++dx) {
                int nx = cx + dx;
                float px = pi.x;
                if (nx < 0) { nx += nc; px = pxp; }
                else if (nx >= nc) { nx -= nc; px = pxm; }
                const uint32_t c = (uint32_t)((nz * nc + ny) * nc + nx);
                const uint32_t s0 = cell_start[c];
                if (s0 == 0xFFFFFFFFu) continue;
                const uint32_t s1 = cell_end[c];
                for (uint32_t j = s0; j < s1; ++j) {
                    const float4 q = cpos[j];
                    const float rx = __fsub_rn(q.x, px), ry = __fsub_rn(q.y, py), rz = __fsub_rn(q.z, pz);
                    const float d2 = fmaf(rz, rz, fmaf(ry, ry, fmaf(rx, rx, 1.0e-30f)));
                    const float inv = rsqrt_approx(d2);
                    const float p1 = fmaf(inv, ncm, c2), p2 = fmaf(inv, c2, nc2);
                    float ti = fmaxf(fminf(p1, p2), 0.0f);
                    float rs = fminf(im - inv, 0.0f);
                    if (RCUT) {
                        if (!(d2 < r2)) { ti = 0.0f; rs = 0.0f; }
                    }
                    const float s = fmaf(arow[f2u(q.w)], ti, rs);
                    ax = fmaf(rx, s, ax);
                    ay = fmaf(ry, s, ay);
                    az = fmaf(rz, s, az);
                }
            }
        }
    }
    frc[vals_sorted[k]] = make_float4(ax, ay, az, 0.f);
}
