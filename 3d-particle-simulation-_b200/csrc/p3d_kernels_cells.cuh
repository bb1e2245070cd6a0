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
            // the three cells of this row: fetch their ranges together (independent loads in flight)
            uint32_t s0[3], s1[3];
            float pxs[3];
#pragma unroll
            for (int dx = -1; dx <= 1; ++dx) {
                int nx = cx + dx;
                float px = pi.x;
                if (nx < 0) { nx += nc; px = pxp; }
                else if (nx >= nc) { nx -= nc; px = pxm; }
                const uint32_t c = (uint32_t)((nz * nc + ny) * nc + nx);
                s0[dx + 1] = __ldg(cell_start + c);
                s1[dx + 1] = __ldg(cell_end + c);
                pxs[dx + 1] = px;
            }
#pragma unroll
            for (int t = 0; t < 3; ++t) {
                if (s0[t] == 0xFFFFFFFFu) continue;
                const float px = pxs[t];
                for (uint32_t j = s0[t]; j < s1[t]; ++j) {
                    const float4 q = __ldg(cpos + j);
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

// ---------------------------------------------------------------------------------------------
// K5 — optional "faithful" correction: reproduces the reference's bucket double-visit quirk
// (SURVEY.md Appendix B.1).  The reference hashes the 27 cells around cell_coord(position + offset)
// into N buckets (src/lib.rs:195-202) and scans each hit bucket's whole list; a neighbour q is
// therefore visited once for every one of the 27 cells whose bucket equals the bucket of q's own
// cell — usually once, sometimes twice when two cells collide modulo N.  The ideal force counts
// every in-range (particle, image) pair once, so the reference's result is
//     ideal + sum over in-range pairs of (multiplicity - 1) * contribution.
// This kernel adds that sum.  Restated from the definitions: SipHash-1-3 with zero keys over three
// little-endian i64 words (Rust DefaultHasher, src/lib.rs:46-52), `(v / r) as isize` (src/lib.rs:37-43).
__device__ __forceinline__ uint64_t rotl64_dev(uint64_t x, int b) { return (x << b) | (x >> (64 - b)); }

#define P3D_SIPROUND(v0, v1, v2, v3)                                                    \
    do {                                                                                \
        v0 += v1; v1 = rotl64_dev(v1, 13); v1 ^= v0; v0 = rotl64_dev(v0, 32);           \
        v2 += v3; v3 = rotl64_dev(v3, 16); v3 ^= v2;                                    \
        v0 += v3; v3 = rotl64_dev(v3, 21); v3 ^= v0;                                    \
        v2 += v1; v1 = rotl64_dev(v1, 17); v1 ^= v2; v2 = rotl64_dev(v2, 32);           \
    } while (0)

__device__ __forceinline__ uint64_t hash_cell_dev(long long x, long long y, long long z) {
    uint64_t v0 = 0x736f6d6570736575ULL, v1 = 0x646f72616e646f6dULL;
    uint64_t v2 = 0x6c7967656e657261ULL, v3 = 0x7465646279746573ULL;
    uint64_t m;
    m = (uint64_t)x; v3 ^= m; P3D_SIPROUND(v0, v1, v2, v3); v0 ^= m;
    m = (uint64_t)y; v3 ^= m; P3D_SIPROUND(v0, v1, v2, v3); v0 ^= m;
    m = (uint64_t)z; v3 ^= m; P3D_SIPROUND(v0, v1, v2, v3); v0 ^= m;
    m = (uint64_t)24 << 56; v3 ^= m; P3D_SIPROUND(v0, v1, v2, v3); v0 ^= m;  // length byte of the 24-byte message
    v2 ^= 0xff;
    P3D_SIPROUND(v0, v1, v2, v3);
    P3D_SIPROUND(v0, v1, v2, v3);
    P3D_SIPROUND(v0, v1, v2, v3);
    return v0 ^ v1 ^ v2 ^ v3;
}

// Rust `(v / r) as isize`: IEEE divide, truncate toward zero, saturate, NaN -> 0 (cvt.rzi.s64.f32 does all three).
__device__ __forceinline__ long long ref_cell_axis(float v, float r) { return __float2ll_rz(__fdiv_rn(v, r)); }

template <bool RCUT>
__global__ void __launch_bounds__(128) k_quirk_correction(const float4 *__restrict__ cpos,
                                                          const uint32_t *__restrict__ keys_sorted,
                                                          const uint32_t *__restrict__ vals_sorted,
                                                          const uint32_t *__restrict__ cell_start,
                                                          const uint32_t *__restrict__ cell_end, int i_begin,
                                                          int i_end, CellGrid g, float4 *__restrict__ frc,
                                                          DevParams P, const float *__restrict__ matrix,
                                                          const int *__restrict__ flags, unsigned long long n_hash) {
    if (flags[0] != 0) return;  // out-of-box input: not supported by the correction (documented)
    extern __shared__ float smat_dyn[];
    for (int k = threadIdx.x; k < P.T * P.T; k += blockDim.x) smat_dyn[k] = matrix[k];
    __syncthreads();
    const int k = i_begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= i_end) return;
    const uint32_t key = keys_sorted[k];
    const int nc = g.nc;
    if (key >= (uint32_t)(nc * nc * nc)) return;
    const float4 pi = cpos[k];
    const int cx = (int)(key % (uint32_t)nc), cy = (int)((key / (uint32_t)nc) % (uint32_t)nc),
              cz = (int)(key / (uint32_t)(nc * nc));
    const float *arow = smat_dyn + f2u(pi.w) * (uint32_t)P.T;
    const float c2 = P.c2, ncm = -P.c2 * P.m, nc2 = -P.c2, im = P.inv_m, r2 = P.r2;
    const float pxm = __fadd_rn(pi.x, -P.W), pxp = __fadd_rn(pi.x, P.W);
    const float pym = __fadd_rn(pi.y, -P.W), pyp = __fadd_rn(pi.y, P.W);
    const float pzm = __fadd_rn(pi.z, -P.W), pzp = __fadd_rn(pi.z, P.W);
    float ax = 0.f, ay = 0.f, az = 0.f;
    bool any = false;
#pragma unroll 1
    for (int dz = -1; dz <= 1; ++dz) {
        int nz = cz + dz;
        float pz = pi.z;
        if (nz < 0) { nz += nc; pz = pzp; } else if (nz >= nc) { nz -= nc; pz = pzm; }
#pragma unroll 1
        for (int dy = -1; dy <= 1; ++dy) {
            int ny = cy + dy;
            float py = pi.y;
            if (ny < 0) { ny += nc; py = pyp; } else if (ny >= nc) { ny -= nc; py = pym; }
#pragma unroll 1
            for (int dx = -1; dx <= 1; ++dx) {
                int nx = cx + dx;
                float px = pi.x;
                if (nx < 0) { nx += nc; px = pxp; } else if (nx >= nc) { nx -= nc; px = pxm; }
                const uint32_t c = (uint32_t)((nz * nc + ny) * nc + nx);
                const uint32_t s0 = cell_start[c];
                if (s0 == 0xFFFFFFFFu) continue;
                const uint32_t s1 = cell_end[c];
                for (uint32_t j = s0; j < s1; ++j) {
                    const float4 q = cpos[j];
                    const float rx = __fsub_rn(q.x, px), ry = __fsub_rn(q.y, py), rz = __fsub_rn(q.z, pz);
                    // the reference's own cutoff test, in its arithmetic (src/lib.rs:213-220)
                    const float d2e = __fadd_rn(__fadd_rn(__fmul_rn(rx, rx), __fmul_rn(ry, ry)), __fmul_rn(rz, rz));
                    if (!(d2e > 0.0f && d2e < r2)) continue;
                    const float inv = rsqrt_approx(d2e);
                    const float p1 = fmaf(inv, ncm, c2), p2 = fmaf(inv, c2, nc2);
                    const float ti = fmaxf(fminf(p1, p2), 0.0f);
                    const float rs = fminf(im - inv, 0.0f);
                    const float s = fmaf(arow[f2u(q.w)], ti, rs);
                    if (s == 0.0f) continue;
                    // multiplicity: how many of the 27 cells around cell_coord(position + offset) share q's bucket
                    const uint64_t bq = hash_cell_dev(ref_cell_axis(q.x, P.r), ref_cell_axis(q.y, P.r),
                                                      ref_cell_axis(q.z, P.r)) % n_hash;
                    const long long c0x = ref_cell_axis(px, P.r), c0y = ref_cell_axis(py, P.r),
                                    c0z = ref_cell_axis(pz, P.r);
                    int mult = 0;
                    for (int ex = -1; ex <= 1; ++ex)
                        for (int ey = -1; ey <= 1; ++ey)
                            for (int ez = -1; ez <= 1; ++ez)
                                mult += (hash_cell_dev(c0x + ex, c0y + ey, c0z + ez) % n_hash) == bq;
                    if (mult != 1) {
                        const float w = (float)(mult - 1) * s;
                        ax = fmaf(rx, w, ax);
                        ay = fmaf(ry, w, ay);
                        az = fmaf(rz, w, az);
                        any = true;
                    }
                }
            }
        }
    }
    if (any) atomicAdd(frc + vals_sorted[k], make_float4(ax, ay, az, 0.0f));
}
