// p3d_kernels_cells.cuh — uniform-grid (cell list) force path: SURVEY.md §8f row 1, the GPU
// analogue of the reference's spatial hash (src/lib.rs:135-236).
//
// The reference hashes cells of edge r into N buckets and scans 27 images x 27 cells.  Here the box
// [-W/2, W/2]^3 is cut into nc^3 cells of edge W/nc >= reach = min(r, max(1, m)) — beyond `reach` the force
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

// Cell of a coordinate, taken modulo the box: positions outside [-W/2, W/2) (callers may pass anything, and a
// fast particle can leave the box, src/lib.rs:74-92 wraps only once) land in the cell of their periodic
// image, so that every pair the reference can reach through its offsets -W, 0, +W sits in adjacent cells.
__device__ __forceinline__ int cell_axis(float x, const CellGrid g) {
    const float W = 2.0f * g.half;
    float u = x + g.half;
    // in-box positions (x == +W/2 included: walls clamp particles exactly there) keep their own cell, so that
    // the image of a neighbour follows from the neighbour cell; only outside positions are wrapped
    if (!(u >= 0.0f && u <= W)) u = u - W * floorf(u / W);
    int c = (int)floorf(u * g.inv_cs);
    return min(max(c, 0), g.nc - 1);  // x == +W/2 lands in the last cell; NaN lands in cell 0
}

// ---------------------------------------------------------------------------------------------
// Sort by cell: a hand-written counting sort (count -> scan -> scatter -> order), the GPU analogue of the
// reference's own counting sort over hash buckets (src/lib.rs:135-164: count, prefix sum, fill).
//
// `gate` (nullable): every kernel of the pipeline returns at once unless gate[0] == gate_value.  The all-pairs
// path queues the pipeline behind its out-of-box flag, so that a device-resident run can hand a step with a
// particle outside the box to the cell list without the host looking at the flag.
#define P3D_GATED(gate, gate_value) \
    if ((gate) != nullptr && (gate)[0] != (gate_value)) return

// Pass 1.  keys[s] = linear cell index of slot s (ghosts: nc^3, a bin of their own after the last cell);
// rank[s] = arrival order inside the cell (atomic: arbitrary, fixed up by pass 4); count[c] = population.
__global__ void __launch_bounds__(256) k_cell_count(const float4 *__restrict__ pos, int n_slots, CellGrid g,
                                                    uint32_t *__restrict__ keys, uint32_t *__restrict__ rank,
                                                    uint32_t *__restrict__ count, int *__restrict__ flag_to_clear,
                                                    const int *__restrict__ gate, int gate_value) {
    P3D_GATED(gate, gate_value);
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s == 0 && flag_to_clear) *flag_to_clear = 0;
    if (s >= n_slots) return;
    const float4 p = pos[s];
    uint32_t key = (uint32_t)(g.nc * g.nc * g.nc);
    if (f2u(p.w) != P3D_GHOST_ID) {
        const int cx = cell_axis(p.x, g), cy = cell_axis(p.y, g), cz = cell_axis(p.z, g);
        key = (uint32_t)((cz * g.nc + cy) * g.nc + cx);
    }
    keys[s] = key;
    if (P3D_CELL_OK(key)) rank[s] = atomicAdd(count + key, 1u);
}

// Pass 2: exclusive prefix sum of count[0 .. L) -> off[0 .. L) in three launches (tile sums, scan of the tile
// sums by one CTA, per-tile scan + offset).  off[c] = first sorted index of cell c; with L = nc^3 + 1 the last
// entry, off[nc^3], is where the ghosts start.
constexpr int kScanThreads = 1024;
constexpr int kScanTile = 4 * kScanThreads;  // four consecutive counts per thread

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t *warp_tot /* [32] shared */, uint32_t &total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t a = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += a;
    }
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        const uint32_t w = warp_tot[lane];
        uint32_t winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t a = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += a;
        }
        warp_tot[lane] = winc - w;                    // exclusive prefix of the warp totals
        if (lane == 31) warp_tot[32] = winc;          // block total
    }
    __syncthreads();
    total = warp_tot[32];
    return warp_tot[warp] + inc - v;
}

__global__ void __launch_bounds__(kScanThreads) k_scan_tile_sums(const uint32_t *__restrict__ in, int L,
                                                                 uint32_t *__restrict__ tile_sum,
                                                                 const int *__restrict__ gate, int gate_value) {
    P3D_GATED(gate, gate_value);
    __shared__ uint32_t warp_tot[33];
    const int base = blockIdx.x * kScanTile + 4 * threadIdx.x;
    uint32_t v = 0u;
#pragma unroll
    for (int k = 0; k < 4; ++k)
        if (base + k < L) v += in[base + k];
    uint32_t total;
    block_exclusive_scan(v, warp_tot, total);
    if (threadIdx.x == 0) tile_sum[blockIdx.x] = total;
}

__global__ void __launch_bounds__(kScanThreads) k_scan_top(uint32_t *__restrict__ tile_sum, int n_tiles,
                                                           const int *__restrict__ gate, int gate_value) {
    P3D_GATED(gate, gate_value);
    __shared__ uint32_t warp_tot[33];
    uint32_t carry = 0u;
    for (int base = 0; base < n_tiles; base += kScanThreads) {
        const int t = base + threadIdx.x;
        const uint32_t v = t < n_tiles ? tile_sum[t] : 0u;
        uint32_t total;
        const uint32_t ex = block_exclusive_scan(v, warp_tot, total);
        if (t < n_tiles) tile_sum[t] = carry + ex;
        carry += total;
        __syncthreads();  // warp_tot is reused by the next trip
    }
}

__global__ void __launch_bounds__(kScanThreads) k_scan_apply(const uint32_t *__restrict__ in, int L,
                                                             const uint32_t *__restrict__ tile_off,
                                                             uint32_t *__restrict__ out,
                                                             const int *__restrict__ gate, int gate_value) {
    P3D_GATED(gate, gate_value);
    __shared__ uint32_t warp_tot[33];
    const int base = blockIdx.x * kScanTile + 4 * threadIdx.x;
    uint32_t c[4], v = 0u;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        c[k] = (base + k < L) ? in[base + k] : 0u;
        v += c[k];
    }
    uint32_t total;
    uint32_t at = tile_off[blockIdx.x] + block_exclusive_scan(v, warp_tot, total);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (base + k < L) out[base + k] = at;
        at += c[k];
    }
}

// Pass 3: slot s goes to position off[cell] + arrival rank (any order inside a cell).
__global__ void __launch_bounds__(256) k_cell_scatter(int n_slots, const uint32_t *__restrict__ keys,
                                                      const uint32_t *__restrict__ rank,
                                                      const uint32_t *__restrict__ cell_off,
                                                      uint32_t *__restrict__ members,
                                                      const int *__restrict__ gate, int gate_value) {
    P3D_GATED(gate, gate_value);
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_slots) return;
    const uint32_t c = keys[s];
    if (!P3D_CELL_OK(c)) return;
    const uint32_t dst = cell_off[c] + rank[s];
    if (P3D_SLOT_OK(dst)) members[dst] = (uint32_t)s;
}

// Pass 4: make the order inside every cell the CALLER order (a stable sort by the caller's particle index, hence the
// same forces bit for bit on every run, on every rank and whatever the slot permutation, see reslot_by_cell), and
// gather keys / slots / positions into cell order.  A slot's rank is the number of members of its cell with a
// smaller caller index (caller_of == nullptr: slot == caller index): a handful of reads at ordinary densities.  Cells with more than
// kStableMax members (and the ghost bin, whose entries are all alike) keep the arrival order: the force on a
// particle is then still exact to rounding, only the summation order of such a cell may differ between runs.
constexpr uint32_t kStableMax = 2048u;

__global__ void __launch_bounds__(256) k_cell_order(const float4 *__restrict__ pos, int n_slots,
                                                    const uint32_t *__restrict__ keys, const uint32_t *__restrict__ rank,
                                                    const uint32_t *__restrict__ cell_off,
                                                    const uint32_t *__restrict__ members, uint32_t n_cells,
                                                    const uint32_t *__restrict__ caller_of,
                                                    uint32_t *__restrict__ keys_sorted, uint32_t *__restrict__ vals_sorted,
                                                    float4 *__restrict__ cpos,
                                                    const int *__restrict__ gate, int gate_value) {
    P3D_GATED(gate, gate_value);
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_slots) return;
    const uint32_t c = keys[s];
    if (!P3D_CELL_OK(c)) return;
    const uint32_t lo = cell_off[c];
    uint32_t r = rank[s];
    if (c < n_cells) {
        const uint32_t hi = P3D_SLOT_END_OK(cell_off[c + 1]) ? cell_off[c + 1] : lo;
        if (hi - lo > 1u && hi - lo <= kStableMax) {
            r = 0u;
            if (caller_of) {
                const uint32_t me = caller_of[s];
                for (uint32_t j = lo; j < hi; ++j) {
                    const uint32_t o = members[j];
                    if (P3D_SLOT_OK(o)) r += caller_of[o] < me;
                }
            } else {
                for (uint32_t j = lo; j < hi; ++j) r += members[j] < (uint32_t)s;
            }
        }
    }
    const uint32_t dst = lo + r;
    if (!P3D_SLOT_OK(dst)) return;
    keys_sorted[dst] = c;
    vals_sorted[dst] = (uint32_t)s;
    cpos[dst] = pos[s];
}

// Re-slotting (reslot_by_cell): sorted position k takes over the particle of old slot vals_sorted[k].  Positions were
// already gathered by k_cell_order (cpos); this moves the velocities and rewrites both index tables.
__global__ void __launch_bounds__(256) k_reslot(int n_slots, int n, const uint32_t *__restrict__ vals_sorted,
                                                const float4 *__restrict__ vel, const uint32_t *__restrict__ caller_of,
                                                float4 *__restrict__ vel_new, uint32_t *__restrict__ caller_new,
                                                uint32_t *__restrict__ slot_of) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_slots) return;
    const uint32_t old = vals_sorted[k];
    if (!P3D_SLOT_OK(old)) return;
    vel_new[k] = vel[old];
    const uint32_t caller = caller_of ? caller_of[old] : (old < (uint32_t)n ? old : P3D_GHOST_ID);
    caller_new[k] = caller;
    if (caller != P3D_GHOST_ID && caller < (uint32_t)n) slot_of[caller] = (uint32_t)k;
}

// One candidate of the cell list: the reference's relative position (src/lib.rs:211-212) and the
// branch-free force law of k_force_pair.
// GENERAL (some particle is outside the box): the image is not implied by the neighbour cell any more, so
// each axis takes the nearest of the reference's three candidates `other - (position + offset)`,
// offset in {-W, 0, +W} (src/lib.rs:190-192,211-212) — pairs that would need a larger offset are not
// reachable in the reference either and fail the distance test here too.
template <bool RCUT, bool GENERAL>
__device__ __forceinline__ void cell_pair(const float4 q, float px, float py, float pz, const float *sx3,
                                          const float *sy3, const float *sz3, const float *arow, float c2,
                                          float ncm, float nc2, float im, float r2, float &ax, float &ay, float &az) {
    float rx, ry, rz;
    if (GENERAL) {
        rx = nearest_image3(q.x, sx3[0], sx3[2], sx3[1]);
        ry = nearest_image3(q.y, sy3[0], sy3[2], sy3[1]);
        rz = nearest_image3(q.z, sz3[0], sz3[2], sz3[1]);
    } else {
        rx = __fsub_rn(q.x, px); ry = __fsub_rn(q.y, py); rz = __fsub_rn(q.z, pz);
    }
    const float d2 = fmaf(rz, rz, fmaf(ry, ry, fmaf(rx, rx, 1.0e-30f)));
    const float inv = rsqrt_approx(d2);
    const float p1 = fmaf(inv, ncm, c2), p2 = fmaf(inv, c2, nc2);
    float ti = fmaxf(fminf(p1, p2), 0.0f);
    float rs = fminf(im - inv, 0.0f);
    if (RCUT || GENERAL) {  // (in GENERAL mode a candidate may be a far image: the cutoff must be explicit)
        if (!(d2 < r2)) { ti = 0.0f; rs = 0.0f; }
    }
    if (GENERAL) {
        // a NaN or infinite coordinate (callers may pass anything) fails `d2 > 0 && d2 < r^2` in the reference
        // (src/lib.rs:216-220) and contributes nothing: the particle is inert.  0 * NaN would poison the sum.
        if (!(d2 < r2)) { rx = 0.0f; ry = 0.0f; rz = 0.0f; }
    }
    const float s = fmaf(arow[f2u(q.w)], ti, rs);
    ax = fmaf(rx, s, ax);
    ay = fmaf(ry, s, ay);
    az = fmaf(rz, s, az);
}

// One thread per particle in cell order.  i_begin/i_end shard the sorted range across GPUs.
//
// The candidates of a particle are up to 18 contiguous runs of the sorted array: for each of the 9 (dy,dz) rows the
// cells cx-1..cx+1 that do not wrap form ONE run (cells are x-fastest), plus one single-cell run when the row wraps
// through an x face.  ncu on the earlier per-run loops: instruction-issue bound, 12-15 of 32 lanes active, and most
// of the instructions were RUN SWITCHES executed for a few lanes at a time (every lane exhausts its ~3-candidate
// runs on different trips).  Hence two phases:
//   A  all lanes walk the 9 rows together (uniform control flow: the offsets of a row are loaded by every lane at the
//      same time) and write the sorted indices of their candidates into a per-thread list in shared memory;
//   B  one flat loop over the list: one candidate per trip, the next one already in flight; a warp iterates
//      max-over-lanes(total candidates) times, ~40 at density 1 against a mean of 27, with no control flow inside.
// Lanes whose list would overflow (dense regions: long runs amortise the switches anyway) and the GENERAL variant
// take the direct per-run loop instead.  The candidate ORDER is the same on both paths (the nine rows z-major, then
// the nine cells behind the x face for edge cells), so the forces do not depend on which path a lane takes.
constexpr int kCellThreads = 128;
// 48 entries (24 KB per CTA) with 8 resident CTAs per SM measured best on the B200 at N = 1M, density 1 (candidates per
// particle ~ Poisson(27)): 0.179 ms per step against 0.184 (64 entries, 6 CTAs) and 0.207 (40 entries, 10 CTAs, spills)
constexpr int kCellList = 48;                 // list entries per thread
constexpr int kCellMinBlocks = 8;
constexpr uint32_t kCellWrapBit = 0x80000000u;  // list entry: candidate seen through the x face ...
constexpr int kCellImgY = 29, kCellImgZ = 27;   // ... and 2 bits each for the y / z image (0: none, 1: +W, 2: -W)
constexpr uint32_t kCellIndexMask = (1u << 27) - 1u;  // sorted index (n_slots < 2^27)
constexpr int kCellEdgeReserve = 12;            // list entries kept free for an edge cell's nine wrap cells

template <bool RCUT, bool GENERAL>
__global__ void __launch_bounds__(kCellThreads, kCellMinBlocks) k_force_cells(const float4 *__restrict__ cpos,
                                                              const uint32_t *__restrict__ keys_sorted,
                                                              const uint32_t *__restrict__ vals_sorted,
                                                              const uint32_t *__restrict__ cell_off, int n_slots,
                                                              int i_begin, int i_end, CellGrid g,
                                                              float4 *__restrict__ frc, DevParams P,
                                                              const float *__restrict__ matrix,
                                                              const int *__restrict__ flags) {
    if ((flags[0] != 0) != GENERAL) return;  // the in-box and the general variant are both launched; one runs
                                             // (the all-pairs path queues only the general one: its fallback)
    extern __shared__ float smat_dyn[];
    __shared__ uint32_t cand[GENERAL ? 1 : kCellList][kCellThreads];
    for (int k = threadIdx.x; k < P.T * P.T; k += blockDim.x) smat_dyn[k] = matrix[k];
    __syncthreads();
    const int nc = g.nc;
    const int t = threadIdx.x;
    const float c2 = P.c2, ncm = -P.c2 * P.m, nc2 = -P.c2, im = P.inv_m, r2 = P.r2;
    // One particle per thread; the (rarely running) general variant is launched with a small grid and strides.
    for (int k = i_begin + blockIdx.x * blockDim.x + threadIdx.x; k < i_end; k += gridDim.x * blockDim.x) {
        const uint32_t key = keys_sorted[k];
        if (key >= (uint32_t)(nc * nc * nc)) continue;  // ghost
        const float4 pi = cpos[k];
        const int cx = (int)(key % (uint32_t)nc), cy = (int)((key / (uint32_t)nc) % (uint32_t)nc),
                  cz = (int)(key / (uint32_t)(nc * nc));
        const float *arow = smat_dyn + f2u(pi.w) * (uint32_t)P.T;
        // `position + offset` for offset = +W / -W (src/lib.rs:190-192), rounded like the reference
        const float sx3[3] = {pi.x, __fadd_rn(pi.x, P.W), __fadd_rn(pi.x, -P.W)};
        const float sy3[3] = {pi.y, __fadd_rn(pi.y, P.W), __fadd_rn(pi.y, -P.W)};
        const float sz3[3] = {pi.z, __fadd_rn(pi.z, P.W), __fadd_rn(pi.z, -P.W)};
        const bool x_edge = (cx == 0) || (cx == nc - 1);
        const int x0 = max(cx - 1, 0), x1 = min(cx + 1, nc - 1);
        const int xw = (cx == 0) ? nc - 1 : 0;             // the cell behind the x face (edge cells only) ...
        const float pxw = (cx == 0) ? sx3[1] : sx3[2];     // ... sees us at x + W resp. x - W
        float ax = 0.f, ay = 0.f, az = 0.f;

        bool direct = GENERAL;
        if (!GENERAL) {
            // ---- phase A: candidate list.  All 18 offsets of the 9 rows are loaded BEFORE any of them is used (one
            // memory latency instead of nine); rows behind a y/z face carry their image in the entry's flag bits. ----
            uint32_t lo[9], hi[9], img[9];
            int total = 0;
#pragma unroll
            for (int row9 = 0; row9 < 9; ++row9) {
                const int dz = row9 / 3 - 1, dy = row9 % 3 - 1;
                int nz = cz + dz, ny = cy + dy;
                uint32_t im9 = 0u;
                if (dz < 0 && nz < 0) { nz += nc; im9 |= 1u << kCellImgZ; }
                if (dz > 0 && nz >= nc) { nz -= nc; im9 |= 2u << kCellImgZ; }
                if (dy < 0 && ny < 0) { ny += nc; im9 |= 1u << kCellImgY; }
                if (dy > 0 && ny >= nc) { ny -= nc; im9 |= 2u << kCellImgY; }
                const uint32_t row = (uint32_t)((nz * nc + ny) * nc);
                lo[row9] = __ldg(cell_off + row + x0);
                hi[row9] = __ldg(cell_off + row + x1 + 1);
                img[row9] = im9;
            }
#pragma unroll
            for (int row9 = 0; row9 < 9; ++row9) {
                if (!P3D_SLOT_END_OK(hi[row9])) hi[row9] = lo[row9];  // (self-checking build only)
                total += (int)(hi[row9] - lo[row9]);
            }
            direct = total > kCellList - (x_edge ? kCellEdgeReserve : 0);
            int cnt = 0;
            if (!direct) {
#pragma unroll
                for (int row9 = 0; row9 < 9; ++row9)
                    for (uint32_t j = lo[row9]; j < hi[row9]; ++j) cand[cnt++][t] = j | img[row9];
                if (x_edge) {  // the cells behind the x face, after the main runs (same order as the direct loop)
#pragma unroll 1
                    for (int row9 = 0; row9 < 9 && !direct; ++row9) {
                        const int dz = row9 / 3 - 1, dy = row9 - (row9 / 3) * 3 - 1;
                        int nz = cz + dz, ny = cy + dy;
                        uint32_t flags9 = kCellWrapBit;  // (recomputed: indexing img[] by a loop variable would spill it)
                        if (nz < 0) { nz += nc; flags9 |= 1u << kCellImgZ; } else if (nz >= nc) { nz -= nc; flags9 |= 2u << kCellImgZ; }
                        if (ny < 0) { ny += nc; flags9 |= 1u << kCellImgY; } else if (ny >= nc) { ny -= nc; flags9 |= 2u << kCellImgY; }
                        const uint32_t row = (uint32_t)((nz * nc + ny) * nc);
                        uint32_t jw = __ldg(cell_off + row + xw), hw = __ldg(cell_off + row + xw + 1);
                        if (!P3D_SLOT_END_OK(hw)) hw = jw;
                        if (cnt + (int)(hw - jw) > kCellList) { direct = true; break; }
                        for (; jw < hw; ++jw) cand[cnt++][t] = jw | flags9;
                    }
                }
            }
            // ---- phase B: flat loop over the list, the next candidate's load issued before this one is evaluated.
            // Warps without a lane in a y/z face row (nearly all) skip the y/z image decode. ----
            const bool face = cy == 0 || cy == nc - 1 || cz == 0 || cz == nc - 1;  // (y/z images; the x image is one select)
            if (!direct && cnt > 0) {
                uint32_t e = cand[0][t];
                float4 q = __ldg(cpos + (e & kCellIndexMask));
                if (__any_sync(__activemask(), face)) {
                    for (int i = 0; i < cnt; ++i) {
                        const uint32_t e_next = cand[min(i + 1, cnt - 1)][t];
                        const float4 q_next = __ldg(cpos + (e_next & kCellIndexMask));
                        const uint32_t iy = (e >> kCellImgY) & 3u, iz = (e >> kCellImgZ) & 3u;
                        const float px = (e & kCellWrapBit) ? pxw : pi.x;
                        const float py = iy == 0u ? sy3[0] : (iy == 1u ? sy3[1] : sy3[2]);
                        const float pz = iz == 0u ? sz3[0] : (iz == 1u ? sz3[1] : sz3[2]);
                        cell_pair<RCUT, false>(q, px, py, pz, sx3, sy3, sz3, arow, c2, ncm, nc2, im, r2, ax, ay, az);
                        e = e_next;
                        q = q_next;
                    }
                } else {
                    for (int i = 0; i < cnt; ++i) {
                        const uint32_t e_next = cand[min(i + 1, cnt - 1)][t];
                        const float4 q_next = __ldg(cpos + (e_next & kCellIndexMask));
                        const float px = (e & kCellWrapBit) ? pxw : pi.x;
                        cell_pair<RCUT, false>(q, px, pi.y, pi.z, sx3, sy3, sz3, arow, c2, ncm, nc2, im, r2, ax, ay, az);
                        e = e_next;
                        q = q_next;
                    }
                }
            }
        }
        if (direct) {
            // ---- direct per-run loop (dense regions, the general variant): runs 0-8 = the non-wrapping cells of row
            //      (dy,dz), runs 9-17 = the cell behind the x face (edge cells only); four candidates per trip ----
            ax = ay = az = 0.f;
            for (int run = 0; run < (x_edge ? 18 : 9); ++run) {
                const bool wrap = run >= 9;
                const int row9 = wrap ? run - 9 : run;
                const int dz = row9 / 3 - 1, dy = row9 - (row9 / 3) * 3 - 1;
                int nz = cz + dz, ny = cy + dy;
                float pz = sz3[0], py = sy3[0];
                if (nz < 0) { nz += nc; pz = sz3[1]; } else if (nz >= nc) { nz -= nc; pz = sz3[2]; }
                if (ny < 0) { ny += nc; py = sy3[1]; } else if (ny >= nc) { ny -= nc; py = sy3[2]; }
                const uint32_t row = (uint32_t)((nz * nc + ny) * nc);
                const uint32_t c_lo = wrap ? (uint32_t)xw : (uint32_t)x0, c_hi = wrap ? (uint32_t)xw : (uint32_t)x1;
                uint32_t j = __ldg(cell_off + row + c_lo);
                uint32_t hi = __ldg(cell_off + row + c_hi + 1);
                if (!P3D_SLOT_END_OK(hi)) hi = j;  // (self-checking build only)
                const float px = wrap ? pxw : sx3[0];
                for (; j + 4u <= hi; j += 4u) {
                    float4 q[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) q[u] = __ldg(cpos + j + u);
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        cell_pair<RCUT, GENERAL>(q[u], px, py, pz, sx3, sy3, sz3, arow, c2, ncm, nc2, im, r2, ax, ay, az);
                }
                for (; j < hi; ++j)
                    cell_pair<RCUT, GENERAL>(__ldg(cpos + j), px, py, pz, sx3, sy3, sz3, arow, c2, ncm, nc2, im, r2, ax, ay, az);
            }
        }
        if (P3D_SLOT_OK(vals_sorted[k])) frc[vals_sorted[k]] = make_float4(ax, ay, az, 0.f);
    }
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
                                                          const uint32_t *__restrict__ cell_off, int i_begin,
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
                const uint32_t s0 = cell_off[c], s1 = P3D_SLOT_END_OK(cell_off[c + 1]) ? cell_off[c + 1] : s0;
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
    if (any && P3D_SLOT_OK(vals_sorted[k])) atomicAdd(frc + vals_sorted[k], make_float4(ax, ay, az, 0.0f));
}
