// p3d_kernels_pair.cuh — the fast all-pairs force pass (K1) for sm_100a.
//
// Replaces the spatial-hash neighbour walk of src/lib.rs:135-243 with an evaluation of ALL
// particle pairs.  Design:
//   * Slots are grouped in blocks of B = 32*R particles of ONE type.  Every step a partition pass
//     stages each type's particles as [interior ... ghosts ... boundary]; a block is INTERIOR when
//     every member is farther than reach = min(r, max(1, m)) from every face, so that only the offset-0 image
//     of src/lib.rs:177-185 can be in range for any pair that involves it.
//   * k_force_pair visits every unordered block pair {a,b} with at least one INTERIOR member once
//     (circulant schedule: row a takes b = a+o, o = 0..M/2).  Inside a pair the relative position,
//     distance, rsqrt and the type-independent part of the force law are shared by both directions
//     (the attraction matrix is asymmetric, src/bin/main.rs:133-139, so only the scalar differs).
//     A warp keeps R i-particles per lane in registers; 64 j-particles per round travel around
//     the warp with shuffles together with their force accumulators, so neither side ever needs a
//     cross-lane reduction.  All arithmetic is packed FP32x2 (FADD2/FFMA2): two j per instruction.
//   * BOUNDARY x BOUNDARY block pairs (the only ones that can interact through a periodic image)
//     go to k_force_bxb, which reproduces the reference's image arithmetic exactly.
#pragma once
#include "p3d_device.cuh"
#include "p3d_kernels_basic.cuh"

enum : uint8_t { P3D_BLK_INTERIOR = 0, P3D_BLK_BOUNDARY = 1, P3D_BLK_EMPTY = 2 };

// ---------------------------------------------------------------------------------------------
// Partition pass.  Every step each type's live slots are staged as [interior ... ghosts ... boundary]
// in slot order (a STABLE partition: count -> scan -> scatter, no atomics), so that every rank of a
// multi-GPU run derives bit-identical block lists from the same positions.
// CTAs of 128 threads never straddle two types (type segments are multiples of B >= 128).
constexpr int kPartThreads = 128;

__device__ __forceinline__ bool part_is_interior(const float4 p, float interior_limit) {
    return fmaxf(fabsf(p.x), fmaxf(fabsf(p.y), fabsf(p.z))) < interior_limit;
}

// part 1a: per-CTA counts of interior / boundary particles.
__global__ void __launch_bounds__(kPartThreads) k_part_count(const float4 *__restrict__ pos, int n_slots,
                                                             float interior_limit, int2 *__restrict__ cta_cnt,
                                                             int *__restrict__ flag_to_clear) {
    const int s = blockIdx.x * kPartThreads + threadIdx.x;
    if (s == 0) *flag_to_clear = 0;
    bool interior = false, boundary = false;
    if (s < n_slots) {
        const float4 p = pos[s];
        const bool live = f2u(p.w) != P3D_GHOST_ID;
        interior = live && part_is_interior(p, interior_limit);
        boundary = live && !interior;
    }
    const int ni = __syncthreads_count(interior);
    const int nb = __syncthreads_count(boundary);
    if (threadIdx.x == 0) cta_cnt[blockIdx.x] = make_int2(ni, nb);
}

// part 1b: one CTA per type scans the per-CTA counts of that type's segment (exclusive prefix),
// and records the type totals in cnt[2t], cnt[2t+1].
__global__ void __launch_bounds__(1024) k_part_scan(const int2 *__restrict__ cta_cnt, int2 *__restrict__ cta_off,
                                                    const int *__restrict__ seg_start,
                                                    const int *__restrict__ seg_end, int *__restrict__ cnt) {
    const int t = blockIdx.x;
    const int c0 = seg_start[t] / kPartThreads, c1 = seg_end[t] / kPartThreads;
    __shared__ int2 warp_tot[32];
    __shared__ int2 carry;
    if (threadIdx.x == 0) carry = make_int2(0, 0);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int base = c0; base < c1; base += 1024) {
        const int c = base + threadIdx.x;
        int2 v = (c < c1) ? cta_cnt[c] : make_int2(0, 0);
        int2 inc = v;  // inclusive scan inside the warp
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int ax = __shfl_up_sync(0xffffffffu, inc.x, o), ay = __shfl_up_sync(0xffffffffu, inc.y, o);
            if (lane >= o) { inc.x += ax; inc.y += ay; }
        }
        if (lane == 31) warp_tot[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            int2 w = warp_tot[lane];
            int2 winc = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int ax = __shfl_up_sync(0xffffffffu, winc.x, o), ay = __shfl_up_sync(0xffffffffu, winc.y, o);
                if (lane >= o) { winc.x += ax; winc.y += ay; }
            }
            warp_tot[lane] = make_int2(winc.x - w.x, winc.y - w.y);  // exclusive prefix of the warp totals
        }
        __syncthreads();
        const int2 wbase = warp_tot[warp];
        const int2 cbase = carry;
        if (c < c1) cta_off[c] = make_int2(cbase.x + wbase.x + inc.x - v.x, cbase.y + wbase.y + inc.y - v.y);
        __syncthreads();
        if (threadIdx.x == 1023) carry = make_int2(cbase.x + wbase.x + inc.x, cbase.y + wbase.y + inc.y);
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        cnt[2 * t] = carry.x;
        cnt[2 * t + 1] = carry.y;
    }
}

// part 1c: scatter.  Interior particles are packed upward from the region start, boundary particles
// downward from its end, both in slot order.
__global__ void __launch_bounds__(kPartThreads) k_part_scatter(
    const float4 *__restrict__ pos, int n_slots, int B, const uint8_t *__restrict__ seg_type,
    const int *__restrict__ seg_start, const int *__restrict__ seg_end, const int2 *__restrict__ cta_off,
    float4 *__restrict__ spos, float *__restrict__ sx, float *__restrict__ sy, float *__restrict__ sz,
    uint32_t *__restrict__ sidx, float interior_limit) {
    const int s = blockIdx.x * kPartThreads + threadIdx.x;
    bool interior = false, boundary = false;
    float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
    if (s < n_slots) {
        p = pos[s];
        const bool live = f2u(p.w) != P3D_GHOST_ID;
        interior = live && part_is_interior(p, interior_limit);
        boundary = live && !interior;
    }
    const unsigned mi = __ballot_sync(0xffffffffu, interior);
    const unsigned mb = __ballot_sync(0xffffffffu, boundary);
    __shared__ int2 wcnt[kPartThreads / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) wcnt[warp] = make_int2(__popc(mi), __popc(mb));
    __syncthreads();
    int wi = 0, wb = 0;
    for (int w = 0; w < warp; ++w) { wi += wcnt[w].x; wb += wcnt[w].y; }
    if (s >= n_slots) return;
    const int t = seg_type[(blockIdx.x * kPartThreads) / B];
    const int2 off = cta_off[blockIdx.x];
    const unsigned lt = (1u << lane) - 1u;
    int dst = -1;
    if (interior) dst = seg_start[t] + off.x + wi + __popc(mi & lt);
    else if (boundary) dst = seg_end[t] - 1 - (off.y + wb + __popc(mb & lt));
    if (dst >= 0 && P3D_SLOT_OK(dst)) {
        spos[dst] = p;
        sx[dst] = p.x; sy[dst] = p.y; sz[dst] = p.z;
        sidx[dst] = (uint32_t)s;
    }
}

// Partition pass, part 2: ghosts into the gap between the two lists, and the class of each block.
__global__ void __launch_bounds__(256) k_part_fill(int n_slots, int B, const uint8_t *__restrict__ seg_type,
                                                   const int *__restrict__ seg_start,
                                                   const int *__restrict__ seg_end, const int *__restrict__ cnt,
                                                   float4 *__restrict__ spos, float *__restrict__ sx,
                                                   float *__restrict__ sy, float *__restrict__ sz,
                                                   uint32_t *__restrict__ sidx, uint8_t *__restrict__ bclass) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n_slots) return;
    const int blk = q / B;
    const int t = seg_type[blk];
    const int lo = seg_start[t] + cnt[2 * t];    // first staged index after the interior list
    const int hi = seg_end[t] - cnt[2 * t + 1];  // first staged index of the boundary list
    if (q >= lo && q < hi) {
        spos[q] = make_float4(P3D_GHOST_COORD, P3D_GHOST_COORD, P3D_GHOST_COORD, u2f(P3D_GHOST_ID));
        sx[q] = P3D_GHOST_COORD; sy[q] = P3D_GHOST_COORD; sz[q] = P3D_GHOST_COORD;
        sidx[q] = P3D_GHOST_ID;
    }
    if (q == blk * B) {
        const int b0 = q, b1 = q + B;
        // EMPTY: ghosts only.  INTERIOR: interior entries (and possibly trailing ghosts) but no
        // boundary entry.  BOUNDARY: holds at least one boundary entry (k_force_bxb walks exactly
        // the blocks from hi / B on, i.e. those with b1 > hi).
        uint8_t c;
        if (b0 >= lo && b1 <= hi) c = P3D_BLK_EMPTY;
        else if (b1 <= hi) c = P3D_BLK_INTERIOR;
        else c = P3D_BLK_BOUNDARY;
        bclass[blk] = c;
    }
}

// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float rsqrt_approx(float x) {
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float2 dup2(float x) { return make_float2(x, x); }
__device__ __forceinline__ float2 shfl2(float2 v, int src) {
    v.x = __shfl_sync(0xffffffffu, v.x, src);
    v.y = __shfl_sync(0xffffffffu, v.y, src);
    return v;
}

// Constants of the branch-free force law, each duplicated into both halves of a packed register.
// With inv = 1/d:   f(d)/d = min(1/m - inv, 0) + a * max(0, min(c2 - c2*m*inv, c2*inv - c2))
// which equals src/lib.rs:55-67 divided by d for every d in (0, r) when r >= 1 (the triangle
// 1 - |2d-1-m|/(1-m) is min(2(d-m), 2(1-d))/(1-m), and dividing by d > 0 commutes with min/max).
struct PairConsts {
    float2 c2;     // 2/(1-m)
    float2 ncm;    // -c2*m
    float2 nc2;    // -c2
    float2 im;     // 1/m (or +inf)
    float2 nim;    // -1/m
    float2 neg1;   // -1
    float2 tiny;   // 1e-30: keeps rsqrt finite at d2 == 0 (self pair, coincident particles)
    float c2m;     // c2*m: folded into the matrix scalars when m > 0 (see pair_group, MPOS)
};

__device__ __forceinline__ PairConsts make_pair_consts(const DevParams &P) {
    PairConsts c;
    c.c2 = dup2(P.c2);
    c.ncm = dup2(-P.c2 * P.m);
    c.nc2 = dup2(-P.c2);
    c.im = dup2(P.inv_m);
    c.nim = dup2(-P.inv_m);
    c.neg1 = dup2(-1.0f);
    c.c2m = P.c2 * P.m;
    c.tiny = dup2(1.0e-30f);
    return c;
}

// G i-particles against the lane's two j-particles (one packed register pair per coordinate), evaluated
// STAGE BY STAGE across the G particles: all relative positions, all squared distances, all rsqrt, all
// force-law scalars, then the accumulations (i-side particle by particle, then the three dependent j-side
// chains).  Same arithmetic per pair as a particle-by-particle loop; the staged order is what lets ptxas
// issue the j-side chains and their shuffles early and fill the shuffle latency with the long run of
// independent i-side FFMA2 (6 % faster than the per-particle order on B200).
// The i-position enters as a broadcast scalar operand of FADD2 (SASS: `-R.F32`), so it needs no
// duplication.  Per i-particle: 16 packed FP32 instructions (17 without MPOS) + 2 MUFU.RSQ + 6 FMNMX for
// four ordered interactions.
// Source-order knobs of pair_group (tools/pair_search.py explores them on the device: same arithmetic, other
// operand slots / instruction order in the source, which is all the influence the source has on ptxas' register
// allocation and operand reuse).  The defaults are the shipped variant.
#ifndef P3D_PG_FADD_SWAP
#define P3D_PG_FADD_SWAP 0   // 1: i-position first in the FADD2
#endif
#ifndef P3D_PG_D2_ORDER
#define P3D_PG_D2_ORDER 0    // component order of the d^2 chain: 0 xyz 1 xzy 2 yxz 3 yzx 4 zxy 5 zyx (changes rounding)
#endif
#ifndef P3D_PG_LAW_SWAP
#define P3D_PG_LAW_SWAP 0    // bit 0: constant first in rs, bit 1: constant first in p2
#endif
#ifndef P3D_PG_RS_FADD
#define P3D_PG_RS_FADD 0     // 1: rs = im - inv as FADD2 instead of FFMA2 with -1
#endif
#ifndef P3D_PG_ACCI_SWAP
#define P3D_PG_ACCI_SWAP 0   // 1: relative position first in the i-side accumulations
#endif
#ifndef P3D_PG_ACCJ_SWAP
#define P3D_PG_ACCJ_SWAP 0
#endif
#ifndef P3D_PG_ACCI_ORDER
#define P3D_PG_ACCI_ORDER 0  // component order of the i-side accumulations (as P3D_PG_D2_ORDER; same bits)
#endif
#ifndef P3D_PG_ACCJ_ORDER
#define P3D_PG_ACCJ_ORDER 0
#endif
#ifndef P3D_PG_STAGE
#define P3D_PG_STAGE 0       // 0: all i-side accumulations, then the j-side chains; 1: j-side first; 2: per particle i then j
#endif
#ifndef P3D_PG_S_SPLIT
#define P3D_PG_S_SPLIT 0     // 1: all sij, then all sji
#endif
#ifndef P3D_PG_IMM
#define P3D_PG_IMM 0         // 1: the law's constants as compile-time immediates (default scene: m = 0.3) - probe only
#endif

template <int K>
__device__ __forceinline__ void pg_perm(int &a, int &b, int &c) {
    constexpr int P[6][3] = {{0, 1, 2}, {0, 2, 1}, {1, 0, 2}, {1, 2, 0}, {2, 0, 1}, {2, 1, 0}};
    a = P[K][0]; b = P[K][1]; c = P[K][2];
}
#define P3D_FMA2(SWAP, s, d, acc) ((SWAP) ? __ffma2_rn((d), (s), (acc)) : __ffma2_rn((s), (d), (acc)))

template <int G, bool RCUT, bool MPOS>
__device__ __forceinline__ void pair_group(const float2 jx, const float2 jy, const float2 jz, const float *nix,
                                           const float *niy, const float *niz, const PairConsts &c,
                                           const float2 aij, const float2 aji, const float r2, float2 *aix,
                                           float2 *aiy, float2 *aiz, float2 &ajx, float2 &ajy, float2 &ajz) {
    float2 d[3][G], sij[G], sji[G];
    float2 d2[G], inv[G], rs[G], ti[G];
#if P3D_PG_IMM
    const float2 k_im = dup2(1.0f / 0.3f), k_nim = dup2(-1.0f / 0.3f), k_neg1 = dup2(-1.0f), k_tiny = dup2(1.0e-30f);
#else
    const float2 k_im = c.im, k_nim = c.nim, k_neg1 = c.neg1, k_tiny = c.tiny;
#endif
#pragma unroll
    for (int g = 0; g < G; ++g) {
        // other.position - position (src/lib.rs:211-212, offset 0)
        d[0][g] = P3D_PG_FADD_SWAP ? __fadd2_rn(dup2(nix[g]), jx) : __fadd2_rn(jx, dup2(nix[g]));
        d[1][g] = P3D_PG_FADD_SWAP ? __fadd2_rn(dup2(niy[g]), jy) : __fadd2_rn(jy, dup2(niy[g]));
        d[2][g] = P3D_PG_FADD_SWAP ? __fadd2_rn(dup2(niz[g]), jz) : __fadd2_rn(jz, dup2(niz[g]));
    }
    {
        int a, b, e;
        pg_perm<P3D_PG_D2_ORDER>(a, b, e);
#pragma unroll
        for (int g = 0; g < G; ++g) {
            d2[g] = __ffma2_rn(d[a][g], d[a][g], k_tiny);
            d2[g] = __ffma2_rn(d[b][g], d[b][g], d2[g]);
            d2[g] = __ffma2_rn(d[e][g], d[e][g], d2[g]);
        }
    }
#pragma unroll
    for (int g = 0; g < G; ++g) inv[g] = make_float2(rsqrt_approx(d2[g].x), rsqrt_approx(d2[g].y));
#pragma unroll
    for (int g = 0; g < G; ++g) {
        // u = 1/m - 1/d
        if (P3D_PG_RS_FADD) rs[g] = __fadd2_rn(k_im, make_float2(-inv[g].x, -inv[g].y));
        else rs[g] = (P3D_PG_LAW_SWAP & 1) ? __ffma2_rn(k_neg1, inv[g], k_im) : __ffma2_rn(inv[g], k_neg1, k_im);
        if (MPOS) {
            // (1/d - 1) / m
            const float2 p2 = (P3D_PG_LAW_SWAP & 2) ? __ffma2_rn(k_im, inv[g], k_nim) : __ffma2_rn(inv[g], k_im, k_nim);
            ti[g] = make_float2(fmaxf(fminf(rs[g].x, p2.x), 0.0f), fmaxf(fminf(rs[g].y, p2.y), 0.0f));
        } else {
            const float2 p1 = __ffma2_rn(inv[g], c.ncm, c.c2);   // c2 * (1 - m/d)
            const float2 p2 = __ffma2_rn(inv[g], c.c2, c.nc2);   // c2 * (1/d - 1)
            ti[g] = make_float2(fmaxf(fminf(p1.x, p2.x), 0.0f), fmaxf(fminf(p1.y, p2.y), 0.0f));
        }
        rs[g] = make_float2(fminf(rs[g].x, 0.0f), fminf(rs[g].y, 0.0f));
        if (RCUT) {  // r < max(1, m): src/lib.rs:216-220 cuts inside the force range
            if (!(d2[g].x < r2)) { ti[g].x = 0.0f; rs[g].x = 0.0f; }
            if (!(d2[g].y < r2)) { ti[g].y = 0.0f; rs[g].y = 0.0f; }
        }
    }
    if (P3D_PG_S_SPLIT) {
#pragma unroll
        for (int g = 0; g < G; ++g) sij[g] = __ffma2_rn(aij, ti[g], rs[g]);   // f(d; A[i][j]) / d
#pragma unroll
        for (int g = 0; g < G; ++g) sji[g] = __ffma2_rn(aji, ti[g], rs[g]);   // f(d; A[j][i]) / d
    } else {
#pragma unroll
        for (int g = 0; g < G; ++g) {
            sij[g] = __ffma2_rn(aij, ti[g], rs[g]);
            sji[g] = __ffma2_rn(aji, ti[g], rs[g]);
        }
    }
    float2 *ai[3] = {aix, aiy, aiz};
    float2 *aj[3] = {&ajx, &ajy, &ajz};
    int ia, ib, ic, ja, jb, jc;
    pg_perm<P3D_PG_ACCI_ORDER>(ia, ib, ic);
    pg_perm<P3D_PG_ACCJ_ORDER>(ja, jb, jc);
    // acc += rel / d * f (src/lib.rs:231); the j side is negated when flushed: rel_ji = -rel_ij.
    // (scalar pair first by default: the product commutes, the operand slots ptxas fills do not — 1.6 % on the B200)
#define P3D_PG_ISIDE(g)                                                          \
    do {                                                                         \
        ai[ia][g] = P3D_FMA2(P3D_PG_ACCI_SWAP, sij[g], d[ia][g], ai[ia][g]);     \
        ai[ib][g] = P3D_FMA2(P3D_PG_ACCI_SWAP, sij[g], d[ib][g], ai[ib][g]);     \
        ai[ic][g] = P3D_FMA2(P3D_PG_ACCI_SWAP, sij[g], d[ic][g], ai[ic][g]);     \
    } while (0)
#define P3D_PG_JSIDE(g)                                                          \
    do {                                                                         \
        *aj[ja] = P3D_FMA2(P3D_PG_ACCJ_SWAP, sji[g], d[ja][g], *aj[ja]);         \
        *aj[jb] = P3D_FMA2(P3D_PG_ACCJ_SWAP, sji[g], d[jb][g], *aj[jb]);         \
        *aj[jc] = P3D_FMA2(P3D_PG_ACCJ_SWAP, sji[g], d[jc][g], *aj[jc]);         \
    } while (0)
    if (P3D_PG_STAGE == 0) {
#pragma unroll
        for (int g = 0; g < G; ++g) P3D_PG_ISIDE(g);
#pragma unroll
        for (int g = 0; g < G; ++g) P3D_PG_JSIDE(g);
    } else if (P3D_PG_STAGE == 1) {
#pragma unroll
        for (int g = 0; g < G; ++g) P3D_PG_JSIDE(g);
#pragma unroll
        for (int g = 0; g < G; ++g) P3D_PG_ISIDE(g);
    } else {
#pragma unroll
        for (int g = 0; g < G; ++g) {
            P3D_PG_ISIDE(g);
            P3D_PG_JSIDE(g);
        }
    }
#undef P3D_PG_ISIDE
#undef P3D_PG_JSIDE
}

__device__ __forceinline__ void atomic_add_f3(float4 *dst, float x, float y, float z) {
    // sm_90+ vector atomic: one 16-byte RED per particle
    atomicAdd(dst, make_float4(x, y, z, 0.0f));
}

// K1 (fast): symmetric block-pair force pass.  grid.x = n_rows * splits, block = NW warps.
// Row a (row_begin + k*row_stride: the multi-GPU shard takes every world-th row) owns the block
// pairs {a, a+o mod M}, o in [0, M/2]; the warps of the `splits` CTAs of a row interleave over o.
template <int R, bool RCUT, int MINB, int NW, bool MPOS>
__global__ void __launch_bounds__(32 * NW, MINB)
k_force_pair(const float *__restrict__ sx, const float *__restrict__ sy, const float *__restrict__ sz,
             const uint32_t *__restrict__ sidx,
             const uint8_t *__restrict__ bclass, const uint8_t *__restrict__ btype, int M, int row_begin,
             int row_stride, int splits, float4 *__restrict__ frc, DevParams P,
             const float *__restrict__ matrix, const int *__restrict__ flags) {
    if (flags[0] != 0) return;  // some particle is outside the box: the reference-order kernel runs instead
    constexpr int B = 32 * R;
    constexpr int ROUNDS = B / 64;
    const int lane = threadIdx.x & 31;
    // NW == 1: one warp per CTA makes every per-block-pair quantity (b, types, matrix entries)
    // provably warp-uniform for the compiler, so it can live in uniform registers
    const int warp = (NW == 1) ? 0 : (threadIdx.x >> 5);
    constexpr int nw = NW;
    const int row = row_begin + (blockIdx.x / splits) * row_stride;
    const int split = blockIdx.x % splits;
    if (row >= M) return;
    const uint8_t ca = bclass[row];
    if (ca == P3D_BLK_EMPTY) return;
    const int ta = btype[row];
    const PairConsts c = make_pair_consts(P);

    float nix[R], niy[R], niz[R];
    float2 aix[R], aiy[R], aiz[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int s = row * B + r * 32 + lane;
        nix[r] = -sx[s]; niy[r] = -sy[s]; niz[r] = -sz[s];
        aix[r] = aiy[r] = aiz[r] = make_float2(0.f, 0.f);
    }

    const int omax = M / 2;  // M even: offset M/2 is shared by rows a and a+M/2; the lower row takes it
    const int next = (lane + 1) & 31;
    for (int o = split * nw + warp; o <= omax; o += splits * nw) {
        if (((M & 1) == 0) && o == omax && row >= omax && o != 0) continue;
        int b = row + o;
        if (b >= M) b -= M;
        const uint8_t cb = bclass[b];
        if (cb == P3D_BLK_EMPTY) continue;
        if (ca == P3D_BLK_BOUNDARY && cb == P3D_BLK_BOUNDARY) continue;  // -> k_force_bxb
        const int tb = btype[b];
        const float ascale = MPOS ? c.c2m : 1.0f;
        const float2 aij = dup2(matrix[ta * P.T + tb] * ascale);
        const float2 aji = dup2(matrix[tb * P.T + ta] * ascale);
        const bool diag = (o == 0);
#pragma unroll 1
        for (int round = 0; round < ROUNDS; ++round) {
            const int base = b * B + round * 64 + 2 * lane;  // two consecutive j per lane: native register pairs
            float2 jx = *reinterpret_cast<const float2 *>(sx + base);
            float2 jy = *reinterpret_cast<const float2 *>(sy + base);
            float2 jz = *reinterpret_cast<const float2 *>(sz + base);
            float2 ajx = make_float2(0.f, 0.f), ajy = ajx, ajz = ajx;
#pragma unroll 1
            for (int step = 0; step < 32; ++step) {
                pair_group<R, RCUT, MPOS>(jx, jy, jz, nix, niy, niz, c, aij, aji, P.r2, aix, aiy, aiz, ajx, ajy, ajz);
                jx = shfl2(jx, next); jy = shfl2(jy, next); jz = shfl2(jz, next);
                ajx = shfl2(ajx, next); ajy = shfl2(ajy, next); ajz = shfl2(ajz, next);
            }
            // 32 rotations bring every j (and its accumulator) back to its home lane
            if (!diag) {
                const uint32_t s0 = sidx[base], s1 = sidx[base + 1];
                if (s0 != P3D_GHOST_ID && P3D_SLOT_OK(s0)) atomic_add_f3(frc + s0, -ajx.x, -ajy.x, -ajz.x);
                if (s1 != P3D_GHOST_ID && P3D_SLOT_OK(s1)) atomic_add_f3(frc + s1, -ajx.y, -ajy.y, -ajz.y);
            }
        }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const uint32_t s = sidx[row * B + r * 32 + lane];
        if (s != P3D_GHOST_ID && P3D_SLOT_OK(s))
            atomic_add_f3(frc + s, aix[r].x + aix[r].y, aiy[r].x + aiy[r].y, aiz[r].x + aiz[r].y);
    }
}

// Nearest of the two images of the i-particle that can be in range of an in-box q: offset 0 and
// offset -sign(p)*W.  `pa` is the reference's rounded `position + offset` (src/lib.rs:190-192,211-212).
__device__ __forceinline__ float nearest_image2(float q, float p0, float pa) {
    const float r0 = __fsub_rn(q, p0), ra = __fsub_rn(q, pa);
    return (fabsf(ra) < fabsf(r0)) ? ra : r0;
}

// K1 (boundary x boundary): ordered pairs between BOUNDARY-class blocks — the only pairs that can
// interact through a periodic image.  One thread per i-particle, j-tiles of boundary blocks staged in
// shared memory.  Relative positions follow the reference's image arithmetic exactly
// (`other.position - (position + offset)` with the f32 rounding of position + offset); the force law
// is the same branch-free rsqrt form as k_force_pair.  All particles are inside the box here (the
// flag check), so per axis only offsets 0 and -sign(p)*W can be within reach.
// Two passes per four j: the distance through the nearest image only needs MAGNITUDES
// (min(|q - p|, |q - (p + offset)|) per axis: two FADD and one FMNMX), so the signed image selection, the
// rsqrt, the matrix lookup and the accumulation run only when some lane of the warp has one of its four
// pairs within reach (the reference's own `d2 < r^2` test, src/lib.rs:216-220; beyond reach the force is
// exactly zero) — about 7 % of the warp iterations at density 1.
// grid = (rows of this shard, jsplit): CTA (x, y) takes every jsplit-th boundary block.
template <int B, bool RCUT>
__global__ void __launch_bounds__(B) k_force_bxb(const float4 *__restrict__ spos, const uint32_t *__restrict__ sidx,
                                                 const uint8_t *__restrict__ bclass, int M, int row_begin,
                                                 int row_stride, const int *__restrict__ seg_start,
                                                 const int *__restrict__ seg_end, const int *__restrict__ cnt,
                                                 float4 *__restrict__ frc, DevParams P,
                                                 const float *__restrict__ matrix, const int *__restrict__ flags) {
    if (flags[0] != 0) return;
    const int row = row_begin + blockIdx.x * row_stride;
    if (row >= M || bclass[row] != P3D_BLK_BOUNDARY) return;
    extern __shared__ float4 sm_dyn[];
    float4 *tile = sm_dyn;
    float *smat = reinterpret_cast<float *>(sm_dyn + B);
    for (int k = threadIdx.x; k < P.T * P.T; k += B) smat[k] = matrix[k];

    // ghost i-slots (1e15 away from everything) walk the loop too, so that warp votes see all 32 lanes
    const float4 pi = spos[row * B + threadIdx.x];
    const uint32_t idi = f2u(pi.w);
    const bool live = idi != P3D_GHOST_ID;
    // position + offset for the one non-zero offset per axis that can matter
    const float pxa = __fadd_rn(pi.x, pi.x > 0.f ? -P.W : P.W);
    const float pya = __fadd_rn(pi.y, pi.y > 0.f ? -P.W : P.W);
    const float pza = __fadd_rn(pi.z, pi.z > 0.f ? -P.W : P.W);
    const uint32_t mrow = live ? idi * (uint32_t)P.T : 0u;
    const float c2 = P.c2, ncm = -P.c2 * P.m, nc2 = -P.c2, im = P.inv_m, r2 = P.r2;
    const float reach2 = P.reach * P.reach * 1.000001f;  // a hair wide: pairs at the very edge take the exact path
    float ax = 0.f, ay = 0.f, az = 0.f;
    int visit = 0;
    for (int t = 0; t < P.T; ++t) {
        const int hi = seg_end[t] - cnt[2 * t + 1];  // first boundary entry of type t
        const int b_first = hi / B;                  // the block holding it
        const int b_last = seg_end[t] / B;
        for (int b = b_first; b < b_last; ++b) {
            if (bclass[b] != P3D_BLK_BOUNDARY) continue;
            if ((visit++ % (int)gridDim.y) != (int)blockIdx.y) continue;
            __syncthreads();
            tile[threadIdx.x] = spos[b * B + threadIdx.x];
            __syncthreads();
            const float *arow = smat + mrow;
#pragma unroll 2
            for (int k0 = 0; k0 < B; k0 += 4) {
                bool near = false;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float4 q = tile[k0 + k];
                    const float mx = fminf(fabsf(__fsub_rn(q.x, pi.x)), fabsf(__fsub_rn(q.x, pxa)));
                    const float my = fminf(fabsf(__fsub_rn(q.y, pi.y)), fabsf(__fsub_rn(q.y, pya)));
                    const float mz = fminf(fabsf(__fsub_rn(q.z, pi.z)), fabsf(__fsub_rn(q.z, pza)));
                    near |= fmaf(mz, mz, fmaf(my, my, mx * mx)) < reach2;
                }
                if (!__any_sync(0xffffffffu, near)) continue;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float4 q = tile[k0 + k];
                    const float rx = nearest_image2(q.x, pi.x, pxa);
                    const float ry = nearest_image2(q.y, pi.y, pya);
                    const float rz = nearest_image2(q.z, pi.z, pza);
                    const float d2 = fmaf(rz, rz, fmaf(ry, ry, fmaf(rx, rx, 1.0e-30f)));
                    const float inv = rsqrt_approx(d2);
                    const float p1 = fmaf(inv, ncm, c2), p2 = fmaf(inv, c2, nc2);
                    float ti = fmaxf(fminf(p1, p2), 0.0f);
                    float rs = fminf(im - inv, 0.0f);
                    if (RCUT) {
                        if (!(d2 < r2)) { ti = 0.0f; rs = 0.0f; }
                    }
                    // ghosts (type id 0xFFFFFFFF) sit 1e15 away: ti = rs = 0, so any matrix entry will do
                    const uint32_t tj = f2u(q.w);
                    const float a = arow[tj < (uint32_t)P.T ? tj : 0u];
                    const float s = fmaf(a, ti, rs);
                    ax = fmaf(rx, s, ax);
                    ay = fmaf(ry, s, ay);
                    az = fmaf(rz, s, az);
                }
            }
        }
    }
    if (live && P3D_SLOT_OK(sidx[row * B + threadIdx.x])) atomic_add_f3(frc + sidx[row * B + threadIdx.x], ax, ay, az);
}
