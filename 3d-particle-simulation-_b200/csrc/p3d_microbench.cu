// p3d_microbench.cu — FP32-pipe microbenchmarks.  They measure what one SM sub-partition can
// issue per clock for the instruction kinds the force kernel is made of, so that the roofline
// denominator (148 SMs x 128 lanes x clock) and the packed-FP32 assumption are checked on the
// actual device instead of taken from a data sheet.
#include <cuda_runtime.h>

#include <cstdio>

#include "p3d.h"

namespace {

__device__ __forceinline__ float rsq(float x) {
    float y;
    asm volatile("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

constexpr int kChains = 8;    // independent accumulators per thread
constexpr int kInner = 64;    // unrolled body repeats

// kind 0: scalar FFMA, 16 lane-FMAs per chain-iteration pair
__global__ void __launch_bounds__(256) mb_ffma(float *out, int iters, float a, float b) {
    float x[2 * kChains];
#pragma unroll
    for (int k = 0; k < 2 * kChains; ++k) x[k] = (float)(threadIdx.x + k);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < kInner; ++u)
#pragma unroll
            for (int k = 0; k < 2 * kChains; ++k) x[k] = fmaf(x[k], a, b);
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 2 * kChains; ++k) s += x[k];
    if (s == 123.456f) out[0] = s;
}

// kind 1: packed FFMA2
__global__ void __launch_bounds__(256) mb_ffma2(float *out, int iters, float a, float b) {
    float2 x[kChains];
    const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
#pragma unroll
    for (int k = 0; k < kChains; ++k) x[k] = make_float2((float)(threadIdx.x + k), (float)k);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < kInner; ++u)
#pragma unroll
            for (int k = 0; k < kChains; ++k) x[k] = __ffma2_rn(x[k], a2, b2);
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < kChains; ++k) s += x[k].x + x[k].y;
    if (s == 123.456f) out[0] = s;
}

// kind 2: the pair kernel's mix per pair-pack: 17 packed FP32 + 2 MUFU.RSQ + 6 FMNMX
__global__ void __launch_bounds__(128) mb_mix(float *out, int iters, float a, float b) {
    constexpr int CH = 4;
    float2 x[CH], y[CH];
    const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
#pragma unroll
    for (int k = 0; k < CH; ++k) { x[k] = make_float2((float)(threadIdx.x + k), (float)k + 1.f); y[k] = x[k]; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int k = 0; k < CH; ++k) {
                float2 d = __fadd2_rn(x[k], a2);                               // 3 FADD2
                float2 e = __fadd2_rn(y[k], a2);
                float2 f = __fadd2_rn(d, b2);
                float2 d2 = __ffma2_rn(d, d, b2);                               // 3 FFMA2
                d2 = __ffma2_rn(e, e, d2);
                d2 = __ffma2_rn(f, f, d2);
                const float2 inv = make_float2(rsq(d2.x), rsq(d2.y));          // 2 MUFU
                const float2 p1 = __ffma2_rn(inv, a2, b2), p2 = __ffma2_rn(inv, b2, a2);  // 2
                float2 ti = make_float2(fmaxf(fminf(p1.x, p2.x), 0.f), fmaxf(fminf(p1.y, p2.y), 0.f));  // 4 FMNMX
                float2 rs = __ffma2_rn(inv, a2, a2);                            // 1
                rs = make_float2(fminf(rs.x, 0.f), fminf(rs.y, 0.f));           // 2 FMNMX
                const float2 s1 = __ffma2_rn(a2, ti, rs), s2 = __ffma2_rn(b2, ti, rs);    // 2
                x[k] = __ffma2_rn(d, s1, x[k]);                                 // 6
                y[k] = __ffma2_rn(e, s1, y[k]);
                x[k] = __ffma2_rn(f, s1, x[k]);
                y[k] = __ffma2_rn(d, s2, y[k]);
                x[k] = __ffma2_rn(e, s2, x[k]);
                y[k] = __ffma2_rn(f, s2, y[k]);
            }
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < CH; ++k) s += x[k].x + x[k].y + y[k].x + y[k].y;
    if (s == 123.456f) out[0] = s;
}

// kind 3: FFMA2 with shuffles at the pair kernel's rate (12 SHFL per 4 pair-packs = 68 FFMA2)
__global__ void __launch_bounds__(128) mb_shfl(float *out, int iters, float a, float b) {
    float2 x[kChains];
    float r[12];
    const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
    const int next = (threadIdx.x + 1) & 31;
#pragma unroll
    for (int k = 0; k < kChains; ++k) x[k] = make_float2((float)(threadIdx.x + k), (float)k);
#pragma unroll
    for (int k = 0; k < 12; ++k) r[k] = (float)(threadIdx.x * 3 + k);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int v = 0; v < 8; ++v)
#pragma unroll
                for (int k = 0; k < kChains; ++k) x[k] = __ffma2_rn(x[k], a2, b2);  // 64 (+4 below)
#pragma unroll
            for (int k = 0; k < 4; ++k) x[k] = __ffma2_rn(x[k], a2, make_float2(r[k], r[k + 4]));
#pragma unroll
            for (int k = 0; k < 12; ++k) r[k] = __shfl_sync(0xffffffffu, r[k], next);
        }
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < kChains; ++k) s += x[k].x + x[k].y;
#pragma unroll
    for (int k = 0; k < 12; ++k) s += r[k];
    if (s == 123.456f) out[0] = s;
}


// Generic instruction-mix probe: per chain and body NF2 packed FFMA2, NFS scalar FFMA, NMN FMNMX,
// NMU MUFU.RSQ, NSH SHFL.  CH independent chains per thread.
template <int NF2, int NFS, int NMN, int NMU, int NSH>
__global__ void __launch_bounds__(128) mb_generic(float *out, int iters, float a, float b) {
    constexpr int CH = 4;
    float2 f2[CH];
    float g[CH], s[CH], u[CH], h[CH];
    const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
    const int next = (threadIdx.x + 1) & 31;
#pragma unroll
    for (int k = 0; k < CH; ++k) {
        f2[k] = make_float2((float)(threadIdx.x + k), (float)k);
        g[k] = (float)(threadIdx.x * 2 + k);
        s[k] = (float)(threadIdx.x * 3 + k);
        u[k] = (float)(threadIdx.x + k + 1);
        h[k] = (float)(threadIdx.x * 5 + k);
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int rep = 0; rep < 4; ++rep)
#pragma unroll
            for (int k = 0; k < CH; ++k) {
#pragma unroll
                for (int j = 0; j < NF2; ++j) f2[k] = __ffma2_rn(f2[k], a2, b2);
#pragma unroll
                for (int j = 0; j < NFS; ++j) g[k] = fmaf(g[k], a, b);
#pragma unroll
                for (int j = 0; j < NMN; ++j)
                    s[k] = (j & 1) ? fminf(s[k], s[(k + 1) % CH]) : fmaxf(s[k], s[(k + 2) % CH]);
#pragma unroll
                for (int j = 0; j < NMU; ++j) u[k] = rsq(u[k]);
#pragma unroll
                for (int j = 0; j < NSH; ++j) h[k] = __shfl_sync(0xffffffffu, h[k], next);
            }
    }
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < CH; ++k) acc += f2[k].x + f2[k].y + g[k] + s[k] + u[k] + h[k];
    if (acc == 123.456f) out[0] = acc;
}

// Register-file operand-bandwidth probes for FFMA2.  OPS: 0: d = a*b + d, three distinct register
// pairs; 1: d = a*a + d, two distinct pairs; 2: d = a*s + d with s a broadcast scalar register;
// 3: as 0 but consecutive instructions share `b` (operand-reuse cache); 4: 12 FFMA2 + 12 scalar FFMA
// on disjoint data (does the scalar FFMA find a free FMA datapath beside FFMA2?).
template <int OPS>
__global__ void __launch_bounds__(128) mb_rf(float *out, int iters, float a, float b) {
    constexpr int CH = 8;
    float2 d[CH], x[CH], y[CH];
    float g[12];
#pragma unroll
    for (int k = 0; k < CH; ++k) {
        d[k] = make_float2((float)(threadIdx.x + k), (float)k);
        x[k] = make_float2(a + 1e-3f * (threadIdx.x + k), a - 1e-3f * k);
        y[k] = make_float2(b + 1e-3f * (threadIdx.x + 2 * k), b - 2e-3f * k);
    }
#pragma unroll
    for (int k = 0; k < 12; ++k) g[k] = (float)(threadIdx.x + k);
    const float sc = a + 1e-4f * threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int rep = 0; rep < 8; ++rep) {
            if (OPS == 0) {
#pragma unroll
                for (int k = 0; k < CH; ++k) d[k] = __ffma2_rn(x[k], y[(k + rep) % CH], d[k]);
            } else if (OPS == 1) {
#pragma unroll
                for (int k = 0; k < CH; ++k) d[k] = __ffma2_rn(x[(k + rep) % CH], x[(k + rep) % CH], d[k]);
            } else if (OPS == 2) {
#pragma unroll
                for (int k = 0; k < CH; ++k) d[k] = __ffma2_rn(x[(k + rep) % CH], make_float2(sc, sc), d[k]);
            } else if (OPS == 3) {
#pragma unroll
                for (int k = 0; k < CH; ++k) d[k] = __ffma2_rn(x[k], y[rep % CH], d[k]);
            } else {
#pragma unroll
                for (int k = 0; k < CH; ++k) d[k] = __ffma2_rn(d[k], make_float2(a, a), make_float2(b, b));
#pragma unroll
                for (int k = 0; k < 4; ++k) d[k] = __ffma2_rn(d[k], make_float2(a, a), make_float2(b, b));
#pragma unroll
                for (int k = 0; k < 12; ++k) g[k] = fmaf(g[k], a, b);
            }
        }
    }
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < CH; ++k) acc += d[k].x + d[k].y;
#pragma unroll
    for (int k = 0; k < 12; ++k) acc += g[k];
    if (acc == 123.456f) out[0] = acc;
}

}  // namespace

extern "C" int p3d_microbench(int device, int kind, int iters, double out[4]) {
    if (!out || iters <= 0 || kind < 0 || kind > 17) return P3D_ERR_INVALID;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) {
        cudaGetLastError();
        return P3D_ERR_NO_DEVICE;
    }
    if (cudaSetDevice(device) != cudaSuccess) return P3D_ERR_CUDA;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return P3D_ERR_CUDA;
    int clock_khz = 0;
    cudaDeviceGetAttribute(&clock_khz, cudaDevAttrClockRate, device);
    float *d = nullptr;
    if (cudaMalloc(&d, 256) != cudaSuccess) return P3D_ERR_CUDA;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int sms = prop.multiProcessorCount;
    int threads = 256, ctas_per_sm = 4;
    double lane_fma_per_thread = 0;
    if (kind == 0) lane_fma_per_thread = (double)iters * kInner * 2 * kChains;
    if (kind == 1) lane_fma_per_thread = (double)iters * kInner * kChains * 2;
    if (kind == 2) { threads = 128; ctas_per_sm = 4; lane_fma_per_thread = (double)iters * 8 * 4 * 17 * 2; }
    if (kind == 3) { threads = 128; ctas_per_sm = 4; lane_fma_per_thread = (double)iters * 8 * 68 * 2; }
    if (kind >= 4) { threads = 128; ctas_per_sm = 4; lane_fma_per_thread = (double)iters * 16; }  // bodies per thread
    if (kind >= 13) lane_fma_per_thread = (double)iters * 8;  // reps per thread (8 FFMA2 each; kind 17: 12 FFMA2 + 12 FFMA)
    const int grid = sms * ctas_per_sm;
    for (int rep = 0; rep < 2; ++rep) {  // first launch warms up
        cudaEventRecord(e0);
        if (kind == 0) mb_ffma<<<grid, threads>>>(d, iters, 0.999f, 0.001f);
        if (kind == 1) mb_ffma2<<<grid, threads>>>(d, iters, 0.999f, 0.001f);
        if (kind == 2) mb_mix<<<grid, threads>>>(d, iters, 0.999f, 0.001f);
        if (kind == 3) mb_shfl<<<grid, threads>>>(d, iters, 0.999f, 0.001f);
        // kinds >= 4 report thread-bodies per second in out[0] (16 bodies per thread per iteration)
        if (kind == 4) mb_generic<0, 0, 6, 0, 0><<<grid, threads>>>(d, iters, 0.999f, 0.001f);    // FMNMX only
        if (kind == 5) mb_generic<0, 0, 0, 2, 0><<<grid, threads>>>(d, iters, 0.999f, 0.001f);    // MUFU only
        if (kind == 6) mb_generic<0, 0, 0, 0, 3><<<grid, threads>>>(d, iters, 0.999f, 0.001f);    // SHFL only
        if (kind == 7) mb_generic<17, 0, 6, 0, 0><<<grid, threads>>>(d, iters, 0.999f, 0.001f);   // FFMA2 + FMNMX
        if (kind == 8) mb_generic<17, 0, 0, 2, 0><<<grid, threads>>>(d, iters, 0.999f, 0.001f);   // FFMA2 + MUFU
        if (kind == 9) mb_generic<0, 34, 6, 2, 0><<<grid, threads>>>(d, iters, 0.999f, 0.001f);   // scalar mix
        if (kind == 10) mb_generic<17, 0, 6, 2, 0><<<grid, threads>>>(d, iters, 0.999f, 0.001f);  // packed mix
        if (kind == 11) mb_generic<17, 0, 6, 2, 3><<<grid, threads>>>(d, iters, 0.999f, 0.001f);  // packed mix + SHFL
        if (kind == 13) mb_rf<0><<<grid, threads>>>(d, iters, 0.999f, 0.001f);
        if (kind == 14) mb_rf<1><<<grid, threads>>>(d, iters, 0.999f, 0.001f);
        if (kind == 15) mb_rf<2><<<grid, threads>>>(d, iters, 0.999f, 0.001f);
        if (kind == 16) mb_rf<3><<<grid, threads>>>(d, iters, 0.999f, 0.001f);
        if (kind == 17) mb_rf<4><<<grid, threads>>>(d, iters, 0.999f, 0.001f);
        if (kind == 12) mb_generic<17, 0, 0, 0, 0><<<grid, threads>>>(d, iters, 0.999f, 0.001f);  // FFMA2 only
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) { cudaFree(d); return P3D_ERR_CUDA; }
    }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    out[0] = lane_fma_per_thread * (double)grid * threads / (ms * 1e-3);
    out[1] = ms;
    out[2] = sms;
    out[3] = clock_khz / 1000.0;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    return cudaGetLastError() == cudaSuccess ? P3D_OK : P3D_ERR_CUDA;
}
