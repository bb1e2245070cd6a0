// pair_probe.cu — times ONE build of k_force_pair (the knobs of pair_group are -D macros) on synthetic type-pure
// blocks, without the engine around it.  MEASUREMENT TOOL for tools/pair_search/search.py; not part of the product.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I<csrc> -DP3D_PG_...=... -DPROBE_R=8 -DPROBE_MINB=12 \
//        pair_probe.cu -o probe && ./probe [n=262144] [reps=4]
// Prints: ms (best of reps), a checksum of the forces (all variants of the same arithmetic must agree to rounding).
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "p3d_kernels_pair.cuh"

#ifndef PROBE_R
#define PROBE_R 8
#endif
#ifndef PROBE_MINB
#define PROBE_MINB 12
#endif

#define CK(x)                                                                     \
    do {                                                                          \
        cudaError_t e_ = (x);                                                     \
        if (e_ != cudaSuccess) {                                                  \
            fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_));              \
            return 2;                                                             \
        }                                                                         \
    } while (0)

int main(int argc, char **argv) {
    const int n = argc > 1 ? atoi(argv[1]) : 262144;
    const int reps = argc > 2 ? atoi(argv[2]) : 4;
    constexpr int R = PROBE_R, B = 32 * R;
    const int M = n / B, ns = M * B;
    const float W = cbrtf((float)ns);  // density 1
    std::vector<float> hx(ns), hy(ns), hz(ns);
    std::vector<uint32_t> hidx(ns);
    std::vector<uint8_t> hclass(M, P3D_BLK_INTERIOR), htype(M);
    uint64_t s = 42;
    auto rnd = [&]() {
        s += 0x9E3779B97F4A7C15ULL;
        uint64_t z = s;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
        z ^= z >> 31;
        return (float)(z >> 40) * (1.0f / 16777216.0f);
    };
    for (int i = 0; i < ns; ++i) {
        hx[i] = (rnd() - 0.5f) * W; hy[i] = (rnd() - 0.5f) * W; hz[i] = (rnd() - 0.5f) * W;
        hidx[i] = (uint32_t)i;
    }
    for (int b = 0; b < M; ++b) htype[b] = (uint8_t)(b * 5 / M);  // five type segments, like the engine's layout
    const float mat[25] = {0.5f, 1.0f, -0.5f, 0.0f, -1.0f, 1.0f, 1.0f, 1.0f, 0.0f, -1.0f, 0.0f, 0.0f, 0.5f,
                           1.5f, -1.0f, 0.0f, 0.0f, 0.0f, 0.0f, -1.0f, 1.0f, 1.0f, 1.0f, 1.0f, 0.5f};
    DevParams P{};
    P.W = W; P.half = W / 2; P.r = 2.0f; P.r2 = 4.0f; P.m = 0.3f; P.kf = 1.0f; P.coef = 0.97f; P.T = 5;
    P.inv_m = 1.0f / 0.3f; P.c2 = 2.0f / 0.7f; P.rcut = 0; P.reach = 1.0f;
    float *sx, *sy, *sz, *dmat;
    uint32_t *sidx;
    uint8_t *bclass, *btype;
    float4 *frc;
    int *flags;
    CK(cudaMalloc(&sx, ns * 4)); CK(cudaMalloc(&sy, ns * 4)); CK(cudaMalloc(&sz, ns * 4));
    CK(cudaMalloc(&sidx, ns * 4)); CK(cudaMalloc(&bclass, M)); CK(cudaMalloc(&btype, M));
    CK(cudaMalloc(&frc, (size_t)ns * 16)); CK(cudaMalloc(&dmat, 100)); CK(cudaMalloc(&flags, 16));
    CK(cudaMemcpy(sx, hx.data(), ns * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(sy, hy.data(), ns * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(sz, hz.data(), ns * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(sidx, hidx.data(), ns * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(bclass, hclass.data(), M, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(btype, htype.data(), M, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dmat, mat, 100, cudaMemcpyHostToDevice));
    CK(cudaMemset(flags, 0, 16));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int rows = M, offsets = M / 2 + 1;
    const long long want = (128LL * 16 * prop.multiProcessorCount + rows - 1) / rows;
    const int splits = (int)(want < 1 ? 1 : (want > offsets ? offsets : want));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int r = 0; r <= reps; ++r) {
        CK(cudaMemset(frc, 0, (size_t)ns * 16));
        CK(cudaEventRecord(e0));
        k_force_pair<R, false, PROBE_MINB, 1, true><<<rows * splits, 32>>>(sx, sy, sz, sidx, bclass, btype, M, 0, 1, splits,
                                                                        frc, P, dmat, flags);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaGetLastError());
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (r > 0 && ms < best) best = ms;
    }
    std::vector<float4> hf(ns);
    CK(cudaMemcpy(hf.data(), frc, (size_t)ns * 16, cudaMemcpyDeviceToHost));
    double sum = 0, l1 = 0;
    for (int i = 0; i < ns; ++i) { sum += hf[i].x + 2.0 * hf[i].y + 3.0 * hf[i].z; l1 += fabs(hf[i].x) + fabs(hf[i].y) + fabs(hf[i].z); }
    const double pairs = (double)ns * ns;
    printf("ms %.4f  Tint/s %.4f  frac20 %.4f  checksum %.6e  l1 %.8e  n %d R %d minb %d\n", best, pairs / best / 1e9,
           pairs * 20 / (best * 1e-3) / (prop.multiProcessorCount * 128 * 2 * 1.965e9), sum, l1, ns, R, PROBE_MINB);
    return 0;
}
