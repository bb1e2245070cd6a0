/* p3d_microbench.h — FP32-pipe microbenchmarks (libp3d_microbench.so).  MEASUREMENT INFRASTRUCTURE, not part of
 * the product library: it makes the roofline denominator of bench.py defensible and backs the cost
 * model in DESIGN.md §4.  The reference has no counterpart. */
#ifndef P3D_MICROBENCH_H
#define P3D_MICROBENCH_H
#ifdef __cplusplus
extern "C" {
#endif
/* kind 0: dependent-chain-free scalar FFMA; 1: packed FFMA2; 2: the pair kernel's instruction mix
 * (17 FFMA2/FADD2 : 2 MUFU.RSQ : 6 FMNMX); 3: FFMA2 with the pair kernel's shuffle rate (12 SHFL per 68 FFMA2).
 * kinds 4..17 are instruction-mix and register-operand-bandwidth probes used in DESIGN.md §5 (out[0] = thread-bodies/s).
 * out[0] = FP32 lane-FMAs per second (an FFMA2 counts 2 per lane), out[1] = kernel ms,
 * out[2] = SM count, out[3] = max SM clock in MHz as reported by the driver. */
int p3d_microbench(int device, int kind, int iters, double out[4]);
#ifdef __cplusplus
}
#endif
#endif
