/* p3d.h — C ABI of the B200-native step engine for the `particle_3d` crate.
 *
 * The reference (navpreett/3D-Particle-Simulation-) has NO FFI/plugin layer: its boundary is
 * the Rust public API  `Particle` (src/lib.rs:12-17), `Particles` (src/lib.rs:20-33) and
 * `Particles::update(&mut self, ts: f32) -> Vec<Particle>` (src/lib.rs:130).  This header is
 * the boundary a `build.rs` + `extern "C"` shim inside that crate would bind (the shim is shown
 * in INTEGRATION.md).  Plain pointers and sizes only; no torch / C++ types.
 *
 * There is no CPU fallback: every compute entry point returns P3D_ERR_NO_DEVICE / P3D_ERR_CUDA
 * when the GPU path cannot run.
 */
#ifndef P3D_H
#define P3D_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define P3D_ABI_VERSION 2

/* ---- error codes (the Rust shim turns non-zero into the panic the reference would raise) ---- */
#define P3D_OK 0
#define P3D_ERR_WORLD_TOO_SMALL 1 /* replaces assert!(world_size >= 2*radius), src/lib.rs:132 */
#define P3D_ERR_BAD_ID 2          /* replaces the slice-index panic at src/lib.rs:225-228 (id >= id_count) */
#define P3D_ERR_CUDA 3            /* any CUDA runtime failure; text in p3d_last_error() */
#define P3D_ERR_INVALID 4         /* null pointer, n mismatch, unsupported id_count, bad option */
#define P3D_ERR_NO_DEVICE 5       /* no CUDA device / not sm_100 */

#define P3D_MAX_TYPES 64 /* id_count limit of this engine (reference: unbounded u32) */

/* ---- data crossing the boundary ---- */

/* Replaces `struct Particle` (src/lib.rs:12-17): position, velocity (cgmath::Vector3<f32>), id.
 * 28 bytes, align 4.  The Rust struct is not #[repr(C)]; the shim adds it. */
typedef struct p3d_particle {
    float px, py, pz;
    float vx, vy, vz;
    uint32_t id;
} p3d_particle;

/* Replaces the scalar fields of `struct Particles` (src/lib.rs:20-33) read by update().
 * Passed on EVERY call: the caller may change any field between steps (src/bin/main.rs:263-359). */
typedef struct p3d_params {
    float world_size;              /* src/lib.rs:21 */
    float coefficient;             /* :27 drag */
    float interaction_force;       /* :28 */
    float min_pull_ratio;          /* :29 */
    float particle_effect_radius;  /* :30 */
    float accel[3];                /* :32 acceleration */
    uint32_t walls;                /* :31 bool */
    uint32_t id_count;             /* :24 */
    const float *attraction_matrix;/* :25 id_count*id_count floats, [self_id*id_count + other_id] */
} p3d_params;

typedef struct p3d_engine p3d_engine; /* opaque; owns device buffers, stream, events */

/* ---- engine lifetime ---- */
int p3d_abi_version(void);
/* Creates an engine on CUDA device `device`.  Not thread-safe: one caller at a time per handle
 * (mirrors `&mut self` at src/lib.rs:130). */
int p3d_create(int device, p3d_engine **out);
/* One handle that drives n_dev (1..8) devices of one node from the calling thread - the form the Rust shim uses
 * (`Particles::update` has a single caller and a single thread, src/bin/main.rs:199).  No IPC, no
 * NCCL, no second process: the devices reach each other's buffers by peer access, cross-device CUDA events are the
 * two barriers of a step.  p3d_upload / p3d_update split the caller's array so that every device moves 1/n_dev of
 * it over its own PCIe link and gathers the rest over NVLink; every device holds the whole state after each step.
 * Every whole-state call works on the handle (p3d_update, p3d_upload, p3d_step, p3d_download, p3d_download_forces,
 * p3d_download_render, p3d_diagnostics, p3d_sync, options, counters); the p3d_shard_*, p3d_ipc_*, p3d_set_stream,
 * p3d_device_buffer and per-kernel timing calls return P3D_ERR_INVALID on it.  A device may be listed more than once
 * (the members then share it; exercises the same code on a one-GPU box).
 * What is sharded: the all-pairs kernel (P3D_FORCE_PAIR) always; the cell-list and exact kernels only from 8M
 * particles (environment P3D_MULTI_CELLS_MIN), because below that one device steps faster than the host can issue
 * a sharded step - such an upload lives on the first device alone and behaves exactly like a one-device engine. */
int p3d_create_multi(const int *devices, int n_dev, p3d_engine **out);
void p3d_destroy(p3d_engine *eng);
/* Message for the last non-zero return on this thread. */
const char *p3d_last_error(void);

/* ---- the drop-in call: replaces Particles::update (src/lib.rs:130-272) ----
 * in/out are HOST arrays of n particles (may alias).  After the call out[i] is the updated
 * particle i (same index order, src/lib.rs:171-173,268); `in` is the state the reference leaves
 * in `past_particles` (src/lib.rs:167).  Synchronous.  n may differ from the previous call. */
int p3d_update(p3d_engine *eng, const p3d_params *prm, float ts,
               const p3d_particle *in, p3d_particle *out, size_t n);

/* ---- device-resident stepping (benches, headless runs; state stays in HBM between steps) ---- */
int p3d_upload(p3d_engine *eng, const p3d_particle *in, size_t n, uint32_t id_count);
/* n_steps x update(ts) without host round trips; asynchronous on the engine stream. */
int p3d_step(p3d_engine *eng, const p3d_params *prm, float ts, int n_steps);
int p3d_download(p3d_engine *eng, p3d_particle *out, size_t n);
int p3d_sync(p3d_engine *eng);
/* Caller index -> slot of the resident layout (identity unless the pair kernel's type-grouped layout is active);
 * slot / (n_slots / world) is the rank that integrates the particle.  For tests and the bench's parity sample. */
int p3d_slot_of(p3d_engine *eng, uint32_t *out, size_t n);
/* The resident state in the layout of the reference app's render storage buffer:
 * WGSL `struct Particles { world_size: f32, length: u32, particles: array<Particle> }` with 32-byte particles
 * (position vec3 @0, velocity vec3 @16, id u32 @28; src/bin/particles.wgsl:1-12).  Replaces the per-frame CPU
 * serialisation by encase at src/bin/main.rs:440-448; `out` receives 16 + 32*n bytes ready for queue.write_buffer. */
int p3d_download_render(p3d_engine *eng, float world_size, void *out, size_t out_bytes, size_t n);
/* total_force of src/lib.rs:177-243 from the most recent step, in particle index order (n*3). */
int p3d_download_forces(p3d_engine *eng, float *out_xyz, size_t n);
/* out[0]=sum 0.5*|v|^2, out[1..3]=sum v, out[4]=max |v|^2, out[5]=count, out[6]=sum |p|^2, out[7]=0 */
int p3d_diagnostics(p3d_engine *eng, double out[8]);

/* ---- options ---- */
enum p3d_option {
    P3D_OPT_FORCE_KERNEL = 0, /* see p3d_force_kernel */
    P3D_OPT_TIMING = 1,       /* 1: record CUDA events around each kernel of a step */
    P3D_OPT_GRAPH = 2,        /* 1 (default): untimed p3d_step runs of >= 6 steps replay a two-step CUDA graph */
    P3D_OPT_BLOCK_SORT = 3,   /* 1: re-partition interior/boundary blocks every step (fast path) */
    P3D_OPT_FAITHFUL = 5,     /* 1: reproduce the reference's bucket double-visit quirk: after the
                                 ideal force pass a correction kernel adds (multiplicity - 1) x contribution for every
                                 in-range pair whose bucket is hit by more than one of the 27 hashed cells
                                 (src/lib.rs:195-206).  Needs in-box positions and a box of at least three cells. */
    P3D_OPT_BLOCK_SIZE = 4    /* particles per block of the pair kernel: 0 = auto (256 from 65,536 particles, else 128), 128 (R=4) or 256 (R=8); applies at the next upload */
};
enum p3d_force_kernel {
    P3D_FORCE_AUTO = 0,      /* CELLS for n >= 192 (all-pairs when the box is narrower than three cells), else REFERENCE_ORDER */
    P3D_FORCE_REFERENCE_ORDER = 1, /* one thread per particle, exact sqrt/div, all three images per axis */
    P3D_FORCE_PAIR = 2,      /* symmetric block-pair kernel, packed FP32x2, rsqrt (all N^2 pairs); a step whose input has a
                              * particle outside the box is evaluated by the exact kernel (1) instead or, from 32,768
                              * particles, by the cell list (3), queued behind a device-side flag: same forces */
    P3D_FORCE_CELLS = 3      /* uniform-grid cell list: the GPU analogue of the reference's spatial hash
                                (src/lib.rs:135-236); same results, O(N * neighbours) work */
};
int p3d_set_option(p3d_engine *eng, int option, int value);
int p3d_get_option(p3d_engine *eng, int option, int *value);

/* Per-kernel device times of the most recent p3d_step/p3d_update (needs P3D_OPT_TIMING=1), CUDA
 * events on the engine stream, milliseconds summed over the timed steps of that call:
 * [0]=whole force pass [1]=integrate kernel [2]=pack (upload side) [3]=unpack (download side)
 * [4]=partition kernels + force memset [5]=h2d copy [6]=d2h copy [7]=force+integrate
 * [8]=pair kernel (the boundary-x-boundary kernel runs beside it on an auxiliary stream) [9]=what remains of the
 * boundary-x-boundary kernel after the pair kernel ended (+ out-of-box fallback) [10]=timed steps [11]=0 */
int p3d_get_timing(p3d_engine *eng, float ms[12]);
/* Launch counters since creation: [0]=kernels launched, [1]=force kernels, [2]=integrate kernels. */
int p3d_get_counters(p3d_engine *eng, uint64_t out[4]);

/* ---- plumbing for one-process-per-GPU drivers (torch.distributed owns the collective) ---- */
/* Run every subsequent launch on this cudaStream_t (e.g. torch's current stream); NULL = own stream
 * (pass cudaStreamLegacy, (cudaStream_t)0x1, to name the legacy default stream). */
int p3d_set_stream(p3d_engine *eng, void *cuda_stream);
enum p3d_buffer {
    P3D_BUF_POS = 0,      /* float4 {x,y,z,id bits} per slot, current positions */
    P3D_BUF_POS_NEXT = 1, /* float4 per slot, written by integrate */
    P3D_BUF_VEL = 2,      /* float4 {vx,vy,vz,0} per slot */
    P3D_BUF_FORCE = 3,    /* float4 {fx,fy,fz,0} per slot */
    P3D_BUF_AOS = 4       /* staging array of the sharded upload: 7 words per caller index (p3d_particle), element
                             count = world * ceil(n / world) particles; valid after p3d_upload_part */
};
/* Device pointer + element count (slots, >= n because type segments are padded). */
int p3d_device_buffer(p3d_engine *eng, int which, void **dev_ptr, size_t *n_slots);
/* Shard = the slot range this rank integrates and, for the force pass, its share of the work (block rows of the
 * pair kernel, cell-sorted index range of the cell list, slot range of the reference-order kernel).
 * world==1 restores single-GPU behaviour.  Call BEFORE p3d_upload: the slot layout is padded so that all ranks own
 * equally many slots; changing `world` afterwards drops the resident state (upload again).  While world > 1 the
 * whole-step calls (p3d_update, p3d_step) return P3D_ERR_INVALID: a step then needs the driver's collectives
 * between the p3d_shard_* calls below. */
int p3d_set_shard(p3d_engine *eng, int rank, int world);
int p3d_shard_range(p3d_engine *eng, size_t *slot_begin, size_t *slot_end);
/* One step split at the collectives a multi-GPU driver inserts:
 *   p3d_shard_force     -> PARTIAL forces into P3D_BUF_FORCE: every kernel zeroes the buffer and then adds / scatters
 *                          its share, so a rank's own slots receive contributions computed on other ranks
 *   [driver: sum-reduce P3D_BUF_FORCE across ranks - REQUIRED for P3D_FORCE_PAIR, P3D_FORCE_CELLS, P3D_OPT_FAITHFUL
 *    and P3D_FORCE_AUTO; only P3D_FORCE_REFERENCE_ORDER without the faithful option writes complete forces for
 *    exactly the rank's own slot range and may skip it]
 *   p3d_shard_integrate -> integrates [slot_begin,slot_end) into P3D_BUF_POS_NEXT / P3D_BUF_VEL
 *   [driver: all-gather P3D_BUF_POS_NEXT (and P3D_BUF_VEL, if any rank is to download the whole state)]
 *   p3d_shard_commit    -> swaps POS/POS_NEXT */
int p3d_shard_force(p3d_engine *eng, const p3d_params *prm);
int p3d_shard_integrate(p3d_engine *eng, const p3d_params *prm, float ts);
int p3d_shard_commit(p3d_engine *eng);

/* Sharded host <-> device traffic for one-process-per-GPU drivers: every rank moves only ITS part of the caller's
 * array over its own PCIe link.
 *   p3d_upload_part   -> callers [i_begin, i_end) of n into the staging array (P3D_BUF_AOS); synchronous.  All parts
 *                        of one upload name the same n and id_count; the staging array may move when n grows, so
 *                        query P3D_BUF_AOS after the call, not before
 *   [driver: all-gather P3D_BUF_AOS across ranks (equal parts of ceil(n / world) particles)]
 *   p3d_upload_commit -> layout + pack from the staging array (what p3d_upload does after its copy)
 *   p3d_download_part -> callers [i_begin, i_end) of the resident state (every rank holds the whole state after a
 *                        fused step, or after the driver gathered positions and velocities) */
int p3d_upload_part(p3d_engine *eng, const p3d_particle *part, size_t i_begin, size_t i_end, size_t n, uint32_t id_count);
int p3d_upload_commit(p3d_engine *eng);
int p3d_download_part(p3d_engine *eng, p3d_particle *out_part, size_t i_begin, size_t i_end);

/* Fused variant of the step's second half over NVLink peer memory (one kernel = reduce-scatter of the
 * partial forces + integrate + all-gather of the new positions and velocities; no NCCL on the data path):
 *   p3d_ipc_export  -> four 64-byte CUDA IPC handles (force buffer, both position buffers, velocities) of this rank
 *   [driver: all-gather the handles]
 *   p3d_ipc_import  -> opens every peer's buffers (world <= 8, same node, after p3d_upload on all ranks)
 *   per step: p3d_shard_force; [barrier]; p3d_shard_integrate_fused; [barrier]; p3d_shard_commit */
#define P3D_IPC_HANDLES 4
int p3d_ipc_export(p3d_engine *eng, unsigned char *handles /* P3D_IPC_HANDLES * 64 bytes */);
int p3d_ipc_import(p3d_engine *eng, int world, const unsigned char *all_handles /* world * P3D_IPC_HANDLES * 64 bytes */);
int p3d_ipc_close(p3d_engine *eng);
int p3d_shard_integrate_fused(p3d_engine *eng, const p3d_params *prm, float ts);

/* Self-checking builds (-DP3D_BOUNDS_CHECK: every data-dependent slot / cell index in the kernels is compared
 * with its extent, violations are counted and the access skipped): number of violations since the library was
 * loaded.  A product build stores UINT64_MAX.  (No reference counterpart: Rust's slice indexing panics,
 * src/lib.rs:225-228; compute-sanitizer is unavailable on the target pool.) */
int p3d_debug_bounds_violations(p3d_engine *eng, unsigned long long *count);

/* ---- seeded scenes (host only; restates the binary-private generator, src/bin/main.rs:60-87,
 *      and the default scene constants, src/bin/main.rs:123-148) ---- */
/* Fills prm with the default scene; matrix25 receives the 5x5 default matrix and prm points at it. */
void p3d_scene_default_params(p3d_params *prm, float matrix25[25]);
/* Uniform positions in [-W/2, W/2]^3, zero velocity, ids uniform in 0..id_count (splitmix64, seeded). */
void p3d_scene_uniform(uint64_t seed, size_t n, float world_size, uint32_t id_count, p3d_particle *out);
/* Plummer-like cloud: radius from the Plummer CDF with scale a, isotropic, rejected outside the box. */
void p3d_scene_plummer(uint64_t seed, size_t n, float world_size, float scale_a, uint32_t id_count,
                       p3d_particle *out);

#ifdef __cplusplus
}
#endif
#endif /* P3D_H */
