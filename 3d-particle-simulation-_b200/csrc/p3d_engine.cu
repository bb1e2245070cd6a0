// p3d_engine.cu — host side of the C ABI declared in include/p3d.h: device-buffer ownership,
// type-sorted slot layout, per-step launch sequence, timing.  No CPU compute fallback exists:
// every compute entry point needs the CUDA device.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <limits>
#include <memory>
#include <string>
#include <vector>

#include "p3d.h"
#include "p3d_kernels_basic.cuh"
#include "p3d_kernels_pair.cuh"
#include "p3d_kernels_cells.cuh"


static_assert(sizeof(p3d_particle) == 28, "boundary struct must be 28 bytes (src/lib.rs:12-17)");
static_assert(sizeof(AosParticle) == 28, "device view of the boundary struct");

namespace {

thread_local std::string g_last_error;

int fail(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}

#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(P3D_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, \
                        __LINE__);                                                                 \
    } while (0)

constexpr int kMaxTimedSteps = 512;
constexpr int kEv = 5;  // events per timed step
constexpr int kRefTile = 128;       // threads per CTA of the reference-order kernel
constexpr int kOutOfBoxCellsMin = 32768;  // all-pairs path, out-of-box input: gated cell list from this n, exact kernel below
constexpr int kPinWords = 128;     // pinned scratch words per engine (>= P3D_MAX_TYPES + 2)
constexpr int kCellsAutoMin = 192;  // P3D_FORCE_AUTO uses the cell list from this n (its ~12 launches cost ~35 us),
                                    // the single-launch reference-order kernel below

template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t cap = 0;
    // Grow-only (contents are not preserved).  The new block is allocated BEFORE the old one is freed, so a
    // failed grow leaves the buffer as it was (and the engine's n / n_slots still describe valid memory).
    int ensure(size_t n) {
        if (n <= cap) return P3D_OK;
        const size_t want = n + n / 8 + 64;  // slack: the host may add particles (main.rs:274-279)
        T *fresh = nullptr;
        CU(cudaMalloc(&fresh, want * sizeof(T)));
        if (p) cudaFree(p);
        p = fresh;
        cap = want;
        return P3D_OK;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};


// ---- pageable host memory ----
// The caller's arrays are ordinary heap memory in the reference's world (`Vec<Particle>`, src/lib.rs:22-23).  A
// cudaMemcpy from / to pageable memory is staged by the driver through one thread at ~14 GB/s (measured: 2.1 ms per
// 29 MB at N = 1M, against 0.55 ms from pinned memory).  Large pageable transfers therefore go through two pinned
// staging buffers of the engine: a few host threads copy chunk k+1 between the caller's array and the staging
// buffer while the DMA engine moves chunk k.
constexpr size_t kStageChunk = 4u << 20;    // bytes per staging buffer
constexpr size_t kStageMin = 512u << 10;    // smaller pageable transfers take the plain path
constexpr size_t kStageSlice = 256u << 10;  // bytes per unit of work of a host thread

class HostCopyPool {
   public:
    static HostCopyPool &get() {
        static HostCopyPool pool;
        return pool;
    }
    // memcpy(dst, src, bytes) split over the pool's threads and the calling thread; returns when done.
    void copy(void *dst, const void *src, size_t bytes) {
        if (workers_.empty() || bytes <= kStageSlice) {
            std::memcpy(dst, src, bytes);
            return;
        }
        // Every copy is its own Job object: a worker that wakes up late still holds the finished job (whose slices
        // are exhausted) and can never touch the fields or counters of the next one.
        auto job = std::make_shared<Job>();
        job->dst = static_cast<char *>(dst);
        job->src = static_cast<const char *>(src);
        job->bytes = bytes;
        job->n_slices = (bytes + kStageSlice - 1) / kStageSlice;
        {
            std::lock_guard<std::mutex> lk(m_);
            current_ = job;
            generation_.fetch_add(1, std::memory_order_release);
        }
        cv_work_.notify_all();
        run(*job);
        std::unique_lock<std::mutex> lk(m_);
        cv_done_.wait(lk, [&] { return job->done.load() == job->n_slices; });
    }

   private:
    struct Job {
        char *dst = nullptr;
        const char *src = nullptr;
        size_t bytes = 0, n_slices = 0;
        std::atomic<size_t> next{0}, done{0};
    };
    HostCopyPool() {
        int n = 3;
        if (const char *env = std::getenv("P3D_HOST_THREADS")) n = std::max(0, std::min(15, std::atoi(env) - 1));
        const unsigned hw = std::thread::hardware_concurrency();
        if (hw && (unsigned)n + 1 > hw) n = (int)hw - 1;
        for (int k = 0; k < n; ++k) workers_.emplace_back([this] { loop(); });
    }
    ~HostCopyPool() {
        {
            std::lock_guard<std::mutex> lk(m_);
            stop_ = true;
        }
        cv_work_.notify_all();
        for (auto &t : workers_) t.join();
    }
    void run(Job &job) {
        for (;;) {
            const size_t k = job.next.fetch_add(1);
            if (k >= job.n_slices) return;
            const size_t off = k * kStageSlice;
            std::memcpy(job.dst + off, job.src + off, std::min(kStageSlice, job.bytes - off));
            if (job.done.fetch_add(1) + 1 == job.n_slices) {
                std::lock_guard<std::mutex> lk(m_);
                cv_done_.notify_all();
            }
        }
    }
    void loop() {
        uint64_t seen = 0;
        for (;;) {
            // The chunks of one transfer arrive ~100 us apart: spin that long for the next one before going to
            // sleep (a condition-variable wake-up per chunk costs as much as copying it).
            const auto t0 = std::chrono::steady_clock::now();
            bool got = false;
            while (std::chrono::steady_clock::now() - t0 < std::chrono::microseconds(300)) {
                if (generation_.load(std::memory_order_acquire) != seen) { got = true; break; }
            }
            std::shared_ptr<Job> job;
            {
                std::unique_lock<std::mutex> lk(m_);
                if (!got) cv_work_.wait(lk, [&] { return stop_ || generation_.load() != seen; });
                if (stop_) return;
                seen = generation_.load();
                job = current_;
            }
            if (job) run(*job);
        }
    }
    std::vector<std::thread> workers_;
    std::mutex m_;
    std::condition_variable cv_work_, cv_done_;
    std::shared_ptr<Job> current_;
    std::atomic<uint64_t> generation_{0};
    bool stop_ = false;
};

bool host_is_pageable(const void *p) {
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) {
        cudaGetLastError();
        return true;
    }
    return attr.type == cudaMemoryTypeUnregistered;
}

}  // namespace

struct p3d_engine {
    int device = 0;
    int sm_count = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    cudaStream_t aux_stream = nullptr;               // boundary-x-boundary kernel runs beside the pair kernel
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;

    // layout
    size_t n = 0;        // live particles
    int n_slots = 0;     // padded slots (multiple of B)
    int B = 128;         // block size in particles (32 * R) of the current layout
    int B_next = 0;      // block size for the next upload (P3D_OPT_BLOCK_SIZE); 0 = by particle count
    int M = 0;           // n_slots / B
    uint32_t T = 0;      // id_count the layout was built for
    bool typed = false;  // slots are grouped by type (needed by the pair kernel); false: slot = caller index ...
    bool permuted = false;   // ... unless the identity layout was re-slotted into cell order (reslot_by_cell)
    int steps_since_reslot = 0;
    long long steps_since_upload = 0;  // resident steps taken on the current upload
    std::vector<int> seg_start_h, seg_end_h;

    DevBuf<float4> pos[2], vel, frc, spos;
    DevBuf<uint32_t> slot_of, sidx, type_cnt;  // type_cnt: per-CTA type counts, offsets and totals of the layout sort
    DevBuf<uint8_t> seg_type, bclass;
    DevBuf<int> seg_start, seg_end, cnt;
    DevBuf<int2> cta_cnt, cta_off;
    // cell-list path
    DevBuf<uint32_t> ckeys[2], cvals[2], crank, cell_cnt, cell_off, scan_tiles;
    DevBuf<uint32_t> caller_of, caller_tmp;  // slot -> caller index of a re-slotted identity layout
    DevBuf<float4> cpos, vel_tmp;
    DevBuf<float> aos, fout, sx, sy, sz;
    DevBuf<float4> render;
    DevBuf<float> matrix;
    DevBuf<int> flags;      // [0],[1]: out-of-box flags (double-buffered by step parity)
    DevBuf<double> diag;
    int cur = 0;            // which pos buffer is current
    int parity = 0;         // which flag word describes the current positions
    bool upload_timed = false;  // ev_call[0..1] were recorded by the upload in flight
    size_t n_staged = 0;    // particles in the AoS staging array `aos` awaiting p3d_upload_commit (sharded upload)
    uint32_t T_staged = 0;
    unsigned char *stage_pin[2] = {nullptr, nullptr};  // pinned staging buffers for pageable caller memory (lazy)
    cudaEvent_t stage_ev[2] = {nullptr, nullptr};      // the DMA out of / into each buffer has completed
    uint32_t *host_pin = nullptr;  // pinned scratch (kPinWords words): per-type totals and flags come back without
                                   // the implicit synchronisation of a copy into pageable memory

    std::vector<uint8_t> seg_type_h;

    // options
    int opt_force = P3D_FORCE_AUTO;
    int opt_timing = 0;
    int opt_block_sort = 1;
    int opt_faithful = 0;    // K5: add the reference's bucket double-visit contributions
    int opt_graph = 1;       // replay device-resident multi-step runs through a CUDA graph (two steps per graph)

    // CUDA graph of two consecutive steps (returns cur/parity to their starting values)
    cudaGraphExec_t graph_exec = nullptr;
    std::vector<unsigned char> graph_key;
    uint64_t graph_launches[3] = {0, 0, 0};  // kernel / force / integrate launches inside one replay
    uint64_t layout_version = 0;

    // sharding
    int rank = 0, world = 1;
    // A handle made by p3d_create_multi drives one member engine per device (rank g of members.size()) from the
    // calling thread; the members' peer tables point straight at each other's buffers (one process, peer access).
    std::vector<p3d_engine *> members;
    std::vector<cudaEvent_t> ev_bar;   // one per member: recorded at each cross-device barrier
    bool is_member = false;
    bool solo = false;                 // handle: the resident upload lives on member 0 alone (see multi_upload)
    size_t multi_cells_min = 8u << 20; // handle: particle count from which the cell-list / exact kernels are sharded
    // peer memory (CUDA IPC or, inside a multi-device handle, plain peer access):
    // [rank][0]=frc, [1]=pos[0], [2]=pos[1], [3]=vel; own entries are the local pointers
    void *peer_ptr[8][4] = {};
    bool peer_open[8] = {};
    bool peers_ready = false;
    bool ipc_exported = false;  // frc/pos were handed to peers: they must not be reallocated until p3d_ipc_close

    // timing
    std::vector<cudaEvent_t> ev;  // kEv per timed step: start, after partition, after pair, after force, after integrate
    int timed_steps = 0;
    cudaEvent_t ev_call[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    float last_ms[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    cudaEvent_t *step_ev = nullptr;  // events of the step being recorded (null: untimed)
    uint64_t counters[4] = {0, 0, 0, 0};
};

namespace {

// Chunk of a staged transfer: a quarter of it (so that host copies and DMA overlap even for a few MB), whole slices,
// at most the staging buffer.
size_t stage_chunk_for(size_t bytes) {
    const size_t quarter = (bytes / 4 + kStageSlice - 1) / kStageSlice * kStageSlice;
    return std::min(kStageChunk, std::max<size_t>(2 * kStageSlice, quarter));
}

int ensure_staging(p3d_engine *e) {
    for (int b = 0; b < 2; ++b) {
        if (!e->stage_pin[b]) CU(cudaMallocHost(&e->stage_pin[b], kStageChunk));
        if (!e->stage_ev[b]) CU(cudaEventCreateWithFlags(&e->stage_ev[b], cudaEventDisableTiming));
    }
    return P3D_OK;
}

// Host -> device on the engine stream.  On return the caller's memory has been consumed only if it is pageable
// (staged path); the callers of this function synchronise the stream before they hand control back anyway.
int copy_h2d(p3d_engine *e, void *dst_dev, const void *src_host, size_t bytes) {
    if (bytes < kStageMin || !host_is_pageable(src_host)) {
        CU(cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, e->stream));
        return P3D_OK;
    }
    int rc;
    if ((rc = ensure_staging(e))) return rc;
    HostCopyPool &pool = HostCopyPool::get();
    const size_t chunk = stage_chunk_for(bytes);
    size_t k = 0;
    for (size_t off = 0; off < bytes; off += chunk, ++k) {
        const int b = (int)(k & 1);
        const size_t len = std::min(chunk, bytes - off);
        CU(cudaEventSynchronize(e->stage_ev[b]));  // the DMA that last read this buffer is done (a fresh event is "done")
        pool.copy(e->stage_pin[b], static_cast<const char *>(src_host) + off, len);
        CU(cudaMemcpyAsync(static_cast<char *>(dst_dev) + off, e->stage_pin[b], len, cudaMemcpyHostToDevice, e->stream));
        CU(cudaEventRecord(e->stage_ev[b], e->stream));
    }
    return P3D_OK;
}

// Device -> host on the engine stream.  *done: the data has already arrived (staged path); otherwise the copy is
// merely queued and the caller synchronises the stream.
int copy_d2h(p3d_engine *e, void *dst_host, const void *src_dev, size_t bytes, bool *done) {
    *done = false;
    if (bytes < kStageMin || !host_is_pageable(dst_host)) {
        CU(cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, e->stream));
        return P3D_OK;
    }
    int rc;
    if ((rc = ensure_staging(e))) return rc;
    HostCopyPool &pool = HostCopyPool::get();
    const size_t chunk = stage_chunk_for(bytes);
    const size_t n_chunks = (bytes + chunk - 1) / chunk;
    auto issue = [&](size_t k) -> int {
        const int b = (int)(k & 1);
        const size_t off = k * chunk, len = std::min(chunk, bytes - off);
        CU(cudaMemcpyAsync(e->stage_pin[b], static_cast<const char *>(src_dev) + off, len, cudaMemcpyDeviceToHost, e->stream));
        CU(cudaEventRecord(e->stage_ev[b], e->stream));
        return P3D_OK;
    };
    for (int b = 0; b < 2; ++b) CU(cudaEventSynchronize(e->stage_ev[b]));  // an earlier upload may still be reading them
    for (size_t k = 0; k < std::min<size_t>(2, n_chunks); ++k)
        if ((rc = issue(k))) return rc;
    for (size_t k = 0; k < n_chunks; ++k) {
        const int b = (int)(k & 1);
        const size_t off = k * chunk, len = std::min(chunk, bytes - off);
        CU(cudaEventSynchronize(e->stage_ev[b]));
        pool.copy(static_cast<char *>(dst_host) + off, e->stage_pin[b], len);
        if (k + 2 < n_chunks && (rc = issue(k + 2))) return rc;
    }
    *done = true;
    return P3D_OK;
}

int canonicalise(const p3d_params *prm, DevParams &P) {
    if (!prm) return fail(P3D_ERR_INVALID, "params is null");
    if (prm->id_count == 0 || prm->id_count > P3D_MAX_TYPES)
        return fail(P3D_ERR_INVALID, "id_count %u outside 1..%d", prm->id_count, P3D_MAX_TYPES);
    if (!prm->attraction_matrix) return fail(P3D_ERR_INVALID, "attraction_matrix is null");
    // src/lib.rs:132  assert!(self.world_size >= 2.0 * self.particle_effect_radius)
    if (!(prm->world_size >= 2.0f * prm->particle_effect_radius))
        return fail(P3D_ERR_WORLD_TOO_SMALL, "world_size %g < 2 * particle_effect_radius %g (src/lib.rs:132)",
                    prm->world_size, prm->particle_effect_radius);
    P.W = prm->world_size;
    P.half = prm->world_size * 0.5f;
    P.r = prm->particle_effect_radius;
    P.r2 = P.r * P.r;
    P.m = prm->min_pull_ratio;
    P.kf = prm->interaction_force;
    P.coef = prm->coefficient;
    P.ax = prm->accel[0];
    P.ay = prm->accel[1];
    P.az = prm->accel[2];
    P.walls = prm->walls ? 1 : 0;
    P.T = (int)prm->id_count;
    const float m = P.m;
    P.inv_m = (m > 0.0f) ? 1.0f / m : std::numeric_limits<float>::infinity();
    if (m < 1.0f) {
        P.c2 = 2.0f / (1.0f - m);
    } else {
        P.c2 = 0.0f;
    }
    // The force law is non-zero on (0, 1) — and on (0, m) when m > 1, where the repulsion branch d < m
    // (src/lib.rs:56-58) outlives the attraction branch.  The cutoff d < r (src/lib.rs:216-220) bites when r is smaller.
    // A negative radius (the fields are public) cuts at |r|: the reference compares d^2 with r*r (src/lib.rs:218-219);
    // only the kick (src/lib.rs:246) sees its sign.
    const float law_range = std::max(1.0f, m);
    const float r_abs = std::fabs(P.r);
    P.rcut = (r_abs < law_range) ? 1 : 0;
    P.reach = std::min(r_abs, law_range);
    return P3D_OK;
}

int ensure_common(p3d_engine *e, size_t n, size_t ns) {
    int rc;
    if (e->ipc_exported && ((size_t)ns > e->pos[0].cap || (size_t)ns > e->pos[1].cap || (size_t)ns > e->frc.cap ||
                            (size_t)ns > e->vel.cap))
        return fail(P3D_ERR_INVALID, "upload needs %zu slots but the position/velocity/force buffers are exported to peer GPUs "
                    "(p3d_ipc_export); call p3d_ipc_close on every rank, upload, then export/import again", (size_t)ns);
    if ((rc = e->pos[0].ensure(ns))) return rc;
    if ((rc = e->pos[1].ensure(ns))) return rc;
    if ((rc = e->vel.ensure(ns))) return rc;
    if ((rc = e->frc.ensure(ns))) return rc;
    (void)n;  // (the AoS staging array was sized by stage_input)
    if ((rc = e->matrix.ensure(P3D_MAX_TYPES * P3D_MAX_TYPES))) return rc;
    if ((rc = e->flags.ensure(4))) return rc;
    if ((rc = e->diag.ensure(8))) return rc;
    return P3D_OK;
}

int resolve_force_kernel_for(const p3d_engine *e, size_t n);

// caller index -> slot table, or null when slot == caller index
const uint32_t *slot_map(const p3d_engine *e) { return (e->typed || e->permuted) ? e->slot_of.p : nullptr; }

// A sharded engine computes PARTIAL forces (its block rows) and integrates only its slot range; between the two
// the driver has to sum the forces and afterwards gather the positions.  A whole-step call would silently
// integrate with partial forces, so it is refused.
int fail_sharded(const p3d_engine *e, const char *call) {
    return fail(P3D_ERR_INVALID, "%s on a sharded engine (rank %d of %d): a whole step needs the driver's collectives - use "
                "p3d_shard_force / p3d_shard_integrate[_fused] / p3d_shard_commit, or p3d_set_shard(eng, 0, 1) first",
                call, e->rank, e->world);
}

// ---- upload, phase 1: the caller's particles [i_begin, i_end) of n -> the device-side AoS staging array ----
// Asynchronous on the engine stream.  The array is sized for world * ceil(n / world) particles so that a
// one-process-per-GPU driver can all-gather equal parts into it (P3D_BUF_AOS).
size_t staged_part(size_t n, int world) { return (n + (size_t)world - 1) / (size_t)world; }

int stage_input(p3d_engine *e, const p3d_particle *part, size_t i_begin, size_t i_end, size_t n) {
    if (n > (size_t)0x7fff0000) return fail(P3D_ERR_INVALID, "n too large");
    int rc;
    const size_t cap = std::max<size_t>(staged_part(n, e->world) * (size_t)e->world, 1);
    if ((rc = e->aos.ensure(cap * 7))) return rc;
    if (i_end > i_begin && (rc = copy_h2d(e, e->aos.p + i_begin * 7, part, (i_end - i_begin) * sizeof(p3d_particle))))
        return rc;
    return P3D_OK;
}

// Identity layout (slot = caller index): all the cell-list and reference-order kernels need.  No host
// pass over the particles; ids are validated on the device by k_pack.
int build_layout_identity(p3d_engine *e, size_t n, uint32_t T) {
    e->B = 128;
    const size_t unit = (size_t)e->B * (size_t)e->world;
    const size_t ns = (std::max<size_t>(n, 1) + unit - 1) / unit * unit;
    int rc;
    if ((rc = ensure_common(e, n, ns))) return rc;
    e->n_slots = (int)ns;
    e->M = e->n_slots / e->B;
    e->n = n;
    e->T = T;
    e->typed = false;
    e->permuted = false;
    e->steps_since_upload = 0;
    e->layout_version++;
    CU(cudaMemsetAsync(e->flags.p, 0, 4 * sizeof(int), e->stream));
    return P3D_OK;
}

// Type-grouped layout for the pair kernel.  The counting sort by type id runs on the device (k_type_hist,
// k_type_scan, later k_pack_typed); the host only sees the T per-type totals, from which it derives the
// block-padded segments.
// Part 1 (asynchronous): per-type counts of the staged AoS array and the smallest caller index with a bad id,
// copied into the engine's pinned scratch ([0..T) counts, [T] bad index).  The caller synchronises the stream.
int layout_count_async(p3d_engine *e, size_t n, uint32_t T) {
    static_assert(P3D_MAX_TYPES == kTypeMax, "k_type_* kernels are sized for P3D_MAX_TYPES");
    static_assert(P3D_MAX_TYPES + 2 <= kPinWords, "pinned scratch holds the type totals");
    cudaStream_t st = e->stream;
    int rc;
    const size_t n_ctas = (n + kTypeThreads - 1) / kTypeThreads;
    if ((rc = e->type_cnt.ensure(2 * std::max<size_t>(n_ctas, 1) * T + P3D_MAX_TYPES))) return rc;
    if ((rc = e->flags.ensure(4))) return rc;
    for (uint32_t t = 0; t <= T; ++t) e->host_pin[t] = 0u;
    e->host_pin[T] = 0x7f7f7f7fu;
    if (!n) return P3D_OK;
    uint32_t *cta_cnt = e->type_cnt.p, *cta_off = cta_cnt + n_ctas * T, *total = cta_off + n_ctas * T;
    CU(cudaMemsetAsync(e->flags.p + 2, 0x7f, sizeof(int), st));  // 0x7f7f7f7f: larger than any index
    k_type_hist<<<(unsigned)n_ctas, kTypeThreads, 0, st>>>(e->aos.p, (int)n, T, cta_cnt, e->flags.p + 2);
    k_type_scan<<<T, 1024, 0, st>>>(cta_cnt, cta_off, (int)n_ctas, T, total);
    e->counters[0] += 2;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(e->host_pin, total, T * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(e->host_pin + T, e->flags.p + 2, sizeof(int), cudaMemcpyDeviceToHost, st));
    return P3D_OK;
}

// Part 2 (the stream has been synchronised): segments from the totals, buffers, small tables.
int layout_finish(p3d_engine *e, size_t n, uint32_t T) {
    cudaStream_t st = e->stream;
    int rc;
    const uint32_t *count = e->host_pin;
    if ((size_t)e->host_pin[T] < n) {
        const size_t bad = e->host_pin[T];
        uint32_t id = 0;
        CU(cudaMemcpy(&id, e->aos.p + bad * 7 + 6, sizeof(uint32_t), cudaMemcpyDeviceToHost));
        return fail(P3D_ERR_BAD_ID, "particle %zu has id %u >= id_count %u (src/lib.rs:225-228)", bad, id, T);
    }
    const int B = e->B_next ? e->B_next : (n >= 65536 ? 256 : 128);
    std::vector<int> seg_start(T, 0), seg_end(T, 0);
    size_t at = 0;
    for (uint32_t t = 0; t < T; ++t) {
        seg_start[t] = (int)at;
        at += ((size_t)count[t] + B - 1) / B * B;
        seg_end[t] = (int)at;
    }
    if (at == 0) at = B;  // keep one (ghost) block so kernels always have a valid grid
    {   // equal shards for the multi-GPU all-gather: the last type's region absorbs the padding blocks
        const size_t unit = (size_t)B * (size_t)e->world;
        at = (at + unit - 1) / unit * unit;
        seg_end[T - 1] = (int)at;
    }
    const size_t ns = at;
    const size_t M = ns / B;
    if (e->ipc_exported && (ns > e->pos[0].cap || ns > e->pos[1].cap || ns > e->frc.cap || ns > e->vel.cap))
        return fail(P3D_ERR_INVALID, "upload needs %zu slots but the position/velocity/force buffers are exported to peer GPUs "
                    "(p3d_ipc_export); call p3d_ipc_close on every rank, upload, then export/import again", ns);
    if ((rc = e->pos[0].ensure(ns))) return rc;
    if ((rc = e->pos[1].ensure(ns))) return rc;
    if ((rc = e->vel.ensure(ns))) return rc;
    if ((rc = e->frc.ensure(ns))) return rc;
    if ((rc = e->spos.ensure(ns))) return rc;
    if ((rc = e->sidx.ensure(ns))) return rc;
    if ((rc = e->sx.ensure(ns))) return rc;
    if ((rc = e->sy.ensure(ns))) return rc;
    if ((rc = e->sz.ensure(ns))) return rc;
    if ((rc = e->slot_of.ensure(n ? n : 1))) return rc;
    if ((rc = e->seg_type.ensure(M))) return rc;
    if ((rc = e->bclass.ensure(M))) return rc;
    if ((rc = e->seg_start.ensure(P3D_MAX_TYPES))) return rc;
    if ((rc = e->seg_end.ensure(P3D_MAX_TYPES))) return rc;
    if ((rc = e->cnt.ensure(2 * P3D_MAX_TYPES))) return rc;
    if ((rc = e->cta_cnt.ensure(ns / kPartThreads + 1))) return rc;
    if ((rc = e->cta_off.ensure(ns / kPartThreads + 1))) return rc;
    if ((rc = e->matrix.ensure(P3D_MAX_TYPES * P3D_MAX_TYPES))) return rc;
    if ((rc = e->diag.ensure(8))) return rc;
    // every allocation succeeded: only now does the engine's layout change
    e->typed = true;
    e->permuted = false;
    e->steps_since_upload = 0;
    e->layout_version++;
    e->B = B;
    e->seg_start_h = seg_start;
    e->seg_end_h = seg_end;
    e->n_slots = (int)ns;
    e->M = (int)M;
    e->n = n;
    e->T = T;
    e->seg_type_h.assign(e->M, 0);
    for (uint32_t t = 0; t < T; ++t)
        for (int b = e->seg_start_h[t] / B; b < e->seg_end_h[t] / B; ++b) e->seg_type_h[b] = (uint8_t)t;

    // small host arrays owned by the engine (they stay valid until the next upload, which first drains the stream)
    CU(cudaMemcpyAsync(e->seg_type.p, e->seg_type_h.data(), (size_t)e->M, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(e->seg_start.p, e->seg_start_h.data(), T * sizeof(int), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(e->seg_end.p, e->seg_end_h.data(), T * sizeof(int), cudaMemcpyHostToDevice, st));
    CU(cudaMemsetAsync(e->flags.p, 0, 4 * sizeof(int), st));
    return P3D_OK;
}

int launch_pack(p3d_engine *e, size_t n) {
    cudaStream_t st = e->stream;
    const int ns = e->n_slots;
#ifdef P3D_BOUNDS_CHECK
    k_debug_publish<<<1, 1, 0, st>>>((unsigned int)ns, 0xFFFFFFFFu);
#endif
    e->cur = 0;
    e->parity = 0;
    k_fill_ghosts<<<(ns + 255) / 256, 256, 0, st>>>(e->pos[0].p, e->pos[1].p, e->vel.p, e->frc.p, ns);
    e->counters[0]++;
    if (n && e->typed) {
        const uint32_t *cta_off = e->type_cnt.p + (n + kTypeThreads - 1) / kTypeThreads * e->T;
        k_pack_typed<<<(unsigned)((n + kTypeThreads - 1) / kTypeThreads), kTypeThreads, 0, st>>>(
            e->aos.p, (int)n, e->T, e->seg_start.p, cta_off, e->pos[0].p, e->vel.p, e->slot_of.p);
        e->counters[0]++;
    } else if (n) {
        k_pack<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(e->aos.p, nullptr, e->pos[0].p, e->vel.p, (int)n, e->T,
                                                            e->flags.p + 2);
        e->counters[0]++;
    }
    CU(cudaGetLastError());
    return P3D_OK;
}

int resolve_force_kernel_for(const p3d_engine *e, size_t n) {
    // AUTO: the cell list (same results, O(N * neighbours) work; falls back to all-pairs inside
    // launch_force when the box is narrower than three cells); tiny systems take the exact kernel.
    if (e->opt_force == P3D_FORCE_AUTO) return n >= (size_t)kCellsAutoMin ? P3D_FORCE_CELLS : P3D_FORCE_REFERENCE_ORDER;
    return e->opt_force;
}
int resolve_force_kernel(const p3d_engine *e) {
    return resolve_force_kernel_for(e, e->n);
}

size_t ref_smem(int tile, int T) { return (size_t)tile * sizeof(float4) + (size_t)T * T * sizeof(float); }

// Grid for the cell list: cells of edge W/nc >= reach, at most ~8 cells per particle so sparse scenes stay cheap.
// g.nc < 3: the box is narrower than three cells, no cell list can be built.
void cell_grid_for(const p3d_engine *e, const DevParams &P, CellGrid &g) {
    const float reach = P.reach * 1.001f + 1.0e-4f;
    long long nc = (long long)std::floor((double)P.W / (double)reach);
    const long long cap = (long long)std::cbrt(8.0 * (double)std::max<size_t>(e->n, 64)) + 1;
    nc = std::min<long long>(std::min<long long>(nc, cap), 1024);
    g.nc = (int)std::max<long long>(nc, 0);
    g.inv_cs = (float)((double)nc / (double)P.W);
    g.half = P.half;
}

// Sorts the slots by grid cell with a counting sort (k_cell_count, three scan kernels, k_cell_scatter,
// k_cell_order: the GPU analogue of src/lib.rs:135-164).  gate/gate_value: see P3D_GATED.
int build_cells(p3d_engine *e, const DevParams &P, const float4 *pos, int *flag_to_clear, CellGrid &g,
                const int *gate = nullptr, int gate_value = 0) {
    cudaStream_t st = e->stream;
    const int ns = e->n_slots;
    cell_grid_for(e, P, g);
    const long long nc = g.nc;
    if (nc < 3) {
        if (flag_to_clear) CU(cudaMemsetAsync(flag_to_clear, 0, sizeof(int), st));
        return P3D_OK;
    }
    const size_t ncell = (size_t)(nc * nc * nc);
    if ((size_t)ns >= (size_t)kCellIndexMask) return fail(P3D_ERR_INVALID, "the cell list handles up to 2^27 slots");
    const int L = (int)ncell + 1;                       // + the ghost bin
    const int tiles = (L + kScanTile - 1) / kScanTile;
    int rc;
    for (int k = 0; k < 2; ++k) {
        if ((rc = e->ckeys[k].ensure((size_t)ns))) return rc;
        if ((rc = e->cvals[k].ensure((size_t)ns))) return rc;
    }
    if ((rc = e->crank.ensure((size_t)ns))) return rc;
    if ((rc = e->cpos.ensure((size_t)ns))) return rc;
    if ((rc = e->cell_cnt.ensure(ncell + 2))) return rc;
    if ((rc = e->cell_off.ensure(ncell + 2))) return rc;
    if ((rc = e->scan_tiles.ensure((size_t)tiles + 1))) return rc;
#ifdef P3D_BOUNDS_CHECK
    k_debug_publish<<<1, 1, 0, st>>>(0xFFFFFFFFu, (unsigned int)ncell);
#endif
    const unsigned grid = (unsigned)((ns + 255) / 256);
    CU(cudaMemsetAsync(e->cell_cnt.p, 0, (size_t)L * sizeof(uint32_t), st));
    k_cell_count<<<grid, 256, 0, st>>>(pos, ns, g, e->ckeys[0].p, e->crank.p, e->cell_cnt.p, flag_to_clear, gate, gate_value);
    k_scan_tile_sums<<<tiles, kScanThreads, 0, st>>>(e->cell_cnt.p, L, e->scan_tiles.p, gate, gate_value);
    k_scan_top<<<1, kScanThreads, 0, st>>>(e->scan_tiles.p, tiles, gate, gate_value);
    k_scan_apply<<<tiles, kScanThreads, 0, st>>>(e->cell_cnt.p, L, e->scan_tiles.p, e->cell_off.p, gate, gate_value);
    k_cell_scatter<<<grid, 256, 0, st>>>(ns, e->ckeys[0].p, e->crank.p, e->cell_off.p, e->cvals[0].p, gate, gate_value);
    k_cell_order<<<grid, 256, 0, st>>>(pos, ns, e->ckeys[0].p, e->crank.p, e->cell_off.p, e->cvals[0].p, (uint32_t)ncell,
                                       e->permuted ? e->caller_of.p : nullptr, e->ckeys[1].p, e->cvals[1].p, e->cpos.p,
                                       gate, gate_value);
    e->counters[0] += 6;
    CU(cudaGetLastError());
    return P3D_OK;
}

// Re-slots an identity layout into cell order (device-resident cell-list runs).  The caller's particles arrive in
// arbitrary spatial order, so every per-step pass of the cell sort (the count atomics, the scatter, the gather) and
// the force kernel's final scatter would touch memory at random; once the SLOTS follow the cells, and for as long as
// the particles have not drifted far, those accesses are nearly sequential.  Pure data movement: positions,
// velocities and the caller<->slot tables are permuted in place (same buffers: pointers held by CUDA graphs stay
// valid), and the order of the candidates inside a cell is the caller order either way (k_cell_order), so the
// forces - and with them every later state - are bit for bit those of the un-permuted layout.
constexpr int kReslotMin = 32768;   // particles from which re-slotting pays
constexpr int kReslotEvery = 32;    // steps between two re-slots of a long run

int reslot_by_cell(p3d_engine *e, const DevParams &P) {
    cudaStream_t st = e->stream;
    const int ns = e->n_slots;
    CellGrid g;
    int rc;
    if ((rc = e->vel_tmp.ensure((size_t)ns))) return rc;
    if ((rc = e->caller_of.ensure((size_t)ns))) return rc;
    if ((rc = e->caller_tmp.ensure((size_t)ns))) return rc;
    if ((rc = e->slot_of.ensure(std::max<size_t>(e->n, 1)))) return rc;
    if ((rc = build_cells(e, P, e->pos[e->cur].p, e->flags.p + 3, g))) return rc;
    if (g.nc < 3) return P3D_OK;
    k_reslot<<<(unsigned)((ns + 255) / 256), 256, 0, st>>>(ns, (int)e->n, e->cvals[1].p, e->vel.p,
                                                          e->permuted ? e->caller_of.p : nullptr, e->vel_tmp.p,
                                                          e->caller_tmp.p, e->slot_of.p);
    e->counters[0]++;
    CU(cudaGetLastError());
    const size_t bytes = (size_t)ns * sizeof(float4);
    CU(cudaMemcpyAsync(e->pos[e->cur].p, e->cpos.p, bytes, cudaMemcpyDeviceToDevice, st));      // cpos[k] = pos[old slot of k]
    CU(cudaMemcpyAsync(e->pos[e->cur ^ 1].p, e->cpos.p, bytes, cudaMemcpyDeviceToDevice, st));  // ghosts line up in both
    CU(cudaMemcpyAsync(e->vel.p, e->vel_tmp.p, bytes, cudaMemcpyDeviceToDevice, st));
    CU(cudaMemcpyAsync(e->caller_of.p, e->caller_tmp.p, (size_t)ns * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
    e->permuted = true;
    e->steps_since_reslot = 0;
    return P3D_OK;
}

bool reslot_applies(const p3d_engine *e) {
    return !e->typed && e->world == 1 && e->n >= (size_t)kReslotMin &&
           resolve_force_kernel(e) == P3D_FORCE_CELLS;
}

// K5: adds the reference's double-visit contributions to the ideal forces already in frc.
int launch_quirk(p3d_engine *e, const DevParams &P, const CellGrid &g, const int *flag_cur) {
    const int ns = e->n_slots;
    const int per = (ns + e->world - 1) / e->world;
    const int i0 = std::min(ns, e->rank * per), i1 = std::min(ns, i0 + per);
    if (i1 > i0 && e->n > 0) {
        const size_t sm = (size_t)P.T * P.T * sizeof(float);
        if (P.rcut)
            k_quirk_correction<true><<<(i1 - i0 + 127) / 128, 128, sm, e->stream>>>(
                e->cpos.p, e->ckeys[1].p, e->cvals[1].p, e->cell_off.p, i0, i1, g, e->frc.p, P, e->matrix.p, flag_cur,
                (unsigned long long)e->n);
        else
            k_quirk_correction<false><<<(i1 - i0 + 127) / 128, 128, sm, e->stream>>>(
                e->cpos.p, e->ckeys[1].p, e->cvals[1].p, e->cell_off.p, i0, i1, g, e->frc.p, P, e->matrix.p, flag_cur,
                (unsigned long long)e->n);
        e->counters[0]++;
    }
    CU(cudaGetLastError());
    return P3D_OK;
}

// K5 after a force kernel that did not build the cell list itself.
int launch_quirk_standalone(p3d_engine *e, const DevParams &P, const float4 *pos, const int *flag_cur) {
    CellGrid g;
    int rc;
    int *scratch_flag = e->flags.p + 3;  // build_cells clears a flag word; the step flags are already set
    if ((rc = build_cells(e, P, pos, scratch_flag, g))) return rc;
    if (g.nc < 3) return P3D_OK;  // box narrower than three cells: the correction is not available
    return launch_quirk(e, P, g, flag_cur);
}

// Force pass for the rows / slots of this shard.  Leaves total_force (src/lib.rs:177-243) in frc.
int launch_force(p3d_engine *e, const DevParams &P) {
    cudaStream_t st = e->stream;
    const int ns = e->n_slots;
    const int kind = resolve_force_kernel(e);
    int *flag_cur = e->flags.p + e->parity;
    int *flag_next = e->flags.p + (e->parity ^ 1);
    const float4 *pos = e->pos[e->cur].p;
    if (kind == P3D_FORCE_REFERENCE_ORDER) {
        CU(cudaMemsetAsync(flag_next, 0, sizeof(int), st));
        if (e->step_ev) {
            CU(cudaEventRecord(e->step_ev[1], st));
            CU(cudaEventRecord(e->step_ev[2], st));
        }
        // shard: contiguous slot range
        const int per = ((e->M + e->world - 1) / e->world) * e->B;
        const int i0 = std::min(ns, e->rank * per), i1 = std::min(ns, i0 + per);
        if (e->world > 1) CU(cudaMemsetAsync(e->frc.p, 0, (size_t)ns * sizeof(float4), st));
        if (i1 > i0) {
            k_force_ref<kRefTile><<<(i1 - i0 + kRefTile - 1) / kRefTile, kRefTile, ref_smem(kRefTile, P.T), st>>>(
                pos, ns, i0, i1, e->frc.p, P, e->matrix.p, flag_cur, -1);
            e->counters[0]++;
            e->counters[1]++;
        }
        CU(cudaGetLastError());
        return e->opt_faithful ? launch_quirk_standalone(e, P, pos, flag_cur) : P3D_OK;
    }
    if (kind == P3D_FORCE_CELLS) {
        CellGrid g;
        int rc;
        if ((rc = build_cells(e, P, pos, flag_next, g))) return rc;
        if (g.nc >= 3) {
            if (e->step_ev) CU(cudaEventRecord(e->step_ev[1], st));
            if (e->world > 1) CU(cudaMemsetAsync(e->frc.p, 0, (size_t)ns * sizeof(float4), st));
            const int per = (ns + e->world - 1) / e->world;
            const int i0 = std::min(ns, e->rank * per), i1 = std::min(ns, i0 + per);
            if (i1 > i0) {
                const size_t sm = (size_t)P.T * P.T * sizeof(float);
                const unsigned cg = (unsigned)((i1 - i0 + kCellThreads - 1) / kCellThreads);
#define P3D_CELL_ARGS e->cpos.p, e->ckeys[1].p, e->cvals[1].p, e->cell_off.p, ns, i0, i1, g, e->frc.p, P, e->matrix.p, flag_cur
                // in-box positions (flag clear): the image follows from the neighbour cell ...
                if (P.rcut) k_force_cells<true, false><<<cg, kCellThreads, sm, st>>>(P3D_CELL_ARGS);
                else        k_force_cells<false, false><<<cg, kCellThreads, sm, st>>>(P3D_CELL_ARGS);
                // ... some particle outside the box (flag set): nearest of the reference's three images per axis.  Rare,
                // so it gets a small grid (it strides over the particles) and costs nothing when it does not run.
                k_force_cells<true, true><<<std::min(cg, (unsigned)(8 * e->sm_count)), kCellThreads, sm, st>>>(P3D_CELL_ARGS);
#undef P3D_CELL_ARGS
                e->counters[0] += 2;
                e->counters[1]++;
            }
            if (e->step_ev) CU(cudaEventRecord(e->step_ev[2], st));
            CU(cudaGetLastError());
            if (e->opt_faithful) return launch_quirk(e, P, g, flag_cur);
            return P3D_OK;
        }
        // box narrower than three cells: the all-pairs path below handles it
    }
    // --- pair path ---
    if (!e->typed) {
        if (kind != P3D_FORCE_CELLS)
            return fail(P3D_ERR_INVALID, "the pair kernel needs the type-grouped layout: set the force kernel before p3d_upload");
        // cell list requested but the box is narrower than three cells: all pairs with the exact kernel
        CU(cudaMemsetAsync(flag_next, 0, sizeof(int), st));
        if (e->step_ev) {
            CU(cudaEventRecord(e->step_ev[1], st));
            CU(cudaEventRecord(e->step_ev[2], st));
        }
        const int per = ((e->M + e->world - 1) / e->world) * e->B;
        const int i0 = std::min(ns, e->rank * per), i1 = std::min(ns, i0 + per);
        if (e->world > 1) CU(cudaMemsetAsync(e->frc.p, 0, (size_t)ns * sizeof(float4), st));
        if (i1 > i0) {
            k_force_ref<kRefTile><<<(i1 - i0 + kRefTile - 1) / kRefTile, kRefTile, ref_smem(kRefTile, P.T), st>>>(
                pos, ns, i0, i1, e->frc.p, P, e->matrix.p, flag_cur, -1);
            e->counters[0]++;
            e->counters[1]++;
        }
        CU(cudaGetLastError());
        return P3D_OK;  // (K5 needs the cell list, which this box is too narrow for)
    }
    const int B = e->B;
    const float margin = std::max(1.0e-3f, 1.0e-5f * P.W);
    const float interior_limit = e->opt_block_sort ? (P.half - P.reach - margin) : -1.0f;
    const int part_ctas = ns / kPartThreads;  // n_slots is a multiple of B >= 128
    k_part_count<<<part_ctas, kPartThreads, 0, st>>>(pos, ns, interior_limit, e->cta_cnt.p, flag_next);
    k_part_scan<<<P.T, 1024, 0, st>>>(e->cta_cnt.p, e->cta_off.p, e->seg_start.p, e->seg_end.p, e->cnt.p);
    k_part_scatter<<<part_ctas, kPartThreads, 0, st>>>(pos, ns, B, e->seg_type.p, e->seg_start.p, e->seg_end.p,
                                                       e->cta_off.p, e->spos.p, e->sx.p, e->sy.p, e->sz.p,
                                                       e->sidx.p, interior_limit);
    k_part_fill<<<(ns + 255) / 256, 256, 0, st>>>(ns, B, e->seg_type.p, e->seg_start.p, e->seg_end.p, e->cnt.p,
                                                 e->spos.p, e->sx.p, e->sy.p, e->sz.p, e->sidx.p, e->bclass.p);
    CU(cudaMemsetAsync(e->frc.p, 0, (size_t)ns * sizeof(float4), st));
    e->counters[0] += 4;
    if (e->step_ev) CU(cudaEventRecord(e->step_ev[1], st));

    const int M = e->M;
    const int rows = (M - e->rank + e->world - 1) / e->world;  // rows rank, rank+world, ...
    if (rows > 0) {
        // One warp per CTA: every per-block-pair scalar is then provably warp-uniform and lives in
        // uniform registers (no register-file reads in the FFMA2 stream).  About 2048 CTAs per SM
        // (~170 waves of the 12 resident one-warp CTAs of the R = 8 kernel, ~128 waves of the 16 of the R = 4 kernel)
        // keep the tail of the last wave small; never more warps than offsets in a row.
        const int offsets = M / 2 + 1;
        int splits = (int)std::min<long long>(offsets, std::max<long long>(1, (128LL * 16 * e->sm_count + rows - 1) / rows));
        const dim3 grid((unsigned)rows * (unsigned)splits);
        // boundary x boundary kernel: forked onto the auxiliary stream so that it runs beside the pair
        // kernel (both only add into frc).  Its j-blocks are split over enough CTAs to fill the machine;
        // only BOUNDARY rows do work and their number is known on the device only, so it is estimated from
        // the boundary shell volume of a uniform cloud (a clustered cloud has fewer and merely over-splits).
        const double shell = 1.0 - std::pow(std::max(0.0, 1.0 - 2.0 * (double)P.reach / (double)P.W), 3.0);
        const long long est_rows = std::max<long long>(P.T, (long long)(shell * rows) + P.T);
        const int jsplit = (int)std::max<long long>(1, std::min<long long>(64, (8LL * e->sm_count + est_rows - 1) / est_rows));
        const dim3 bgrid((unsigned)rows, (unsigned)jsplit);
        cudaStream_t ax = e->aux_stream;
        CU(cudaEventRecord(e->ev_fork, st));
        CU(cudaStreamWaitEvent(ax, e->ev_fork, 0));
#define P3D_BXB_ARGS e->spos.p, e->sidx.p, e->bclass.p, M, e->rank, e->world, e->seg_start.p, e->seg_end.p, e->cnt.p, e->frc.p, P, e->matrix.p, flag_cur
        if (B == 128) {
            if (P.rcut) k_force_bxb<128, true><<<bgrid, 128, ref_smem(128, P.T), ax>>>(P3D_BXB_ARGS);
            else        k_force_bxb<128, false><<<bgrid, 128, ref_smem(128, P.T), ax>>>(P3D_BXB_ARGS);
        } else {
            if (P.rcut) k_force_bxb<256, true><<<bgrid, 256, ref_smem(256, P.T), ax>>>(P3D_BXB_ARGS);
            else        k_force_bxb<256, false><<<bgrid, 256, ref_smem(256, P.T), ax>>>(P3D_BXB_ARGS);
        }
#undef P3D_BXB_ARGS
        CU(cudaEventRecord(e->ev_join, ax));
        const float *sx = e->sx.p, *sy = e->sy.p, *sz = e->sz.p;
#define P3D_PAIR_ARGS sx, sy, sz, e->sidx.p, e->bclass.p, e->seg_type.p, M, e->rank, e->world, splits, e->frc.p, P, e->matrix.p, flag_cur
        // MPOS: min_pull_ratio > 0 lets the kernel fold c2*m into the matrix scalars (one FFMA2 fewer per pack)
        const bool mpos = P.m > 0.0f && P.m < 1.0f;
#define P3D_PAIR_LAUNCH(R_, RC_, MP_) k_force_pair<R_, RC_, (R_ == 8 ? 12 : 16), 1, MP_><<<grid, 32, 0, st>>>(P3D_PAIR_ARGS)
        if (B == 128) {
            if (P.rcut) { if (mpos) P3D_PAIR_LAUNCH(4, true, true); else P3D_PAIR_LAUNCH(4, true, false); }
            else        { if (mpos) P3D_PAIR_LAUNCH(4, false, true); else P3D_PAIR_LAUNCH(4, false, false); }
        } else {
            if (P.rcut) { if (mpos) P3D_PAIR_LAUNCH(8, true, true); else P3D_PAIR_LAUNCH(8, true, false); }
            else        { if (mpos) P3D_PAIR_LAUNCH(8, false, true); else P3D_PAIR_LAUNCH(8, false, false); }
        }
#undef P3D_PAIR_LAUNCH
#undef P3D_PAIR_ARGS
        if (e->step_ev) CU(cudaEventRecord(e->step_ev[2], st));
        CU(cudaStreamWaitEvent(st, e->ev_join, 0));  // join
        e->counters[0] += 2;
        e->counters[1] += 2;
    }
    else if (e->step_ev) CU(cudaEventRecord(e->step_ev[2], st));
    // Out-of-box input (flag set: the pair kernels above returned at once).  The same images are evaluated by the
    // cell list's general variant, O(N x neighbours), queued BEHIND the flag on the device (no host round trip, so a
    // device-resident multi-step run is covered too); small systems and boxes narrower than three cells take the
    // exact all-pairs kernel instead.
    {
        CellGrid g;
        cell_grid_for(e, P, g);
        if (e->n >= (size_t)kOutOfBoxCellsMin && g.nc >= 3) {
            int rc;
            if ((rc = build_cells(e, P, pos, nullptr, g, flag_cur, 1))) return rc;
            const int per = (ns + e->world - 1) / e->world;
            const int i0 = std::min(ns, e->rank * per), i1 = std::min(ns, i0 + per);
            if (i1 > i0) {
                k_force_cells<true, true><<<std::min((unsigned)((i1 - i0 + kCellThreads - 1) / kCellThreads),
                                                     (unsigned)(8 * e->sm_count)), kCellThreads,
                                            (size_t)P.T * P.T * sizeof(float), st>>>(
                    e->cpos.p, e->ckeys[1].p, e->cvals[1].p, e->cell_off.p, ns, i0, i1, g, e->frc.p, P, e->matrix.p, flag_cur);
                e->counters[0]++;
            }
        } else {
            const int per = ((M + e->world - 1) / e->world) * B;
            const int i0 = std::min(ns, e->rank * per), i1 = std::min(ns, i0 + per);
            if (i1 > i0) {
                k_force_ref<kRefTile><<<(i1 - i0 + kRefTile - 1) / kRefTile, kRefTile, ref_smem(kRefTile, P.T), st>>>(
                    pos, ns, i0, i1, e->frc.p, P, e->matrix.p, flag_cur, 1);
                e->counters[0]++;
            }
        }
    }
    CU(cudaGetLastError());
    return e->opt_faithful ? launch_quirk_standalone(e, P, pos, flag_cur) : P3D_OK;
}

int launch_integrate(p3d_engine *e, const DevParams &P, float ts) {
    const int ns = e->n_slots;
    int s0 = 0, s1 = ns;
    if (e->world > 1) {
        const int per = ((e->M + e->world - 1) / e->world) * e->B;
        s0 = std::min(ns, e->rank * per);
        s1 = std::min(ns, s0 + per);
    }
    if (s1 > s0) {
        k_integrate<<<(s1 - s0 + 255) / 256, 256, 0, e->stream>>>(e->pos[e->cur].p, e->pos[e->cur ^ 1].p, e->vel.p,
                                                                  e->frc.p, s0, s1, P, ts,
                                                                  e->flags.p + (e->parity ^ 1));
        e->counters[0]++;
        e->counters[2]++;
    }
    CU(cudaGetLastError());
    return P3D_OK;
}

int upload_matrix(p3d_engine *e, const p3d_params *prm) {
    const size_t bytes = (size_t)prm->id_count * prm->id_count * sizeof(float);
    // pageable source: staged synchronously by the runtime, so the caller's array may change afterwards
    CU(cudaMemcpyAsync(e->matrix.p, prm->attraction_matrix, bytes, cudaMemcpyHostToDevice, e->stream));
    return P3D_OK;
}

int ensure_events(p3d_engine *e, int steps) {
    const size_t need = (size_t)std::min(steps, kMaxTimedSteps) * kEv;
    while (e->ev.size() < need) {
        cudaEvent_t x;
        CU(cudaEventCreate(&x));
        e->ev.push_back(x);
    }
    for (int k = 0; k < 6; ++k)
        if (!e->ev_call[k]) CU(cudaEventCreate(&e->ev_call[k]));
    return P3D_OK;
}

int check_box_now(p3d_engine *e, const DevParams &P) {
    int *flag_cur = e->flags.p + e->parity;
    CU(cudaMemsetAsync(flag_cur, 0, sizeof(int), e->stream));
    k_check_box<<<(e->n_slots + 255) / 256, 256, 0, e->stream>>>(e->pos[e->cur].p, e->n_slots, P.half, flag_cur);
    e->counters[0]++;
    CU(cudaGetLastError());
    return P3D_OK;
}

int one_step(p3d_engine *e, const DevParams &P, float ts, bool timed, int s) {
    int rc;
    e->step_ev = timed ? &e->ev[(size_t)kEv * s] : nullptr;
    if (timed) CU(cudaEventRecord(e->step_ev[0], e->stream));
    if ((rc = launch_force(e, P))) { e->step_ev = nullptr; return rc; }
    if (timed) CU(cudaEventRecord(e->step_ev[3], e->stream));
    if ((rc = launch_integrate(e, P, ts))) { e->step_ev = nullptr; return rc; }
    if (timed) {
        CU(cudaEventRecord(e->step_ev[4], e->stream));
        e->timed_steps = s + 1;
    }
    e->step_ev = nullptr;
    e->cur ^= 1;
    e->parity ^= 1;
    return P3D_OK;
}

void drop_graph(p3d_engine *e) {
    if (e->graph_exec) cudaGraphExecDestroy(e->graph_exec);
    e->graph_exec = nullptr;
    e->graph_key.clear();
}

// Everything a captured pair of steps depends on; any change forces a re-capture.
std::vector<unsigned char> make_graph_key(const p3d_engine *e, const DevParams &P, float ts) {
    std::vector<unsigned char> k(sizeof(DevParams) + sizeof(float) + 9 * sizeof(uint64_t));
    unsigned char *p = k.data();
    std::memcpy(p, &P, sizeof(DevParams)); p += sizeof(DevParams);
    std::memcpy(p, &ts, sizeof(float)); p += sizeof(float);
    const uint64_t v[9] = {e->layout_version, (uint64_t)e->cur, (uint64_t)e->parity, (uint64_t)resolve_force_kernel(e),
                           (uint64_t)e->opt_faithful, (uint64_t)e->opt_block_sort, (uint64_t)(uintptr_t)e->stream,
                           (uint64_t)e->rank * 64 + (uint64_t)e->world, (uint64_t)e->permuted};
    std::memcpy(p, v, sizeof(v));
    return k;
}

// Captures two consecutive steps on the engine stream.  Must run after at least one ordinary step with
// the same configuration, so that every lazily allocated buffer already exists (no cudaMalloc in capture).
int capture_two_steps(p3d_engine *e, const DevParams &P, float ts) {
    drop_graph(e);
    const uint64_t before[3] = {e->counters[0], e->counters[1], e->counters[2]};
    if (cudaStreamBeginCapture(e->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
        cudaGetLastError();
        return P3D_ERR_CUDA;  // e.g. the legacy default stream cannot be captured: caller steps normally
    }
    int rc = one_step(e, P, ts, false, 0);
    if (!rc) rc = one_step(e, P, ts, false, 0);
    cudaGraph_t graph = nullptr;
    const cudaError_t ce = cudaStreamEndCapture(e->stream, &graph);
    for (int k = 0; k < 3; ++k) {
        e->graph_launches[k] = e->counters[k] - before[k];
        e->counters[k] = before[k];  // nothing ran yet
    }
    if (rc || ce != cudaSuccess || !graph) {
        cudaGetLastError();
        if (graph) cudaGraphDestroy(graph);
        return rc ? rc : P3D_ERR_CUDA;
    }
    const cudaError_t ie = cudaGraphInstantiate(&e->graph_exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ie != cudaSuccess) {
        cudaGetLastError();
        e->graph_exec = nullptr;
        return P3D_ERR_CUDA;
    }
    e->graph_key = make_graph_key(e, P, ts);
    return P3D_OK;
}

int run_steps(p3d_engine *e, const p3d_params *prm, const DevParams &P, float ts, int n_steps) {
    int rc;
    if ((rc = upload_matrix(e, prm))) return rc;
    if ((rc = check_box_now(e, P))) return rc;  // world_size may have changed since the last call
    e->timed_steps = 0;
    if (e->opt_timing && (rc = ensure_events(e, n_steps))) return rc;
    // device-resident cell-list runs keep their slots in cell order (see reslot_by_cell)
    // (not for the single step of a p3d_update, whose upload is fresh every call; but a caller that keeps the state
    // resident and steps it one step per call - a render loop - gets it from its third step on)
    const bool reslot = reslot_applies(e) && (n_steps >= 2 || e->steps_since_upload >= 2);
    auto maybe_reslot = [&]() -> int {
        if (reslot && (!e->permuted || e->steps_since_reslot >= kReslotEvery)) return reslot_by_cell(e, P);
        return P3D_OK;
    };
    if ((rc = maybe_reslot())) return rc;
    int s = 0;
    // Long untimed runs replay a two-step CUDA graph: a step is 8-15 launches, and at small N their
    // launch latency is the whole step time.
    if (e->opt_graph && !e->opt_timing && n_steps >= 6) {
        for (; s < 2; ++s)
            if ((rc = one_step(e, P, ts, false, s))) return rc;  // allocates every lazily created buffer
        e->steps_since_reslot += 2;
        if (!e->graph_exec || e->graph_key != make_graph_key(e, P, ts)) {
            if (capture_two_steps(e, P, ts) != P3D_OK) drop_graph(e);  // fall back to ordinary launches
        }
        if (e->graph_exec) {
            for (; s + 2 <= n_steps; s += 2) {
                if ((rc = maybe_reslot())) return rc;  // same buffers, same pointers: the graph stays valid
                CU(cudaGraphLaunch(e->graph_exec, e->stream));
                for (int k = 0; k < 3; ++k) e->counters[k] += e->graph_launches[k];
                e->steps_since_reslot += 2;
            }
        }
    }
    for (; s < n_steps; ++s) {
        const bool timed = e->opt_timing && s < kMaxTimedSteps;
        if ((rc = maybe_reslot())) return rc;
        if ((rc = one_step(e, P, ts, timed, s))) return rc;
        e->steps_since_reslot++;
    }
    e->steps_since_upload += n_steps;
    return P3D_OK;
}

}  // namespace

// =============================================================================================
// multi-device handle (p3d_create_multi): defined at the end of this file
static int multi_upload(p3d_engine *grp, const p3d_particle *in, size_t n, uint32_t id_count);
static int multi_step(p3d_engine *grp, const p3d_params *prm, float ts, int n_steps);
static int multi_download(p3d_engine *grp, p3d_particle *out, size_t n);
static int multi_download_forces(p3d_engine *grp, float *out_xyz, size_t n);
static int multi_not_supported(const char *call) {
    return fail(P3D_ERR_INVALID, "%s is not available on a multi-device handle (p3d_create_multi): it drives its devices "
                "itself - use p3d_upload / p3d_step / p3d_update / p3d_download", call);
}

extern "C" {

int p3d_abi_version(void) { return P3D_ABI_VERSION; }

int p3d_debug_bounds_violations(p3d_engine *e, unsigned long long *count) {
    if (!e || !count) return fail(P3D_ERR_INVALID, "null argument");
#ifdef P3D_BOUNDS_CHECK
    CU(cudaSetDevice(e->device));
    CU(cudaStreamSynchronize(e->stream));
    CU(cudaMemcpyFromSymbol(count, g_p3d_oob, sizeof(unsigned long long)));
#else
    *count = ~0ull;  // not a self-checking build
#endif
    return P3D_OK;
}

const char *p3d_last_error(void) { return g_last_error.c_str(); }

int p3d_create(int device, p3d_engine **out) {
    if (!out) return fail(P3D_ERR_INVALID, "out is null");
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
        cudaGetLastError();
        return fail(P3D_ERR_NO_DEVICE, "no CUDA device: this engine has no CPU fallback");
    }
    if (device < 0 || device >= count) return fail(P3D_ERR_NO_DEVICE, "device %d not in 0..%d", device, count - 1);
    CU(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(P3D_ERR_NO_DEVICE, "device %d is sm_%d%d; kernels are built for sm_100a only", device, prop.major,
                    prop.minor);
    cudaStream_t stream = nullptr, aux = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    int prio_lo = 0, prio_hi = 0;
    // highest priority for the auxiliary stream: its few CTAs are placed as soon as slots free up, instead of
    // queueing behind the tens of thousands of pair-kernel CTAs
    cudaError_t ce = cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking);
    if (ce == cudaSuccess) ce = cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    if (ce == cudaSuccess) ce = cudaStreamCreateWithPriority(&aux, cudaStreamNonBlocking, prio_hi);
    if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming);
    if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming);
    uint32_t *pin = nullptr;
    if (ce == cudaSuccess) ce = cudaMallocHost(&pin, kPinWords * sizeof(uint32_t));
    if (ce != cudaSuccess) {  // nothing half-made is left behind
        if (ev_join) cudaEventDestroy(ev_join);
        if (ev_fork) cudaEventDestroy(ev_fork);
        if (aux) cudaStreamDestroy(aux);
        if (stream) cudaStreamDestroy(stream);
        return fail(P3D_ERR_CUDA, "creating the engine's streams/events failed: %s", cudaGetErrorString(ce));
    }
    p3d_engine *e = new p3d_engine();
    e->host_pin = pin;
    e->aux_stream = aux;
    e->ev_fork = ev_fork;
    e->ev_join = ev_join;
    e->device = device;
    e->sm_count = prop.multiProcessorCount;
    e->own_stream = stream;
    e->stream = stream;
    *out = e;
    return P3D_OK;
}

void p3d_destroy(p3d_engine *e) {
    if (!e) return;
    if (!e->members.empty()) {
        for (p3d_engine *m : e->members) {
            cudaSetDevice(m->device);
            cudaStreamSynchronize(m->stream);
        }
        for (size_t g = 0; g < e->ev_bar.size(); ++g) {
            cudaSetDevice(e->members[g % e->members.size()]->device);
            cudaEventDestroy(e->ev_bar[g]);
        }
        for (p3d_engine *m : e->members) p3d_destroy(m);
        delete e;
        return;
    }
    cudaSetDevice(e->device);
    cudaStreamSynchronize(e->stream);
    p3d_ipc_close(e);
    for (auto &b : e->pos) b.release();
    e->vel.release(); e->frc.release(); e->spos.release();
    e->slot_of.release(); e->sidx.release(); e->type_cnt.release();
    e->seg_type.release(); e->bclass.release();
    e->seg_start.release(); e->seg_end.release(); e->cnt.release(); e->cta_cnt.release(); e->cta_off.release();
    for (auto &b : e->ckeys) b.release();
    for (auto &b : e->cvals) b.release();
    e->cell_off.release(); e->cpos.release(); e->crank.release(); e->cell_cnt.release(); e->scan_tiles.release(); e->caller_of.release(); e->caller_tmp.release(); e->vel_tmp.release();
    e->aos.release(); e->fout.release(); e->render.release(); e->sx.release(); e->sy.release(); e->sz.release(); e->matrix.release(); e->flags.release(); e->diag.release();
    drop_graph(e);
    for (auto x : e->ev) cudaEventDestroy(x);
    for (auto x : e->ev_call) if (x) cudaEventDestroy(x);
    if (e->ev_fork) cudaEventDestroy(e->ev_fork);
    if (e->ev_join) cudaEventDestroy(e->ev_join);
    if (e->aux_stream) cudaStreamDestroy(e->aux_stream);
    if (e->own_stream) cudaStreamDestroy(e->own_stream);
    if (e->host_pin) cudaFreeHost(e->host_pin);
    for (int b = 0; b < 2; ++b) {
        if (e->stage_ev[b]) cudaEventDestroy(e->stage_ev[b]);
        if (e->stage_pin[b]) cudaFreeHost(e->stage_pin[b]);
    }
    delete e;
}

int p3d_set_stream(p3d_engine *e, void *cuda_stream) {
    if (!e) return fail(P3D_ERR_INVALID, "engine is null");
    if (!e->members.empty()) return multi_not_supported("p3d_set_stream");
    e->stream = cuda_stream ? (cudaStream_t)cuda_stream : e->own_stream;
    return P3D_OK;
}

int p3d_set_option(p3d_engine *e, int option, int value) {
    if (!e) return fail(P3D_ERR_INVALID, "engine is null");
    if (option == P3D_OPT_TIMING && value && !e->members.empty()) return multi_not_supported("per-kernel timing (P3D_OPT_TIMING)");
    for (p3d_engine *m : e->members) {  // a multi-device handle keeps its members' options identical
        const int rc = p3d_set_option(m, option, value);
        if (rc) return rc;
    }
    switch (option) {
        case P3D_OPT_FORCE_KERNEL:
            if (value < P3D_FORCE_AUTO || value > P3D_FORCE_CELLS) return fail(P3D_ERR_INVALID, "bad force kernel %d", value);
            e->opt_force = value;
            return P3D_OK;
        case P3D_OPT_TIMING: e->opt_timing = value ? 1 : 0; return P3D_OK;
        case P3D_OPT_BLOCK_SORT: e->opt_block_sort = value ? 1 : 0; return P3D_OK;
        case P3D_OPT_FAITHFUL: e->opt_faithful = value ? 1 : 0; return P3D_OK;
        case P3D_OPT_GRAPH: e->opt_graph = value ? 1 : 0; return P3D_OK;
        case P3D_OPT_BLOCK_SIZE:
            if (value != 0 && value != 128 && value != 256) return fail(P3D_ERR_INVALID, "block size must be 0 (auto), 128 or 256");
            e->B_next = value;
            return P3D_OK;
        default: return fail(P3D_ERR_INVALID, "unknown option %d", option);
    }
}

int p3d_get_option(p3d_engine *e, int option, int *value) {
    if (!e || !value) return fail(P3D_ERR_INVALID, "null argument");
    switch (option) {
        case P3D_OPT_FORCE_KERNEL: *value = e->opt_force; return P3D_OK;
        case P3D_OPT_TIMING: *value = e->opt_timing; return P3D_OK;
        case P3D_OPT_BLOCK_SORT: *value = e->opt_block_sort; return P3D_OK;
        case P3D_OPT_BLOCK_SIZE: *value = e->B_next; return P3D_OK;
        case P3D_OPT_FAITHFUL: *value = e->opt_faithful; return P3D_OK;
        case P3D_OPT_GRAPH: *value = e->opt_graph; return P3D_OK;
        default: return fail(P3D_ERR_INVALID, "unknown option %d", option);
    }
}

// Upload, phase 2: layout + pack from the staged AoS array.  `async_counts_done`: the pair layout's count kernels
// were already queued (multi-device handle: all members count concurrently).  Returns with the stream idle.
static int upload_commit(p3d_engine *e, size_t n, uint32_t id_count, bool counts_queued) {
    int rc;
    const bool pair = resolve_force_kernel_for(e, n) == P3D_FORCE_PAIR;
    if (pair) {
        if (!counts_queued && (rc = layout_count_async(e, n, id_count))) return rc;
        CU(cudaStreamSynchronize(e->stream));
        if ((rc = layout_finish(e, n, id_count))) return rc;
    } else {
        if ((rc = build_layout_identity(e, n, id_count))) return rc;
    }
    if ((rc = launch_pack(e, n))) return rc;
    const bool timed = e->upload_timed;
    e->upload_timed = false;
    if (timed) CU(cudaEventRecord(e->ev_call[2], e->stream));
    // The same sync brings back the device-side id check of the identity layout.
    e->host_pin[kPinWords - 1] = 0u;
    CU(cudaMemcpyAsync(e->host_pin + kPinWords - 1, e->flags.p + 2, sizeof(int), cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    if (timed) {
        CU(cudaEventElapsedTime(&e->last_ms[5], e->ev_call[0], e->ev_call[1]));
        CU(cudaEventElapsedTime(&e->last_ms[2], e->ev_call[1], e->ev_call[2]));
    }
    if (!pair && e->host_pin[kPinWords - 1]) {
        e->n = 0;
        e->n_slots = 0;
        return fail(P3D_ERR_BAD_ID, "a particle has id >= id_count %u (src/lib.rs:225-228)", id_count);
    }
    return P3D_OK;
}

static int upload_args_ok(p3d_engine *e, const void *in, size_t n, uint32_t id_count) {
    if (!e) return fail(P3D_ERR_INVALID, "engine is null");
    if (n && !in) return fail(P3D_ERR_INVALID, "in is null");
    if (id_count == 0 || id_count > P3D_MAX_TYPES)
        return fail(P3D_ERR_INVALID, "id_count %u outside 1..%d", id_count, P3D_MAX_TYPES);
    return P3D_OK;
}

int p3d_upload(p3d_engine *e, const p3d_particle *in, size_t n, uint32_t id_count) {
    int rc;
    if ((rc = upload_args_ok(e, in, n, id_count))) return rc;
    if (!e->members.empty()) return multi_upload(e, in, n, id_count);
    CU(cudaSetDevice(e->device));
    e->upload_timed = e->opt_timing != 0;
    if (e->upload_timed) {
        if ((rc = ensure_events(e, 1))) return rc;
        CU(cudaEventRecord(e->ev_call[0], e->stream));
    }
    // The pair kernel needs type-grouped slots (a counting sort by type id, done on the device); every other
    // kernel runs on the identity layout.  Either way there is no host pass over the particles.
    if ((rc = stage_input(e, in, 0, n, n))) return rc;
    if (e->upload_timed) CU(cudaEventRecord(e->ev_call[1], e->stream));
    e->n_staged = 0;
    e->T_staged = 0;  // (a plain upload discards whatever p3d_upload_part had staged)
    // `in` may be pageable or reused by the caller: upload_commit returns with the stream idle, i.e. after the copy
    return upload_commit(e, n, id_count, false);
}

// Sharded upload for one-process-per-GPU drivers: every rank copies only ITS part of the caller's array over its
// own PCIe link, the driver all-gathers the staging array over NVLink, then every rank builds the layout.
int p3d_upload_part(p3d_engine *e, const p3d_particle *part, size_t i_begin, size_t i_end, size_t n, uint32_t id_count) {
    int rc;
    if ((rc = upload_args_ok(e, part, i_end > i_begin ? 1 : 0, id_count))) return rc;
    if (!e->members.empty() || e->is_member) return fail(P3D_ERR_INVALID, "p3d_upload_part on a multi-device handle: use p3d_upload");
    if (i_begin > i_end || i_end > n) return fail(P3D_ERR_INVALID, "bad part [%zu, %zu) of %zu", i_begin, i_end, n);
    if (e->T_staged != 0 && (e->n_staged != n || e->T_staged != id_count))
        return fail(P3D_ERR_INVALID, "p3d_upload_part: an upload of %zu particles (id_count %u) is being staged; commit it "
                    "before staging one of %zu (id_count %u)", e->n_staged, e->T_staged, n, id_count);
    CU(cudaSetDevice(e->device));
    e->upload_timed = e->opt_timing != 0;
    if (e->upload_timed) {
        if ((rc = ensure_events(e, 1))) return rc;
        CU(cudaEventRecord(e->ev_call[0], e->stream));
    }
    if ((rc = stage_input(e, part, i_begin, i_end, n))) return rc;
    if (e->upload_timed) CU(cudaEventRecord(e->ev_call[1], e->stream));
    e->n_staged = n;
    e->T_staged = id_count;
    CU(cudaStreamSynchronize(e->stream));  // `part` may be pageable or reused by the caller
    return P3D_OK;
}

int p3d_upload_commit(p3d_engine *e) {
    if (!e) return fail(P3D_ERR_INVALID, "engine is null");
    if (e->T_staged == 0) return fail(P3D_ERR_INVALID, "nothing staged: call p3d_upload_part first");
    CU(cudaSetDevice(e->device));
    const size_t n = e->n_staged;
    const uint32_t T = e->T_staged;
    e->n_staged = 0;
    e->T_staged = 0;
    return upload_commit(e, n, T, false);
}

int p3d_step(p3d_engine *e, const p3d_params *prm, float ts, int n_steps) {
    if (!e) return fail(P3D_ERR_INVALID, "engine is null");
    if (n_steps < 0) return fail(P3D_ERR_INVALID, "n_steps < 0");
    if (!e->members.empty()) return multi_step(e, prm, ts, n_steps);
    DevParams P;
    int rc;
    if ((rc = canonicalise(prm, P))) return rc;
    if (e->n_slots == 0) return fail(P3D_ERR_INVALID, "no particles uploaded");
    if (prm->id_count != e->T)  // (also for an empty upload: the partition kernels index per-type tables of size T)
        return fail(P3D_ERR_INVALID, "id_count %u differs from the uploaded layout (%u): upload again", prm->id_count,
                    e->T);
    if (e->world > 1) return fail_sharded(e, "p3d_step");
    CU(cudaSetDevice(e->device));
    return run_steps(e, prm, P, ts, n_steps);
}

int p3d_sync(p3d_engine *e) {
    if (!e) return fail(P3D_ERR_INVALID, "engine is null");
    for (p3d_engine *m : e->members) {
        const int rc = p3d_sync(m);
        if (rc) return rc;
    }
    if (!e->members.empty()) return P3D_OK;
    CU(cudaSetDevice(e->device));
    CU(cudaStreamSynchronize(e->stream));
    return P3D_OK;
}

// Callers [i_begin, i_end) of the resident state -> host (out_part[0] = particle i_begin).
static int download_range(p3d_engine *e, p3d_particle *out_part, size_t i_begin, size_t i_end, bool sync) {
    CU(cudaSetDevice(e->device));
    int rc;
    const size_t cnt = i_end - i_begin;
    if (e->opt_timing) {
        if ((rc = ensure_events(e, 1))) return rc;
        CU(cudaEventRecord(e->ev_call[3], e->stream));
    }
    if (cnt) {
        k_unpack<<<(unsigned)((cnt + 255) / 256), 256, 0, e->stream>>>(e->pos[e->cur].p, e->vel.p,
                                                                       slot_map(e), e->aos.p,
                                                                       (int)i_begin, (int)i_end);
        e->counters[0]++;
        CU(cudaGetLastError());
    }
    if (e->opt_timing) CU(cudaEventRecord(e->ev_call[4], e->stream));
    bool arrived = false;
    if (cnt && (rc = copy_d2h(e, out_part, e->aos.p + i_begin * 7, cnt * sizeof(p3d_particle), &arrived))) return rc;
    if (e->opt_timing) CU(cudaEventRecord(e->ev_call[5], e->stream));
    if (!sync) return P3D_OK;
    CU(cudaStreamSynchronize(e->stream));
    if (e->opt_timing) {
        CU(cudaEventElapsedTime(&e->last_ms[3], e->ev_call[3], e->ev_call[4]));
        CU(cudaEventElapsedTime(&e->last_ms[6], e->ev_call[4], e->ev_call[5]));
    }
    return P3D_OK;
}

int p3d_download(p3d_engine *e, p3d_particle *out, size_t n) {
    if (!e) return fail(P3D_ERR_INVALID, "engine is null");
    if (!e->members.empty()) return multi_download(e, out, n);
    if (n != e->n) return fail(P3D_ERR_INVALID, "n=%zu but %zu particles are resident", n, e->n);
    if (n && !out) return fail(P3D_ERR_INVALID, "out is null");
    return download_range(e, out, 0, n, true);
}

int p3d_download_part(p3d_engine *e, p3d_particle *out_part, size_t i_begin, size_t i_end) {
    if (!e) return fail(P3D_ERR_INVALID, "engine is null");
    if (!e->members.empty() || e->is_member) return fail(P3D_ERR_INVALID, "p3d_download_part on a multi-device handle: use p3d_download");
    if (i_begin > i_end || i_end > e->n) return fail(P3D_ERR_INVALID, "bad part [%zu, %zu) of %zu resident particles", i_begin, i_end, e->n);
    if (i_end > i_begin && !out_part) return fail(P3D_ERR_INVALID, "out is null");
    return download_range(e, out_part, i_begin, i_end, true);
}

int p3d_download_render(p3d_engine *e, float world_size, void *out, size_t out_bytes, size_t n) {
    if (!e) return fail(P3D_ERR_INVALID, "engine is null");
    if (!e->members.empty()) {  // every device holds the whole state after a step
        int rc = p3d_sync(e);
        return rc ? rc : p3d_download_render(e->members[0], world_size, out, out_bytes, n);
    }
    if (n != e->n) return fail(P3D_ERR_INVALID, "n=%zu but %zu particles are resident", n, e->n);
    if (!out || out_bytes < 16 + 32 * n) return fail(P3D_ERR_INVALID, "render buffer needs %zu bytes", 16 + 32 * n);
    // header of WGSL `struct Particles` (src/bin/particles.wgsl:8-12): world_size f32 @0, length u32 @4,
    // the runtime array starts at its 16-byte alignment
    unsigned char *o = static_cast<unsigned char *>(out);
    std::memset(o, 0, 16);
    std::memcpy(o, &world_size, 4);
    const uint32_t len = (uint32_t)n;
    std::memcpy(o + 4, &len, 4);
    if (!n) return P3D_OK;
    CU(cudaSetDevice(e->device));
    int rc;
    if ((rc = e->render.ensure(2 * n))) return rc;
    k_unpack_render<<<(unsigned)((n + 255) / 256), 256, 0, e->stream>>>(e->pos[e->cur].p, e->vel.p,
                                                                        slot_map(e),
                                                                        e->render.p, (int)n);
    e->counters[0]++;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(o + 16, e->render.p, 32 * n, cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    return P3D_OK;
}

int p3d_download_forces(p3d_engine *e, float *out_xyz, size_t n) {
    if (!e) return fail(P3D_ERR_INVALID, "engine is null");
    if (!e->members.empty()) return multi_download_forces(e, out_xyz, n);
    if (n != e->n) return fail(P3D_ERR_INVALID, "n=%zu but %zu particles are resident", n, e->n);
    if (!n) return P3D_OK;
    if (!out_xyz) return fail(P3D_ERR_INVALID, "out is null");
    CU(cudaSetDevice(e->device));
    int rc;
    if ((rc = e->fout.ensure(n * 3))) return rc;
    k_unpack_forces<<<(unsigned)((n + 255) / 256), 256, 0, e->stream>>>(e->frc.p, slot_map(e),
                                                                        e->fout.p, (int)n);
    e->counters[0]++;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(out_xyz, e->fout.p, n * 3 * sizeof(float), cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    return P3D_OK;
}

int p3d_update(p3d_engine *e, const p3d_params *prm, float ts, const p3d_particle *in, p3d_particle *out, size_t n) {
    if (!e) return fail(P3D_ERR_INVALID, "engine is null");
    DevParams P;
    int rc;
    if ((rc = canonicalise(prm, P))) return rc;  // src/lib.rs:132 comes first in the reference too
    if (n == 0) return P3D_OK;                   // src/lib.rs:135-171 with an empty Vec is a no-op
    if (!in || !out) return fail(P3D_ERR_INVALID, "in/out is null");
    if (!e->members.empty()) {
        if ((rc = multi_upload(e, in, n, prm->id_count))) return rc;
        if ((rc = multi_step(e, prm, ts, 1))) return rc;
        return multi_download(e, out, n);
    }
    if (e->world > 1) return fail_sharded(e, "p3d_update");
    CU(cudaSetDevice(e->device));
    if ((rc = p3d_upload(e, in, n, prm->id_count))) return rc;
    if ((rc = run_steps(e, prm, P, ts, 1))) return rc;
    if ((rc = p3d_download(e, out, n))) return rc;
    return P3D_OK;
}

int p3d_diagnostics(p3d_engine *e, double out[8]) {
    if (!e || !out) return fail(P3D_ERR_INVALID, "null argument");
    if (!e->members.empty()) {
        int rc = p3d_sync(e);
        return rc ? rc : p3d_diagnostics(e->members[0], out);
    }
    if (e->n_slots == 0) return fail(P3D_ERR_INVALID, "no particles uploaded");
    CU(cudaSetDevice(e->device));
    CU(cudaMemsetAsync(e->diag.p, 0, 8 * sizeof(double), e->stream));
    const int blocks = std::min(4 * e->sm_count, (e->n_slots + 255) / 256);
    k_diag<<<blocks, 256, 0, e->stream>>>(e->pos[e->cur].p, e->vel.p, e->n_slots, e->diag.p);
    e->counters[0]++;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(out, e->diag.p, 8 * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    return P3D_OK;
}

int p3d_get_timing(p3d_engine *e, float ms[12]) {
    if (!e || !ms) return fail(P3D_ERR_INVALID, "null argument");
    if (!e->members.empty()) return multi_not_supported("p3d_get_timing");
    CU(cudaSetDevice(e->device));
    CU(cudaStreamSynchronize(e->stream));
    float acc[4] = {0.f, 0.f, 0.f, 0.f};  // partition, pair, rest of force, integrate
    for (int s = 0; s < e->timed_steps; ++s) {
        for (int k = 0; k < 4; ++k) {
            float a = 0.f;
            CU(cudaEventElapsedTime(&a, e->ev[(size_t)kEv * s + k], e->ev[(size_t)kEv * s + k + 1]));
            acc[k] += a;
        }
    }
    e->last_ms[0] = acc[0] + acc[1] + acc[2];
    e->last_ms[1] = acc[3];
    e->last_ms[4] = acc[0];
    e->last_ms[7] = acc[0] + acc[1] + acc[2] + acc[3];
    e->last_ms[8] = acc[1];
    e->last_ms[9] = acc[2];
    e->last_ms[10] = (float)e->timed_steps;
    std::memcpy(ms, e->last_ms, sizeof(e->last_ms));
    return P3D_OK;
}

int p3d_get_counters(p3d_engine *e, uint64_t out[4]) {
    if (!e || !out) return fail(P3D_ERR_INVALID, "null argument");
    if (!e->members.empty()) {  // launches on all devices
        for (int k = 0; k < 4; ++k) out[k] = 0;
        for (const p3d_engine *m : e->members)
            for (int k = 0; k < 4; ++k) out[k] += m->counters[k];
        return P3D_OK;
    }
    std::memcpy(out, e->counters, sizeof(e->counters));
    return P3D_OK;
}

int p3d_device_buffer(p3d_engine *e, int which, void **dev_ptr, size_t *n_slots) {
    if (!e || !dev_ptr) return fail(P3D_ERR_INVALID, "null argument");
    if (!e->members.empty()) return multi_not_supported("p3d_device_buffer");
    if (which == P3D_BUF_AOS) {  // staging array of the sharded upload: 7 words per caller index
        const size_t n = e->n_staged ? e->n_staged : e->n;
        if (!e->aos.p || n == 0) return fail(P3D_ERR_INVALID, "no particles staged (p3d_upload_part) or uploaded");
        *dev_ptr = e->aos.p;
        if (n_slots) *n_slots = staged_part(n, e->world) * (size_t)e->world;
        return P3D_OK;
    }
    if (e->n_slots == 0) return fail(P3D_ERR_INVALID, "no particles uploaded");
    switch (which) {
        case P3D_BUF_POS: *dev_ptr = e->pos[e->cur].p; break;
        case P3D_BUF_POS_NEXT: *dev_ptr = e->pos[e->cur ^ 1].p; break;
        case P3D_BUF_VEL: *dev_ptr = e->vel.p; break;
        case P3D_BUF_FORCE: *dev_ptr = e->frc.p; break;
        default: return fail(P3D_ERR_INVALID, "unknown buffer %d", which);
    }
    if (n_slots) *n_slots = (size_t)e->n_slots;
    return P3D_OK;
}

int p3d_set_shard(p3d_engine *e, int rank, int world) {
    if (!e) return fail(P3D_ERR_INVALID, "engine is null");
    if (world < 1 || rank < 0 || rank >= world) return fail(P3D_ERR_INVALID, "bad shard %d/%d", rank, world);
    if (!e->members.empty() || e->is_member) return multi_not_supported("p3d_set_shard");
    if (world != e->world && e->n_slots != 0) {
        // the slot layout was padded to B * world at upload time: with another world the force / integrate splits
        // would disagree and the shards become unequal.  The resident state is dropped; upload again.
        CU(cudaSetDevice(e->device));
        CU(cudaStreamSynchronize(e->stream));
        e->n = 0;
        e->n_slots = 0;
        e->M = 0;
        e->layout_version++;
    }
    e->rank = rank;
    e->world = world;
    return P3D_OK;
}

int p3d_shard_range(p3d_engine *e, size_t *slot_begin, size_t *slot_end) {
    if (!e || !slot_begin || !slot_end) return fail(P3D_ERR_INVALID, "null argument");
    if (!e->members.empty()) return multi_not_supported("p3d_shard_range");
    const int per = ((e->M + e->world - 1) / e->world) * e->B;
    const int s0 = std::min(e->n_slots, e->rank * per);
    *slot_begin = (size_t)s0;
    *slot_end = (size_t)std::min(e->n_slots, s0 + per);
    return P3D_OK;
}

int p3d_shard_force(p3d_engine *e, const p3d_params *prm) {
    if (!e) return fail(P3D_ERR_INVALID, "engine is null");
    if (!e->members.empty()) return multi_not_supported("p3d_shard_force");
    DevParams P;
    int rc;
    if ((rc = canonicalise(prm, P))) return rc;
    if (e->n_slots == 0) return fail(P3D_ERR_INVALID, "no particles uploaded");
    if (prm->id_count != e->T)
        return fail(P3D_ERR_INVALID, "id_count %u differs from the uploaded layout (%u): upload again", prm->id_count, e->T);
    CU(cudaSetDevice(e->device));
    if ((rc = upload_matrix(e, prm))) return rc;
    // the out-of-box flag written by this rank's integrate covers only its own shard; after the
    // driver's all-gather every rank re-derives it from the full position array
    if ((rc = check_box_now(e, P))) return rc;
    return launch_force(e, P);
}

int p3d_shard_integrate(p3d_engine *e, const p3d_params *prm, float ts) {
    if (!e) return fail(P3D_ERR_INVALID, "engine is null");
    if (!e->members.empty()) return multi_not_supported("p3d_shard_integrate");
    DevParams P;
    int rc;
    if ((rc = canonicalise(prm, P))) return rc;
    CU(cudaSetDevice(e->device));
    return launch_integrate(e, P, ts);
}

// ---- peer memory for the fused integrate kernel ----
constexpr int kPeerBufs = 4;  // frc, pos[0], pos[1], vel

int p3d_ipc_export(p3d_engine *e, unsigned char *handles /* 4 * 64 bytes */) {
    if (!e || !handles) return fail(P3D_ERR_INVALID, "null argument");
    if (!e->members.empty() || e->is_member) return fail(P3D_ERR_INVALID, "a multi-device handle needs no IPC");
    if (e->n_slots == 0) return fail(P3D_ERR_INVALID, "no particles uploaded");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    CU(cudaSetDevice(e->device));
    void *ptrs[kPeerBufs] = {e->frc.p, e->pos[0].p, e->pos[1].p, e->vel.p};
    for (int k = 0; k < kPeerBufs; ++k) {
        cudaIpcMemHandle_t h;
        CU(cudaIpcGetMemHandle(&h, ptrs[k]));
        std::memcpy(handles + 64 * k, &h, 64);
    }
    e->ipc_exported = true;
    return P3D_OK;
}

int p3d_ipc_close(p3d_engine *e) {
    if (!e) return fail(P3D_ERR_INVALID, "engine is null");
    if (e->is_member) return P3D_OK;  // peer tables of a multi-device handle are plain pointers
    cudaSetDevice(e->device);
    for (int g = 0; g < 8; ++g) {
        if (e->peer_open[g])
            for (int k = 0; k < kPeerBufs; ++k)
                if (e->peer_ptr[g][k]) cudaIpcCloseMemHandle(e->peer_ptr[g][k]);
        e->peer_open[g] = false;
        for (int k = 0; k < kPeerBufs; ++k) e->peer_ptr[g][k] = nullptr;
    }
    e->peers_ready = false;
    e->ipc_exported = false;
    return P3D_OK;
}

int p3d_ipc_import(p3d_engine *e, int world, const unsigned char *all_handles /* world * 4 * 64 bytes */) {
    if (!e || !all_handles) return fail(P3D_ERR_INVALID, "null argument");
    if (!e->members.empty() || e->is_member) return fail(P3D_ERR_INVALID, "a multi-device handle needs no IPC");
    if (world != e->world || world > 8) return fail(P3D_ERR_INVALID, "world %d does not match the shard (%d) or exceeds 8", world, e->world);
    CU(cudaSetDevice(e->device));
    const bool exported = e->ipc_exported;
    p3d_ipc_close(e);
    e->ipc_exported = exported;
    for (int g = 0; g < world; ++g) {
        if (g == e->rank) {
            e->peer_ptr[g][0] = e->frc.p;
            e->peer_ptr[g][1] = e->pos[0].p;
            e->peer_ptr[g][2] = e->pos[1].p;
            e->peer_ptr[g][3] = e->vel.p;
            continue;
        }
        for (int k = 0; k < kPeerBufs; ++k) {
            cudaIpcMemHandle_t h;
            std::memcpy(&h, all_handles + (size_t)(g * kPeerBufs + k) * 64, 64);
            CU(cudaIpcOpenMemHandle(&e->peer_ptr[g][k], h, cudaIpcMemLazyEnablePeerAccess));
        }
        e->peer_open[g] = true;
    }
    e->peers_ready = true;
    return P3D_OK;
}

static int launch_integrate_fused(p3d_engine *e, const DevParams &P, float ts) {
    const int ns = e->n_slots;
    const int per = ((e->M + e->world - 1) / e->world) * e->B;
    const int s0 = std::min(ns, e->rank * per), s1 = std::min(ns, s0 + per);
    PeerTable pt;
    for (int g = 0; g < 8; ++g) {
        pt.frc[g] = (const float4 *)e->peer_ptr[g][0];
        pt.pos_next[g] = (float4 *)e->peer_ptr[g][1 + (e->cur ^ 1)];
        pt.vel[g] = (float4 *)e->peer_ptr[g][3];
    }
    if (s1 > s0) {
        k_integrate_fused<<<(s1 - s0 + 255) / 256, 256, 0, e->stream>>>(e->pos[e->cur].p, e->vel.p, pt, e->world, s0, s1,
                                                                        P, ts, e->flags.p + (e->parity ^ 1));
        e->counters[0]++;
        e->counters[2]++;
    }
    CU(cudaGetLastError());
    return P3D_OK;
}

int p3d_shard_integrate_fused(p3d_engine *e, const p3d_params *prm, float ts) {
    if (!e) return fail(P3D_ERR_INVALID, "engine is null");
    if (!e->members.empty()) return fail(P3D_ERR_INVALID, "p3d_shard_* on a multi-device handle: use p3d_step / p3d_update");
    if (!e->peers_ready) return fail(P3D_ERR_INVALID, "peer buffers not imported (p3d_ipc_import)");
    DevParams P;
    int rc;
    if ((rc = canonicalise(prm, P))) return rc;
    CU(cudaSetDevice(e->device));
    return launch_integrate_fused(e, P, ts);
}

int p3d_shard_commit(p3d_engine *e) {
    if (!e) return fail(P3D_ERR_INVALID, "engine is null");
    if (!e->members.empty()) return multi_not_supported("p3d_shard_commit");
    e->cur ^= 1;
    e->parity ^= 1;
    return P3D_OK;
}


// Caller index -> slot of the current layout (identity unless the pair kernel's type-grouped layout is active).
// Lets a test or bench pick particles from every type segment / every rank's slot range.
int p3d_slot_of(p3d_engine *e, uint32_t *out, size_t n) {
    if (!e) return fail(P3D_ERR_INVALID, "engine is null");
    if (!e->members.empty()) e = e->members[0];  // all devices hold the same layout
    if (n != e->n) return fail(P3D_ERR_INVALID, "n=%zu but %zu particles are resident", n, e->n);
    if (!n) return P3D_OK;
    if (!out) return fail(P3D_ERR_INVALID, "out is null");
    CU(cudaSetDevice(e->device));
    if (slot_map(e)) {
        CU(cudaMemcpyAsync(out, e->slot_of.p, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, e->stream));
        CU(cudaStreamSynchronize(e->stream));
    } else {
        for (size_t i = 0; i < n; ++i) out[i] = (uint32_t)i;
    }
    return P3D_OK;
}

// ---- one handle, several devices of one node (p3d_create(devices[], n_dev)) ----
// One member engine per entry of `devices` (rank g of n_dev); the calling thread drives them all.  The same
// device may be listed more than once (the members then share it: useful on a one-GPU box, no speed-up).
int p3d_create_multi(const int *devices, int n_dev, p3d_engine **out) {
    if (!out) return fail(P3D_ERR_INVALID, "out is null");
    *out = nullptr;
    if (!devices || n_dev < 1 || n_dev > 8) return fail(P3D_ERR_INVALID, "p3d_create_multi needs 1..8 devices");
    if (n_dev == 1) return p3d_create(devices[0], out);
    p3d_engine *grp = new p3d_engine();
    grp->device = devices[0];
    int rc = P3D_OK;
    for (int g = 0; g < n_dev && !rc; ++g) {
        p3d_engine *m = nullptr;
        rc = p3d_create(devices[g], &m);
        if (rc) break;
        m->rank = g;
        m->world = n_dev;
        m->is_member = true;
        grp->members.push_back(m);
    }
    // peer access between every pair of distinct devices (k_integrate_fused loads and stores peer memory directly)
    for (int g = 0; g < n_dev && !rc; ++g) {
        for (int h = 0; h < n_dev && !rc; ++h) {
            if (devices[h] == devices[g]) continue;
            int can = 0;
            cudaError_t ce = cudaDeviceCanAccessPeer(&can, devices[g], devices[h]);
            if (ce != cudaSuccess || !can) {
                cudaGetLastError();
                rc = fail(P3D_ERR_CUDA, "device %d cannot access device %d's memory (no NVLink / PCIe peer path)", devices[g], devices[h]);
                break;
            }
            cudaSetDevice(devices[g]);
            ce = cudaDeviceEnablePeerAccess(devices[h], 0);
            if (ce == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); ce = cudaSuccess; }
            if (ce != cudaSuccess) rc = fail(P3D_ERR_CUDA, "cudaDeviceEnablePeerAccess(%d -> %d): %s", devices[g], devices[h], cudaGetErrorString(ce));
        }
    }
    // two events per member: the barrier after the force pass and the one after the fused integrate
    for (int k = 0; k < 2 * n_dev && !rc; ++k) {
        cudaEvent_t ev = nullptr;
        cudaSetDevice(devices[k % n_dev]);
        if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) {
            rc = fail(P3D_ERR_CUDA, "creating the barrier events failed");
            break;
        }
        grp->ev_bar.push_back(ev);
    }
    if (rc) {
        const std::string keep = g_last_error;
        if (grp->members.empty()) delete grp;
        else p3d_destroy(grp);  // destroys the members made so far and the handle
        g_last_error = keep;
        return rc;
    }
    grp->world = n_dev;
    if (const char *env = std::getenv("P3D_MULTI_CELLS_MIN")) grp->multi_cells_min = (size_t)std::strtoull(env, nullptr, 10);
    *out = grp;
    return P3D_OK;
}

}  // extern "C"

// Cross-device barrier on the members' streams: every stream continues only after every other stream reached this
// point.  set = 0 / 1 selects the event set (after the force pass / after the fused integrate).
static int multi_barrier(p3d_engine *grp, int set) {
    const int G = (int)grp->members.size();
    for (int g = 0; g < G; ++g) {
        CU(cudaSetDevice(grp->members[g]->device));
        CU(cudaEventRecord(grp->ev_bar[set * G + g], grp->members[g]->stream));
    }
    for (int g = 0; g < G; ++g) {
        CU(cudaSetDevice(grp->members[g]->device));
        for (int h = 0; h < G; ++h)
            if (h != g) CU(cudaStreamWaitEvent(grp->members[g]->stream, grp->ev_bar[set * G + h], 0));
    }
    return P3D_OK;
}

static int multi_upload(p3d_engine *grp, const p3d_particle *in, size_t n, uint32_t id_count) {
    const int G = (int)grp->members.size();
    int rc;
    grp->n = 0;
    grp->n_slots = 0;
    // Only the all-pairs kernel has enough work per step to be worth several devices at ordinary sizes: a cell-list
    // step at N = 1M is 0.15 ms on one device, less than the host needs to issue the launches and cross-device
    // barriers of a sharded one (and every device would sort all N cells redundantly).  Below multi_cells_min
    // particles (8M; P3D_MULTI_CELLS_MIN) such an upload therefore lives on member 0 alone, which then steps it like
    // a one-device engine (CUDA graphs, cell-ordered slots); from there on the work is sharded like the pair kernel's.
    p3d_engine *m0 = grp->members[0];
    for (int g = 0; g < G; ++g) {  // no step of the previous state may still be running
        CU(cudaSetDevice(grp->members[g]->device));
        CU(cudaStreamSynchronize(grp->members[g]->stream));
    }
    grp->solo = resolve_force_kernel_for(m0, n) != P3D_FORCE_PAIR && n < grp->multi_cells_min;
    m0->world = grp->solo ? 1 : G;
    if (grp->solo) {
        CU(cudaSetDevice(m0->device));
        if ((rc = stage_input(m0, in, 0, n, n))) return rc;
        if ((rc = upload_commit(m0, n, id_count, false))) return rc;
        grp->n = n;
        grp->n_slots = m0->n_slots;
        grp->T = id_count;
        return P3D_OK;
    }
    const size_t per = staged_part(n, G);
    // every member copies ITS part of the caller's array over its own PCIe link ...
    for (int g = 0; g < G; ++g) {
        p3d_engine *m = grp->members[g];
        CU(cudaSetDevice(m->device));
        CU(cudaStreamSynchronize(m->stream));  // no step of the previous state may still be reading the buffers
        const size_t c0 = std::min(n, per * g), c1 = std::min(n, c0 + per);
        if ((rc = stage_input(m, in ? in + c0 : nullptr, c0, c1, n))) return rc;
    }
    // ... and pushes it into every other member's staging array over NVLink (all-gather by peer copies)
    for (int g = 0; g < G; ++g) {
        p3d_engine *m = grp->members[g];
        CU(cudaSetDevice(m->device));
        const size_t c0 = std::min(n, per * g), c1 = std::min(n, c0 + per);
        for (int h = 0; h < G && c1 > c0; ++h) {
            if (h == g) continue;
            p3d_engine *o = grp->members[h];
            CU(cudaMemcpyPeerAsync(o->aos.p + c0 * 7, o->device, m->aos.p + c0 * 7, m->device,
                                   (c1 - c0) * sizeof(p3d_particle), m->stream));
        }
    }
    if ((rc = multi_barrier(grp, 0))) return rc;
    // every member builds the same layout from the same array; the count kernels of all members run concurrently
    const bool pair = resolve_force_kernel_for(grp->members[0], n) == P3D_FORCE_PAIR;
    if (pair)
        for (int g = 0; g < G; ++g) {
            CU(cudaSetDevice(grp->members[g]->device));
            if ((rc = layout_count_async(grp->members[g], n, id_count))) return rc;
        }
    for (int g = 0; g < G; ++g) {
        CU(cudaSetDevice(grp->members[g]->device));
        if ((rc = upload_commit(grp->members[g], n, id_count, pair))) return rc;
    }
    // peer tables: plain pointers (one process, peer access enabled); refreshed because an upload may reallocate
    for (int g = 0; g < G; ++g) {
        p3d_engine *m = grp->members[g];
        for (int h = 0; h < G; ++h) {
            p3d_engine *o = grp->members[h];
            m->peer_ptr[h][0] = o->frc.p;
            m->peer_ptr[h][1] = o->pos[0].p;
            m->peer_ptr[h][2] = o->pos[1].p;
            m->peer_ptr[h][3] = o->vel.p;
        }
        m->peers_ready = true;
    }
    grp->n = n;
    grp->n_slots = grp->members[0]->n_slots;
    grp->T = id_count;
    return P3D_OK;
}

static int multi_step(p3d_engine *grp, const p3d_params *prm, float ts, int n_steps) {
    const int G = (int)grp->members.size();
    DevParams P;
    int rc;
    if ((rc = canonicalise(prm, P))) return rc;
    if (grp->n_slots == 0) return fail(P3D_ERR_INVALID, "no particles uploaded");
    if (grp->solo) return p3d_step(grp->members[0], prm, ts, n_steps);
    if (prm->id_count != grp->T)
        return fail(P3D_ERR_INVALID, "id_count %u differs from the uploaded layout (%u): upload again", prm->id_count, grp->T);
    for (int g = 0; g < G; ++g) {
        CU(cudaSetDevice(grp->members[g]->device));
        if ((rc = upload_matrix(grp->members[g], prm))) return rc;
    }
    for (int s = 0; s < n_steps; ++s) {
        // force pass: every device evaluates its share (block rows / cell-sorted range) -> PARTIAL forces for all slots
        for (int g = 0; g < G; ++g) {
            p3d_engine *m = grp->members[g];
            CU(cudaSetDevice(m->device));
            // the out-of-box flag written by a member's integrate covers only its own slots
            if ((rc = check_box_now(m, P))) return rc;
            if ((rc = launch_force(m, P))) return rc;
        }
        if ((rc = multi_barrier(grp, 0))) return rc;  // every device's partial forces are complete
        // reduce-scatter(forces) + integrate + all-gather(positions, velocities) in one kernel per device
        for (int g = 0; g < G; ++g) {
            CU(cudaSetDevice(grp->members[g]->device));
            if ((rc = launch_integrate_fused(grp->members[g], P, ts))) return rc;
        }
        if ((rc = multi_barrier(grp, 1))) return rc;  // every device's peer stores have landed
        for (int g = 0; g < G; ++g) {
            grp->members[g]->cur ^= 1;
            grp->members[g]->parity ^= 1;
        }
    }
    return P3D_OK;
}

static int multi_download(p3d_engine *grp, p3d_particle *out, size_t n) {
    const int G = (int)grp->members.size();
    if (n != grp->n) return fail(P3D_ERR_INVALID, "n=%zu but %zu particles are resident", n, grp->n);
    if (n && !out) return fail(P3D_ERR_INVALID, "out is null");
    int rc;
    if (grp->solo) return download_range(grp->members[0], out, 0, n, true);
    const size_t per = staged_part(n, G);
    // every device holds the whole state; each serves its part of the caller's array over its own PCIe link
    for (int g = 0; g < G; ++g) {
        const size_t c0 = std::min(n, per * g), c1 = std::min(n, c0 + per);
        if ((rc = download_range(grp->members[g], out + c0, c0, c1, false))) return rc;
    }
    for (int g = 0; g < G; ++g) {
        CU(cudaSetDevice(grp->members[g]->device));
        CU(cudaStreamSynchronize(grp->members[g]->stream));
    }
    return P3D_OK;
}

static int multi_download_forces(p3d_engine *grp, float *out_xyz, size_t n) {
    const int G = (int)grp->members.size();
    if (n != grp->n) return fail(P3D_ERR_INVALID, "n=%zu but %zu particles are resident", n, grp->n);
    if (!n) return P3D_OK;
    if (!out_xyz) return fail(P3D_ERR_INVALID, "out is null");
    int rc;
    if (grp->solo) return p3d_download_forces(grp->members[0], out_xyz, n);
    if ((rc = p3d_sync(grp))) return rc;
    p3d_engine *e = grp->members[0];
    CU(cudaSetDevice(e->device));
    if ((rc = e->fout.ensure(n * 3))) return rc;
    PeerForces pf;
    for (int g = 0; g < 8; ++g) pf.frc[g] = g < G ? grp->members[g]->frc.p : nullptr;
    k_unpack_forces_sum<<<(unsigned)((n + 255) / 256), 256, 0, e->stream>>>(pf, G, slot_map(e),
                                                                            e->fout.p, (int)n);
    e->counters[0]++;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(out_xyz, e->fout.p, n * 3 * sizeof(float), cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    return P3D_OK;
}
