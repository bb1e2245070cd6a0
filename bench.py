#!/usr/bin/env python
"""bench.py — the hot path of particle_3d (Particles::update, src/lib.rs:130-272) on B200.

  python bench.py --gpus N --steps K --warmup W            (N>1: launched by torch.distributed.run)
  python bench.py --impl reference --gpus N --steps K --warmup W

Workload (BASELINE.json config 4, the one its metric is quoted on): N = 1,048,576 particles,
uniform cloud, W = 101.6 (density 1), default scene constants (src/bin/main.rs:133-148),
ts = 1/60, seed 42.  A "step" is one update(): all-pairs force pass + fused integration.
`value` = pair interactions per second (N^2 ordered pairs per step, the 20-flop convention's unit)
with the state resident in HBM; `e2e` = the same through p3d_update with host buffers.
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "3d-particle-simulation-_b200"))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

N_DEFAULT = 1_048_576
W_DEFAULT = 101.6
TS = float(np.float32(1.0 / 60.0))  # src/bin/main.rs:164,194
SEED = 42
FLOP_PER_INTERACTION = 20  # north_star's convention
METRIC = "pair_interactions_per_s"
UNIT = "interactions/s"

# stdout carries exactly ONE JSON line: everything else a library prints there (NCCL's version banner,
# for one) is sent to stderr by pointing fd 1 at fd 2 for the duration of the run.
_JSON_FD = None


def _capture_stdout():
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)


def emit(line: dict):
    sys.stdout.flush()
    data = (json.dumps(line) + "\n").encode()
    os.write(_JSON_FD if _JSON_FD is not None else 1, data)


# ------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock / throttle reasons DURING the timed region (B200_PROFILING.md recipe): an NVML polling
    thread (5 ms period, so that even a 50 ms multi-GPU step is sampled); nvidia-smi -lms as fallback."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        if vis and all(v.strip().isdigit() for v in vis.split(",")):
            gpu_index = int(vis.split(",")[gpu_index])
        self.gpu = gpu_index
        self.p = None
        self.thread = None
        self.rows = []  # (sm_mhz, power_w, reasons bitmask)
        self.summary = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}

    def _poll(self, nv, h):
        while not self._stop.is_set():
            try:
                self.rows.append((float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)),
                                  nv.nvmlDeviceGetPowerUsage(h) / 1000.0,
                                  int(nv.nvmlDeviceGetCurrentClocksEventReasons(h))))
            except Exception:
                pass
            self._stop.wait(0.005)

    def __enter__(self):
        try:
            import pynvml as nv

            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.gpu)
            self._nv, self._h = nv, h
            self._max = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            self._stop = threading.Event()
            self.thread = threading.Thread(target=self._poll, args=(nv, h), daemon=True)
            self.thread.start()
            return self
        except Exception:
            self.thread = None
        try:
            self.p = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.p = None
        return self

    def __exit__(self, *a):
        if self.thread:
            self._stop.set()
            self.thread.join(timeout=2)
            nv = self._nv
            if self.rows:
                clk = sorted(r[0] for r in self.rows)
                bits = 0
                for r in self.rows:
                    bits |= r[2]
                names = (("hw_slowdown", nv.nvmlClocksEventReasonHwSlowdown), ("hw_thermal_slowdown", nv.nvmlClocksEventReasonHwThermalSlowdown),
                         ("sw_thermal_slowdown", nv.nvmlClocksEventReasonSwThermalSlowdown), ("sw_power_cap", nv.nvmlClocksEventReasonSwPowerCap))
                self.summary = {"sm_mhz": clk[len(clk) // 2], "sm_min_mhz": clk[0], "sm_max_mhz": self._max,
                                "reasons": sorted(n for n, b in names if bits & b), "samples": len(self.rows),
                                "power_w_max": max(r[1] for r in self.rows), "source": "NVML, 5 ms period"}
            return
        if not self.p:
            return
        self.p.terminate()
        try:
            out = self.p.communicate(timeout=5)[0]
        except Exception:
            return
        rows = [r.split(", ") for r in out.strip().splitlines() if r.count(",") >= 7]
        if not rows:
            return
        clk = sorted(float(r[1]) for r in rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[k] for r in rows for k in range(4) if r[4 + k].strip().lower().startswith("active")})
        self.summary = {"sm_mhz": clk[len(clk) // 2], "sm_max_mhz": float(rows[0][2]), "reasons": reasons,
                        "samples": len(rows), "power_w_max": max(float(r[3]) for r in rows), "source": "nvidia-smi -lms 100"}


def weak_particles(gpus: int) -> int:
    """BASELINE.json config 5, weak scaling: the all-pairs work per GPU, N^2 / G, is held at the 1-GPU value, so
    N_G = 1,048,576 * sqrt(G), rounded to a whole number of 256-particle blocks per GPU."""
    g = max(1, int(gpus))
    unit = 256 * g
    return int(round(N_DEFAULT * (g ** 0.5) / unit)) * unit


def multi_gpu_roofline(n: int, world: int, ms_per_step: float, sm_count: int, sm_max_mhz: float, phases=None) -> dict:
    """FP32 roofline of a sharded run (BASELINE.json configs[3] and [4] ask for the fraction at every GPU count).
    The job evaluates N^2 ordered pairs per step on `world` GPUs, so the denominator is world x the per-GPU FP32 FMA
    peak.  `achieved` divides by the WHOLE timed step (slowest rank; force pass + exchange + integrate), which is
    what the job delivers; the force pass alone, from the per-rank CUDA-event diagnostic, is reported beside it."""
    peak_gpu = sm_count * 128 * 2 * float(sm_max_mhz) * 1e6 / 1e12
    flops = float(n) * n * FLOP_PER_INTERACTION
    achieved = flops / (ms_per_step * 1e-3) / 1e12
    out = {
        "kernel": "whole sharded step on the slowest rank: force pass (k_force_pair + k_force_bxb + partition) + "
                  "exchange/integrate",
        "bound": "fp32_fma", "achieved": achieved, "peak": peak_gpu * world, "unit": "TFLOP/s",
        "frac": achieved / (peak_gpu * world),
        "peak_source": f"{world} GPUs x {sm_count} SMs x 128 lanes x 2 flop x {float(sm_max_mhz):.0f} MHz",
        "algorithmic_flops_per_step": flops, "flop_per_interaction": FLOP_PER_INTERACTION,
        "per_gpu": {"achieved": achieved / world, "peak": peak_gpu},
        "traffic": None,
    }
    force = [float(x) for x in (phases or {}).get("force", []) if x and x > 0]
    if force:
        slowest = max(force)
        out["force_pass"] = {"ms_slowest_rank": slowest, "achieved": flops / (slowest * 1e-3) / 1e12,
                             "frac": flops / (slowest * 1e-3) / 1e12 / (peak_gpu * world),
                             "what": "each rank's share of the block rows, CUDA events on its own stream (diagnostic steps)"}
    return out


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return None


def workload(n: int, W: float, cloud: str = "uniform", impl: str = "engine"):
    """Seeded inputs.  The engine arm uses the product's generator (p3d_scene_*), the reference arm the oracle's own
    restatement of it (byte-identical: tests/test_oracle_scene.py), so that `--impl reference` never loads libp3d.so."""
    if impl == "reference":
        from oracle import oracle as O

        prm = O.default_params_dict()
        prm["world_size"] = W
        if cloud == "plummer":
            return prm, O.scene_plummer(W, n, W / 6, seed=SEED)
        return prm, O.scene_uniform(W, n, seed=SEED)
    import particle_3d as p3

    prm = p3.default_params_dict()
    prm["world_size"] = W
    if cloud == "plummer":  # BASELINE.json's clustered cloud: Plummer radius CDF, scale a = W/6, truncated to the box
        return prm, p3.generate_plummer(W, n, W / 6, seed=SEED)
    return prm, p3.generate_particles(W, n, seed=SEED)


def config_dict(n, W, cloud="uniform", weak=False):
    """Identical in both arms (the driver compares the two dicts): the workload only, nothing about how it is run."""
    what = "uniform cloud" if cloud == "uniform" else "clustered (Plummer a=W/6) cloud"
    c = {"workload": f"N={n} {what}, W={W} (density 1), default scene constants (main.rs:133-148), ts=1/60, seed {SEED}; "
                     "BASELINE.json configs[3]",
         "n_particles": n, "world_size": W, "cloud": cloud, "algorithm": "all-pairs (N^2 ordered pairs per step)",
         "cache": "state is ~100 MB (< L2), so a 512 MiB buffer is overwritten between timed steps to flush L2"}
    if weak:
        c["weak_scaling"] = ("N = 1,048,576 * sqrt(GPUs) (BASELINE.json configs[4]): the pair count per GPU is that of "
                             "the 1-GPU run")
    return c


PARITY_TOL = 1e-5      # north_star: relative tolerance after one step (floors as in tests/helpers.py)
PARITY_SAMPLE = 16384  # particles compared with the oracle, drawn from every rank's slot range


def parity_sample_errors(out_sample, ref_sample, world_size):
    """tests/helpers.py parity metric on a sample: |dv| / max(|v|, v_rms), |dp| / max(|p|, W/2), both over PARITY_TOL."""
    def v3(a, k):
        return np.stack([a[k + "x"], a[k + "y"], a[k + "z"]], 1).astype(np.float64)

    v, vr, p, pr = v3(out_sample, "v"), v3(ref_sample, "v"), v3(out_sample, "p"), v3(ref_sample, "p")
    vrms = float(np.sqrt((vr ** 2).sum(1).mean())) if len(vr) else 0.0
    dv = np.linalg.norm(v - vr, axis=1) / np.maximum(np.linalg.norm(vr, axis=1), max(vrms, 1e-30))
    dp = np.linalg.norm(p - pr, axis=1) / np.maximum(np.linalg.norm(pr, axis=1), world_size / 2)
    return float(dv.max() / PARITY_TOL), float(dp.max() / PARITY_TOL)


# ------------------------------------------------------------------------------------------
def host_threads() -> int:
    """All host cores this process may use.  torch.distributed.run exports OMP_NUM_THREADS=1 for N>1,
    which must not throttle the CPU arm: the thread count is passed to the oracle explicitly."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def cpu_reference_sample(prm, parts, target_s: float, nthreads: int = 0):
    """Times the CPU restatement of src/lib.rs (oracle, 'port') on a bounded sample of one step:
    the counting sort over all N plus the per-particle pass for the first S particles."""
    from oracle import oracle as O

    if nthreads <= 0:
        nthreads = host_threads()

    n = len(parts)
    probe = min(n, 4096)
    _, st = O.update_sample(prm, TS, parts, 0, probe, mode=O.FAITHFUL, nthreads=nthreads)
    per_particle = max(st["t_force_s"] / probe, 1e-9)
    sample = int(min(n, max(probe, target_s / per_particle)))
    _, st = O.update_sample(prm, TS, parts, 0, sample, mode=O.FAITHFUL, nthreads=nthreads)
    step_s = st["t_build_s"] + st["t_force_s"] * (n / sample)
    return {"step_s": step_s, "sample": sample, "build_s": st["t_build_s"], "force_sample_s": st["t_force_s"],
            "cores": nthreads,
            "candidates_per_particle": st["candidates"] / sample}


def run_reference(args):
    """--impl reference: the reference's own CPU algorithm (spatial hash, 27x27 cell walk) on all
    host threads.  The Rust crate cannot be built in this image (no cargo/rustc), so this is the C
    restatement under oracle/ ('port').  Rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n, W = args.n, args.world_size
    prm, parts = workload(n, W, args.cloud, impl="reference")
    per_step_budget = max(1.0, min(20.0, 120.0 / max(1, args.steps + args.warmup)))
    times, last = [], None
    for s in range(args.warmup + args.steps):
        last = cpu_reference_sample(prm, parts, per_step_budget)
        if s >= args.warmup:
            times.append(last["step_s"])
    step_s = float(np.mean(times))
    value = float(n) * n / step_s
    sample_txt = (f"per step: counting sort over all {n} particles + per-particle pass for the first {last['sample']} "
                  f"particles, scaled by N/sample (faithful mode, {last['cores']} OpenMP threads)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": step_s * 1e3, "higher_is_better": True,
        "scaling": "weak" if args.weak else "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config_dict(n, W, args.cloud, args.weak),
        "steps_per_s": 1.0 / step_s, "cpu_cores": last["cores"],
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": last["cores"], "kind": "port", "sample": sample_txt,
                         "ms_per_step": step_s * 1e3,
                         "candidate_pairs_per_s": last["candidates_per_particle"] * n / step_s},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "all-pairs-equivalent N^2/t of the reference's cell-list algorithm; the Rust crate itself cannot be "
                "built here (no cargo/rustc), this is the C restatement oracle/p3d_oracle.c",
    }
    emit(line)


# ------------------------------------------------------------------------------------------
def run_engine(args):
    import torch

    import particle_3d as p3
    from particle_3d import _abi
    from particle_3d.sharded import ShardedStepper, engine_tensors, exchange_peer_handles, part_range, sharded_upload

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist  # noqa: F811

        dist.init_process_group(backend="nccl", device_id=torch.device(f"cuda:{local}"))

    n, W = args.n, args.world_size
    prm, parts = workload(n, W, args.cloud)
    P = p3.Engine.make_params(**prm)
    eng = p3.Engine(local)
    eng.set_option(_abi.OPT_FORCE_KERNEL, _abi.FORCE_PAIR if n >= 4096 else _abi.FORCE_REFERENCE_ORDER)
    eng.set_option(_abi.OPT_BLOCK_SIZE, args.block)
    eng.set_option(_abi.OPT_TIMING, 1)
    stream = torch.cuda.Stream(device=local)  # every launch, copy and collective of the bench runs on this stream
    torch.cuda.set_stream(stream)
    eng.set_stream(stream.cuda_stream)
    eng.set_shard(rank, world)
    eng.upload(parts, prm["id_count"])
    flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=f"cuda:{local}")

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    stepper = None
    fused = world > 1 and not args.no_fused

    def make_stepper():
        if fused:
            exchange_peer_handles(eng, dist, world)
            bar = torch.zeros(1, device=f"cuda:{local}")
            return ShardedStepper(eng, dist, rank, world, lambda: engine_tensors(eng, local), fused=True, barrier_tensor=bar)
        return ShardedStepper(eng, dist, rank, world, lambda: engine_tensors(eng, local))

    if world > 1:
        stepper = make_stepper()

    def one_step():
        if stepper:
            stepper.step(P, TS, 1)
        else:
            eng.step(P, TS, 1)

    # ---------------- device-resident leg: `value` ----------------
    for _ in range(args.warmup):
        one_step()
        flush.zero_()
    c0 = eng.counters()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kern = {"force": 0.0, "pair": 0.0, "bxb": 0.0, "integrate": 0.0, "partition": 0.0, "steps": 0}
    with ClockSampler(local) as clocks:
        # the sampler starts (nvmlInit: milliseconds, different on every rank) BEFORE the barrier, so that all ranks
        # enter the timed region together; its samples from before the barrier are dropped
        barrier()
        clocks.rows.clear()
        ev0.record(stream)
        for _ in range(args.steps):
            one_step()
            if world == 1:
                t = eng.timing()  # CUDA events recorded by the engine on this stream (syncs the stream)
                for k in ("force", "pair", "bxb", "integrate", "partition"):
                    kern[k] += t[k]
                kern["steps"] += 1
            flush.zero_()
        ev1.record(stream)
        barrier()
    ms = ev0.elapsed_time(ev1)
    if world > 1:
        tt = torch.tensor([ms], device=f"cuda:{local}")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt.item())
    c1 = eng.counters()
    clock_summary = clocks.summary
    if world > 1:  # every rank sampled its own GPU: report the slowest median and the union of the reasons
        allc = [None] * world
        dist.all_gather_object(allc, clocks.summary)
        meds = [c.get("sm_mhz") for c in allc]
        clock_summary = dict(allc[0], per_rank_sm_mhz=meds,
                             sm_mhz=min((m for m in meds if m is not None), default=None),
                             reasons=sorted({r for c in allc for r in c.get("reasons", [])}),
                             samples=min(c.get("samples", 0) for c in allc),
                             power_w_max=max((c.get("power_w_max") or 0.0) for c in allc))
    ms_per_step = ms / args.steps
    value = float(n) * n / (ms_per_step * 1e-3)
    launches = (c1["kernels"] - c0["kernels"]) + args.steps  # + the L2-flush fill kernel per step

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak" if args.weak else "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "force_kernel": "k_force_pair: all N^2 pairs (north_star); the cell list is reported separately",
        "config": config_dict(n, W, args.cloud, args.weak),
        "engine": {"parallelism": (f"block rows sharded over {world} GPUs; " + ("per step ONE fused kernel does reduce-scatter(forces) + integrate + all-gather(positions, velocities) over NVLink peer memory, NCCL only as two 4-byte barrier all-reduces" if fused else "per step all-reduce(forces) + all-gather(positions, velocities) over NCCL")) if world > 1 else "1 GPU",
                   "block": args.block},
        "steps_per_s": 1e3 / ms_per_step,
        "gpu_launches": int(launches),
        "clocks": clock_summary,
    }

    # ---------------- parity of THIS run's path against the CPU oracle (outside the timed region) ----------------
    # Re-upload the seed-42 state, take ONE step through the same (sharded) path the timed region used, download,
    # and compare a sample drawn from every rank's slot range with the oracle (ideal mode; src/lib.rs:167-171,
    # 268-271: same index order, same state).  A broken peer pointer or a stale barrier would leave every timing
    # unchanged; this is what catches it.  The run FAILS (exit code 1) when the sample is out of tolerance.
    def reupload(src):
        if world > 1:
            sharded_upload(eng, dist, rank, world, local, src, prm["id_count"])
            stepper.reset()
        else:
            eng.upload(src, prm["id_count"])

    eng.set_option(_abi.OPT_TIMING, 0)
    reupload(parts)
    one_step()
    torch.cuda.synchronize()
    state = eng.download()  # every rank holds the whole state after a step
    parity = None
    digest = __import__("hashlib").sha256(state.tobytes()).hexdigest()
    digests = [digest]
    if world > 1:
        digests = [None] * world
        dist.all_gather_object(digests, digest)
    if rank == 0 and not args.no_parity:
        from oracle import oracle as O

        slot = eng.slot_of().astype(np.int64)
        s0, s1 = eng.shard_range()
        owner = slot // max(1, s1 - s0)
        rng = np.random.default_rng(20261018)
        per_rank = max(1, PARITY_SAMPLE // world)
        idx = np.concatenate([rng.choice(np.flatnonzero(owner == g), size=min(per_rank, int((owner == g).sum())), replace=False)
                              for g in range(world) if (owner == g).any()])
        t0 = time.perf_counter()
        ref, _ = O.update_indices(prm, TS, parts, idx, mode=O.IDEAL, nthreads=host_threads())
        dv, dp = parity_sample_errors(state[idx], ref, W)
        parity = {"max_dv_over_tol": dv, "max_dp_over_tol": dp, "n_checked": int(idx.size), "mode": "ideal", "tol": PARITY_TOL,
                  "ranks_covered": sorted(int(g) for g in np.unique(owner[idx])),
                  "ids_and_order_exact": bool(np.array_equal(state["id"], parts["id"])),
                  "replicas_identical": len(set(digests)) == 1,
                  "what": f"one step from the seed-{SEED} state through the timed path; {idx.size} particles drawn from every rank's "
                          "slot range vs oracle.update_indices (CPU restatement of src/lib.rs, ideal mode); every rank's "
                          "downloaded copy of the whole state must be byte-identical",
                  "oracle_s": time.perf_counter() - t0}
        parity["ok"] = bool(dv <= 1.0 and dp <= 1.0 and parity["ids_and_order_exact"] and parity["replicas_identical"]
                            and len(parity["ranks_covered"]) == world)
        line["parity"] = parity

    if world == 1 and rank == 0:
        # ---------------- roofline of the dominant kernel (k_force_pair) ----------------
        prop = torch.cuda.get_device_properties(local)
        peaks = measured_peaks()
        sm_max_mhz = (peaks or {}).get("sm_max_mhz") or clock_summary.get("sm_max_mhz") or 1965.0
        fp32_peak_tf = prop.multi_processor_count * 128 * 2 * sm_max_mhz * 1e6 / 1e12
        from tools import microbench
        _, mb = microbench.run(local, 1, 2000)  # packed FFMA2 microbenchmark (libp3d_microbench.so)
        force_ms = kern["force"] / max(1, kern["steps"])
        pair_ms = kern["pair"] / max(1, kern["steps"])
        integ_ms = kern["integrate"] / max(1, kern["steps"])
        flops = float(n) * n * FLOP_PER_INTERACTION
        achieved_tf = flops / (force_ms * 1e-3) / 1e12
        prof = {}
        try:
            prof = json.load(open(os.path.join(ROOT, "profiles", "force_pair_ncu_summary.json")))
        except Exception:
            pass
        line["roofline"] = {
            "kernel": "force pass = k_force_pair (+ k_force_bxb for boundary x boundary blocks, + partition kernels)",
            "bound": "fp32_fma", "achieved": achieved_tf, "peak": fp32_peak_tf, "unit": "TFLOP/s",
            "frac": achieved_tf / fp32_peak_tf,
            "peak_source": f"{prop.multi_processor_count} SMs x 128 lanes x 2 flop x {sm_max_mhz:.0f} MHz (MEASURED_PEAKS.json sm_max_mhz); "
                           "tensor/HBM peaks do not bound this kernel (not a dense contraction)",
            "peak_ffma2_microbench_tflops": mb[0] * 2 / 1e12,
            "algorithmic_flops_per_launch": flops, "flop_per_interaction": FLOP_PER_INTERACTION,
            "frac_note": "`frac` follows north_star's 20-flop-per-ordered-interaction convention.  The kernel evaluates each "
                         "unordered pair once and EXECUTES 16 FP32 flop (8 lane-FMAs) per ordered interaction, so the executed "
                         "fraction of the FMA peak is frac * 16/20; ncu's measured FMA-pipe utilisation is under `ncu`",
            "frac_executed_flops": achieved_tf / fp32_peak_tf * 16.0 / 20.0,
            "kernel_ms": force_ms, "pair_kernel_ms": pair_ms, "bxb_tail_ms": kern["bxb"] / max(1, kern["steps"]),
            "bxb_note": "k_force_bxb runs BESIDE k_force_pair on an auxiliary stream (alone, under ncu: 2.2 ms = 0.7 % of the step, "
                        "profiles/r01_launches_n1048576.csv); pair_kernel_ms therefore contains it and bxb_tail_ms is only what "
                        "remains after the pair kernel ended",
            "partition_ms": kern["partition"] / max(1, kern["steps"]),
            "interactions_per_s_force_pass": float(n) * n / (force_ms * 1e-3),
            "traffic": prof.get("dram_bytes_per_launch"),
            "ncu": prof or None,
        }
        hbm = (peaks or {}).get("hbm_gbs") or 6650.0
        ach = 80.0 * n / (integ_ms * 1e-3) / 1e9
        line["roofline_integrate"] = {
            "kernel": "k_integrate", "bound": "hbm", "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm,
            "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6.65 TB/s",
            "algorithmic_bytes_per_particle": 80, "kernel_ms": integ_ms, "traffic": None,
            "note": f"{80 * n / 1e6:.0f} MB working set; L2 is flushed between steps, but the force pass re-reads positions "
                    "before integrate runs, so part of it is L2-resident",
        }

    # ---------------- k_integrate on a working set larger than L2 (N = 4M: 335 MB) ----------------
    if world == 1 and rank == 0 and not args.no_cells:
        n4, W4 = 4 * 1048576, 161.3
        big = p3.Engine(local)
        big.set_stream(stream.cuda_stream)
        big.set_option(_abi.OPT_FORCE_KERNEL, _abi.FORCE_CELLS)  # identity layout, no host pass
        prm4 = dict(prm, world_size=W4)
        P4 = p3.Engine.make_params(**prm4)
        big.upload(p3.generate_particles(W4, n4, seed=SEED), prm["id_count"])
        big.step(P4, TS, 2)  # non-trivial forces and velocities
        for _ in range(3):
            big.shard_integrate(P4, TS)
        i0, i1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 20
        i0.record(stream)
        for _ in range(reps):
            big.shard_integrate(P4, TS)
        i1.record(stream)
        torch.cuda.synchronize()
        ims = i0.elapsed_time(i1) / reps
        hbm = (measured_peaks() or {}).get("hbm_gbs") or 6650.0
        line["roofline_integrate"]["at_4m_particles"] = {
            "kernel_ms": ims, "achieved": 80.0 * n4 / (ims * 1e-3) / 1e9, "unit": "GB/s",
            "frac": 80.0 * n4 / (ims * 1e-3) / 1e9 / hbm,
            "note": "335 MB working set (> 126 MB L2), 20 back-to-back launches of k_integrate"}
        big.close()

    # ---------------- end-to-end leg: p3d_update with pinned HOST buffers ----------------
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    if world == 1:
        hin = torch.empty(n * 28, dtype=torch.uint8).pin_memory()
        hout = torch.empty(n * 28, dtype=torch.uint8).pin_memory()
        a_in = hin.numpy().view(_abi.PARTICLE)
        a_out = hout.numpy().view(_abi.PARTICLE)
        a_in[:] = parts
        eng.set_option(_abi.OPT_TIMING, 0)
        eng.update_into(P, TS, a_in, a_out)  # warm-up (allocations, first-touch)
        a_in[:] = a_out
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            eng.update_into(P, TS, a_in, a_out)  # synchronous: H2D + step + D2H
            a_in, a_out = a_out, a_in            # src/lib.rs:167 swap
        e2e_s = (time.perf_counter() - t0) / e2e_steps
        line["e2e"] = {"value": float(n) * n / e2e_s, "unit": UNIT, "h2d_bytes_per_step": n * 28,
                       "d2h_bytes_per_step": n * 28, "ms_per_step": e2e_s * 1e3, "steps": e2e_steps,
                       "api": "p3d_update(engine, params, ts, in, out, n) on pinned host arrays; wall clock around the synchronous call"}
    else:
        # multi-GPU e2e: the caller's array is split over the ranks.  Every step each rank uploads ITS 1/G of the
        # state from pinned host memory over its own PCIe link, the staging arrays are all-gathered over NVLink, the
        # type-grouped layout is rebuilt, one sharded step runs, and each rank reads ITS 1/G of the result back.
        c0, c1 = part_range(n, rank, world)
        hin = torch.empty(max(1, c1 - c0) * 28, dtype=torch.uint8).pin_memory()
        hout = torch.empty(max(1, c1 - c0) * 28, dtype=torch.uint8).pin_memory()
        a_in = hin.numpy().view(_abi.PARTICLE)[:c1 - c0]
        a_out = hout.numpy().view(_abi.PARTICLE)[:c1 - c0]
        a_in[:] = parts[c0:c1]

        def e2e_step():
            eng.upload_part(a_in, c0, n, prm["id_count"])
            ptr, cap = eng.device_buffer(_abi.BUF_AOS)
            t = aos_view(ptr, cap)
            per = cap // world
            dist.all_gather_into_tensor(t, t[rank * per * 7:(rank + 1) * per * 7].clone())
            eng.upload_commit(n)
            stepper.reset()
            stepper.step(P, TS, 1)
            eng.download_part_into(a_out, c0)

        aos_cache = {}

        def aos_view(ptr, cap):
            if (ptr, cap) not in aos_cache:
                class _V:
                    __cuda_array_interface__ = {"shape": (cap * 7,), "typestr": "<f4", "data": (ptr, False), "version": 2,
                                                "strides": None}
                aos_cache[(ptr, cap)] = torch.as_tensor(_V(), device=f"cuda:{local}")
            return aos_cache[(ptr, cap)]

        e2e_step()  # warm-up
        a_in[:] = parts[c0:c1]
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
            a_in, a_out = a_out, a_in  # src/lib.rs:167 swap
        barrier()
        e2e_s = (time.perf_counter() - t0) / e2e_steps
        tt = torch.tensor([e2e_s], device=f"cuda:{local}")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_s = float(tt.item())
        line["e2e"] = {"value": float(n) * n / e2e_s, "unit": UNIT, "h2d_bytes_per_step": n * 28,
                       "d2h_bytes_per_step": n * 28, "ms_per_step": e2e_s * 1e3, "steps": e2e_steps,
                       "api": f"per rank: p3d_upload_part (1/{world} of the array, pinned) + NCCL all-gather of the staging array + "
                              "p3d_upload_commit + sharded step + p3d_download_part (1/G of the array); bytes are the job's "
                              "totals over all ranks"}
        reupload(parts)

    # ---------------- multi-GPU: where a step's time goes on every rank (diagnostic, outside the timed region) ----------------
    if world > 1 and fused:
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        acc = [0.0, 0.0, 0.0]
        d_steps = 3
        for _ in range(d_steps):
            barrier()
            evs[0].record(stream)
            eng.shard_force(P)                      # partition + this rank's block rows (pair + boundary kernels)
            evs[1].record(stream)
            dist.all_reduce(stepper._bar)           # waits for the slowest rank's force pass
            evs[2].record(stream)
            eng.shard_integrate_fused(P, TS)        # P2P reduce-scatter + integrate + all-gather
            dist.all_reduce(stepper._bar)
            evs[3].record(stream)
            eng.shard_commit()
            torch.cuda.synchronize()
            for k in range(3):
                acc[k] += evs[k].elapsed_time(evs[k + 1]) / d_steps
        allp = [None] * world
        dist.all_gather_object(allp, acc)
        line["phases_ms_per_rank"] = {
            "what": "3 untimed diagnostic steps, CUDA events on each rank: own force pass | waiting for the slowest rank "
                    "(first barrier) | fused P2P integrate + second barrier",
            "force": [round(a[0], 3) for a in allp], "wait_for_slowest": [round(a[1], 3) for a in allp],
            "exchange_integrate": [round(a[2], 3) for a in allp]}

    if world > 1 and rank == 0:
        try:  # (no collective in here: only rank 0 runs it)
            prop = torch.cuda.get_device_properties(local)
            sm_max_mhz = (measured_peaks() or {}).get("sm_max_mhz") or clock_summary.get("sm_max_mhz") or 1965.0
            line["roofline"] = multi_gpu_roofline(n, world, ms_per_step, prop.multi_processor_count, sm_max_mhz,
                                                  line.get("phases_ms_per_rank"))
        except Exception as ex:  # a reporting extra must never take the bench line down
            line["roofline"] = {"error": repr(ex)}

    # ---------------- the other cloud (north_star: uniform AND clustered clouds at every GPU count) ----------------
    if not args.no_other_cloud and n >= 4096:
        other = "plummer" if args.cloud == "uniform" else "uniform"
        _, parts_o = workload(n, W, other)
        eng.set_option(_abi.OPT_TIMING, 0)
        reupload(parts_o)  # same n: device buffers (and IPC mappings) stay put
        o_steps = max(1, min(args.steps, 3))
        one_step()
        flush.zero_()
        barrier()
        o0, o1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        o0.record(stream)
        for _ in range(o_steps):
            one_step()
            flush.zero_()
        o1.record(stream)
        barrier()
        o_ms = o0.elapsed_time(o1) / o_steps
        if world > 1:
            tt = torch.tensor([o_ms], device=f"cuda:{local}")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            o_ms = float(tt.item())
        line["other_cloud"] = {"cloud": other, "what": f"same N, W and constants on a {'clustered (Plummer a=W/6, truncated to the box)' if other == 'plummer' else 'uniform'} cloud",
                               "steps": o_steps, "warmup": 1, "ms_per_step": o_ms, "steps_per_s": 1e3 / o_ms,
                               "value": float(n) * n / (o_ms * 1e-3), "unit": UNIT}
        reupload(parts)

    # ---------------- "next" row: the cell-list path on the same workload ----------------
    # Reported as effective steps/s only: it evaluates ~30 candidates per particle instead of N, so it
    # must never be read against the FP32 roofline.
    if world == 1 and rank == 0 and not args.no_cells:
        eng.set_option(_abi.OPT_FORCE_KERNEL, _abi.FORCE_CELLS)
        eng.set_option(_abi.OPT_TIMING, 0)
        eng.upload(parts, prm["id_count"])
        eng.step(P, TS, 5)
        eng.sync()
        c_steps = 50
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        eng.step(P, TS, c_steps)
        e1.record(stream)
        torch.cuda.synchronize()
        c_ms = e0.elapsed_time(e1) / c_steps
        a_in[:] = parts
        eng.update_into(P, TS, a_in, a_out)
        t0 = time.perf_counter()
        for _ in range(5):
            eng.update_into(P, TS, a_in, a_out)
            a_in, a_out = a_out, a_in
        c_e2e = (time.perf_counter() - t0) / 5
        line["cell_list"] = {
            "what": "P3D_FORCE_CELLS: GPU uniform grid, the analogue of the reference's spatial hash (src/lib.rs:135-236); "
                    "same results within the parity tolerance; NOT all-pairs, so no roofline fraction",
            "ms_per_step": c_ms, "steps_per_s": 1e3 / c_ms,
            "e2e_ms_per_step": c_e2e * 1e3, "e2e_steps_per_s": 1.0 / c_e2e,
            "all_pairs_equivalent_interactions_per_s": float(n) * n / (c_ms * 1e-3),
        }
        eng.set_option(_abi.OPT_FORCE_KERNEL, _abi.FORCE_PAIR)

    # ---------------- the other BASELINE.json configs, for the record (parity-test cases, not bench lines) ----------------
    if world == 1 and rank == 0 and not args.no_cells:
        extras = {}
        for tag, n_x, W_x, plummer in (("config1_default_scene_n1000", 1000, 10.0, False),
                                       ("config2_uniform_n16384", 16384, 25.4, False),
                                       ("config3_plummer_n262144", 262144, 64.0, True)):
            prm_x = dict(prm, world_size=W_x)
            P_x = p3.Engine.make_params(**prm_x)
            parts_x = p3.generate_plummer(W_x, n_x, W_x / 6, seed=SEED) if plummer else p3.generate_particles(W_x, n_x, seed=SEED)
            row = {"n_particles": n_x, "world_size": W_x, "cloud": "plummer a=W/6" if plummer else "uniform"}
            for name, kernel in (("all_pairs", _abi.FORCE_PAIR if n_x >= 4096 else _abi.FORCE_REFERENCE_ORDER), ("cell_list", _abi.FORCE_CELLS)):
                ex = p3.Engine(local)
                ex.set_stream(stream.cuda_stream)
                ex.set_option(_abi.OPT_FORCE_KERNEL, kernel)
                ex.upload(parts_x, prm["id_count"])
                ex.step(P_x, TS, 6)
                ex.sync()
                reps = 20 if n_x <= 16384 or name == "cell_list" else 6
                x0, x1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                x0.record(stream)
                ex.step(P_x, TS, reps)
                x1.record(stream)
                torch.cuda.synchronize()
                ms_x = x0.elapsed_time(x1) / reps
                row[name] = {"ms_per_step": ms_x, "steps_per_s": 1e3 / ms_x}
                if name == "all_pairs":
                    row[name]["interactions_per_s"] = float(n_x) * n_x / (ms_x * 1e-3)
                ex.close()
            extras[tag] = row
        line["other_configs"] = extras

    # ---------------- CPU baseline (rank 0, N=1 only) ----------------
    if world == 1 and rank == 0 and not args.no_cpu:
        r = cpu_reference_sample(prm, parts, args.cpu_seconds)
        line["cpu_baseline"] = {
            "value": float(n) * n / r["step_s"], "unit": UNIT, "cores": r["cores"], "kind": "port",
            "sample": f"one step: counting sort over all {n} particles ({r['build_s']*1e3:.0f} ms) + per-particle pass for the first "
                      f"{r['sample']} particles ({r['force_sample_s']:.2f} s), scaled by N/sample; C restatement of src/lib.rs "
                      f"(spatial hash, faithful mode), {r['cores']} OpenMP threads",
            "ms_per_step": r["step_s"] * 1e3, "steps_per_s": 1.0 / r["step_s"],
            "candidate_pairs_per_s": r["candidates_per_particle"] * n / r["step_s"],
        }
        line["cpu_cores"] = r["cores"]
    if rank == 0:
        emit(line)
    if world > 1:
        dist.barrier()
        eng.ipc_close()
        dist.barrier()
    eng.close()
    if world > 1:
        dist.destroy_process_group()
    if rank == 0 and parity is not None and not parity["ok"]:
        sys.stderr.write(f"bench.py: PARITY FAILED: {json.dumps(parity)}\n")
        sys.exit(1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--particles", "--n", dest="n", type=int, default=N_DEFAULT,
                    help="particle count (use --particles under torch.distributed.run: its parser trips over --n)")
    ap.add_argument("--world-size", type=float, default=None, help="box edge W (default: density 1)")
    ap.add_argument("--block", type=int, default=0, choices=[0, 128, 256], help="0 = engine default")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="CPU baseline sample budget")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle parity step (it is outside the timed region)")
    ap.add_argument("--no-fused", action="store_true", help="multi-GPU: NCCL all-reduce + all-gather instead of the fused P2P kernel")
    ap.add_argument("--no-cells", action="store_true", help="skip the cell-list (SURVEY §8f-1) section")
    ap.add_argument("--cloud", default="uniform", choices=["uniform", "plummer"],
                    help="cloud the headline numbers are measured on (the other one is reported in `other_cloud`)")
    ap.add_argument("--no-other-cloud", action="store_true", help="skip the `other_cloud` section")
    ap.add_argument("--weak", action="store_true",
                    help="weak scaling (BASELINE.json config 5): N = 1,048,576 * sqrt(gpus), so that the pair count per GPU "
                         "stays that of the 1-GPU run; density 1")
    args = ap.parse_args()
    if args.weak:
        args.n = weak_particles(args.gpus)
        args.world_size = None if args.gpus == 1 else round(float(args.n) ** (1.0 / 3.0), 1)
    if args.world_size is None:
        args.world_size = W_DEFAULT if args.n == N_DEFAULT else round(float(args.n) ** (1.0 / 3.0), 1)
    if args.warmup < 3:
        args.warmup = 3  # timing rules: at least 3 warm-up steps
    _capture_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_engine(args)


if __name__ == "__main__":
    main()
