"""The multi-device handle (p3d_create_multi) on BASELINE.json config 4: ONE process, ONE engine handle over G devices,
the unchanged p3d_update / p3d_step calls.  Prints one JSON line: device-resident ms/step (wall clock around p3d_step +
p3d_sync, the handle has no per-kernel timing), end-to-end ms/step through p3d_update with pinned host arrays, and the
oracle parity sample of bench.py.  Builder-run evidence next to `bench.py --gpus G` (one process per GPU over NCCL).
Usage: python tools/multi_handle_bench.py G [n=1048576] [steps=5]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "3d-particle-simulation-_b200"))
sys.path.insert(0, ROOT)
import torch  # pinned host memory only

import bench
import particle_3d as p3
from particle_3d import _abi
from oracle import oracle as O

G = int(sys.argv[1]) if len(sys.argv) > 1 else 2
n = int(sys.argv[2]) if len(sys.argv) > 2 else bench.N_DEFAULT
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
W = bench.W_DEFAULT if n == bench.N_DEFAULT else round(float(n) ** (1.0 / 3.0), 1)
TS = bench.TS
prm, parts = bench.workload(n, W)
P = p3.Engine.make_params(**prm)
devs = list(range(G)) if torch.cuda.device_count() >= G else [0] * G
eng = p3.Engine(devs)
eng.set_option(_abi.OPT_FORCE_KERNEL, _abi.FORCE_PAIR)

# device-resident
eng.upload(parts, prm["id_count"])
eng.step(P, TS, 3)
eng.sync()
t0 = time.perf_counter()
eng.step(P, TS, steps)
eng.sync()
dev_ms = (time.perf_counter() - t0) / steps * 1e3

# end to end: p3d_update on pinned host arrays
hin = torch.empty(n * 28, dtype=torch.uint8).pin_memory()
hout = torch.empty(n * 28, dtype=torch.uint8).pin_memory()
a_in, a_out = hin.numpy().view(_abi.PARTICLE), hout.numpy().view(_abi.PARTICLE)
a_in[:] = parts
eng.update_into(P, TS, a_in, a_out)
a_in[:] = parts
t0 = time.perf_counter()
for _ in range(steps):
    eng.update_into(P, TS, a_in, a_out)
    a_in, a_out = a_out, a_in
e2e_ms = (time.perf_counter() - t0) / steps * 1e3

# parity: one p3d_update from the seed state against the oracle, sample over every device's slot range
a_in[:] = parts
eng.update_into(P, TS, a_in, a_out)
slot = eng.slot_of().astype(np.int64)
n_slots = int(slot.max()) + 1
per = -(-n_slots // G)
owner = np.minimum(slot // per, G - 1)
rng = np.random.default_rng(20261018)
idx = np.concatenate([rng.choice(np.flatnonzero(owner == g), size=min(bench.PARITY_SAMPLE // G, int((owner == g).sum())), replace=False)
                      for g in range(G)])
ref, _ = O.update_indices(prm, TS, parts, idx, mode=O.IDEAL)
dv, dp = bench.parity_sample_errors(a_out[idx], ref, W)
c = eng.counters()
print(json.dumps({"what": "p3d_create_multi: one process, one handle", "devices": devs, "n_particles": n, "world_size": W, "steps": steps,
                  "device_resident_ms_per_step": dev_ms, "e2e_ms_per_step": e2e_ms,
                  "e2e_api": "p3d_update(handle, params, ts, in, out, n), pinned host arrays, wall clock",
                  "interactions_per_s": float(n) * n / (dev_ms * 1e-3), "e2e_interactions_per_s": float(n) * n / (e2e_ms * 1e-3),
                  "parity": {"max_dv_over_tol": dv, "max_dp_over_tol": dp, "n_checked": int(idx.size), "mode": "ideal",
                             "devices_covered": sorted(int(g) for g in np.unique(owner[idx])),
                             "ids_and_order_exact": bool(np.array_equal(a_out["id"], parts["id"])), "ok": bool(dv <= 1 and dp <= 1)},
                  "kernel_launches": c["kernels"]}))
eng.close()
