"""Builds many variants of k_force_pair (the P3D_PG_* source-order knobs of pair_group, R, resident CTAs) with nvcc
in parallel and times each on the device with pair_probe.  Run it ON THE GPU BOX (it has nvcc):
    python tools/pair_search/search.py [n_random=120] [seed=1] [n=262144]
Writes gpurun_out/pair_search_<seed>.txt (sorted) — copied to profiles/ by the builder."""
import concurrent.futures as cf
import itertools
import os
import random
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
CSRC = os.path.join(ROOT, "3d-particle-simulation-_b200", "csrc")
SRC = os.path.join(ROOT, "tools", "pair_search", "pair_probe.cu")
OUT = os.path.join(ROOT, "build", "scratch", "pair_search")
os.makedirs(OUT, exist_ok=True)
n_random = int(sys.argv[1]) if len(sys.argv) > 1 else 120
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
n = int(sys.argv[3]) if len(sys.argv) > 3 else 262144

KNOBS = {"P3D_PG_FADD_SWAP": (0, 1), "P3D_PG_D2_ORDER": (0, 1, 2, 3, 4, 5), "P3D_PG_LAW_SWAP": (0, 1, 2, 3), "P3D_PG_RS_FADD": (0, 1),
         "P3D_PG_ACCI_SWAP": (0, 1), "P3D_PG_ACCJ_SWAP": (0, 1), "P3D_PG_ACCI_ORDER": (0, 1, 2, 3, 4, 5),
         "P3D_PG_ACCJ_ORDER": (0, 1, 2, 3, 4, 5), "P3D_PG_STAGE": (0, 1, 2), "P3D_PG_S_SPLIT": (0, 1), "P3D_PG_IMM": (0, 1),
         "PROBE_MINB": (11, 12, 13), "PROBE_R": (8,)}
base = {k: v[0] for k, v in KNOBS.items()}
base["PROBE_MINB"] = 12
variants = [dict(base)]
for k, vals in KNOBS.items():           # every single-knob change from the shipped variant
    for v in vals:
        if v != base[k]:
            variants.append(dict(base, **{k: v}))
variants += [dict(base, PROBE_R=6, PROBE_MINB=m) for m in (12, 14, 16)] + [dict(base, PROBE_R=4, PROBE_MINB=16)]
rng = random.Random(seed)
for _ in range(n_random):
    variants.append({k: rng.choice(v) for k, v in KNOBS.items()})
uniq, seen = [], set()
for v in variants:
    key = tuple(sorted(v.items()))
    if key not in seen:
        seen.add(key)
        uniq.append(v)


def build(iv):
    i, v = iv
    exe = os.path.join(OUT, f"probe_{seed}_{i}")
    cmd = ["nvcc", "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-I", CSRC] + \
          [f"-D{k}={val}" for k, val in v.items()] + [SRC, "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    return i, exe if r.returncode == 0 else None, r.stderr[-300:]


with cf.ThreadPoolExecutor(max_workers=max(2, (os.cpu_count() or 4) - 1)) as ex:
    built = list(ex.map(build, enumerate(uniq)))
rows = []
for i, exe, err in built:
    if not exe:
        rows.append((1e9, f"BUILD FAILED {uniq[i]} {err}"))
        continue
    r = subprocess.run([exe, str(n), "4"], capture_output=True, text=True)
    line = r.stdout.strip() or ("RUN FAILED " + r.stderr[-200:])
    try:
        ms = float(line.split()[1])
    except Exception:
        ms = 1e9
    diff = {k: v for k, v in uniq[i].items() if v != base.get(k)}
    rows.append((ms, f"{line}  | {diff if diff else 'SHIPPED'}"))
    os.remove(exe)
rows.sort(key=lambda x: x[0])
path = os.path.join(ROOT, "gpurun_out", f"pair_search_{seed}.txt")
os.makedirs(os.path.dirname(path), exist_ok=True)
with open(path, "w") as f:
    for ms, line in rows:
        f.write(line + "\n")
print("\n".join(l for _, l in rows[:25]))
print("...")
print("\n".join(l for _, l in rows[-5:]))
shipped = [l for _, l in rows if l.endswith("SHIPPED")]
print("shipped:", shipped)
