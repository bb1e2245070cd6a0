"""Developer GPU probe: parity vs oracle + kernel timings + FP32 microbench.  Not part of the product."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "3d-particle-simulation-_b200"))
sys.path.insert(0, ROOT)
import particle_3d as p3
from particle_3d import _abi
from oracle import oracle as O
from tools import microbench as _mb


def parity(n, W, kernel, steps=1, seed=42, plummer=False, block=0, **over):
    prm = p3.default_params_dict()
    prm["world_size"] = W
    prm.update(over)
    parts = p3.generate_plummer(W, n, W / 6, seed) if plummer else p3.generate_particles(W, n, seed)
    eng = p3.Engine(0)
    eng.set_option(_abi.OPT_FORCE_KERNEL, kernel)
    eng.set_option(_abi.OPT_BLOCK_SIZE, block)
    P = p3.Engine.make_params(**prm)
    cur = parts.copy()
    ref = parts.copy()
    for _ in range(steps):
        cur = eng.update(P, 1 / 60, cur)
        ref = O.update(prm, 1 / 60, ref, mode=O.IDEAL)["out"]
    v = np.stack([cur["vx"], cur["vy"], cur["vz"]], 1).astype(np.float64)
    vr = np.stack([ref["vx"], ref["vy"], ref["vz"]], 1).astype(np.float64)
    p = np.stack([cur["px"], cur["py"], cur["pz"]], 1).astype(np.float64)
    pr = np.stack([ref["px"], ref["py"], ref["pz"]], 1).astype(np.float64)
    vrms = np.sqrt((vr ** 2).sum(1).mean())
    dv = np.linalg.norm(v - vr, axis=1) / np.maximum(np.linalg.norm(vr, axis=1), vrms)
    dp = np.linalg.norm(p - pr, axis=1) / np.maximum(np.linalg.norm(pr, axis=1), W / 2)
    print(f"parity n={n} W={W} kernel={kernel} steps={steps} {over}: max dv={dv.max():.3e} (n>1e-5: {(dv>1e-5).sum()}) "
          f"max dp={dp.max():.3e} ids_ok={np.array_equal(cur['id'], ref['id'])}", flush=True)
    eng.close()


class Clocks:
    """Samples nvidia-smi SM clock / power / throttle reasons while a region runs."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active"

    def __enter__(self):
        import subprocess
        self.p = subprocess.Popen(["nvidia-smi", "-i", "0", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50"],
                                  stdout=subprocess.PIPE, text=True)
        return self

    def __exit__(self, *a):
        self.p.terminate()
        out = self.p.communicate()[0].strip().splitlines()
        rows = [r.split(", ") for r in out if r]
        clk = sorted(float(r[0]) for r in rows)
        pw = [float(r[2]) for r in rows]
        self.summary = dict(samples=len(rows), sm_mhz_median=clk[len(clk) // 2] if clk else None,
                            sm_mhz_min=clk[0] if clk else None, sm_max=float(rows[0][1]) if rows else None,
                            power_max=max(pw) if pw else None, reasons=sorted(set(r[3] for r in rows)))


def timing(n, W, kernel, steps=3, plummer=False, block=0, tune=0):
    prm = p3.default_params_dict()
    prm["world_size"] = W
    parts = p3.generate_plummer(W, n, W / 6, 42) if plummer else p3.generate_particles(W, n, 42)
    eng = p3.Engine(0)
    eng.set_option(_abi.OPT_FORCE_KERNEL, kernel)
    eng.set_option(_abi.OPT_TIMING, 1)
    eng.set_option(_abi.OPT_BLOCK_SIZE, block)
    P = p3.Engine.make_params(**prm)
    eng.upload(parts, 5)
    eng.step(P, 1 / 60, 1)
    eng.sync()
    with Clocks() as ck:
        t0 = time.time()
        eng.step(P, 1 / 60, steps)
        eng.sync()
        wall = (time.time() - t0) / steps
    t = eng.timing()
    f_ms = t["force"] / steps
    i_ms = t["integrate"] / steps
    pairs = float(n) * n
    print(f"timing n={n} kernel={kernel}: force {f_ms:.3f} ms  integrate {i_ms*1e3:.1f} us  wall/step {wall*1e3:.3f} ms  "
          f"{pairs/f_ms/1e9:.2f} T interactions/s  = {pairs*20/(f_ms*1e-3)/74.5e12*100:.1f}% of 74.5 TF  "
          f"integrate {80*n/i_ms/1e6:.0f} GB/s  B={block} tune={tune} pair={t["pair"]/steps:.3f}ms bxb={t["bxb"]/steps:.3f}ms part={t["partition"]/steps*1e3:.0f}us  clocks={ck.summary}", flush=True)
    eng.close()


def micro():
    L = _mb.load()
    for kind, name in ((0, "FFMA"), (1, "FFMA2"), (2, "pair-mix"), (3, "FFMA2+SHFL")):
        out = (C.c_double * 4)()
        rc = L.p3d_microbench(0, kind, 2000, out)
        print(f"micro {name}: rc={rc} {out[0]/1e12:.2f} T lane-FMA/s  ({out[0]*2/1e12:.1f} TFLOP/s)  {out[1]:.2f} ms  "
              f"sms={int(out[2])} clk={out[3]:.0f} MHz  -> per SM per clk @max: {out[0]/out[2]/(out[3]*1e6):.1f}", flush=True)


def micro2():
    L = _mb.load()
    names = {4: "6 FMNMX", 5: "2 MUFU.RSQ", 6: "3 SHFL", 7: "17 FFMA2 + 6 FMNMX", 8: "17 FFMA2 + 2 MUFU",
             9: "34 FFMA + 6 FMNMX + 2 MUFU", 10: "17 FFMA2 + 6 FMNMX + 2 MUFU", 11: "17 FFMA2 + 6 FMNMX + 2 MUFU + 3 SHFL",
             12: "17 FFMA2"}
    for kind in sorted(names):
        out = (C.c_double * 4)()
        rc = L.p3d_microbench(0, kind, 1000, out)
        warp_bodies_per_s = out[0] / 32
        cyc = out[3] * 1e6 * out[2] * 4 / warp_bodies_per_s
        print(f"micro2 [{names[kind]}]: rc={rc} {out[1]:.2f} ms -> {cyc:.1f} SMSP-cycles per warp-body (at max clock)", flush=True)


def micro3():
    L = _mb.load()
    names = {13: "8 FFMA2 d=a*b+d (3 distinct pairs)", 14: "8 FFMA2 d=a*a+d (2 distinct)", 15: "8 FFMA2 d=a*s+d (pair, scalar, pair)",
             16: "8 FFMA2 d=a*b+d, b shared by consecutive instrs", 17: "12 FFMA2 + 12 FFMA (disjoint)"}
    for kind in sorted(names):
        out = (C.c_double * 4)()
        rc = L.p3d_microbench(0, kind, 4000, out)
        cyc = out[3] * 1e6 * out[2] * 4 / (out[0] / 32)
        print(f"micro3 [{names[kind]}]: rc={rc} {out[1]:.2f} ms -> {cyc:.2f} SMSP-cycles per body", flush=True)


def micro4():
    L = _mb.load()
    names = {0: "F2 stream only (17 F2)", 1: "+MUFU", 2: "+FMNMX", 3: "+MUFU +FMNMX (= kernel body)", 4: "F2 only, no j-side (13 F2)",
             7: "kernel body, no j-side", 8: "F2 only, no i-side (14 F2 + 1 FADD2)", 11: "kernel body, no i-side", 12: "F2 only, no accumulation (11 F2)",
             16: "F2 only, i-positions as pairs", 19: "kernel body, i-positions as pairs", 32: "F2 only, no FADD2 (14 F2)", 35: "kernel body, no FADD2"}
    for v in sorted(names):
        out = (C.c_double * 4)()
        rc = L.p3d_microbench(0, 18 + v, 4000, out)
        cyc = out[3] * 1e6 * out[2] * 4 / (out[0] / 32)
        print(f"micro4 [{names[v]:40s}] rc={rc} -> {cyc:6.2f} SMSP-cycles per pair-pack", flush=True)


if __name__ == "__main__":
    what = sys.argv[1:] or ["parity", "timing", "micro"]
    if "micro5" in what:
        L = _mb.load()
        for kind, name in ((90, "two-phase G=2, F2 only (16 F2)"), (91, "two-phase G=4, F2 only"), (92, "two-phase G=8, F2 only"),
                           (93, "two-phase G=2, full body"), (94, "two-phase G=4, full body"), (95, "two-phase G=8, full body")):
            out = (C.c_double * 4)()
            rc = L.p3d_microbench(0, kind, 4000, out)
            cyc = out[3] * 1e6 * out[2] * 4 / (out[0] / 32)
            print(f"micro5 [{name:34s}] rc={rc} -> {cyc:6.2f} SMSP-cycles per pair-pack", flush=True)
    if "micro4" in what:
        micro4()
    if "micro3" in what:
        micro3()
    if "micro" in what:
        micro()
        micro2()
    if "parity" in what:
        parity(1000, 10.0, _abi.FORCE_REFERENCE_ORDER)
        parity(1000, 10.0, _abi.FORCE_PAIR)
        parity(1000, 10.0, _abi.FORCE_REFERENCE_ORDER, steps=10)
        parity(16384, 25.4, _abi.FORCE_PAIR)
        parity(16384, 25.4, _abi.FORCE_PAIR, block=256)
        parity(16384, 25.4, _abi.FORCE_REFERENCE_ORDER)
        parity(16384, 25.4, _abi.FORCE_PAIR, walls=True, acceleration=(0.0, -1.0, 0.0))
        parity(16384, 25.4, _abi.FORCE_PAIR, particle_effect_radius=0.8)
        parity(20000, 64.0, _abi.FORCE_PAIR, plummer=True)
    if "oob" in what:  # one particle outside the box: the general cell-list variant instead of an O(N^2) fallback
        prm = p3.default_params_dict(); prm["world_size"] = 101.6
        parts = p3.generate_particles(101.6, 1048576, 42); parts["px"][7] += 101.6
        eng = p3.Engine(0); eng.set_option(_abi.OPT_FORCE_KERNEL, _abi.FORCE_CELLS); eng.set_option(_abi.OPT_TIMING, 1)
        P = p3.Engine.make_params(**prm); eng.upload(parts, 5); eng.step(P, 0.0, 1); eng.sync(); eng.step(P, 0.0, 5); t = eng.timing()
        print(f"oob n=1048576 cells-general: force {t['force']/5:.3f} ms/step", flush=True)
        ref = O.update(prm, 0.0, parts, mode=O.IDEAL, want_force=True)["force"]
        f = eng.download_forces(); print("oob max |dF|", np.abs(f - ref).max(), flush=True); eng.close()
    if "graph" in what:
        for n, W in ((1000, 10.0), (16384, 25.4), (262144, 64.0), (1048576, 101.6)):
            for kernel in (_abi.FORCE_CELLS, _abi.FORCE_REFERENCE_ORDER) if n == 1000 else (_abi.FORCE_CELLS,):
                for graph in (0, 1):
                    prm = p3.default_params_dict(); prm["world_size"] = W
                    eng = p3.Engine(0); eng.set_option(_abi.OPT_FORCE_KERNEL, kernel); eng.set_option(_abi.OPT_GRAPH, graph)
                    P = p3.Engine.make_params(**prm)
                    eng.upload(p3.generate_particles(W, n, 42), 5)
                    eng.step(P, 1 / 60, 10); eng.sync()
                    t0 = time.time(); eng.step(P, 1 / 60, 400); eng.sync(); dt = (time.time() - t0) / 400
                    print(f"graph={graph} n={n} kernel={kernel}: {dt*1e6:.1f} us/step", flush=True)
                    eng.close()
    if "cells1m" in what:
        timing(1048576, 101.6, _abi.FORCE_CELLS, steps=5)
    if "cells" in what:
        timing(1000, 10.0, _abi.FORCE_CELLS, steps=50)
        timing(16384, 25.4, _abi.FORCE_CELLS, steps=50)
        timing(262144, 64.0, _abi.FORCE_CELLS, steps=20)
        timing(262144, 64.0, _abi.FORCE_CELLS, steps=20, plummer=True)
        timing(1048576, 101.6, _abi.FORCE_CELLS, steps=20)
        timing(4194304, 161.3, _abi.FORCE_CELLS, steps=10)
    if "tune" in what:
        for blk, tune in ((128, 0), (256, 0)):
            timing(262144, 64.0, _abi.FORCE_PAIR, steps=4, block=blk, tune=tune)
    if "timing" in what:
        timing(1000, 10.0, _abi.FORCE_REFERENCE_ORDER, steps=20)
        timing(16384, 25.4, _abi.FORCE_REFERENCE_ORDER, steps=5)
        timing(16384, 25.4, _abi.FORCE_PAIR, steps=20)
        timing(16384, 25.4, _abi.FORCE_PAIR, steps=20, block=256)
        timing(262144, 64.0, _abi.FORCE_PAIR, steps=3)
        timing(262144, 64.0, _abi.FORCE_PAIR, steps=3, block=256)
        timing(1048576, 101.6, _abi.FORCE_PAIR, steps=4)
        timing(1048576, 101.6, _abi.FORCE_PAIR, steps=4, block=256)
