"""Condenses `ncu -i X.ncu-rep --page raw --csv` exports into profiles/force_pair_ncu_summary.json,
the file bench.py reads for `roofline.traffic` and `roofline.ncu`.  Not part of the product.

  ncu -i gpurun_out/r01_force_pair_n1m.ncu-rep --page raw --csv > profiles/r01_kernels_n1048576_raw.csv
  ncu -i gpurun_out/r01_kernels_n262144.ncu-rep --page raw --csv > profiles/r01_kernels_n262144_raw.csv
  ncu -i gpurun_out/r01_cells_n1m.ncu-rep        --page raw --csv > profiles/r01_cells_n1048576_raw.csv
  ncu -i gpurun_out/r01_layout_n1m.ncu-rep       --page raw --csv > profiles/r01_layout_n1048576_raw.csv
  python tools/ncu_summary.py
"""
import csv
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PROF = os.path.join(ROOT, "profiles")

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}
STALLS = ("math_pipe_throttle", "not_selected", "wait", "dispatch_stall", "long_scoreboard", "short_scoreboard", "mio_throttle")


def kernels(path):
    rows = list(csv.reader(open(path)))
    hdr, unit = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}

    def val(r, name, scale=False):
        i = ix[name]
        v = float(r[i].replace(",", "")) if r[i] not in ("", "n/a") else 0.0
        return v * UNIT.get(unit[i], 1.0) if scale else v

    out = []
    for r in rows[2:]:
        k = {
            "kernel": r[ix["Kernel Name"]].split("(")[0].replace("void ", ""),
            "duration_ms": val(r, "gpu__time_duration.sum", True),
            "dram_bytes_per_launch": val(r, "dram__bytes_read.sum", True) + val(r, "dram__bytes_write.sum", True),
            "fma_pipe_active_pct": val(r, "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
            "issue_slots_busy_pct": val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
            "alu_pipe_pct": val(r, "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
            "xu_pipe_pct": val(r, "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
            "registers_per_thread": val(r, "launch__registers_per_thread"),
            "warps_active_pct": val(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
            "dram_throughput_pct": val(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
            "stall_cycles_per_issue": {s: round(val(r, f"smsp__average_warps_issue_stalled_{s}_per_issue_active.ratio"), 3)
                                       for s in STALLS},
        }
        out.append(k)
    return out


def main():
    main_csv = "r02_kernels_n1048576_raw.csv" if os.path.exists(os.path.join(PROF, "r02_kernels_n1048576_raw.csv")) \
        else "r01_kernels_n1048576_raw.csv"
    top = kernels(os.path.join(PROF, main_csv))[0]
    summary = {"source": f"profiles/{main_csv} (ncu --set full --clock-control none --import-source on, one launch, N=1,048,576)",
               "n_particles": 1048576}
    summary.update(top)
    for key, name in (("other_kernels_n262144", "r01_kernels_n262144_raw.csv"),
                      ("cell_list_kernels_n1048576", "r02_cells_kernel_n1048576_raw.csv"),
                      ("cell_list_kernels_n1048576_round1", "r01_cells_n1048576_raw.csv"),
                      ("layout_kernels_n1048576", "r01_layout_n1048576_raw.csv")):
        p = os.path.join(PROF, name)
        if os.path.exists(p):
            summary[key] = {"source": f"profiles/{name}", "kernels": kernels(p)}
    json.dump(summary, open(os.path.join(PROF, "force_pair_ncu_summary.json"), "w"), indent=1)
    print(json.dumps({k: v for k, v in summary.items() if not isinstance(v, dict) or k == "stall_cycles_per_issue"}, indent=1))


if __name__ == "__main__":
    main()
