#!/usr/bin/env python
"""Summarise an `ncu --page raw --csv` export: per kernel duration, FP64 pipe utilisation,
executed FP64 flop, DRAM traffic, occupancy and the main stall reasons.
usage: ncu -i prof.ncu-rep --page raw --csv > raw.csv ; python tools/ncu_summary.py raw.csv [units_per_launch]"""
import csv
import json
import sys


def base_name(kernel_name):
    """'void band_kernel<(bool)1>(const double *, ...)' -> 'band_kernel'"""
    import re
    m = re.match(r"(?:void\s+)?(?:[A-Za-z_0-9]+::)*([A-Za-z_0-9]+)", kernel_name.strip())
    return m.group(1) if m else kernel_name


def main(path, units=1e6):
    rows = list(csv.reader(open(path)))
    hdr = rows[0]

    def g(r, name, default=None):
        if name not in hdr:
            return default
        v = r[hdr.index(name)].replace(",", "")
        try:
            return float(v)
        except ValueError:
            return default

    out = {}
    for r in rows[2:]:
        name = base_name(r[hdr.index("Kernel Name")])
        cyc = g(r, "smsp__cycles_elapsed.avg")
        d = {k: g(r, f"smsp__sass_thread_inst_executed_op_{k}_pred_on.sum.per_cycle_elapsed") * cyc
             for k in ("dadd", "dmul", "dfma")}
        flop = d["dadd"] + d["dmul"] + 2 * d["dfma"]
        t_ms = g(r, "gpu__time_duration.sum")
        unit = rows[1][hdr.index("gpu__time_duration.sum")]
        if unit == "us":
            t_ms /= 1e3
        elif unit == "ns":
            t_ms /= 1e6
        elif unit in ("s", "second"):
            t_ms *= 1e3
        rd = g(r, "dram__bytes_read.sum")
        wr = g(r, "dram__bytes_write.sum")
        scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
        rd *= scale[rows[1][hdr.index("dram__bytes_read.sum")]]
        wr *= scale[rows[1][hdr.index("dram__bytes_write.sum")]]
        info = {
            "duration_ms": t_ms,
            "registers": g(r, "launch__registers_per_thread"),
            "warps_active_pct": g(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
            "fp64_pipe_pct": g(r, "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
            "issue_active_pct": g(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
            "fma_pipe_pct": g(r, "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
            "xu_pipe_pct": g(r, "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
            "alu_pipe_pct": g(r, "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
            "lsu_pipe_pct": g(r, "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
            "threads_per_inst": g(r, "smsp__thread_inst_executed_per_inst_executed.ratio"),
            "warp_inst": g(r, "smsp__inst_executed.sum"),
            "fp64_thread_inst_per_unit": {k: v / units for k, v in d.items()},
            "fp64_flop_per_unit": flop / units,
            "achieved_fp64_tflops": flop / (t_ms * 1e-3) / 1e12,
            "dram_read_bytes": rd, "dram_write_bytes": wr,
            "dram_pct": g(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
            "stalls_per_issue": {h.split("stalled_")[1].split("_per")[0]: float(r[i]) for i, h in enumerate(hdr)
                                 if h.startswith("smsp__average_warps_issue_stalled") and h.endswith(
                                     "_per_issue_active.ratio") and float(r[i]) > 0.1},
        }
        out[name] = info
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 1e6)
