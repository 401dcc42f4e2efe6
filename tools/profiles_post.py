#!/usr/bin/env python
"""Turn gpurun_out/<tag>_* (written by tools/collect_profiles.sh on the GPU box) into the tracked evidence
under profiles/:
  <tag>_bench.json            the bench line measured in the same gpurun call
  <tag>_launches.csv          ncu launch list (gpu__time_duration.sum per launch)
  <tag>_counts.json           executed FP64 / FP32 / SFU instructions and DRAM bytes per simulation, per config
  flop_per_sample.json        executed flop per simulation, per config and kernel (read by bench.py)
  dram_traffic.json           dram bytes read+written per simulation, per config and kernel (read by bench.py)
  <tag>_ncu_<name>.json       per-kernel metrics of the --set full captures (pipes, issue, occupancy, stalls)
usage: python tools/profiles_post.py r02"""
import csv
import io
import json
import re
import subprocess
import sys
from contextlib import redirect_stdout
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "tools"))
sys.path.insert(0, str(ROOT))
import ncu_summary  # noqa: E402

FAMILY = {"lidf_kernel": "lidf_kernel", "geometry_kernel": "geometry_kernel", "band_kernel": "band_kernel",
          "geometry_kernel_f32": "geometry_kernel_f32", "band_kernel_f32": "band_kernel_f32"}


def read_counts(path):
    """ncu --csv log with one row per (launch, metric) -> {kernel: {metric: summed value, 'launches': k}}"""
    rows = list(csv.reader(open(path)))
    h = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[h]
    out = {}
    seen = set()
    for r in rows[h + 1:]:
        if len(r) != len(hdr):
            continue
        d = dict(zip(hdr, r))
        k = ncu_summary.base_name(d["Kernel Name"])
        v = float(d["Metric Value"].replace(",", ""))
        unit = d["Metric Unit"]
        scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1e-3, "ns": 1e-6, "ms": 1.0, "s": 1e3}.get(unit, 1.0)
        e = out.setdefault(k, {})
        e[d["Metric Name"]] = e.get(d["Metric Name"], 0.0) + v * scale
        if (d["ID"], k) not in seen:
            seen.add((d["ID"], k))
            e["launches"] = e.get("launches", 0) + 1
    return out


def main(tag):
    go, prof = ROOT / "gpurun_out", ROOT / "profiles"
    prof.mkdir(exist_ok=True)
    import bench
    line = json.loads((go / f"{tag}_bench.json").read_text().strip().splitlines()[-1])
    (prof / f"{tag}_bench.json").write_text(json.dumps(line, indent=1) + "\n")
    build = line["roofline"].get("kernel_build") or bench.kernel_source_sha()
    if build != bench.kernel_source_sha():
        print("WARNING: the collected profile belongs to build", build, "but the tree holds", bench.kernel_source_sha())

    # ---- counters per config -> flop_per_sample.json / dram_traffic.json
    flop, traffic, counts = {"src_sha": build}, {}, {}
    S = "smsp__sass_thread_inst_executed_op_%s_pred_on.sum"
    for f in sorted(go.glob(f"{tag}_counts_cfg*_*.csv")):
        m = re.match(rf"{tag}_counts_cfg(\d)_(fp\d\d)\.csv", f.name)
        cfg, prec = int(m.group(1)), m.group(2)
        n = 262144 if cfg != 4 else 32768          # tools/profile_step.py defaults
        c = read_counts(f)
        key = f"cfg{cfg}" + ("" if prec == "fp64" else "_fp32")
        counts[key] = {}
        for k, e in c.items():
            d64 = (e.get(S % "dadd", 0) + e.get(S % "dmul", 0) + 2 * e.get(S % "dfma", 0)) / n
            f32 = (e.get(S % "fadd", 0) + e.get(S % "fmul", 0) + 2 * e.get(S % "ffma", 0)) / n
            # xu pipe: warp-level instructions; lanes active per instruction from the ratio metric (averaged)
            lanes = e.get("smsp__thread_inst_executed_per_inst_executed.ratio", 32.0 * e["launches"]) / e["launches"]
            mufu = e.get("smsp__inst_executed_pipe_xu.sum", 0) * lanes / n
            by = (e.get("dram__bytes_read.sum", 0) + e.get("dram__bytes_write.sum", 0)) / n
            counts[key][k] = {"launches_per_step": e["launches"], "fp64_flop": d64, "fp32_flop": f32, "mufu": mufu,
                              "dram_bytes": by, "ms_per_step_under_ncu": e.get("gpu__time_duration.sum", 0.0),
                              "fp64_inst": {x: e.get(S % x, 0) / n for x in ("dadd", "dmul", "dfma")}}
        if prec == "fp64":
            flop[key] = {k: v["fp64_flop"] for k, v in counts[key].items() if k in bench.KERNELS}
            traffic[key] = {k: v["dram_bytes"] for k, v in counts[key].items() if k in bench.KERNELS}
        else:
            flop[key] = {k: {"fp32_flop": v["fp32_flop"], "mufu": v["mufu"], "fp64_flop": v["fp64_flop"]}
                         for k, v in counts[key].items()}
    (prof / f"{tag}_counts.json").write_text(json.dumps(counts, indent=1) + "\n")
    (prof / "flop_per_sample.json").write_text(json.dumps(flop, indent=1) + "\n")
    (prof / "dram_traffic.json").write_text(json.dumps(traffic, indent=1) + "\n")

    # ---- launch list: keep id, kernel, grid, block, duration
    lf = go / f"{tag}_launches.csv"
    if lf.exists():
        rows = list(csv.reader(open(lf)))
        h = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
        hdr = rows[h]
        keep = ["ID", "Kernel Name", "Grid Size", "Block Size", "Metric Name", "Metric Unit", "Metric Value"]
        dur = {}          # (kernel, grid) -> durations; the full-batch launches are those with the largest grid
        with open(prof / f"{tag}_launches.csv", "w", newline="") as f:
            w = csv.writer(f)
            w.writerow(keep)
            for r in rows[h + 1:]:
                if len(r) != len(hdr):
                    continue
                d = dict(zip(hdr, r))
                d["Kernel Name"] = ncu_summary.base_name(d["Kernel Name"])
                w.writerow([d[k] for k in keep])
                blocks = 1
                for x in re.findall(r"\d+", d["Grid Size"]):
                    blocks *= int(x)
                dur.setdefault((d["Kernel Name"], blocks), []).append(float(d["Metric Value"].replace(",", "")))
        fp64 = {}
        for k in bench.KERNELS:
            grids = [g for (kk, g) in dur if kk == k]
            if grids:
                v = dur[(k, max(grids))]
                fp64[k] = sum(v) / len(v)
        tot = sum(fp64.values())
        print("ncu launch-list shares:", {k: round(v / tot, 3) for k, v in fp64.items()})
        km = line["roofline"]["kernel_ms"]
        print("CUDA-event shares:     ", {k: round(v / sum(km.values()), 3) for k, v in km.items()})

    # ---- full captures
    for rep in sorted(go.glob(f"{tag}_prof_*.ncu-rep")):
        name = rep.stem[len(tag) + 6:]
        raw = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True,
                             check=True).stdout
        tmp = go / f"{tag}_raw_{name}.csv"
        tmp.write_text(raw)
        buf = io.StringIO()
        units = {"cfg2": 1e6, "cfg3": 1e6, "cfg2_fp32": 1e6, "srf": 131072, "spectrum": 4096, "lut": 20000}.get(name, 1e6)
        with redirect_stdout(buf):
            ncu_summary.main(str(tmp), units)
        summ = json.loads(buf.getvalue())
        (prof / f"{tag}_ncu_{name}.json").write_text(json.dumps(summ, indent=1) + "\n")
        print(name, {k: {kk: (round(v[kk], 3) if isinstance(v[kk], float) else v[kk]) for kk in
                         ("duration_ms", "fp64_pipe_pct", "issue_active_pct", "warps_active_pct", "dram_pct")}
                     for k, v in summ.items()})


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "r02")
