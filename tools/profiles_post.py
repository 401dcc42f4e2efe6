#!/usr/bin/env python
"""Turn gpurun_out/<tag>_* (written by tools/collect_profiles.sh on the GPU box) into the
tracked evidence under profiles/:
  <tag>_launches.csv          ncu launch list (gpu__time_duration.sum per launch)
  <tag>_ncu_summary.json      per-kernel metrics of the --set full capture
  <tag>_bench.json            the bench line measured in the same gpurun call
  flop_per_sample.json        executed FP64 flop per simulation and kernel (read by bench.py)
  dram_traffic.json           dram bytes read+written per launch and kernel (read by bench.py)
usage: python tools/profiles_post.py r01"""
import csv
import io
import json
import subprocess
import sys
from contextlib import redirect_stdout
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "tools"))
import ncu_summary  # noqa: E402


def main(tag):
    go, prof = ROOT / "gpurun_out", ROOT / "profiles"
    prof.mkdir(exist_ok=True)
    raw = subprocess.run(["ncu", "-i", str(go / f"{tag}_prof.ncu-rep"), "--page", "raw", "--csv"],
                         capture_output=True, text=True, check=True).stdout
    tmp = go / f"{tag}_raw.csv"
    tmp.write_text(raw)
    buf = io.StringIO()
    with redirect_stdout(buf):
        ncu_summary.main(str(tmp), 1e6)
    summ = json.loads(buf.getvalue())
    (prof / f"{tag}_ncu_summary.json").write_text(json.dumps(summ, indent=1) + "\n")
    flop = {k: v["fp64_flop_per_unit"] for k, v in summ.items()}
    sys.path.insert(0, str(ROOT))
    import bench
    # the counts belong to the kernel build that ran under ncu: take its identity from the bench line
    # written by the same gpurun call, not from whatever library is in the tree now
    plain = json.loads((go / f"{tag}_bench_plain.json").read_text().strip().splitlines()[-1])
    flop["src_sha"] = plain["roofline"].get("kernel_build") or bench.kernel_source_sha()
    if flop["src_sha"] != bench.kernel_source_sha():
        print("WARNING: the collected profile belongs to build", flop["src_sha"], "but the tree holds",
              bench.kernel_source_sha())
    (prof / "flop_per_sample.json").write_text(json.dumps(flop, indent=1) + "\n")
    traffic = {k: v["dram_read_bytes"] + v["dram_write_bytes"] for k, v in summ.items()}
    (prof / "dram_traffic.json").write_text(json.dumps(traffic, indent=1) + "\n")
    # launch list: keep id, kernel, grid, block, duration
    rows = list(csv.reader(open(go / f"{tag}_launches.csv")))
    h = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[h]
    keep = ["ID", "Kernel Name", "Grid Size", "Block Size", "Metric Name", "Metric Unit", "Metric Value"]
    with open(prof / f"{tag}_launches.csv", "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(keep)
        for r in rows[h + 1:]:
            d = dict(zip(hdr, r))
            d["Kernel Name"] = ncu_summary.base_name(d["Kernel Name"])
            w.writerow([d[k] for k in keep])
    for name in (f"{tag}_bench.json",):
        line = (go / name).read_text().strip().splitlines()[-1]
        (prof / name).write_text(json.dumps(json.loads(line), indent=1) + "\n")
    # shares: launch list vs CUDA events
    dur = {}
    for r in rows[h + 1:]:
        d = dict(zip(hdr, r))
        k = ncu_summary.base_name(d["Kernel Name"])
        if k.endswith("_kernel") and "fma_chain" not in k:
            dur.setdefault(k, []).append(float(d["Metric Value"].replace(",", "")))
    tot = sum(sum(v) / len(v) for v in dur.values())
    print("ncu launch-list shares:", {k: round(sum(v) / len(v) / tot, 3) for k, v in dur.items()})
    b = json.loads((prof / f"{tag}_bench.json").read_text())
    km = b["roofline"]["kernel_ms"]
    print("CUDA-event shares:     ", {k: round(v / sum(km.values()), 3) for k, v in km.items()})
    print(json.dumps({k: {kk: v[kk] for kk in ("duration_ms", "fp64_pipe_pct", "issue_active_pct", "warps_active_pct",
                                                 "threads_per_inst", "fp64_flop_per_unit", "achieved_fp64_tflops")}
                      for k, v in summ.items()}, indent=1))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "r01")
