#!/usr/bin/env python
"""Benchmark of the SPART forward hot path (BASELINE.json metric: simulations / second).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--n SAMPLES]

One "step" = one pass of the hot path over one batch of synthetic parameter sets of
BASELINE.json configs[1]: 1M-sample PROSPECT-5D + SAILH look-up table, Sentinel2A-MSI bands,
fixed geometry (sza 40, vza 0, raa 0), FP64.  With N > 1 (launched by torchrun, one rank per
GPU) every rank processes its own 1M-sample shard (weak scaling, no data-path collective);
the timed region is bracketed by a barrier + synchronize and the reported time is the max
over ranks.

Keys of the JSON line:
  value        whole-job simulations/s with parameters already resident in HBM
  e2e          the same metric through the public host-buffer API (run_batch_params on pinned
               host arrays): H2D of the parameters and D2H of the result inside the timed region
  roofline     the dominant kernel against the FP64 pipe (the path is FP64-arithmetic bound,
               not HBM bound: ~0.75 KB of HBM traffic vs ~1e5 FP64 flop per simulation);
               `peak` is a DFMA-chain micro-benchmark measured live in this run.
               roofline_hbm gives the same kernel against the measured HBM copy bandwidth.
  cpu_baseline the NumPy oracle port (oracle/spart_oracle.py) on all host cores, bounded sample
"""
import argparse
import json
import multiprocessing as mp
import os
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT / "spart-python_b200"))

SENSOR = "Sentinel2A-MSI"
CONFIG_ID = 2
N_DEFAULT = 1_000_000
WORKLOAD = "configs[1]: 1M-sample PROSPECT-5D+SAILH LUT, Sentinel2A-MSI (13 bands), fixed geometry, FP64"

# Algorithmic FP64 work per simulation of this workload, in flop (FMA = 2), counted from the
# executed SASS of each kernel (ncu smsp__sass_thread_inst_executed_op_{dadd,dmul,dfma}_pred_on,
# profiles/r01_*; see DESIGN.md "Roofline").  Updated whenever a kernel changes.
FLOP_PER_SAMPLE = {"lidf_kernel": 0.0, "geometry_kernel": 0.0, "band_kernel": 0.0}
# Algorithmic HBM bytes per simulation and kernel: leaf angles read 2 parameter rows and write
# 12 cumulative values; geometry reads those 12 + 12 parameter rows and writes the 32-double
# record; the band kernel reads the record + 12 parameter rows and writes 13 x 3 results.
BYTES_PER_SAMPLE = {"lidf_kernel": (2 + 12) * 8, "geometry_kernel": (12 + 12 + 32) * 8,
                    "band_kernel": (32 + 12) * 8 + 13 * 3 * 8}


def kernel_source_sha():
    """Identity of the build the executed-flop counts in profiles/flop_per_sample.json belong to: a
    hash of the SASS of the three band-path kernels (so that adding or editing unrelated kernels does
    not invalidate the counts); falls back to a hash of the CUDA sources without cuobjdump."""
    import hashlib
    import re
    import shutil
    import subprocess
    h = hashlib.sha256()
    lib = ROOT / "spart-python_b200" / "spart_b200" / "lib" / "libspart_b200.so"
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    try:
        txt = subprocess.run([cuobjdump, "-sass", str(lib)], capture_output=True, text=True, check=True).stdout
        keep = False
        for line in txt.splitlines():
            m = re.search(r"Function : (\S+)", line)
            if m:
                keep = bool(re.search(r"lidf_kernel|geometry_kernelPK|band_kernelILb1EEvPK", m.group(1)))
            if keep:
                h.update(re.sub(r"/\*[0-9a-f]{16}\*/", "", line).encode())     # drop the encoding column
        return "sass:" + h.hexdigest()[:16]
    except Exception:
        for f in sorted((ROOT / "spart-python_b200" / "csrc").glob("*")):
            if f.suffix in (".cu", ".cuh", ".h"):
                h.update(f.read_bytes())
        return "src:" + h.hexdigest()[:16]


def load_flop_counts():
    """Returns a note when the stored counts belong to another kernel build (roofline.achieved is
    then still computed, but flagged)."""
    p = ROOT / "profiles" / "flop_per_sample.json"
    if not p.exists():
        return "no profiles/flop_per_sample.json"
    d = json.loads(p.read_text())
    sha = d.pop("src_sha", None)
    FLOP_PER_SAMPLE.update(d)
    if sha != kernel_source_sha():
        return f"flop counts measured on kernel sources {sha}, current sources {kernel_source_sha()}"
    return None


# ----------------------------------------------------------------------------- helpers
def synthetic_params_torch(n, seed, device):
    """Synthetic config-2 parameter block [27, n] generated on the device (distributions of
    SURVEY.md section 8(d); same ranges as oracle.synthetic_params(config=2))."""
    import torch
    g = torch.Generator(device=device).manual_seed(seed)
    u = lambda lo, hi: torch.rand(n, generator=g, device=device, dtype=torch.float64) * (hi - lo) + lo
    P = torch.empty((27, n), dtype=torch.float64, device=device)
    P[0] = u(5, 80); P[1] = u(0.002, 0.02); P[2] = u(0.005, 0.05); P[3] = u(0, 0.5)
    P[4] = u(1, 25); P[5] = u(0, 10); P[6] = u(1, 3); P[7] = 0.0; P[8] = 0.0
    P[9] = u(0.2, 0.8); P[10] = u(0, 25); P[11] = u(90, 115)
    smp = u(5, 55)
    dry = u(0, 1) < 0.05
    P[12] = torch.where(dry, u(0, 5), smp); P[13] = 25.0; P[14] = 0.015
    P[15] = u(0.1, 8); P[16] = u(-0.5, 0.5); P[17] = u(-0.5, 0.5); P[18] = u(0.01, 0.2)
    P[19] = 40.0; P[20] = 0.0; P[21] = 0.0
    P[22] = u(0.05, 0.6); P[23] = u(0.25, 0.45); P[24] = u(0.5, 4); P[25] = u(900, 1030)
    P[26] = torch.randint(1, 366, (n,), generator=g, device=device).to(torch.float64)
    return P


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons with NVML during the timed region."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.samples, self.reasons, self.max_mhz = index, False, [], set(), None

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {
                nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
            }
            while not self.stop_flag:
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
                time.sleep(0.002)
        except Exception as e:  # NVML missing: report that instead of failing the bench
            self.reasons.add(f"nvml_unavailable:{type(e).__name__}")

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


# ------------------------------------------------------------------- CPU baseline (oracle)
def _oracle_worker(args):
    n, seed = args
    sys.path.insert(0, str(ROOT / "oracle"))
    import spart_oracle as so
    P = so.synthetic_params(n, CONFIG_ID, seed=seed)
    out = so.spart_bands(P, SENSOR)
    return float(out[0, 0, 0])


def time_oracle(total, cores, chunk=512, pool=None):
    """Times the NumPy oracle over `total` samples split into `chunk`-sample tasks on `cores`
    worker processes; returns (simulations/s, samples actually run)."""
    tasks = [(chunk, 1000 + i) for i in range(max(1, total // chunk))]
    own = pool is None
    if own:
        pool = mp.get_context("fork").Pool(cores)
        pool.map(_oracle_worker, [(8, 1)] * cores)           # import + table load outside timing
    t0 = time.perf_counter()
    pool.map(_oracle_worker, tasks, chunksize=1)
    dt = time.perf_counter() - t0
    if own:
        pool.close()
    return len(tasks) * chunk / dt, len(tasks) * chunk


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = host_cores()
    per_step = 512 * cores * 2                      # bounded sample per step (a few seconds)
    pool = mp.get_context("fork").Pool(cores)
    pool.map(_oracle_worker, [(8, 1)] * cores)      # imports and table loads happen before timing
    for _ in range(args.warmup):
        time_oracle(512 * cores, cores, pool=pool)
    t0 = time.perf_counter()
    done = 0
    for _ in range(args.steps):
        _, n = time_oracle(per_step, cores, pool=pool)
        done += n
    dt = time.perf_counter() - t0
    pool.close()
    value = done / dt
    line = {
        "impl": "reference", "metric": "SPART simulations/sec (full RT + SRF)", "value": value,
        "unit": "simulations/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "samples_per_step": per_step},
        "cpu_baseline": {"value": value, "unit": "simulations/s", "cores": cores, "kind": "port",
                         "sample": f"{per_step} samples/step of the same synthetic distribution, NumPy oracle port "
                                   f"(oracle/spart_oracle.py) in {cores} processes; the unmodified reference cannot "
                                   "travel to the GPU box (measured 2.19 simulations/s/core in the build container)"},
        "e2e": {"value": value, "unit": "simulations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------ our arm
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import spart_b200

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # keep this rank (and the pinned host buffers it first-touches) on the CPUs next to its GPU;
    # the original affinity is restored before the CPU baseline leg
    full_affinity = os.sched_getaffinity(0) if hasattr(os, "sched_getaffinity") else None
    try:
        import pynvml
        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local))
    except Exception:
        pass
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    n = args.n
    eng = spart_b200.default_engine(dev)
    _, st = eng.sensor(SENSOR)
    nb = st.n_bands
    params = synthetic_params_torch(n, 20261018 + CONFIG_ID + 7919 * rank, dev)
    out = torch.empty((n, nb, 3), dtype=torch.float64, device=dev)
    flop_note = load_flop_counts()

    # --- resident-input timing -------------------------------------------------------
    for _ in range(args.warmup):
        eng.forward_bands(params, SENSOR, out=out, uniform_geometry=True)
    barrier()
    eng.profile_enable(SENSOR, True)
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = eng.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        eng.forward_bands(params, SENSOR, out=out, uniform_geometry=True)
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    launches = eng.launch_count() - launches0
    sampler.stop_flag = True
    sampler.join(2)
    prof = eng.profile_read(SENSOR)
    eng.profile_enable(SENSOR, False)
    value = world * n * args.steps / (ms * 1e-3)

    # --- FP32 mode on the same batch (reported beside the FP64 headline) ------------------
    out32 = torch.empty_like(out)
    for _ in range(args.warmup):
        eng.forward_bands(params, SENSOR, out=out32, uniform_geometry=True, precision="fp32")
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(args.steps):
        eng.forward_bands(params, SENSOR, out=out32, uniform_geometry=True, precision="fp32")
    f1.record()
    barrier()
    ms32 = max_over_ranks(f0.elapsed_time(f1))
    err32 = float(((out32 - out).abs() / out.abs()).max().item())

    # --- final gather of the per-rank results (the only collective of the path) -----------
    gather = None
    if world > 1:
        from spart_b200.distributed import gather_results
        for _ in range(2):
            gather_results(out, world * n, dst=0)
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        full = gather_results(out, world * n, dst=0)
        g1.record()
        barrier()
        gms = max_over_ranks(g0.elapsed_time(g1))
        gather = {"ms": gms, "bytes_to_root": (world - 1) * n * nb * 3 * 8,
                  "gb_per_s": (world - 1) * n * nb * 3 * 8 / (gms * 1e-3) / 1e9, "op": "nccl gather to rank 0"}
        del full

    # --- end to end through the public host-buffer API ---------------------------------
    host_in = torch.empty((27, n), dtype=torch.float64).pin_memory()
    host_in.copy_(params)
    host_out = torch.empty((n, nb, 3), dtype=torch.float64).pin_memory()
    e2e_steps = max(1, min(args.steps, 5))
    for _ in range(2):
        spart_b200.run_batch_params(host_in, SENSOR, out=host_out, uniform_geometry=True)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        spart_b200.run_batch_params(host_in, SENSOR, out=host_out, uniform_geometry=True)
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_value = world * n * e2e_steps / e2e_s
    assert torch.equal(host_out.to(dev), out), "host-buffer path disagrees with the device path"

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # --- roofline of the dominant kernel ----------------------------------------------
    peaks = eng.measure_peaks()
    calls = max(prof["calls"], 1)
    kern_ms = {"lidf_kernel": prof["lidf_ms"] / calls, "geometry_kernel": prof["geometry_ms"] / calls,
               "band_kernel": prof["band_ms"] / calls}
    dominant = max(kern_ms, key=kern_ms.get)
    mp_file = ROOT / "MEASURED_PEAKS.json"
    hbm_peak, hbm_src = 6650.0, "fallback (B200_PROFILING.md)"
    if mp_file.exists():
        hbm_peak, hbm_src = float(json.loads(mp_file.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    flop = FLOP_PER_SAMPLE[dominant] * n
    ach_tf = flop / (kern_ms[dominant] * 1e-3) / 1e12 if flop else None
    roofline = {
        "kernel": dominant, "bound": "fp64", "achieved": ach_tf, "peak": peaks["fp64_tflops"], "unit": "TFLOP/s",
        "frac": (ach_tf / peaks["fp64_tflops"]) if ach_tf else None, "traffic": None,
        "peak_source": "DFMA-chain micro-benchmark measured live in this run (spart_measure_peaks)",
        "kernel_ms": kern_ms, "share_of_step": kern_ms[dominant] / (ms / args.steps),
        "flop_per_simulation": FLOP_PER_SAMPLE, "flop_count_note": flop_note,
        "kernel_build": kernel_source_sha(),
    }
    nbytes = BYTES_PER_SAMPLE[dominant] * n
    ach_gb = nbytes / (kern_ms[dominant] * 1e-3) / 1e9
    roofline_hbm = {"kernel": dominant, "bound": "hbm", "achieved": ach_gb, "peak": hbm_peak, "unit": "GB/s",
                    "frac": ach_gb / hbm_peak, "traffic": None, "peak_source": hbm_src}
    tr = ROOT / "profiles" / "dram_traffic.json"
    if tr.exists():
        t = json.loads(tr.read_text())
        roofline["traffic"] = roofline_hbm["traffic"] = t.get(dominant)

    # --- CPU baseline: the oracle port on all host cores, bounded sample -----------------
    if full_affinity is not None:
        os.sched_setaffinity(0, full_affinity)
    cores = host_cores()
    cpu = None
    if not args.no_cpu:
        rate, ran = time_oracle(512 * cores * 16, cores)     # ~25 core-seconds of oracle work
        cpu = {"value": rate, "unit": "simulations/s", "cores": cores, "kind": "port",
               "sample": f"{ran} samples of the same synthetic distribution, NumPy oracle port "
                         f"(oracle/spart_oracle.py), {cores} processes"}

    line = {
        "metric": "SPART simulations/sec (full RT + SRF)", "value": value, "unit": "simulations/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "samples_per_gpu_per_step": n, "sensor": SENSOR, "bands": nb,
                   "l2": "working set per step (216 MB params + 352 MB workspace + 312 MB output) exceeds the 126 MB L2"},
        "e2e": {"value": e2e_value, "unit": "simulations/s", "h2d_bytes_per_step": 27 * 8 * n,
                "d2h_bytes_per_step": nb * 3 * 8 * n, "steps": e2e_steps,
                "api": "spart_b200.run_batch_params(pinned host [27,n], uniform_geometry=True) -> pinned host [n,13,3]"},
        "gpu_launches": launches,
        "clocks": sampler.summary(),
        "roofline": roofline, "roofline_hbm": roofline_hbm,
        "cpu_baseline": cpu,
        "peaks": peaks,
        "fp32_mode": {"value": world * n * args.steps / (ms32 * 1e-3), "unit": "simulations/s",
                      "ms_per_step": ms32 / args.steps, "max_rel_err_vs_fp64": err32},
        "gather": gather,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=N_DEFAULT, help="samples per GPU per step")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
