set -x
python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r02_t3.log
python tools/e2e_sweep.py > gpurun_out/r02_e2e_sweep.log 2>&1
SPART_HOST_THREADS=16 python tools/e2e_sweep.py 2>&1 | grep pageable > gpurun_out/r02_e2e_sweep_t16.log
python - > gpurun_out/r02_srf_time.log 2>&1 <<'PY'
import sys, time, torch
sys.path.insert(0, "spart-python_b200"); sys.path.insert(0, ".")
import bench, spart_b200 as sb
dev = torch.device("cuda", 0)
eng = sb.default_engine(dev)
for sensor, cfg in (("TerraAqua-MODIS", 3), ("LANDSAT8-OLI", 3), ("Sentinel2A-MSI", 2)):
    n = 200_000
    P = bench.synthetic_params_torch(n, cfg, 5, dev)
    f = lambda: eng.forward_bands(P, sensor, band_mode="srf")
    for _ in range(2): f()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(sensor, "srf mode", n, "samples", round(ms, 2), "ms ->", round(n / ms / 1e3, 2), "M simulations/s")
PY
cat gpurun_out/r02_t3.log gpurun_out/r02_e2e_sweep.log gpurun_out/r02_e2e_sweep_t16.log gpurun_out/r02_srf_time.log | grep -v "^{" 
