python -m pytest tests -m gpu -q -x 2>&1 | tail -4
python tools/lidf_parity_scale.py 1000000 > gpurun_out/r02_lidf_parity_v2.json 2>/dev/null; cat gpurun_out/r02_lidf_parity_v2.json
