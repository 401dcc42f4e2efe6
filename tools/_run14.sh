bash tools/kbench_all.sh
python tools/lidf_parity_scale.py 1000000 > gpurun_out/r02_lidf_parity_v3.json 2> gpurun_out/r02_lidf_parity_v3.err; cat gpurun_out/r02_lidf_parity_v3.json
SPART_B200_LIB=build/alt/lib_r20.so python tools/lidf_parity_scale.py 1000000 2>/dev/null
python -m pytest tests -m gpu -q -x -k "leafangles or lidf or large_batch or golden or edge or grids" 2>&1 | tail -4
