set -x
python -m pytest tests -m gpu -q -k "leafangles or threshold or golden" 2>&1 | tail -4 > gpurun_out/r02_t6.log
python tools/kbench.py > gpurun_out/r02_kbench.log 2>&1
for v in a2_0 a2_as1 a2_b4 a2_b10; do SPART_B200_LIB=build/alt/lib_$v.so python tools/kbench.py >> gpurun_out/r02_kbench.log 2>&1; done
python tools/split_test.py > gpurun_out/r02_split.log 2>&1
python tools/lidf_parity_scale.py 1000000 > gpurun_out/r02_lidf_parity.json 2> gpurun_out/r02_lidf_parity.err
cat gpurun_out/r02_t6.log gpurun_out/r02_kbench.log gpurun_out/r02_split.log gpurun_out/r02_lidf_parity.json
