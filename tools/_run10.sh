python tools/kbench.py > gpurun_out/r02_kbench2.log 2>&1
for v in a2_mb6 a2_mb5 a0_mb7 a2_mb7; do SPART_B200_LIB=build/alt/lib_$v.so python tools/kbench.py >> gpurun_out/r02_kbench2.log 2>&1; done
grep '^{' gpurun_out/r02_kbench2.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['lib'], round(d['lidf_ms'],4), round(d['geometry_ms'],4), round(d['band_ms'],4), d['checksum'])
"
