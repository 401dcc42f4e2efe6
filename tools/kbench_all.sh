#!/bin/bash
# kbench over the default library and every build/alt variant; prints one line per library
out=gpurun_out/kbench_all.log
python tools/kbench.py "$@" > $out 2>&1
for f in build/alt/lib_*.so; do SPART_B200_LIB=$f python tools/kbench.py "$@" >> $out 2>&1; done
grep '^{' $out | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('%-28s lidf %.4f geo %.4f band %.4f sum %.12g'%(d['lib'], d['lidf_ms'], d['geometry_ms'], d['band_ms'], d['checksum']))
"
grep -i "error\|Traceback" $out | head -5
