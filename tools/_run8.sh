set -x
python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/r02_t8.log
python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err
cat gpurun_out/r02_t8.log; tail -3 gpurun_out/r02b_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02b_bench.json').read().strip().splitlines()[-1])
print('value',d['value'],'fp32',d['fp32_mode']['value'],d['fp32_mode']['kernel_ms'])
print('lut',json.dumps(d['lut_retrieval']))
PY
