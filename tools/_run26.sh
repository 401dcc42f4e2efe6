for t in 4 8 12 16; do echo -n "threads $t: "; SPART_HOST_THREADS=$t python - <<'PY'
import os, sys, time
sys.path.insert(0,'spart-python_b200'); sys.path.insert(0,'.')
import torch, numpy as np, bench, spart_b200 as sb
dev=torch.device('cuda',0); n=1_000_000; cfg=bench.CONFIGS[2]
P=bench.synthetic_params_torch(n,2,1,dev)
hin=np.ascontiguousarray(P.cpu().numpy()); hout=np.empty((n,13,3))
call=lambda: sb.run_batch_params(hin,"Sentinel2A-MSI",out=hout,broadcast_rows=cfg["bcast"])
for _ in range(2): call()
t0=time.perf_counter()
for _ in range(5): call()
dt=(time.perf_counter()-t0)/5
print(round(dt*1e3,2),'ms', round(n/dt/1e6,1),'M/s')
PY
done
