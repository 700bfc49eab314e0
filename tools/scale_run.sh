#!/bin/bash
# usage: tools/scale_run.sh "<gpu counts>" ; runs both bench arms for each N (driver-style launch), prints compact lines
python -m hicom_b200.build 2>/dev/null
for N in $1; do
  for wl in ${WL:-c2 c4}; do
    if [ "$N" = "1" ]; then
      timeout 300 python bench.py --gpus 1 --steps 50 --warmup 5 --workload $wl --no-cpu-baseline 2>gpurun_out/scale_${wl}_$N.err | tail -1 > gpurun_out/scale_${wl}_$N.json
    else
      timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 50 --warmup 5 --workload $wl --no-cpu-baseline 2>gpurun_out/scale_${wl}_$N.err | tail -1 > gpurun_out/scale_${wl}_$N.json
    fi
    python - <<PY
import json
try:
    d=json.load(open("gpurun_out/scale_${wl}_$N.json")); print("$wl N=$N", round(d["value"]), "frames/s", round(d["ms_per_step"],3), "ms", d["config"]["launch"], d["clocks"]["reasons"])
except Exception as e:
    print("$wl N=$N FAILED", e); print(open("gpurun_out/scale_${wl}_$N.err").read()[-600:])
PY
  done
done
