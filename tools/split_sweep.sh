#!/bin/bash
# sweep HICOM_SM_SPLIT over the headline workloads (graph replay, no extras)
for wl in $WLS; do
  for L in $SPLITS; do
    HICOM_SM_SPLIT=$L python bench.py --workload $wl --steps 10 --no-cpu-baseline --no-parity --no-sustained --no-c4 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$wl split=$L', round(d['value']), round(d['ms_per_step'],4), round(d['executed_tflops']))
"
  done
done
