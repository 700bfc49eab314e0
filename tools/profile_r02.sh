#!/bin/bash
# Round-2 evidence in one gpurun call (outputs under gpurun_out/; summaries are copied into profiles/ afterwards).
# Numbers printed under ncu are never used as bench values.
set -u
O=gpurun_out
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_launches_c2.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-graph --no-c4 --no-sustained --no-parity --no-traffic > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
  -k "regex:local_attend_v2|tc_gemm_kernel<.int.288|tc_gemm_kernel<.int.144" -s 0 -c 3 -o $O/r02_prof_top3 \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-graph --no-c4 --no-sustained --no-parity --no-traffic > $O/r02_ncu_top3.log 2>&1
ncu -i $O/r02_prof_top3.ncu-rep --page raw --csv > $O/r02_prof_top3_raw.csv 2>/dev/null
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file $O/r02_launches_train.csv python tools/train_profile.py coarse > /dev/null 2>&1
timeout 300 python tools/autograd_round.py --skip-tests > $O/r02_autograd_bench.txt 2>&1
ls -la $O/*.ncu-rep | tail -3
