#!/bin/bash
# One gpurun call that regenerates the round's evidence under gpurun_out/ (copy the summaries into profiles/ afterwards):
#   bench lines (ours, reference arm, three guide modes), the ncu launch list and one `ncu --set full` capture of the
#   three dominant kernels.  Numbers printed under ncu are never used as bench values.
set -u
O=gpurun_out
timeout 600 python bench.py 2>/dev/null | tail -1 > $O/bench_final.json
timeout 600 python bench.py --impl reference 2>/dev/null | tail -1 > $O/bench_reference_arm.json
: > $O/bench_guide_modes.json
for g in coarse direct none; do timeout 600 python bench.py --use-guide $g 2>/dev/null | tail -1 >> $O/bench_guide_modes.json; done
: > $O/bench_workloads.json
for w in c3 c5 c4; do timeout 600 python bench.py --workload $w --no-cpu-baseline 2>/dev/null | tail -1 >> $O/bench_workloads.json; done
timeout 300 python tools/latency_b1.py > $O/latency_b1.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_final.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-graph > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
  -k "regex:local_attend_v2|tc_gemm_kernel<.int.288|tc_gemm_kernel<.int.144" -s 0 -c 5 -o $O/prof_final \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-graph > $O/ncu_final.log 2>&1
ls -la $O/*.ncu-rep
