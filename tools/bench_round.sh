#!/bin/bash
# Bench lines of the round (no profiler): ours (default), the reference arm, the three guide modes, the other workloads.
set -u
O=gpurun_out
timeout 600 python bench.py 2>/dev/null | tail -1 > $O/bench_final.json
timeout 600 python bench.py --impl reference 2>/dev/null | tail -1 > $O/bench_reference_arm.json
: > $O/bench_guide_modes.json
for g in coarse direct none; do timeout 600 python bench.py --use-guide $g 2>/dev/null | tail -1 >> $O/bench_guide_modes.json; done
: > $O/bench_workloads.json
for w in c3 c5 c4; do timeout 600 python bench.py --workload $w --no-cpu-baseline 2>/dev/null | tail -1 >> $O/bench_workloads.json; done
timeout 300 python tools/latency_b1.py > $O/latency_b1.txt 2>&1
# rows f2 / f3: producer of frames_embed and the training step (tests + CUDA-event timings, one process each)
timeout 300 python tools/producer_round.py > $O/producer_round.log 2>&1
timeout 300 python tools/autograd_round.py > $O/autograd_round.log 2>&1
