timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -x -q 2>&1 | tail -3 > gpurun_out/t.txt
for w in c2; do timeout 300 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$w', round(d['value']), round(d['ms_per_step'],3)); [print('   ',k,v) for k,v in list(d['kernels'].items())[:6]]" >> gpurun_out/t.txt 2>&1; done
timeout 300 python tools/microbench_linear.py 2>&1 | grep "10368" >> gpurun_out/t.txt
