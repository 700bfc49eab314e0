"""Print value, ms/step and the largest kernels of a bench.py JSON line read from stdin:  python bench.py ... | python tools/show_bench.py TAG [kernel-substring]"""
import json, sys
tag = sys.argv[1]; pat = sys.argv[2] if len(sys.argv) > 2 else ""
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print(tag, round(d["value"]), round(d["ms_per_step"], 3))
items = [(k, v) for k, v in d["kernels"].items() if pat in k] if pat else list(d["kernels"].items())[:8]
for k, v in items:
    print("   ", k, round(v["ms_per_step"] * 1e3 / max(1, 1), 1), "us/step", v["launches"])
