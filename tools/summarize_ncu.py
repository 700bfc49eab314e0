"""Summarise ncu outputs into small text files for profiles/.

    python tools/summarize_ncu.py launches <launches.csv> <out.md>      # per-kernel launch list (gpu__time_duration)
    python tools/summarize_ncu.py full <raw.csv from `ncu -i X.ncu-rep --page raw --csv`> <out.md>
"""
import collections
import csv
import re
import sys

FULL_KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__grid_size",
    "lts__t_bytes.sum", "sm__cycles_active.avg",
]


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*", "", name)
    return name.replace("hicom::", "")


def launches(src, dst):
    lines = [l for l in open(src) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1e3 if unit in ("ns", "nsecond") else v * 1e3 if unit in ("ms", "msecond") else v
        agg.setdefault((short(row["Kernel Name"]), row.get("Grid Size", "")), []).append(v)
    total = sum(sum(v) for v in agg.values())
    with open(dst, "w") as f:
        f.write(f"# ncu launch list ({src}) — gpu__time_duration.sum, --clock-control none, cold-cache serialised\n\n")
        f.write("| kernel | grid | launches | avg us | share |\n|---|---|---|---|---|\n")
        for (k, g), v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            f.write(f"| `{k[:80]}` | {g} | {len(v)} | {sum(v) / len(v):.1f} | {100 * sum(v) / total:.1f}% |\n")
        f.write(f"\ntotal {total / 1e3:.2f} ms over {sum(len(v) for v in agg.values())} launches\n")


def full(src, dst):
    rows = list(csv.reader(open(src)))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as f:
        f.write(f"# ncu --set full summary ({src})\n\n")
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            f.write(f"## `{short(d['Kernel Name'])[:90]}` grid {d.get('Grid Size', '')}\n\n")
            for k in FULL_KEYS:
                if k in d:
                    f.write(f"- {k}: {d[k]} {units[hdr.index(k)]}\n")
            f.write("\n")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
