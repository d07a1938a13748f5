"""ncu launch list (--metrics gpu__time_duration.sum --csv) → markdown table of kernels by share of the step.

    python profiles/summarize_launches.py gpurun_out/launches.csv "title" > profiles/rXX_launches_summary.md
"""
import collections
import csv
import sys

path, title = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
rows = [r for r in csv.reader(open(path)) if len(r) > 5]
hdr = rows[0]
i_name, i_val, i_unit = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    if r[hdr.index("Metric Name")] != "gpu__time_duration.sum":
        continue
    v = float(r[i_val].replace(",", ""))
    u = r[i_unit]
    us = v / 1e3 if u in ("ns", "nsecond") else (v * 1e3 if u in ("ms", "msecond") else v)
    k = r[i_name].split("(")[0][:64]
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += us
tot = sum(v[1] for v in agg.values())
n = sum(v[0] for v in agg.values())
print(f"# {title}\n")
print(f"Total kernel time in the step: {tot / 1e3:.2f} ms over {n} launches (cold-cache, serialised under ncu: compare SHARES with bench.py's "
      f"event-timed table, not absolutes).\n")
print("| kernel | launches | us | share |\n|---|---:|---:|---:|")
for k, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{k}` | {c} | {us:.1f} | {100 * us / tot:.1f}% |")
