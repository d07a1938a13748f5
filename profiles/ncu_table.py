"""Per-kernel table of an `ncu --set full` report (averages over the launches of each kernel) -> markdown on stdout, and - with
--json PATH - the per-launch counters bench.py attaches to its roofline objects (profiles/ncu_metrics.json).

    python profiles/ncu_table.py gpurun_out/r2_step_full.ncu-rep "title" --json profiles/ncu_metrics.json > profiles/r2_ncu_full_summary.md
(the first argument may also be the raw-page CSV of a report: `ncu -i REP --page raw --csv > x.csv`)
"""
import collections
import csv
import io
import json
import subprocess
import sys

rep, title = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 and not sys.argv[2].startswith("--") else "")
# a .ncu-rep, or the CSV that `ncu -i REP --page raw --csv` printed on the GPU box (reports of a whole step exceed what comes back)
raw = open(rep).read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
M = {"us": "gpu__time_duration.sum", "rd": "dram__bytes_read.sum", "wr": "dram__bytes_write.sum",
     "dram": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "tens": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
     "issue": "smsp__issue_active.avg.pct_of_peak_sustained_active", "regs": "launch__registers_per_thread", "grid": "launch__grid_size",
     "block": "launch__block_size", "xu": "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
     "l2": "lts__throughput.avg.pct_of_peak_sustained_elapsed"}


def val(r, key):
    i = col.get(M[key])
    if i is None or not r[i]:
        return 0.0
    v = float(r[i].replace(",", ""))
    u = units[i]
    if key == "us":
        v = v / 1e3 if u in ("ns", "nsecond") else (v * 1e3 if u in ("ms", "msecond") else v)
    if key in ("rd", "wr"):
        v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    return v


agg = collections.OrderedDict()
for r in rows[2:]:
    name = r[col["Kernel Name"]].split("(")[0].replace("void ", "").replace("umpr::", "")
    a = agg.setdefault(name, collections.defaultdict(float))
    a["n"] += 1
    for k in M:
        a[k] += val(r, k)
    st = {h.split("issue_stalled_")[1].split("_per")[0]: float(r[i]) for i, h in enumerate(hdr)
          if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("per_issue_active.ratio") and r[i]}
    for k, v in st.items():
        a["stall_" + k] += v
print(f"# {title}\n")
print("`ncu --set full --clock-control none --import-source on` of ONE train step (profiles/prof_step.py: music_full, batch 1024, the native one-call step); "
      "averages over the launches of each kernel inside that step.  Times are cold-cache and serialised under the profiler - "
      "bench.py's CUDA-event table is the timing reference; these counters explain it.\n")
print("| kernel | launches | us / launch | DRAM read+write MB / launch | DRAM % of peak | tensor pipe active % | issue active % | XU % | regs | top stalls (cycles per issue) |")
print("|---|---:|---:|---:|---:|---:|---:|---:|---:|---|")
out = {}
for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
    n = a["n"]
    stalls = sorted(((k[6:], v / n) for k, v in a.items() if k.startswith("stall_")), key=lambda kv: -kv[1])[:3]
    print(f"| `{name}` | {int(n)} | {a['us'] / n:.1f} | {(a['rd'] + a['wr']) / n / 1e6:.1f} | {a['dram'] / n:.1f} | {a['tens'] / n:.1f} | "
          f"{a['issue'] / n:.1f} | {a['xu'] / n:.1f} | {int(a['regs'] / n)} | " + ", ".join(f"{k} {v:.1f}" for k, v in stalls) + " |")
    entry = "umpr_" + name.split("<")[0].replace("_kernel", "")
    entry = {"umpr_coattn_affinity_tc2": "umpr_coattn_fwd_tc"}.get(entry, entry)
    out[entry] = {"dram_bytes_per_launch": int((a["rd"] + a["wr"]) / n), "tensor_pipe_active_pct": round(a["tens"] / n, 2),
                  "dram_pct": round(a["dram"] / n, 2), "ncu_us_per_launch": round(a["us"] / n, 1), "source": rep.split("/")[-1]}
if "--json" in sys.argv:
    path = sys.argv[sys.argv.index("--json") + 1]
    json.dump(out, open(path, "w"), indent=1)
