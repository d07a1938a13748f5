"""Summarise an .ncu-rep here (no GPU): key raw metrics per launch, and the stall-sample distribution of the source page.

    python profiles/ncu_summary.py gpurun_out/prof.ncu-rep [--top 25] [--ranges name:lo:hi,...]
"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 25
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__cycles_elapsed.avg", "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print("==", r[hdr.index("Kernel Name")][:100])
    for k in KEYS:
        if k in hdr:
            print(f"   {k:80s} {r[hdr.index(k)]:>16s} {units[hdr.index(k)]}")
    st = {h.split("issue_stalled_")[1].split("_per")[0]: float(r[i]) for i, h in enumerate(hdr)
          if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("per_issue_active.ratio") and r[i]}
    print("   stalls per issue:", ", ".join(f"{k}={v:.2f}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:7]))
if "--nosrc" in sys.argv:
    sys.exit(0)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
i_src, i_s, i_ex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data = []
for r in rows[2:]:
    if len(r) < len(hdr) or r[0] in ("Address", "Kernel Name"):
        break
    data.append(r)
tot = sum(int(r[i_s]) for r in data)
print(f"-- source page of the first launch: {len(data)} SASS instructions, {tot} samples")
c = collections.Counter()
for r in data:
    for k in stall_cols:
        c[hdr[k]] += int(r[k])
print("   by reason:", ", ".join(f"{k}={v}" for k, v in c.most_common(8)))
if "--ranges" in sys.argv:
    for spec in sys.argv[sys.argv.index("--ranges") + 1].split(","):
        name, lo, hi = spec.split(":")
        sub = data[int(lo):int(hi)]
        cc = collections.Counter()
        for r in sub:
            for k in stall_cols:
                cc[hdr[k]] += int(r[k])
        print(f"   [{name}] samples {sum(int(r[i_s]) for r in sub)} inst {sum(int(r[i_ex]) for r in sub)}:", ", ".join(f"{k}={v}" for k, v in cc.most_common(5)))
for i in sorted(sorted(range(len(data)), key=lambda i: -int(data[i][i_s]))[:top]):
    r = data[i]
    st = sorted(((hdr[k], int(r[k])) for k in stall_cols if int(r[k]) > 0), key=lambda kv: -kv[1])[:2]
    print(f"   {i:5d} {int(r[i_s]):6d} {100 * int(r[i_s]) / max(tot, 1):5.1f}% ex={r[i_ex]:>9s} {r[i_src].strip()[:64]:64s} {st}")
