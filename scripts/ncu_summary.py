#!/usr/bin/env python3
"""Summarise gpurun_out/launches.csv and gpurun_out/trace_full.ncu-rep into profiles/ (tracked)."""
import collections, csv, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
out = os.path.join(ROOT, "profiles")
os.makedirs(out, exist_ok=True)
go = os.path.join(ROOT, "gpurun_out")

rows = [r for r in csv.reader(open(os.path.join(go, "launches.csv"))) if len(r) > 10]
hdr = rows[0]
ki, vi, ii = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("ID")
agg = collections.OrderedDict()
per_launch = []
for r in rows[1:]:
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    k = r[ki]
    per_launch.append((int(r[ii]), k[:90], v))
    a = agg.setdefault(k[:90], [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
with open(os.path.join(out, f"{tag}_launches_summary.txt"), "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)\n")
    f.write("# command: python bench.py --rays $RAYS --steps 2 --warmup 1 --no-e2e --no-cpu   (scripts/profile.sh)\n")
    f.write(f"# total {tot/1e6:.3f} ms over {len(per_launch)} launches\n")
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        f.write(f"{a[1]/1e6:10.3f} ms {a[0]:5d}x {100*a[1]/tot:5.1f}%  {k}\n")
with open(os.path.join(out, f"{tag}_launches.csv"), "w") as f:
    f.write("id,kernel,ns\n")
    for i, k, v in per_launch:
        f.write(f'{i},"{k}",{v:.0f}\n')

rep = os.path.join(go, "trace_full.ncu-rep")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
r = list(csv.reader(io.StringIO(raw)))
h, units, vals = r[0], r[1], r[2]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "l1tex__t_bytes.sum", "smsp__warps_eligible.avg.per_cycle_active", "local_load", "smsp__inst_executed_op_local",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warp_latency_per_inst_issued.ratio", "smsp__average_warps_issue_stalled"]
d = {}
with open(os.path.join(out, f"{tag}_trace_ncu_full.txt"), "w") as f:
    f.write("# ncu --set full --clock-control none --import-source on -k regex:trace_kernel (scripts/profile.sh), one bench launch\n")
    for i, name in enumerate(h):
        if any(name.startswith(w) for w in want):
            f.write(f"{name:80s} {units[i]:>16s} {vals[i]}\n")
            d[name] = (units[i], vals[i])
def num(k):
    u, v = d[k]
    x = float(v.replace(",", ""))
    return x * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "Tbyte": 1e12}.get(u, 1)
traffic = num("dram__bytes_read.sum") + num("dram__bytes_write.sum")
json.dump({"dram_bytes_per_launch": traffic, "dram_bytes_read": num("dram__bytes_read.sum"), "dram_bytes_write": num("dram__bytes_write.sum"),
           "source": f"profiles/{tag}_trace_ncu_full.txt", "kernel_ms_under_ncu": d["gpu__time_duration.sum"][1]},
          open(os.path.join(out, "ncu_traffic.json"), "w"), indent=1)
det = subprocess.run(["ncu", "-i", rep, "--page", "details"], stdout=subprocess.PIPE, text=True).stdout
open(os.path.join(out, f"{tag}_trace_ncu_details.txt"), "w").write(det)
print(open(os.path.join(out, f"{tag}_trace_ncu_full.txt")).read())
print("traffic per launch: %.2f GB" % (traffic / 1e9))
