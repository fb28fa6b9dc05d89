"""Turns the raw ncu outputs brought back in gpurun_out/ into the small, committed summaries of this directory.
usage: python profiles/summarise.py <launches.csv> <prof.ncu-rep> <tag>"""
import collections
import csv
import json
import subprocess
import sys

launch_csv, rep, tag = sys.argv[1:4]
rows = [r for r in csv.reader(open(launch_csv)) if len(r) > 10]
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[1:]:
    name = r[ki].split("(")[0].replace("void ", "")
    try:
        agg.setdefault(name, []).append(float(r[vi].replace(",", "")) / 1e3)
    except ValueError:
        pass
ours = {k: v for k, v in agg.items() if k.startswith("jx::")}
step_kernels = {k: v for k, v in ours.items() if any(s in k for s in ("bv_pass_kernel", "reduce_kernel", "bv_tail_kernel"))}
tot = sum(sum(v[1:]) if "bv_pass" in k else sum(v) for k, v in step_kernels.items())
lines = [f"# ncu launch list ({tag}) -- gpu__time_duration.sum per launch, --clock-control none (cold cache, serialised)", "",
         "| kernel | launches | mean us | share of step kernels |", "|---|---|---|---|"]
for k, v in ours.items():
    vv = v[1:] if "bv_pass" in k and len(v) > 1 else v  # first pass launch is the forward-only init pass
    share = f"{100 * sum(vv) / tot:.1f} %" if k in step_kernels else "-"
    lines.append(f"| `{k}` | {len(v)} | {sum(vv) / len(vv):.1f} | {share} |")

raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
h = rr[0]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__grid_size", "launch__block_size",
        "smsp__average_warp_latency_per_inst_issued.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"]
summ = {}
for w in want:
    if w in h:
        i = h.index(w)
        summ[w] = {"unit": rr[1][i], "values": [r[i] for r in rr[2:]]}


def num(x):
    return float(x.replace(",", ""))


def to_bytes(key):
    unit = summ[key]["unit"].lower()
    scale = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[unit]
    vals = [num(v) * scale for v in summ[key]["values"]]
    return sum(vals) / len(vals)


traffic = to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum")
summ["dram_bytes_per_launch"] = traffic
json.dump(summ, open(f"profiles/ncu_pass_{tag}.json", "w"), indent=1)
lines += ["", f"# ncu --set full, bv_pass_kernel<1> ({tag})", ""]
for k, v in summ.items():
    lines.append(f"- `{k}`: {v if not isinstance(v, dict) else (v['values'], v['unit'])}")
open(f"profiles/ncu_launches_{tag}.md", "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
