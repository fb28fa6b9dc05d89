"""Turns the raw ncu outputs brought back in gpurun_out/ into the small, committed summaries of this directory.

usage: python profiles/summarise.py <launches.csv> <prof.ncu-rep> <tag> <kernel-substring> [step-kernel-substrings,comma-separated]

Writes profiles/ncu_launches_<tag>.md (per-kernel launch counts / mean duration / share of the step kernels from the
`--metrics gpu__time_duration.sum --clock-control none` pass, cold cache and serialised) and profiles/ncu_<kernel>_<tag>.json
(selected metrics of the `--set full` capture of the dominant kernel, incl. DRAM bytes per launch)."""
import collections
import csv
import json
import subprocess
import sys

launch_csv, rep, tag, kern = sys.argv[1:5]
step_subs = sys.argv[5].split(",") if len(sys.argv) > 5 else ["bv_pass_kernel", "reduce_kernel", "bv_tail_kernel"]
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
ours = {k: v for k, v in agg.items() if k.startswith("jx::") or k.startswith("bg::")}


def steady(k, v):
    # the first pass / adam / gemm2 launches belong to the forward-only initialisation
    return v[1:] if len(v) > 1 and any(s in k for s in ("bv_pass", "bt_adam", "bg_gemm_kernel<2>", "bt_reduce2", "bt_tail", "bt_kl")) else v


step_kernels = {k: v for k, v in ours.items() if any(s in k for s in step_subs)}
per_step = {k: sum(steady(k, v)) / len(steady(k, v)) for k, v in step_kernels.items()}
tot = sum(per_step.values())
lines = [f"# ncu launch list ({tag}) -- gpu__time_duration.sum per launch, --clock-control none (cold cache, serialised)", "",
         "| kernel | launches | mean us | share of one step |", "|---|---|---|---|"]
for k, v in ours.items():
    vv = steady(k, v)
    share = f"{100 * per_step[k] / tot:.1f} %" if k in step_kernels else "-"
    lines.append(f"| `{k}` | {len(v)} | {sum(vv) / len(vv):.1f} | {share} |")
lines += ["", f"sum of the step kernels' mean durations: {tot:.1f} us"]

raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
h = rr[0]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__grid_size", "launch__block_size",
        "sm__inst_executed_pipe_tensor.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__average_warp_latency_per_inst_issued.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"]
names = [r[h.index("Kernel Name")] for r in rr[2:]] if "Kernel Name" in h else []
summ = {"kernels": names}
for w in [c for c in h if any(c == x or (x in c and "tensor" in x) for x in want)]:
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
json.dump(summ, open(f"profiles/ncu_{kern}_{tag}.json", "w"), indent=1)
lines += ["", f"# ncu --set full, {kern} ({tag})", ""]
for k, v in summ.items():
    if k == "kernels":
        lines.append(f"- kernels: {sorted(set(n.split('(')[0] for n in v))}")
    elif isinstance(v, dict):
        lines.append(f"- `{k}`: {v['values']} {v['unit']}")
    else:
        lines.append(f"- `{k}`: {v}")
open(f"profiles/ncu_launches_{tag}.md", "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
