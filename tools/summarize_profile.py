"""Turns gpurun_out ncu artefacts into small tracked summaries under profiles/.

  python tools/summarize_profile.py <tag> <launches.csv> [<kernel>=<prof.ncu-rep> ...]
"""
import collections, csv, json, subprocess, sys
from pathlib import Path

root = Path(__file__).resolve().parent.parent
tag, launches = sys.argv[1], sys.argv[2]
reps = dict(a.split("=", 1) for a in sys.argv[3:])
out = [f"# ncu summary {tag}\n"]

rows = list(csv.reader(open(launches)))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
H = rows[hdr]; ki, vi = H.index("Kernel Name"), H.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) <= vi:
        continue
    name = r[ki].split("(")[0].replace("unnamed>::", "")
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += float(r[vi].replace(",", ""))
tot = sum(v[1] for v in agg.values())
out.append("## launch list (`ncu --metrics gpu__time_duration.sum --clock-control none`; cold-cache, serialised: compare shares)\n")
out.append("| kernel | launches | total ms | avg ms | share |\n|---|---|---|---|---|")
for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
    out.append(f"| {k} | {v[0]} | {v[1] / 1e6:.3f} | {v[1] / 1e6 / v[0]:.4f} | {v[1] / tot * 100:.1f}% |")
out.append("")

want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_shared_mem",
        "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warp_latency_issue_stalled_barrier.pct", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
        "lts__t_bytes.sum", "l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_bytes_pipe_lsu_mem_global_op_st.sum"]
traffic = {}
for kern, rep in reps.items():
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(txt.splitlines()))
    r = [x for x in r if x and x[0] in ("ID",) or (x and x[0].isdigit()) or (x and x[0] == "")]
    Hh = r[0]; units = r[1]; vals = r[2]
    out.append(f"## {kern} (`ncu --set full --clock-control none`, one launch: {vals[Hh.index('Kernel Name')][:60]}, grid {vals[Hh.index('Grid Size')]}, block {vals[Hh.index('Block Size')]})\n")
    out.append("| metric | value | unit |\n|---|---|---|")
    got = {}
    for w in want:
        if w in Hh:
            i = Hh.index(w); got[w] = vals[i]
            out.append(f"| {w} | {vals[i]} | {units[i]} |")
    out.append("")
    try:
        def tobytes(name):
            i = Hh.index(name); v = float(vals[i].replace(",", "")); u = units[i].lower()
            return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[u]
        traffic[kern] = int(tobytes("dram__bytes_read.sum") + tobytes("dram__bytes_write.sum"))
    except Exception as e:
        print("traffic parse failed", e)
(root / "profiles" / f"{tag}.md").write_text("\n".join(out) + "\n")
if traffic:
    tp = root / "profiles" / "traffic.json"
    cur = json.loads(tp.read_text()) if tp.exists() else {}
    cur.update(traffic); cur["_source"] = f"dram__bytes_read.sum + dram__bytes_write.sum per launch, {tag}"
    tp.write_text(json.dumps(cur, indent=1) + "\n")
print((root / "profiles" / f"{tag}.md").read_text())
