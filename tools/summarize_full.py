"""Per-kernel summary of a multi-kernel `ncu --set full` report + profiles/traffic.json.
   python tools/summarize_full.py <tag> <report.ncu-rep> "<command line that was profiled>" """
import csv, json, subprocess, sys
from pathlib import Path
root = Path(__file__).resolve().parent.parent
tag, rep, cmd = sys.argv[1], sys.argv[2], sys.argv[3]
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
H, U = rows[0], rows[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_shared_mem",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"] + \
       [f"smsp__average_warps_issue_stalled_{k}_per_issue_active.ratio" for k in
        ("barrier", "long_scoreboard", "short_scoreboard", "wait", "mio_throttle", "lg_throttle", "branch_resolving", "math_pipe_throttle", "not_selected", "no_instruction")]
short = {"k_lz": "lz", "k_stored": "stored", "k_candidates": "cand", "k_info": "info", "k_parse": "parse", "k_huffman": "huff", "k_huffman_lanes": "huff", "k_emit": "emit",
         "k_emit2": "emit", "k_checksums": "cksum", "k_offsets": "offs", "k_fixed": "fixed", "k_gather": "gather"}
out = [f"# ncu --set full, {tag}\n", f"Command: `{cmd}` (after the same command exited 0 without ncu).\n"]
traffic = {}

def tob(name, r):
    i = H.index(name); v = float(r[i].replace(",", "")); u = U[i].lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[u]

for r in rows[2:]:
    if not r or not r[0].isdigit():
        continue
    k = r[H.index("Kernel Name")].replace("unnamed>::", "").split("(")[0].replace("void ", "").split("<")[0].strip()
    out.append(f"## {k}  grid {r[H.index('Grid Size')]} block {r[H.index('Block Size')]}\n")
    out.append("| metric | value | unit |\n|---|---|---|")
    for w in want:
        if w in H:
            out.append(f"| {w} | {r[H.index(w)]} | {U[H.index(w)]} |")
    t = int(tob("dram__bytes_read.sum", r) + tob("dram__bytes_write.sum", r))
    out.append(f"| dram read+write per launch | {t} | byte |\n")
    traffic[short.get(k, k)] = t
(root / "profiles" / f"{tag}_full.md").write_text("\n".join(out) + "\n")
traffic["_source"] = f"dram__bytes_read.sum + dram__bytes_write.sum per launch, profiles/{tag}_full.md"
(root / "profiles" / "traffic.json").write_text(json.dumps(traffic, indent=1) + "\n")
print(json.dumps(traffic))
