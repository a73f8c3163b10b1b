"""Hottest CUDA source lines of one kernel from an .ncu-rep (needs -lineinfo + --import-source on).
   python tools/hot_lines.py <rep> <kernel-regex> [topN]"""
import csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name", f"regex:{kern}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
lines = []
H = None; fname = ""
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]; continue
    if r[0] == "Line No":
        H = r; continue
    if H and r[0].isdigit():
        d = dict(zip(H[4:], r[4:]))
        try:
            int(d["# Samples"]); int(d["Instructions Executed"])
        except (ValueError, KeyError):
            continue
        lines.append((fname, int(r[0]), r[1], d))
tot_s = sum(int(l[3]["# Samples"]) for l in lines) or 1
tot_i = sum(int(l[3]["Instructions Executed"]) for l in lines) or 1
print(f"kernel {kern}: samples {tot_s}, warp instructions {tot_i}")
stalls = ["stall_barrier", "stall_long_sb", "stall_short_sb", "stall_wait", "stall_mio", "stall_math", "stall_branch_resolving", "stall_not_selected", "stall_selected", "stall_lg", "stall_no_inst", "stall_dispatch"]
agg = {s: sum(int(l[3].get(s, 0) or 0) for l in lines) for s in stalls}
print("stall samples:", {k.replace("stall_", ""): f"{v * 100 / tot_s:.1f}%" for k, v in sorted(agg.items(), key=lambda x: -x[1]) if v})
for f, ln, src, d in sorted(lines, key=lambda l: -int(l[3]["# Samples"]))[:top]:
    s = int(d["# Samples"]); i = int(d["Instructions Executed"])
    st = max(stalls, key=lambda k: int(d.get(k, 0) or 0))
    print(f"{s * 100 / tot_s:5.1f}% smp {i * 100 / tot_i:5.1f}% inst  thr/inst {d['Avg. Threads Executed']:>5} {st.replace('stall_', ''):<10} {f}:{ln:<5} {src.strip()[:105]}")

if len(sys.argv) > 4:      # extra: instruction share by source-line ranges "name:lo-hi,..."
    print("\ninstruction / sample share by range")
    for spec in sys.argv[4].split(","):
        name, rng = spec.split(":"); lo, hi = map(int, rng.split("-"))
        i = sum(int(l[3]["Instructions Executed"]) for l in lines if l[0].startswith("zz_kernels") and lo <= l[1] <= hi)
        sm = sum(int(l[3]["# Samples"]) for l in lines if l[0].startswith("zz_kernels") and lo <= l[1] <= hi)
        print(f"  {name:<12} {i * 100 / tot_i:5.1f}% inst  {sm * 100 / tot_s:5.1f}% samples")
    i = sum(int(l[3]["Instructions Executed"]) for l in lines if not l[0].startswith("zz_kernels"))
    print(f"  {'headers':<12} {i * 100 / tot_i:5.1f}% inst")
