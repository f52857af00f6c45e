"""Warp instructions executed and stall samples aggregated per CUDA source line
(needs -lineinfo + --import-source on).   usage: python tools/ncu_lines.py rep kernel_regex [N]"""
import csv, io, subprocess, sys, collections
rep, kre = sys.argv[1], sys.argv[2]
N = int(sys.argv[3]) if len(sys.argv) > 3 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kre}", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
agg = collections.OrderedDict()
fname, hdr = "?", None
for r in rows:
    if not r:
        continue
    if r[0] == "File Name":
        fname = r[1].split("/")[-1]
        continue
    if r[0] == "Line No" and len(r) > 5:
        hdr = {h: i for i, h in enumerate(r) if h not in ("Source",)}
        src_col = 1
        continue
    if hdr is None or len(r) < 8:
        continue
    try:
        ex = int(r[hdr["Instructions Executed"]] or 0)
        smp = int(r[hdr["# Samples"]] or 0)
    except (ValueError, KeyError):
        continue
    key = (fname, r[0])
    a = agg.setdefault(key, [0, 0, r[src_col]])
    a[0] += ex
    a[1] += smp
tot_ex = sum(a[0] for a in agg.values()); tot_s = sum(a[1] for a in agg.values())
print(f"total warp instructions {tot_ex}, samples {tot_s}")
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:N]:
    print(f"{100*a[0]/max(tot_ex,1):5.1f}% inst {100*a[1]/max(tot_s,1):5.1f}% smp  {f}:{ln:>4s}  {a[2].strip()[:110]}")
