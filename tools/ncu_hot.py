"""Top stall lines of a kernel from an .ncu-rep source page (needs -lineinfo + --import-source on).
usage: python tools/ncu_hot.py rep kernel_regex [N]"""
import csv, io, subprocess, sys
rep, kre = sys.argv[1], sys.argv[2]
N = int(sys.argv[3]) if len(sys.argv) > 3 else 25
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kre}", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ci = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[hi + 1:] if len(r) == len(hdr) and r[0] != "Address"]
tot = sum(int(r[ci["# Samples"]] or 0) for r in body)
inst = sum(int(r[ci["Instructions Executed"]] or 0) for r in body)
print(f"total samples {tot}, warp instructions {inst}")
body.sort(key=lambda r: -int(r[ci["# Samples"]] or 0))
for r in body[:N]:
    s = int(r[ci["# Samples"]] or 0)
    st = {k: int(r[ci[k]] or 0) for k in ("stall_long_sb", "stall_short_sb", "stall_wait", "stall_barrier", "stall_mio", "stall_math", "stall_not_selected", "stall_lg", "stall_branch_resolving")}
    top = sorted(st.items(), key=lambda kv: -kv[1])[:2]
    print(f"{100*s/tot:5.1f}%  ex={int(r[ci['Instructions Executed']] or 0):>10d}  {r[ci['Source']][:90]:90s} {top}")
