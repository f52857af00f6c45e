"""Top SASS instructions by stall samples (with neighbours) of one kernel in an .ncu-rep.  usage: ncu_hot.py rep regex [N]"""
import csv, io, subprocess, sys
rep, kre = sys.argv[1], sys.argv[2]
N = int(sys.argv[3]) if len(sys.argv) > 3 else 30
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kre}", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr = None; ins = []
for r in rows:
    if r and "Instructions Executed" in r:
        hdr = {h: i for i, h in enumerate(r)}; continue
    if hdr is None or len(r) < len(hdr): continue
    try:
        ex = int(r[hdr["Instructions Executed"]] or 0); sm = int(r[hdr["# Samples"]] or 0)
    except (ValueError, KeyError):
        continue
    ins.append((r[hdr["Address"]] if "Address" in hdr else "", r[hdr["Source"]].strip(), ex, sm))
tot = sum(i[3] for i in ins)
order = sorted(range(len(ins)), key=lambda i: -ins[i][3])[:N]
for i in sorted(order):
    a, s, ex, sm = ins[i]
    print(f"{i:5d} {100*sm/max(tot,1):5.2f}% smp {ex:10d} ex  {s[:100]}")
