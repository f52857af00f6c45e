"""Warp instructions executed per SASS opcode (and the top individual instructions) of one kernel in an .ncu-rep.
usage: python tools/ncu_opcodes.py rep kernel_regex [N]"""
import collections, csv, io, subprocess, sys
rep, kre = sys.argv[1], sys.argv[2]
N = int(sys.argv[3]) if len(sys.argv) > 3 else 25
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kre}", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr = None
ops = collections.Counter(); stall = collections.Counter(); tot = 0; samples = 0
for r in rows:
    if r and r[0] in ("Address", "#") or (r and "Instructions Executed" in r):
        hdr = {h: i for i, h in enumerate(r)}
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    try:
        ex = int(r[hdr["Instructions Executed"]] or 0)
        sm = int(r[hdr.get("# Samples", hdr.get("Warp Stall Sampling (All Samples)", 0))] or 0)
    except (ValueError, KeyError):
        continue
    src = r[hdr["Source"]].strip()
    parts = src.split()
    op = parts[1] if parts and parts[0].startswith("@") and len(parts) > 1 else (parts[0] if parts else "?")
    op = op.split(".")[0]
    ops[op] += ex; stall[op] += sm; tot += ex; samples += sm
print(f"total {tot} warp instructions, {samples} samples")
for op, c in ops.most_common(N):
    print(f"{op:10s} {c:12d} {100*c/max(tot,1):5.1f}% inst  {100*stall[op]/max(samples,1):5.1f}% samples")
