"""Turn an `ncu --metrics gpu__time_duration.sum --csv` log into the markdown launch list kept under profiles/.

usage: python tools/launch_list.py gpurun_out/launches.csv STEPS > profiles/rN_launch_list.md
Only the last of the STEPS profiled steps is listed (setup kernels and earlier steps are dropped by counting
rpn_select_decode launches).
"""
import csv, sys

def main():
    path, steps = sys.argv[1], int(sys.argv[2])
    rows = []
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        if r["Metric Name"] == "gpu__time_duration.sum":
            v = float(r["Metric Value"].replace(",", ""))
            unit = r["Metric Unit"]
            us = v / 1000.0 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1000.0)
            rows.append((r["Kernel Name"], us))
    starts = [i for i, (k, _) in enumerate(rows) if "rpn_select_decode_kernel" in k]
    last = rows[starts[-1]:] if starts else rows
    tot = sum(u for _, u in last)
    print("| kernel | us | share |\n|---|---:|---:|")
    for k, u in last:
        print(f"| `{k[:100]}` | {u:.1f} | {100*u/tot:.1f}% |")
    print(f"| **total** | {tot:.1f} | 100% |")

if __name__ == "__main__":
    main()
