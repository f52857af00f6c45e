#!/bin/bash
# Round-end measurement pass on ONE B200 (run under gpurun): tests, smoke, every bench line kept under profiles/, the ncu
# launch list and one --set full capture of the two ROIAlign kernels.  Outputs go to gpurun_out/ (scratch); the judged
# copies are made afterwards (profiles/README.md).
set -u
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3 > $O/r2_end_pytest.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/r2_end_smoke.txt 2>&1
timeout 900 python bench.py > $O/r2_end_n1.json 2> $O/r2_end_n1.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $O/r2_end_ref.json 2> $O/r2_end_ref.err
for c in cfg3 cfg4 cfg5; do timeout 900 python bench.py --config $c > $O/r2_end_$c.json 2> $O/r2_end_$c.err; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r2_end_launches.csv python tools/prof_step.py --steps 3 > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"roi_align_bwd_clr_kernel|roi_align_fwd_nhwc_kernel" -s 2 -c 2 -f -o $O/r2_end_roi python tools/prof_step.py --steps 2 > $O/r2_end_ncu.log 2>&1
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_roi_align.py -m gpu -q -x -k "more_rois or persistent or dense_tile or golden_ramp" > $O/r2_end_memcheck.txt 2>&1; echo "memcheck rc=$?" >> $O/r2_end_memcheck.txt
tail -2 $O/r2_end_pytest.txt; cat $O/r2_end_smoke.txt | tail -1; tail -3 $O/r2_end_memcheck.txt
python - <<'PY'
import json
for n in ["n1", "ref", "cfg3", "cfg4", "cfg5"]:
    try:
        d = json.loads(open(f"gpurun_out/r2_end_{n}.json").read().strip().splitlines()[-1])
        print(n, d.get("value"), d.get("unit"), d.get("ms_per_step"), (d.get("roofline") or {}).get("frac"), (d.get("e2e") or {}).get("value"))
    except Exception as e:
        print(n, "ERR", e)
PY
