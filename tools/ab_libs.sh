#!/bin/bash
# A/B of two BUILDS of the library on the same box, interleaved: tools/ab_libs.sh <other.so> [rounds] [ab.py args...]
# (the loader honours OSR_LIB_PATH).  Prints forward / backward / step milliseconds per run.
OTHER=$1; ROUNDS=${2:-3}; shift; shift
for r in $(seq 1 $ROUNDS); do
  for lib in new other; do
    if [ $lib = other ]; then export OSR_LIB_PATH=$PWD/$OTHER; else unset OSR_LIB_PATH; fi
    python tools/ab.py "$@" 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print('$lib', d['setting'], round(d['s3_roialign_fwd'], 4), round(d['s3_roialign_bwd'], 4), round(d['step'], 4))
"
  done
done
