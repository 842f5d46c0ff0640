#!/bin/bash
# Compare launch shapes / item orders over N (tuning hooks of gsn_api.cu).  Run on the GPU box.
for cfg in "narrow cols 2" "wide cols 2" "wide rows 2" "wide rows 4"; do
  set -- $cfg
  echo "== shape=$1 order=$2 group=$3"
  GSN_FORCE_SHAPE=$1 GSN_ITEM_ORDER=$2 GSN_ITEM_GROUP=$3 python tools/sweep.py ${SIZES:-256 320 352 396 416 448 512 768 1024} 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try:
        d=json.loads(l); print(d['N'], round(d['evals_per_s']), round(d['frac'],3))
    except Exception: pass"
done
