#!/bin/bash
# normalize tuning: ops bench with several (elements per inversion, thread target) settings
O=gpurun_out; mkdir -p $O
for cfg in "8 37888" "4 75776" "6 75776" "8 75776" "4 151552" "8 151552"; do
  set -- $cfg
  BGN_NORM_PER_THREAD=$1 BGN_NORM_THREADS=$2 timeout 300 python tools/opsbench.py --reps 2 > $O/norm_$1_$2.json 2>$O/norm_$1_$2.err
  python - <<PY
import json
d=json.load(open("$O/norm_$1_$2.json"))
print("per=$1 threads=$2 " + "  ".join("%s %.3f/%.3f" % (k, d["ops"][k]["ms"], d["ops"][k]["kernel_ms"].get("k_normalize",0)) for k in ("encrypt","eadd_l1","multconst_l1_16bit","blind_l1","multconstpoly_l1")))
PY
done
